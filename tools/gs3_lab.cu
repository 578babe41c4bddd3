// grid_sample_3d laboratory (GPU box, native): experiments that decide what the product kernel looks like, kept out of
// the library.  Compiles the product source (emoportraits_b200/csrc/grid_sample.cu) into this binary and adds
//   * the memory-system denominators the kernel is measured against: DRAM read stream, L2-resident read stream (LDG.128
//     from a 48 MB buffer: what the L2 -> SM fabric delivers), both in TB/s;
//   * variants of the balanced kernel's gather loop: L2 eviction-priority hints (volume evict_last, output evict_first),
//     one warp per voxel row (3 aligned lines per load instruction instead of 4-5), two items in flight per thread;
// every variant is checked bit for bit against the product kernel and timed like bench.py times it (L2 flushed before
// every launch, CUDA events, median of 15).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gs3_lab tools/gs3_lab.cu
#include <stdarg.h>

#include <algorithm>
#include <vector>

#include "../emoportraits_b200/csrc/grid_sample.cu"

// ---------------------------------------------------------------------------------------------------------------
// Kernels that were product code until the processing-order study (sweep_main below) replaced them by the brick kernel with
// flat bricks: the two phases of the brick kernel as device functions, the balanced persistent kernel (one contiguous share
// of the brick-ordered enumeration per CTA) and the sweep kernel (persistent, the enumeration cut into waves).  Kept here so
// that the study stays reproducible.
// ---------------------------------------------------------------------------------------------------------------
namespace emo {
struct LabParams : GS3Params {
  int rounds;                // sweep kernel: number of consecutive waves the enumeration is cut into
  int hint_in;               // 1: volume loads carry L2::evict_last, 0: evict_normal
  long long prefetch_bytes;  // > 0: the CTAs first prefetch this many bytes of `in` into L2 (bulk prefetch, one slice per CTA)
};
// The two device functions below restate the brick kernel's two phases for the balanced variant further down (the
// brick kernel keeps its own inline copy: routing it through these functions changed its register allocation, 40 -> 44,
// i.e. 6 -> 5 resident CTAs per SM).
// phase 1 for one output voxel: sample position -> eight clamped corner offsets (in float4 units) and trilinear
// weights (0 for corners outside the volume: zeros padding) -> shared memory slot `vox`
__device__ __forceinline__ void gs3_setup_voxel(const GS3Params& p, int n, int od, int oh, int ow, int vox,
                                                int (*s_off)[8], float (*s_wgt)[8], long long* s_out) {
  const int c4n = p.C >> 2;
  long long o = -1;
  if (ow < p.Wout && oh < p.Hout && od < p.Dout) {
    float gx, gy, gz;
    sample_coord(p, n, od, oh, ow, gx, gy, gz);
    const Corner8 k = corners(p, gx, gy, gz);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
      const int x = k.x0 + dx, y = k.y0 + dy, z = k.z0 + dz;
      const bool ok = (unsigned)x < (unsigned)p.Win && (unsigned)y < (unsigned)p.Hin && (unsigned)z < (unsigned)p.Din;
      const int xc = min(max(x, 0), p.Win - 1), yc = min(max(y, 0), p.Hin - 1), zc = min(max(z, 0), p.Din - 1);
      s_wgt[vox][j] = ok ? (dx ? k.fx : 1.f - k.fx) * (dy ? k.fy : 1.f - k.fy) * (dz ? k.fz : 1.f - k.fz) : 0.f;
      s_off[vox][j] = ((zc * p.Hin + yc) * p.Win + xc) * c4n;
    }
    o = (long long)n * p.os_n + (long long)od * p.os_d + (long long)oh * p.os_h + (long long)ow * p.os_w;
  }
  s_out[vox] = o;
}

// phase 2: the CTA's threads sweep (voxel, float4-of-channels) items of `nvox` voxels set up in shared memory
template <bool SPLIT>
__device__ __forceinline__ void gs3_gather_items(const GS3Params& p, int n, int nvox, const int (*s_off)[8],
                                                 const float (*s_wgt)[8], const long long* s_out, uint64_t pol_in, uint64_t pol_out) {
  const int c4n = p.C >> 2;
  const int work = nvox * c4n;
  const float4* in4 = (const float4*)p.in + (long long)n * p.Din * p.Hin * p.Win * c4n;
  for (int t = threadIdx.x; t < work; t += blockDim.x) {
    const int vox = t / c4n, c4 = t - vox * c4n;
    const long long ob = s_out[vox];
    if (ob < 0) continue;
    const int4 o0 = *(const int4*)&s_off[vox][0], o1 = *(const int4*)&s_off[vox][4];
    const float4 w0 = *(const float4*)&s_wgt[vox][0], w1 = *(const float4*)&s_wgt[vox][4];
    const float4* base = in4 + c4;
    const float4 v0 = ldg_hint(base + o0.x, pol_in), v1 = ldg_hint(base + o0.y, pol_in), v2 = ldg_hint(base + o0.z, pol_in), v3 = ldg_hint(base + o0.w, pol_in);
    const float4 v4 = ldg_hint(base + o1.x, pol_in), v5 = ldg_hint(base + o1.y, pol_in), v6 = ldg_hint(base + o1.z, pol_in), v7 = ldg_hint(base + o1.w, pol_in);
    float4 acc;
#define EMO_GS_ACC(f) \
  acc.f = fmaf(v7.f, w1.w, fmaf(v6.f, w1.z, fmaf(v5.f, w1.y, fmaf(v4.f, w1.x, \
          fmaf(v3.f, w0.w, fmaf(v2.f, w0.z, fmaf(v1.f, w0.y, v0.f * w0.x)))))));
    EMO_GS_ACC(x) EMO_GS_ACC(y) EMO_GS_ACC(z) EMO_GS_ACC(w)
#undef EMO_GS_ACC
    const long long o = ob + (long long)(c4 * 4) * p.os_c;
    if (p.os_c == 1) {
      if (p.out) stg_hint((float4*)(p.out + o), acc, pol_out);  // the output is not re-read by this kernel
      if (SPLIT) {
        uint2 hi, lo, lo2;
        if (p.out_lo2) {
          split4x3(acc, hi, lo, lo2);
          *(uint2*)(p.out_lo2 + o) = lo2;
        } else {
          split4(acc, hi, lo);
        }
        *(uint2*)(p.out_hi + o) = hi;
        *(uint2*)(p.out_lo + o) = lo;
      }
    } else {
      const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (p.out) p.out[o + j * p.os_c] = a[j];
        if (SPLIT) {
          __nv_bfloat16 h, l, l2;
          split_bf16x3(a[j], h, l, l2);
          if (!p.out_lo2) split_bf16(a[j], h, l);
          p.out_hi[o + j * p.os_c] = h;
          p.out_lo[o + j * p.os_c] = l;
          if (p.out_lo2) p.out_lo2[o + j * p.os_c] = l2;
        }
      }
    }
  }
}


// Balanced variant (the product's choice for single-sample 64^3 lattices until the brick-order study below replaced it;
// profiles/gs3_check_r1.txt).  The brick kernel above launches one CTA
// per 256-voxel brick: 1024 CTAs for a 64^3 lattice against 148 SMs x 6 resident CTAs = 888 slots, so 136 bricks run
// in a second, nearly empty wave whose lone CTA per SM is latency-bound (24 dependent gather rounds).  Here the grid
// is exactly (SMs x resident CTAs) and every CTA takes an equal contiguous share of the brick-ordered voxel
// enumeration, processed in chunks of <= 256 voxels; a chunk never straddles two samples.  Same per-voxel and per-item
// arithmetic as the brick kernel (shared device functions), so the outputs are bit-identical.
template <bool SPLIT>
__global__ void __launch_bounds__(256, 6) gs3_cl_balanced_kernel(const LabParams p) {
  __shared__ __align__(16) int s_off[kBrickVox][8];
  __shared__ __align__(16) float s_wgt[kBrickVox][8];
  __shared__ long long s_out[kBrickVox];
  const int brick_vox = p.bw * p.bh * p.bd;
  const int per_sample = p.bricks_w * p.bricks_h * p.bricks_d * brick_vox;  // padded lattice; N * per_sample < 2^31 (host)
  const long long total = (long long)per_sample * p.N;
  const int v_begin = (int)(total * blockIdx.x / gridDim.x), v_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  for (int base = v_begin; base < v_end;) {
    const int n = base / per_sample;
    const int stop = min(min(base + kBrickVox, v_end), (n + 1) * per_sample);
    const int nvox = stop - base;
    if ((int)threadIdx.x < nvox) {
      const int v = base + (int)threadIdx.x - n * per_sample;
      int b = v / brick_vox;
      const int l = v - b * brick_vox;
      const int bwi = b % p.bricks_w; b /= p.bricks_w;
      const int bhi = b % p.bricks_h; b /= p.bricks_h;
      const int bdi = b;
      const int lw = l % p.bw, lh = (l / p.bw) % p.bh, ld = l / (p.bw * p.bh);
      gs3_setup_voxel(p, n, bdi * p.bd + ld, bhi * p.bh + lh, bwi * p.bw + lw, threadIdx.x, s_off, s_wgt, s_out);
    }
    __syncthreads();
    gs3_gather_items<SPLIT>(p, n, nvox, s_off, s_wgt, s_out, l2_policy_evict_last(), l2_policy_evict_first());
    __syncthreads();  // the next chunk overwrites the shared-memory slots
    base = stop;
  }
}

// Sweep variant: persistent CTAs like the balanced kernel, but the voxel enumeration is cut into `rounds` consecutive waves
// and every wave is shared equally by all CTAs, so that at any moment the whole grid works inside ONE thin slab of the
// lattice.  Why: with one contiguous share per CTA (balanced kernel) or one brick per CTA (brick kernel, 1024 bricks
// against 888 resident CTAs) the whole 64^3 lattice is in flight at once; a warp-field tensor with sigma = 3 voxels of
// jitter re-uses every input line from ~8 output voxels spread over +-10 slices, i.e. from other SMs at other times, and
// the 100 MB volume does not survive in L2 between those uses: ncu counts 256 MB of DRAM reads for a 100 MB volume
// (profiles/prof_gs3_r2.txt).  A slab of (wave + jitter) slices is 30-45 MB and stays resident.  Chunks are staged
// through double-buffered shared-memory slots (one __syncthreads per chunk); per-voxel and per-item arithmetic is the
// brick kernel's (shared device functions): bit-identical outputs.  With prefetch_bytes > 0 (volumes that fit L2) each
// CTA first issues one bulk L2 prefetch of its slice of the volume, so that the cold DRAM fetch of the volume streams at
// full rate beside the first gathers instead of trickling in miss by miss.
static constexpr int kSweepVox = 64;

template <bool SPLIT>
__global__ void __launch_bounds__(256, 6) gs3_cl_sweep_kernel(const LabParams p) {
  __shared__ __align__(16) int s_off[2][kSweepVox][8];
  __shared__ __align__(16) float s_wgt[2][kSweepVox][8];
  __shared__ long long s_out[2][kSweepVox];
  if (p.prefetch_bytes > 0 && threadIdx.x == 0) {
    const long long per = ((p.prefetch_bytes / gridDim.x) + 15) & ~15ll;
    const long long off = per * blockIdx.x;
    if (off < p.prefetch_bytes) {
      const unsigned size = (unsigned)min(per, p.prefetch_bytes - off);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)p.in + off), "r"(size) : "memory");
    }
  }
  const int brick_vox = p.bw * p.bh * p.bd;
  const int per_sample = p.bricks_w * p.bricks_h * p.bricks_d * brick_vox;  // padded lattice; N * per_sample < 2^31 (host)
  const long long total = (long long)per_sample * p.N;
  const uint64_t pol_in = p.hint_in ? l2_policy_evict_last() : l2_policy_evict_normal(), pol_out = l2_policy_evict_first();
  int r = -1, b = 0, e = 0;  // current wave, remaining range of this CTA's share of it
  // next chunk: sample n, first voxel v0 (inside the sample's enumeration), nvox voxels; false when the CTA is done
  auto next = [&](int& n, int& v0, int& nvox) -> bool {
    while (b >= e) {
      if (++r >= p.rounds) return false;
      const long long lo = total * r / p.rounds, w = total * (r + 1) / p.rounds - lo;
      b = (int)(lo + w * blockIdx.x / gridDim.x);
      e = (int)(lo + w * (blockIdx.x + 1) / gridDim.x);
    }
    n = b / per_sample;
    const int stop = min(min(b + kSweepVox, e), (n + 1) * per_sample);
    v0 = b - n * per_sample; nvox = stop - b; b = stop;
    return true;
  };
  auto setup = [&](int n, int v0, int nvox, int buf) {
    if ((int)threadIdx.x < nvox) {
      const int v = v0 + (int)threadIdx.x;
      int bk = v / brick_vox;
      const int l = v - bk * brick_vox;
      const int bwi = bk % p.bricks_w; bk /= p.bricks_w;
      const int bhi = bk % p.bricks_h; bk /= p.bricks_h;
      const int lw = l % p.bw, lh = (l / p.bw) % p.bh, ld = l / (p.bw * p.bh);
      gs3_setup_voxel(p, n, bk * p.bd + ld, bhi * p.bh + lh, bwi * p.bw + lw, threadIdx.x, s_off[buf], s_wgt[buf], s_out[buf]);
    }
  };
  int n, v0, nvox, buf = 0;
  bool have = next(n, v0, nvox);
  if (have) setup(n, v0, nvox, 0);
  __syncthreads();
  while (have) {
    int n2 = 0, v2 = 0, nvox2 = 0;
    const bool have2 = next(n2, v2, nvox2);
    if (have2) setup(n2, v2, nvox2, buf ^ 1);  // its readers (previous iteration) are behind the barrier below
    gs3_gather_items<SPLIT>(p, n, nvox, s_off[buf], s_wgt[buf], s_out[buf], pol_in, pol_out);
    __syncthreads();
    n = n2; v0 = v2; nvox = nvox2; buf ^= 1; have = have2;
  }
}


}  // namespace emo

namespace emo {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return EMO_ERR_CUDA; }
  return EMO_OK;
}
}  // namespace emo

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

using namespace emo;

// ---------------------------------------------------------------------------------------------------------------
__global__ void read_stream_kernel(const float4* __restrict__ p, long long n4, int passes, float* sink) {
  float acc = 0.f;
  for (int k = 0; k < passes; ++k)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(p + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;
}
__global__ void flush_k(float4* buf, long long n4) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) buf[t] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void fill_uniform(float* p, long long n, unsigned seed, float lo, float hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = lo + (hi - lo) * (hash32((unsigned)i * 2654435761u + seed) >> 8) * (1.0f / 16777216.0f);
}
// identity lattice + gaussian-ish jitter (sum of 4 uniforms, sigma = `sigma`), as bench.py's 0.1 * randn
__global__ void fill_grid(float* g, int N, int D, int H, int W, float sigma, unsigned seed) {
  const long long total = (long long)N * D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const float base[3] = {-1.f + 2.f * w / (W - 1), -1.f + 2.f * h / (H - 1), D > 1 ? -1.f + 2.f * d / (D - 1) : 0.f};
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int q = 0; q < 4; ++q) s += (hash32((unsigned)(i * 12 + k * 4 + q) + seed) >> 8) * (1.0f / 16777216.0f) - 0.5f;
      g[i * 3 + k] = base[k] + sigma * s * 1.7320508f;  // var of the sum = 4/12 -> x sqrt(3)
    }
  }
}
__global__ void count_diff(const unsigned* a, const unsigned* b, long long n, unsigned long long* cnt) {
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += a[i] != b[i];
  if (c) atomicAdd(cnt, c);
}

// ---------------------------------------------------------------------------------------------------------------
// gather-loop variants.  MODE bit 0: L2 hints, bit 1: two items in flight per thread, bit 2: one warp per voxel (24 lanes)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint(float4* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
#define GS_ACC(f) acc.f = fmaf(v7.f, w1.w, fmaf(v6.f, w1.z, fmaf(v5.f, w1.y, fmaf(v4.f, w1.x, fmaf(v3.f, w0.w, fmaf(v2.f, w0.z, fmaf(v1.f, w0.y, v0.f * w0.x)))))));

template <int MODE>
__device__ __forceinline__ void one_item(const GS3Params& p, const float4* in4, int vox, int c4, const int (*s_off)[8], const float (*s_wgt)[8],
                                         const long long* s_out, uint64_t pol_in, uint64_t pol_out) {
  const long long ob = s_out[vox];
  if (ob < 0) return;
  const int4 o0 = *(const int4*)&s_off[vox][0], o1 = *(const int4*)&s_off[vox][4];
  const float4 w0 = *(const float4*)&s_wgt[vox][0], w1 = *(const float4*)&s_wgt[vox][4];
  const float4* base = in4 + c4;
  float4 v0, v1, v2, v3, v4, v5, v6, v7;
  if (MODE & 1) {
    v0 = ld_hint(base + o0.x, pol_in); v1 = ld_hint(base + o0.y, pol_in); v2 = ld_hint(base + o0.z, pol_in); v3 = ld_hint(base + o0.w, pol_in);
    v4 = ld_hint(base + o1.x, pol_in); v5 = ld_hint(base + o1.y, pol_in); v6 = ld_hint(base + o1.z, pol_in); v7 = ld_hint(base + o1.w, pol_in);
  } else {
    v0 = __ldg(base + o0.x); v1 = __ldg(base + o0.y); v2 = __ldg(base + o0.z); v3 = __ldg(base + o0.w);
    v4 = __ldg(base + o1.x); v5 = __ldg(base + o1.y); v6 = __ldg(base + o1.z); v7 = __ldg(base + o1.w);
  }
  float4 acc;
  GS_ACC(x) GS_ACC(y) GS_ACC(z) GS_ACC(w)
  float4* dst = (float4*)(p.out + ob + (long long)(c4 * 4));
  if (MODE & 1) st_hint(dst, acc, pol_out);
  else __stcs(dst, acc);
}

template <int MODE>
__global__ void __launch_bounds__(256, 6) gs3_lab_kernel(const GS3Params p) {
  __shared__ __align__(16) int s_off[kBrickVox][8];
  __shared__ __align__(16) float s_wgt[kBrickVox][8];
  __shared__ long long s_out[kBrickVox];
  uint64_t pol_in = 0, pol_out = 0;
  if (MODE & 1) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_in));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_out));
  }
  const int c4n = p.C >> 2;
  const int brick_vox = p.bw * p.bh * p.bd;
  const int per_sample = p.bricks_w * p.bricks_h * p.bricks_d * brick_vox;
  const long long total = (long long)per_sample * p.N;
  const int v_begin = (int)(total * blockIdx.x / gridDim.x), v_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  for (int base = v_begin; base < v_end;) {
    const int n = base / per_sample;
    const int stop = min(min(base + kBrickVox, v_end), (n + 1) * per_sample);
    const int nvox = stop - base;
    if ((int)threadIdx.x < nvox) {
      const int v = base + (int)threadIdx.x - n * per_sample;
      int b = v / brick_vox;
      const int l = v - b * brick_vox;
      const int bwi = b % p.bricks_w; b /= p.bricks_w;
      const int bhi = b % p.bricks_h; b /= p.bricks_h;
      const int bdi = b;
      const int lw = l % p.bw, lh = (l / p.bw) % p.bh, ld = l / (p.bw * p.bh);
      gs3_setup_voxel(p, n, bdi * p.bd + ld, bhi * p.bh + lh, bwi * p.bw + lw, threadIdx.x, s_off, s_wgt, s_out);
    }
    __syncthreads();
    const float4* in4 = (const float4*)p.in + (long long)n * p.Din * p.Hin * p.Win * c4n;
    if (MODE & 4) {
      // one warp per voxel: lanes 0..c4n-1 each own one float4 of the voxel's channel row (three aligned 128-byte lines at C = 96)
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int vox = warp; vox < nvox; vox += 8)
        for (int c4 = lane; c4 < c4n; c4 += 32) one_item<MODE>(p, in4, vox, c4, s_off, s_wgt, s_out, pol_in, pol_out);
    } else if (MODE & 2) {
      const int work = nvox * c4n;
      for (int t = threadIdx.x; t < work; t += 2 * blockDim.x) {
        // both items' loads are issued before either item's arithmetic (the compiler keeps the two chains independent)
        const int t2 = t + blockDim.x;
        const int va = t / c4n, ca = t - va * c4n;
        if (t2 < work) {
          const int vb = t2 / c4n, cb = t2 - vb * c4n;
          const long long oa = s_out[va], obb = s_out[vb];
          if (oa >= 0 && obb >= 0) {
            const int4 a0 = *(const int4*)&s_off[va][0], a1 = *(const int4*)&s_off[va][4], b0 = *(const int4*)&s_off[vb][0], b1 = *(const int4*)&s_off[vb][4];
            const float4* pa = in4 + ca; const float4* pb = in4 + cb;
            const float4 x0 = __ldg(pa + a0.x), x1 = __ldg(pa + a0.y), x2 = __ldg(pa + a0.z), x3 = __ldg(pa + a0.w), x4 = __ldg(pa + a1.x), x5 = __ldg(pa + a1.y), x6 = __ldg(pa + a1.z), x7 = __ldg(pa + a1.w);
            const float4 y0 = __ldg(pb + b0.x), y1 = __ldg(pb + b0.y), y2 = __ldg(pb + b0.z), y3 = __ldg(pb + b0.w), y4 = __ldg(pb + b1.x), y5 = __ldg(pb + b1.y), y6 = __ldg(pb + b1.z), y7 = __ldg(pb + b1.w);
            {
              const float4 w0 = *(const float4*)&s_wgt[va][0], w1 = *(const float4*)&s_wgt[va][4];
              const float4 v0 = x0, v1 = x1, v2 = x2, v3 = x3, v4 = x4, v5 = x5, v6 = x6, v7 = x7;
              float4 acc; GS_ACC(x) GS_ACC(y) GS_ACC(z) GS_ACC(w)
              __stcs((float4*)(p.out + oa + (long long)(ca * 4)), acc);
            }
            {
              const float4 w0 = *(const float4*)&s_wgt[vb][0], w1 = *(const float4*)&s_wgt[vb][4];
              const float4 v0 = y0, v1 = y1, v2 = y2, v3 = y3, v4 = y4, v5 = y5, v6 = y6, v7 = y7;
              float4 acc; GS_ACC(x) GS_ACC(y) GS_ACC(z) GS_ACC(w)
              __stcs((float4*)(p.out + obb + (long long)(cb * 4)), acc);
            }
            continue;
          }
          one_item<MODE>(p, in4, vb, cb, s_off, s_wgt, s_out, pol_in, pol_out);
        }
        one_item<MODE>(p, in4, va, ca, s_off, s_wgt, s_out, pol_in, pol_out);
      }
    } else {
      const int work = nvox * c4n;
      for (int t = threadIdx.x; t < work; t += blockDim.x) {
        const int vox = t / c4n;
        one_item<MODE>(p, in4, vox, t - vox * c4n, s_off, s_wgt, s_out, pol_in, pol_out);
      }
    }
    __syncthreads();
    base = stop;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant (what BASELINE.json's north_star sketches: "TMA-staged feature tiles"), affine lattice only.  One CTA per
// 8 x 8 x 4 output brick: the brick's samples fall into a box of the volume (the bounding box of its eight corner samples,
// +1 for the trilinear footprint); the box is staged into shared memory with TMA, 16 channels at a time (double buffered,
// [z][y][x][16 floats]), zero-filled outside the volume (= zeros padding), and the eight corners of every sample are read
// with LDS.128.  The box extent is fixed by the tensor map: 14 x 14 x 6 voxels covers the 30-degree test lattice.
// L2 -> SM traffic = box / brick = 14*14*6 / 256 = 4.6 x the volume (the gather kernel moves 8 x (1 - L1 hit rate) = 3.2 x).
// ---------------------------------------------------------------------------------------------------------------
#include <cudaTypedefs.h>
static constexpr int SBX = 14, SBY = 14, SBZ = 6, SCH = 16;
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma5(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(s_u32(dst)), "l"(m), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(256, 1) gs3_staged_kernel(const __grid_constant__ CUtensorMap tmap, const GS3Params p) {
  extern __shared__ __align__(128) uint8_t sraw[];
  float* buf[2] = {(float*)sraw, (float*)(sraw + SBX * SBY * SBZ * SCH * 4)};
  __shared__ uint64_t bar[2];
  __shared__ int s_org[3];
  __shared__ int s_loc[256];          // corner (x0,y0,z0) relative to the box, packed 10 bits each
  __shared__ float s_frac[256][3];
  __shared__ int s_bad;
  int b = blockIdx.x;
  const int bwi = b % p.bricks_w; b /= p.bricks_w;
  const int bhi = b % p.bricks_h; b /= p.bricks_h;
  const int bdi = b % p.bricks_d; b /= p.bricks_d;
  const int n = b;
  const int tid = threadIdx.x;
  if (tid == 0) { mb_init(&bar[0], 1); mb_init(&bar[1], 1); s_bad = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // sample position of this thread's voxel
  const int lw = tid % 8, lh = (tid / 8) % 8, ld = tid / 64;
  const int ow = bwi * 8 + lw, oh = bhi * 8 + lh, od = bdi * 4 + ld;
  float gx, gy, gz;
  sample_coord(p, n, od, oh, ow, gx, gy, gz);
  const float ix = ((gx + 1.f) * (float)p.Win - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)p.Hin - 1.f) * 0.5f, iz = ((gz + 1.f) * (float)p.Din - 1.f) * 0.5f;
  const int x0 = (int)floorf(ix), y0 = (int)floorf(iy), z0 = (int)floorf(iz);
  // box origin = minimum corner over the brick (block reduction through shared memory atomics)
  if (tid == 0) { s_org[0] = 1 << 30; s_org[1] = 1 << 30; s_org[2] = 1 << 30; }
  __syncthreads();
  atomicMin(&s_org[0], x0); atomicMin(&s_org[1], y0); atomicMin(&s_org[2], z0);
  __syncthreads();
  const int bx = s_org[0], by = s_org[1], bz = s_org[2];
  const int rx = x0 - bx, ry = y0 - by, rz = z0 - bz;
  if (rx + 1 >= SBX || ry + 1 >= SBY || rz + 1 >= SBZ) s_bad = 1;  // the lattice does not fit the fixed box: results invalid
  s_loc[tid] = rx | (ry << 10) | (rz << 20);
  s_frac[tid][0] = ix - (float)x0; s_frac[tid][1] = iy - (float)y0; s_frac[tid][2] = iz - (float)z0;
  const int nsl = p.C / SCH;
  const uint32_t slice_bytes = SBX * SBY * SBZ * SCH * 4;
  if (tid == 0) {
    mb_expect(&bar[0], slice_bytes);
    tma5(&tmap, &bar[0], buf[0], 0, bx, by, bz, n);
  }
  __syncthreads();
  for (int s = 0; s < nsl; ++s) {
    if (tid == 0 && s + 1 < nsl) {  // next slice into the other buffer (its readers finished at the barrier below)
      mb_expect(&bar[(s + 1) & 1], slice_bytes);
      tma5(&tmap, &bar[(s + 1) & 1], buf[(s + 1) & 1], (s + 1) * SCH, bx, by, bz, n);
    }
    mb_wait(&bar[s & 1], (s >> 1) & 1);
    const float4* src = (const float4*)buf[s & 1];
    for (int t = tid; t < 256 * (SCH / 4); t += 256) {
      const int vox = t / (SCH / 4), q = t % (SCH / 4);
      const int loc = s_loc[vox];
      const int cx = loc & 1023, cy = (loc >> 10) & 1023, cz = loc >> 20;
      const float fx = s_frac[vox][0], fy = s_frac[vox][1], fz = s_frac[vox][2];
      const int base = ((cz * SBY + cy) * SBX + cx) * (SCH / 4) + q;
      const int dxo = SCH / 4, dyo = SBX * (SCH / 4), dzo = SBY * SBX * (SCH / 4);
      const float4 v0 = src[base], v1 = src[base + dxo], v2 = src[base + dyo], v3 = src[base + dyo + dxo];
      const float4 v4 = src[base + dzo], v5 = src[base + dzo + dxo], v6 = src[base + dzo + dyo], v7 = src[base + dzo + dyo + dxo];
      const float4 w0 = make_float4((1 - fx) * (1 - fy) * (1 - fz), fx * (1 - fy) * (1 - fz), (1 - fx) * fy * (1 - fz), fx * fy * (1 - fz));
      const float4 w1 = make_float4((1 - fx) * (1 - fy) * fz, fx * (1 - fy) * fz, (1 - fx) * fy * fz, fx * fy * fz);
      float4 acc;
      GS_ACC(x) GS_ACC(y) GS_ACC(z) GS_ACC(w)
      const int vw = vox % 8, vh = (vox / 8) % 8, vd = vox / 64;
      const long long o = (long long)n * p.os_n + (long long)(bdi * 4 + vd) * p.os_d + (long long)(bhi * 8 + vh) * p.os_h + (long long)(bwi * 8 + vw) * p.os_w + s * SCH + q * 4;
      __stcs((float4*)(p.out + o), acc);
    }
    __syncthreads();
  }
  if (tid == 0 && s_bad) p.out[0] = __int_as_float(0x7fc00000);
}

// ---------------------------------------------------------------------------------------------------------------
// Bulk-copy gather ("TMA-staged feature tiles" in gather form).  The LSU path of the product kernel pays one L1tex wavefront per
// 128-byte line and ~2 clk per additional line inside one LDG.128: ~50 clk per voxel and SM at C = 96, i.e. 46 us for a 64^3
// lattice however the loads are arranged.  Here the eight corner rows of a voxel (C x 4 bytes each, contiguous in a channels-last
// volume) are fetched by cp.async.bulk (1-D bulk copies, global -> shared memory, completion on an mbarrier): the async proxy
// does not go through the L1tex wavefront path.  One producer warp per CTA sets up BV voxels per stage (sample position ->
// clamped corner offsets + trilinear weights -> shared memory) and issues the 8 x BV row copies; BV x C/4 consumer threads
// (one float4 of one voxel each) wait for the stage, blend the eight rows from shared memory in the product kernel's FMA
// order (bit-identical output) and store; NST stages ring.  Chunks of BV voxels go round-robin over the persistent CTAs in
// raster order (x fastest ... z slowest): the grid sweeps the lattice as one thin slab (see the order study).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)), "l"(src), "r"(bytes),
               "r"(s_u32(bar)) : "memory");
}

template <int BV, int NST>
__global__ void __launch_bounds__(32 + BV * 24, 1) gs3_bulk_kernel(const GS3Params p) {
  extern __shared__ __align__(128) uint8_t sraw[];
  const int c4n = p.C >> 2;                       // 24
  const uint32_t row_bytes = (uint32_t)p.C * 4u;  // 384
  float* ring = (float*)sraw;                     // [NST][BV][8][C]
  __shared__ __align__(16) float s_wgt[NST][BV][8];
  __shared__ long long s_out[NST][BV];
  __shared__ int s_offb[BV][8];
  __shared__ uint64_t full_bar[NST], empty_bar[NST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = (BV * c4n + 31) / 32;  // consumer warps
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mb_init(&full_bar[i], 1); mb_init(&empty_bar[i], ncw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long per_sample = (long long)p.Dout * p.Hout * p.Wout;
  const long long total = per_sample * p.N;  // chunks never straddle samples when per_sample % BV == 0 (host checks)
  const long long nchunks = total / BV;
  const size_t stage_floats = (size_t)BV * 8 * p.C;
  if (warp == 0) {
    // ---------------- producer ----------------
    int it = 0;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    long long chunk = blockIdx.x;
    auto coords = [&](long long ch) {  // lanes < BV: sample position of voxel ch * BV + lane (issued one stage ahead of its use)
      if (lane < BV && ch < nchunks) {
        const long long v = ch * BV + lane;
        const int n = (int)(v / per_sample);
        long long r = v - (long long)n * per_sample;
        const int ow = (int)(r % p.Wout); r /= p.Wout;
        const int oh = (int)(r % p.Hout);
        const int od = (int)(r / p.Hout);
        sample_coord(p, n, od, oh, ow, gx, gy, gz);
      }
    };
    coords(chunk);
    for (; chunk < nchunks; chunk += gridDim.x, ++it) {
      const int slot = it % NST;
      const float cx = gx, cy = gy, cz = gz;
      coords(chunk + gridDim.x);  // next stage's grid loads fly during this stage's set-up
      if (it >= NST) mb_wait(&empty_bar[slot], ((it / NST) - 1) & 1);
      const long long v0 = chunk * BV;
      const int n = (int)(v0 / per_sample);
      if (lane < BV) {
        const long long v = v0 + lane;
        long long r = v - (long long)n * per_sample;
        const int ow = (int)(r % p.Wout); r /= p.Wout;
        const int oh = (int)(r % p.Hout);
        const int od = (int)(r / p.Hout);
        const Corner8 k = corners(p, cx, cy, cz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
          const int x = k.x0 + dx, y = k.y0 + dy, z = k.z0 + dz;
          const bool ok = (unsigned)x < (unsigned)p.Win && (unsigned)y < (unsigned)p.Hin && (unsigned)z < (unsigned)p.Din;
          const int xc = min(max(x, 0), p.Win - 1), yc = min(max(y, 0), p.Hin - 1), zc = min(max(z, 0), p.Din - 1);
          s_wgt[slot][lane][j] = ok ? (dx ? k.fx : 1.f - k.fx) * (dy ? k.fy : 1.f - k.fy) * (dz ? k.fz : 1.f - k.fz) : 0.f;
          s_offb[lane][j] = ((zc * p.Hin + yc) * p.Win + xc);
        }
        s_out[slot][lane] = (long long)n * p.os_n + (long long)od * p.os_d + (long long)oh * p.os_h + (long long)ow * p.os_w;
      }
      __syncwarp();
      if (lane == 0) mb_expect(&full_bar[slot], (uint32_t)(BV * 8) * row_bytes);
      __syncwarp();
      const char* vol = (const char*)(p.in + (long long)n * p.Din * p.Hin * p.Win * p.C);
      float* dst = ring + (size_t)slot * stage_floats;
      for (int q = lane; q < BV * 8; q += 32)
        bulk_g2s(dst + (size_t)q * p.C, vol + (long long)(&s_offb[0][0])[q] * row_bytes, row_bytes, &full_bar[slot]);
      __syncwarp();  // s_offb is rewritten by the next stage
    }
  } else if (warp - 1 < ncw) {
    // ---------------- consumers ----------------
    const int t = threadIdx.x - 32;
    const int vox = t / c4n, c4 = t - vox * c4n;
    const bool active = t < BV * c4n;
    uint64_t pol_out;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_out));
    int it = 0;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
      const int slot = it % NST;
      mb_wait(&full_bar[slot], (it / NST) & 1);
      if (active) {
        const float4 w0 = *(const float4*)&s_wgt[slot][vox][0], w1 = *(const float4*)&s_wgt[slot][vox][4];
        const float4* src = (const float4*)(ring + (size_t)slot * stage_floats + (size_t)vox * 8 * p.C) + c4;
        const float4 v0 = src[0 * c4n], v1 = src[1 * c4n], v2 = src[2 * c4n], v3 = src[3 * c4n];
        const float4 v4 = src[4 * c4n], v5 = src[5 * c4n], v6 = src[6 * c4n], v7 = src[7 * c4n];
        float4 acc;
        GS_ACC(x) GS_ACC(y) GS_ACC(z) GS_ACC(w)
        st_hint((float4*)(p.out + s_out[slot][vox] + (long long)(c4 * 4)), acc, pol_out);
      }
      __syncwarp();
      if (lane == 0) mb_arrive(&empty_bar[slot]);
    }
  }
}

struct Case { const char* name; int N, C, D, S; bool affine; };

static int legacy_main() {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const long long flush_bytes = 256ll << 20;
  float4* flush_buf; CK(cudaMalloc(&flush_buf, flush_bytes));
  float* sink; CK(cudaMalloc(&sink, 4));
  // ---- denominators ----
  {
    const long long big = 1ll << 30, small = 48ll << 20;
    float4* buf; CK(cudaMalloc(&buf, big));
    fill_uniform<<<592, 256, 0, st>>>((float*)buf, big / 4, 1u, 0.f, 1.f);
    for (int which = 0; which < 2; ++which) {
      const long long bytes = which ? small : big;
      const int passes = which ? 20 : 1;
      float best = 1e9f;
      for (int r = 0; r < 5; ++r) {
        if (which) read_stream_kernel<<<148 * 8, 256, 0, st>>>(buf, bytes / 16, 1, sink);  // warm L2
        CK(cudaEventRecord(e0, st));
        read_stream_kernel<<<148 * 8, 256, 0, st>>>(buf, bytes / 16, passes, sink);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        best = std::min(best, t);
      }
      printf("%s read stream (LDG.128, %lld MB x %d): %.2f TB/s\n", which ? "L2-resident" : "DRAM", bytes >> 20, passes, bytes * (double)passes / (best * 1e-3) * 1e-12);
    }
    CK(cudaFree(buf));
  }
  unsigned long long* d_cnt; CK(cudaMalloc(&d_cnt, 8));
  const Case cases[] = {{"d64_affine", 1, 96, 64, 64, true}, {"d64_grid", 1, 96, 64, 64, false}, {"d16_affine", 1, 96, 16, 64, true},
                        {"d16_grid", 1, 96, 16, 64, false}, {"d64_b8_grid", 8, 96, 64, 64, false}};
  int bad = 0;
  for (const Case& c : cases) {
    const long long vox = (long long)c.N * c.D * c.S * c.S, n = vox * c.C;
    float *in, *grid = nullptr, *theta = nullptr, *out[2];
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out[0], n * 4)); CK(cudaMalloc(&out[1], n * 4));
    fill_uniform<<<592, 256, 0, st>>>(in, n, 17u, -1.f, 1.f);
    if (c.affine) {
      std::vector<float> t(12 * c.N);
      for (int k = 0; k < c.N; ++k) {
        const float a = 0.5236f;
        const float m[12] = {cosf(a), -sinf(a), 0, 0.2f, sinf(a), cosf(a), 0, 0.2f, 0, 0, 1.f, 0.2f};
        for (int i = 0; i < 12; ++i) t[k * 12 + i] = m[i];
      }
      CK(cudaMalloc(&theta, t.size() * 4));
      CK(cudaMemcpy(theta, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    } else {
      CK(cudaMalloc(&grid, vox * 3 * 4));
      fill_grid<<<592, 256, 0, st>>>(grid, c.N, c.D, c.S, c.S, 0.1f, 99u);
    }
    LabParams p;
    memset(&p, 0, sizeof(p));
    p.in = in; p.grid = grid; p.theta = theta; p.N = c.N; p.C = c.C; p.Din = c.D; p.Hin = c.S; p.Win = c.S; p.Dout = c.D; p.Hout = c.S; p.Wout = c.S;
    p.os_c = 1; p.os_w = c.C; p.os_h = (long long)c.S * c.C; p.os_d = p.os_h * c.S; p.os_n = p.os_d * c.D;
    p.bw = 8; p.bh = 8; p.bd = 4; p.bricks_w = c.S / 8; p.bricks_h = c.S / 8; p.bricks_d = c.D / 4;
    const double bytes = ((double)2 * c.C + (c.affine ? 0 : 3)) * 4.0 * (double)vox;
    const char* names[] = {"product (balanced)", "L2 hints", "2 items in flight", "warp per voxel", "warp per voxel + hints"};
    for (int v = 0; v < 5; ++v) {
      const int slot = v == 0 ? 0 : 1;
      p.out = out[slot];
      std::vector<float> ms;
      for (int r = 0; r < 15; ++r) {
        flush_k<<<148 * 8, 256, 0, st>>>(flush_buf, flush_bytes / 16);
        CK(cudaEventRecord(e0, st));
        const unsigned ctas = 888;
        if (v == 0) gs3_cl_balanced_kernel<false><<<ctas, 256, 0, st>>>(p);
        else if (v == 1) gs3_lab_kernel<1><<<ctas, 256, 0, st>>>(p);
        else if (v == 2) gs3_lab_kernel<2><<<ctas, 256, 0, st>>>(p);
        else if (v == 3) gs3_lab_kernel<4><<<ctas, 256, 0, st>>>(p);
        else gs3_lab_kernel<5><<<ctas, 256, 0, st>>>(p);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        ms.push_back(t);
      }
      std::sort(ms.begin(), ms.end());
      unsigned long long diff = 0;
      if (v > 0) {
        CK(cudaMemsetAsync(d_cnt, 0, 8, st));
        count_diff<<<592, 256, 0, st>>>((const unsigned*)out[0], (const unsigned*)out[1], n, d_cnt);
        CK(cudaMemcpyAsync(&diff, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (diff) ++bad;
        CK(cudaMemsetAsync(out[1], 0xff, n * 4, st));
      }
      printf("%-12s %-24s median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic  diff_words %llu\n", c.name, names[v], ms[7] * 1e3, ms[0] * 1e3,
             bytes / (ms[7] * 1e-3) * 1e-9, diff);
      fflush(stdout);
    }
    if (c.affine) {
      // TMA-staged variant: tensor map over the channels-last volume, box {16 ch, 14, 14, 6, 1}
      PFN_cuTensorMapEncodeTiled encode = nullptr;
      cudaDriverEntryPointQueryResult qres;
      CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
      CUtensorMap tmap;
      cuuint64_t gdim[5] = {(cuuint64_t)c.C, (cuuint64_t)c.S, (cuuint64_t)c.S, (cuuint64_t)c.D, (cuuint64_t)c.N};
      cuuint64_t gstr[4] = {(cuuint64_t)c.C * 4, (cuuint64_t)c.S * c.C * 4, (cuuint64_t)c.S * c.S * c.C * 4, (cuuint64_t)c.D * c.S * c.S * c.C * 4};
      cuuint32_t box[5] = {SCH, SBX, SBY, SBZ, 1}, es[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, in, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { fprintf(stderr, "tensor map encode failed %d\n", (int)r); return 2; }
      const size_t smem = 2 * (size_t)SBX * SBY * SBZ * SCH * 4;
      CK(cudaFuncSetAttribute(gs3_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      p.out = out[1];
      std::vector<float> ms;
      const unsigned bricks = (unsigned)(c.N * (c.D / 4) * (c.S / 8) * (c.S / 8));
      for (int r2 = 0; r2 < 15; ++r2) {
        flush_k<<<148 * 8, 256, 0, st>>>(flush_buf, flush_bytes / 16);
        CK(cudaEventRecord(e0, st));
        gs3_staged_kernel<<<bricks, 256, smem, st>>>(tmap, p);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        ms.push_back(t);
      }
      std::sort(ms.begin(), ms.end());
      // same trilinear sample with another (exact) weight / summation order: compare by value
      std::vector<float> h0(4096), h1(4096);
      CK(cudaMemcpy(h0.data(), out[0] + n / 2, 4096 * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(h1.data(), out[1] + n / 2, 4096 * 4, cudaMemcpyDeviceToHost));
      float maxerr = 0.f, first = 0.f;
      CK(cudaMemcpy(&first, out[1], 4, cudaMemcpyDeviceToHost));
      for (int i = 0; i < 4096; ++i) maxerr = fmaxf(maxerr, fabsf(h0[i] - h1[i]));
      printf("%-12s %-24s median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic  max |diff| on a 4096-element sample %.2e%s\n", c.name,
             "TMA-staged box", ms[7] * 1e3, ms[0] * 1e3, bytes / (ms[7] * 1e-3) * 1e-9, maxerr, first != first ? "  (BOX TOO SMALL)" : "");
    }
    CK(cudaFree(in)); CK(cudaFree(out[0])); CK(cudaFree(out[1]));
    if (grid) CK(cudaFree(grid));
    if (theta) CK(cudaFree(theta));
  }
  printf(bad ? "FAIL: %d variants differ\n" : "OK: all variants bit-identical to the product kernel\n", bad);
  return bad ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Sweep study (round 2, second half): processing ORDER.  Variants: the brick kernel with its standard 8x8x4 bricks (reference
// output) and with small flat bricks in plain blockIdx order (the hardware's CTA dispatch is the sweep), the balanced kernel,
// and the sweep kernel over several enumerations (brick shapes), chunk targets (voxels per CTA and wave), L2 hint on/off,
// bulk L2 prefetch (volumes that fit L2).  Timed after a DIRTY flush (256 MB of zeros written: what bench.py does) and, for
// selected variants, after a CLEAN flush (the same buffer read back afterwards: L2 full of clean lines).
// ---------------------------------------------------------------------------------------------------------------
struct Variant {
  const char* kind;   // "brick", "balanced", "sweep"
  int bw, bh, bd;     // enumeration / brick shape
  int chunk;          // sweep: target voxels per CTA per wave (0: one wave)
  int hint, prefetch, clean;
  int threads;        // brick kernel: threads per CTA (0 = 256)
};

static int sweep_main(int quick) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const long long flush_bytes = 256ll << 20;
  float4* flush_buf; CK(cudaMalloc(&flush_buf, flush_bytes));
  float* sink; CK(cudaMalloc(&sink, 4));
  unsigned long long* d_cnt; CK(cudaMalloc(&d_cnt, 8));
  int sms = 0, occ = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gs3_cl_sweep_kernel<false>, 256, 0));
  printf("SMs %d, resident sweep CTAs per SM %d\n", sms, occ);
  const int slots = sms * occ;
  const Case cases[] = {{"d64_grid", 1, 96, 64, 64, false}, {"d64_affine", 1, 96, 64, 64, true}, {"d16_grid", 1, 96, 16, 64, false},
                        {"d16_affine", 1, 96, 16, 64, true},  {"d64_b8_grid", 8, 96, 64, 64, false}, {"d64_b8_affine", 8, 96, 64, 64, true}};
  std::vector<Variant> vars = {
      {"brick", 8, 8, 4, 0, 1, 0, 0, 0},     // reference output (round-1 brick)
      {"brick", 8, 8, 4, 0, 1, 0, 1, 0},
      {"balanced", 8, 8, 4, 0, 1, 0, 0, 0},
      {"balanced", 8, 8, 4, 0, 1, 0, 1, 0},
      {"brick", 8, 8, 2, 0, 1, 0, 0, 0},
      {"brick", 8, 8, 2, 0, 1, 0, 1, 0},
      {"brick", 8, 8, 2, 0, 1, 0, 0, 128},
      {"brick", 8, 8, 1, 0, 1, 0, 0, 0},
      {"brick", 8, 8, 1, 0, 1, 0, 1, 0},
      {"brick", 8, 8, 1, 0, 1, 0, 0, 128},
      {"brick", 8, 8, 1, 0, 1, 0, 0, 64},
      {"brick", 8, 4, 1, 0, 1, 0, 0, 0},
      {"brick", 8, 4, 1, 0, 1, 0, 0, 128},
      {"brick", 8, 4, 1, 0, 1, 0, 0, 64},
      {"brick", 16, 4, 1, 0, 1, 0, 0, 0},
      {"brick", 16, 8, 1, 0, 1, 0, 0, 0},
      {"brick", 16, 16, 1, 0, 1, 0, 0, 0},
      {"brick", 4, 4, 4, 0, 1, 0, 0, 0},
      {"bulk", 8, 4, 1, 0, 1, 0, 0, 0},      // bulk-copy gather: bw = voxels per stage, bh = ring stages, chunk = CTAs per SM (0 = as many as fit)
      {"bulk", 8, 4, 1, 0, 1, 0, 1, 0},
      {"bulk", 8, 4, 1, 1, 1, 0, 0, 0},
      {"bulk", 8, 3, 1, 0, 1, 0, 0, 0},
      {"bulk", 16, 2, 1, 0, 1, 0, 0, 0},
      {"bulk", 16, 3, 1, 0, 1, 0, 0, 0},
      {"bulk", 16, 4, 1, 0, 1, 0, 0, 0},
      {"bulk", 4, 8, 1, 0, 1, 0, 0, 0},
      {"bulk", 4, 4, 1, 0, 1, 0, 0, 0},
      {"sweep", 8, 8, 4, 0, 1, 0, 0, 0},
      {"sweep", 8, 8, 1, 64, 1, 0, 0, 0},
      {"sweep", 16, 16, 1, 64, 1, 0, 0, 0},
      {"sweep", 16, 16, 1, 32, 1, 0, 0, 0},
      {"sweep", 16, 16, 1, 32, 0, 0, 0, 0},
      {"sweep", 16, 16, 1, 32, 1, 1, 0, 0},
  };
  int bad = 0;
  for (const Case& c : cases) {
    const long long vox = (long long)c.N * c.D * c.S * c.S, n = vox * c.C;
    float *in, *grid = nullptr, *theta = nullptr, *out[2];
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out[0], n * 4)); CK(cudaMalloc(&out[1], n * 4));
    fill_uniform<<<592, 256, 0, st>>>(in, n, 17u, -1.f, 1.f);
    if (c.affine) {
      std::vector<float> t(12 * c.N);
      for (int k = 0; k < c.N; ++k) {
        const float a = 0.5236f;
        const float m[12] = {cosf(a), -sinf(a), 0, 0.2f, sinf(a), cosf(a), 0, 0.2f, 0, 0, 1.f, 0.2f};
        for (int i = 0; i < 12; ++i) t[k * 12 + i] = m[i];
      }
      CK(cudaMalloc(&theta, t.size() * 4));
      CK(cudaMemcpy(theta, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    } else {
      CK(cudaMalloc(&grid, vox * 3 * 4));
      fill_grid<<<592, 256, 0, st>>>(grid, c.N, c.D, c.S, c.S, 0.1f, 99u);
    }
    const double bytes = ((double)2 * c.C + (c.affine ? 0 : 3)) * 4.0 * (double)vox;
    const int reps = c.N > 1 ? 7 : 15;
    for (size_t vi = 0; vi < vars.size(); ++vi) {
      const Variant& v = vars[vi];
      if (v.prefetch && (long long)c.D * c.S * c.S * c.C * 4 * c.N > (48ll << 20)) continue;  // prefetch only for volumes that fit L2
      if (quick && !strcmp(v.kind, "sweep")) continue;
      LabParams p;
      memset(&p, 0, sizeof(p));
      p.in = in; p.grid = grid; p.theta = theta; p.N = c.N; p.C = c.C; p.Din = c.D; p.Hin = c.S; p.Win = c.S; p.Dout = c.D; p.Hout = c.S; p.Wout = c.S;
      p.os_c = 1; p.os_w = c.C; p.os_h = (long long)c.S * c.C; p.os_d = p.os_h * c.S; p.os_n = p.os_d * c.D;
      const bool bulk = !strcmp(v.kind, "bulk");
      if (!bulk) { p.bw = v.bw; p.bh = v.bh; p.bd = v.bd; p.bricks_w = c.S / v.bw; p.bricks_h = c.S / v.bh; p.bricks_d = c.D / v.bd; }
      p.rounds = 1; p.hint_in = v.hint;
      p.prefetch_bytes = v.prefetch ? (long long)c.N * c.D * c.S * c.S * c.C * 4 : 0;
      p.out = out[vi == 0 ? 0 : 1];
      const long long total = vox;
      unsigned ctas = 0;
      size_t bulk_smem = 0;
      int bulk_threads = 0;
      void (*bulk_fn)(const GS3Params) = nullptr;
      if (bulk) {
        if (v.bw == 8 && v.bh == 4) bulk_fn = gs3_bulk_kernel<8, 4>;
        else if (v.bw == 8 && v.bh == 3) bulk_fn = gs3_bulk_kernel<8, 3>;
        else if (v.bw == 16 && v.bh == 2) bulk_fn = gs3_bulk_kernel<16, 2>;
        else if (v.bw == 16 && v.bh == 3) bulk_fn = gs3_bulk_kernel<16, 3>;
        else if (v.bw == 16 && v.bh == 4) bulk_fn = gs3_bulk_kernel<16, 4>;
        else if (v.bw == 4 && v.bh == 8) bulk_fn = gs3_bulk_kernel<4, 8>;
        else if (v.bw == 4 && v.bh == 4) bulk_fn = gs3_bulk_kernel<4, 4>;
        else { fprintf(stderr, "no bulk instantiation %d %d\n", v.bw, v.bh); return 2; }
        bulk_smem = (size_t)v.bh * v.bw * 8 * c.C * 4;
        bulk_threads = 32 + v.bw * (c.C / 4);
        bulk_threads = (bulk_threads + 31) / 32 * 32;
        CK(cudaFuncSetAttribute(bulk_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
        int bocc = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, bulk_fn, bulk_threads, bulk_smem));
        if (v.chunk > 0 && v.chunk < bocc) bocc = v.chunk;
        ctas = (unsigned)(sms * bocc);
        p.rounds = bocc;  // printed in the rounds column: resident CTAs per SM
        if ((vox / c.N) % v.bw) { fprintf(stderr, "lattice not a multiple of the stage\n"); return 2; }
      } else if (!strcmp(v.kind, "brick")) ctas = (unsigned)(c.N * p.bricks_w * p.bricks_h * p.bricks_d);
      else {
        ctas = (unsigned)std::min<long long>(slots, (total + 31) / 32);
        if (v.chunk > 0) p.rounds = (int)std::max<long long>(1, (total + (long long)ctas * v.chunk / 2) / ((long long)ctas * v.chunk));
      }
      std::vector<float> ms;
      for (int r = 0; r < reps; ++r) {
        flush_k<<<148 * 8, 256, 0, st>>>(flush_buf, flush_bytes / 16);
        if (v.clean) read_stream_kernel<<<148 * 8, 256, 0, st>>>(flush_buf, flush_bytes / 16, 1, sink);
        CK(cudaEventRecord(e0, st));
        if (bulk) bulk_fn<<<ctas, bulk_threads, bulk_smem, st>>>(p);
        else if (!strcmp(v.kind, "brick")) gs3_cl_kernel<false><<<ctas, v.threads ? v.threads : 256, 0, st>>>(p);
        else if (!strcmp(v.kind, "balanced")) gs3_cl_balanced_kernel<false><<<ctas, 256, 0, st>>>(p);
        else gs3_cl_sweep_kernel<false><<<ctas, 256, 0, st>>>(p);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        ms.push_back(t);
      }
      std::sort(ms.begin(), ms.end());
      unsigned long long diff = 0;
      if (vi > 0) {
        CK(cudaMemsetAsync(d_cnt, 0, 8, st));
        count_diff<<<592, 256, 0, st>>>((const unsigned*)out[0], (const unsigned*)out[1], n, d_cnt);
        CK(cudaMemcpyAsync(&diff, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (diff) ++bad;
        CK(cudaMemsetAsync(out[1], 0xff, n * 4, st));
      }
      const double med = ms[reps / 2];
      printf("%-13s %-8s enum %2dx%2dx%d thr %3d chunk %3d rounds %3d hint %d prefetch %d %s  median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic  frac(6480) %.3f  diff_words %llu\n",
             c.name, v.kind, v.bw, v.bh, v.bd, v.threads ? v.threads : 256, v.chunk, p.rounds, v.hint, v.prefetch, v.clean ? "clean-flush" : "dirty-flush", med * 1e3, ms[0] * 1e3,
             bytes / (med * 1e-3) * 1e-9, bytes / (med * 1e-3) * 1e-9 / 6480.0, diff);
      fflush(stdout);
    }
    CK(cudaFree(in)); CK(cudaFree(out[0])); CK(cudaFree(out[1]));
    if (grid) CK(cudaFree(grid));
    if (theta) CK(cudaFree(theta));
  }
  printf(bad ? "FAIL: %d variants differ\n" : "OK: all variants bit-identical to the brick kernel\n", bad);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "--legacy")) return legacy_main();
  return sweep_main(argc > 1 && !strcmp(argv[1], "--quick"));
}
