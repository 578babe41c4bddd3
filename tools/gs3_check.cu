// Native (no Python) check of the channels-last grid_sample_3d kernel behind the C-ABI: the brick shape the library picks
// ("auto") and forced shapes (EMO_GS3_BRICK="bw,bh,bd[,threads]", instrumented build) against the round-1 brick of
// 8 x 8 x 4 voxels, bit for bit, plus timings.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gs3_check tools/gs3_check.cu -ldl
//   tools/gs3_check [path/to/libemoport_dbg.so]    (default: emoportraits_b200/csrc/libemoport_dbg.so, the instrumented build:
//                                                   `python -m emoportraits_b200.csrc.build --debug`; the product library has no switches)
//
// For every case: run every shape on the same seeded input, count differing output words on the device (must be 0:
// the brick shape only changes which CTA computes a voxel), then time each over REPS launches with an L2 flush
// (emo_l2_flush, 256 MB) before every launch and CUDA events around the launch only.  Prints one line per case and
// variant: median microseconds and algorithmic GB/s ((2*C*D*H*W [+ 3*D*H*W for a grid tensor]) * 4 B * N).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../include/emoportraits_b200.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void fill_uniform(float* p, long long n, unsigned seed, float lo, float hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = lo + (hi - lo) * (hash32((unsigned)i * 2654435761u + seed) >> 8) * (1.0f / 16777216.0f);
}
// identity lattice (pixel centres, align_corners=False style) + jitter; a few samples pushed far outside
__global__ void fill_grid(float* g, int N, int D, int H, int W, float jitter, unsigned seed) {
  const long long total = (long long)N * D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const float base[3] = {W > 1 ? -1.f + 2.f * w / (W - 1) : 0.f, H > 1 ? -1.f + 2.f * h / (H - 1) : 0.f,
                           D > 1 ? -1.f + 2.f * d / (D - 1) : 0.f};
    for (int k = 0; k < 3; ++k) {
      const unsigned hsh = hash32((unsigned)(i * 3 + k) + seed);
      float v = base[k] + jitter * ((hsh >> 8) * (2.0f / 16777216.0f) - 1.0f);
      if ((hsh & 0x3ff) == 0) v += 3.0f;  // ~0.1 % wild coordinates: zeros padding
      g[i * 3 + k] = v;
    }
  }
}
__global__ void count_diff(const unsigned* a, const unsigned* b, long long n, unsigned long long* cnt) {
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += a[i] != b[i];
  if (c) atomicAdd(cnt, c);
}

typedef int (*gs3_fn)(const emo_grid_sample3d_desc*, void*);
typedef int (*flush_fn)(void*, long long, void*);
typedef const char* (*err_fn)(void);

struct Case {
  const char* name;
  int N, C, Di, Hi, Wi, Do, Ho, Wo;
  bool affine;
  bool split;  // also write bf16 planes, [H][W][D][C] order (the decoder-input path of notebooks/infer.py:627)
};

int main(int argc, char** argv) {
  const char* libpath = argc > 1 ? argv[1] : "emoportraits_b200/csrc/libemoport_dbg.so";  // the instrumented build honours EMO_GS3_BRICK
  void* h = dlopen(libpath, RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen(%s): %s\n", libpath, dlerror()); return 2; }
  gs3_fn gs3 = (gs3_fn)dlsym(h, "emo_grid_sample3d");
  flush_fn l2flush = (flush_fn)dlsym(h, "emo_l2_flush");
  err_fn last_error = (err_fn)dlsym(h, "emo_last_error");
  if (!gs3 || !l2flush || !last_error) { fprintf(stderr, "missing symbols\n"); return 2; }

  const Case cases[] = {
      {"d64_affine", 1, 96, 64, 64, 64, 64, 64, 64, true, false},
      {"d64_grid", 1, 96, 64, 64, 64, 64, 64, 64, false, false},
      {"d16_affine", 1, 96, 16, 64, 64, 16, 64, 64, true, false},
      {"d16_grid", 1, 96, 16, 64, 64, 16, 64, 64, false, false},
      {"d16_affine_split_hwdc", 1, 96, 16, 64, 64, 16, 64, 64, true, true},
      {"ragged_b2", 2, 8, 9, 17, 13, 7, 11, 19, false, false},
      {"ragged_b3_affine", 3, 12, 5, 6, 7, 10, 9, 21, true, false},
      {"d64_b8_affine", 8, 96, 64, 64, 64, 64, 64, 64, true, false},
  };
  const int REPS = 15;
  const long long flush_bytes = 256ll << 20;
  void* flush_buf;
  CK(cudaMalloc(&flush_buf, flush_bytes));
  unsigned long long* d_cnt;
  CK(cudaMalloc(&d_cnt, 8));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int bad = 0;

  for (const Case& c : cases) {
    const long long in_n = (long long)c.N * c.C * c.Di * c.Hi * c.Wi;
    const long long vox = (long long)c.N * c.Do * c.Ho * c.Wo;
    const long long out_n = vox * c.C;
    float *in, *grid = nullptr, *theta = nullptr, *out[2];
    void* planes[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    CK(cudaMalloc(&in, in_n * 4));
    fill_uniform<<<592, 256, 0, st>>>(in, in_n, 17u, -1.f, 1.f);
    if (c.affine) {
      // 30 degree rotation about z, 10 degrees about x, scale 0.9, translation (0.2, -0.1, 0.05): some samples fall outside
      std::vector<float> t(12 * c.N);
      for (int n = 0; n < c.N; ++n) {
        const float a = 0.5236f + 0.1f * n, b = 0.1745f, s = 0.9f;
        const float R[9] = {cosf(a), -sinf(a), 0, sinf(a) * cosf(b), cosf(a) * cosf(b), -sinf(b), sinf(a) * sinf(b), cosf(a) * sinf(b), cosf(b)};
        const float tr[3] = {0.2f, -0.1f, 0.05f};
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) t[n * 12 + i * 4 + j] = s * R[i * 3 + j];
          t[n * 12 + i * 4 + 3] = tr[i];
        }
      }
      CK(cudaMalloc(&theta, t.size() * 4));
      CK(cudaMemcpyAsync(theta, t.data(), t.size() * 4, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
    } else {
      CK(cudaMalloc(&grid, vox * 3 * 4));
      fill_grid<<<592, 256, 0, st>>>(grid, c.N, c.Do, c.Ho, c.Wo, 0.1f, 99u);
    }
    for (int v = 0; v < 2; ++v) {
      CK(cudaMalloc(&out[v], out_n * 4));
      CK(cudaMemsetAsync(out[v], 0xff, out_n * 4, st));
      if (c.split)
        for (int k = 0; k < 2; ++k) {
          CK(cudaMalloc(&planes[v][k], out_n * 2));
          CK(cudaMemsetAsync(planes[v][k], 0xff, out_n * 2, st));
        }
    }
    emo_grid_sample3d_desc d = {};
    d.in = in; d.in_layout = 1;
    d.N = c.N; d.C = c.C; d.Din = c.Di; d.Hin = c.Hi; d.Win = c.Wi;
    d.grid = grid; d.theta = theta;
    d.Dout = c.Do; d.Hout = c.Ho; d.Wout = c.Wo;
    d.os_c = 1;
    if (c.split) {  // [N][H][W][D][C]
      d.os_d = c.C; d.os_w = (long long)c.Do * c.C; d.os_h = d.os_w * c.Wo; d.os_n = d.os_h * c.Ho;
    } else {        // [N][D][H][W][C]
      d.os_w = c.C; d.os_h = (long long)c.Wo * c.C; d.os_d = d.os_h * c.Ho; d.os_n = d.os_d * c.Do;
    }
    const double bytes = ((double)2 * c.C + (c.affine ? 0 : 3)) * 4.0 * (double)vox;
    const char* variants[] = {"brick_8x8x4", "auto", "brick_8x8x2", "brick_8x8x1", "brick_8x8x1_128thr", "brick_4x4x4", "bulk_copy_gather"};
    const char* envs[] = {"8,8,4", nullptr, "8,8,2", "8,8,1", "8,8,1,128", "4,4,4", nullptr};
    for (int v = 0; v < 7; ++v) {
      if (envs[v]) setenv("EMO_GS3_BRICK", envs[v], 1);
      else unsetenv("EMO_GS3_BRICK");
      if (v == 6) setenv("EMO_GS3_BULK", "1", 1);  // falls back to the brick kernel where its preconditions do not hold (ragged cases)
      else unsetenv("EMO_GS3_BULK");
      const int slot = v == 0 ? 0 : 1;
      d.out = out[slot];
      d.out_hi = planes[slot][0]; d.out_lo = planes[slot][1];
      std::vector<float> ms;
      for (int r = 0; r < REPS; ++r) {
        if (l2flush(flush_buf, flush_bytes, st) != 0) { fprintf(stderr, "l2 flush: %s\n", last_error()); return 2; }
        CK(cudaEventRecord(e0, st));
        if (gs3(&d, st) != 0) { fprintf(stderr, "%s/%s: %s\n", c.name, variants[v], last_error()); return 2; }
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        ms.push_back(t);
      }
      std::sort(ms.begin(), ms.end());
      const double med = ms[ms.size() / 2];
      unsigned long long diff = 0;
      if (v > 0) {
        CK(cudaMemsetAsync(d_cnt, 0, 8, st));
        count_diff<<<592, 256, 0, st>>>((const unsigned*)out[0], (const unsigned*)out[1], out_n, d_cnt);
        if (c.split)
          for (int k = 0; k < 2; ++k)
            count_diff<<<592, 256, 0, st>>>((const unsigned*)planes[0][k], (const unsigned*)planes[1][k], out_n / 2, d_cnt);
        CK(cudaMemcpyAsync(&diff, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (diff) ++bad;
        CK(cudaMemsetAsync(out[1], 0xff, out_n * 4, st));  // the next variant must write everything again
        if (c.split) for (int k = 0; k < 2; ++k) CK(cudaMemsetAsync(planes[1][k], 0xff, out_n * 2, st));
      }
      printf("%-22s %-20s median %8.2f us  min %8.2f us  %8.1f GB/s algorithmic  diff_words_vs_8x8x4 %llu\n", c.name,
             variants[v], med * 1e3, ms[0] * 1e3, bytes / (med * 1e-3) * 1e-9, diff);
      fflush(stdout);
    }
    CK(cudaFree(in));
    if (grid) CK(cudaFree(grid));
    if (theta) CK(cudaFree(theta));
    for (int v = 0; v < 2; ++v) {
      CK(cudaFree(out[v]));
      for (int k = 0; k < 2; ++k) if (planes[v][k]) CK(cudaFree(planes[v][k]));
    }
  }
  unsetenv("EMO_GS3_BRICK");
  unsetenv("EMO_GS3_BULK");
  printf(bad ? "FAIL: %d variant runs differ from the 8x8x4 brick\n" : "OK: all brick shapes bit-identical\n", bad);
  return bad ? 1 : 0;
}
