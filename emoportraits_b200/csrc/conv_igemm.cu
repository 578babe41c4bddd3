// Implicit-GEMM convolution (2-D / 3-D) for sm_100a: TMA-fed, tcgen05.mma with TMEM accumulators.
//
// Replaces the F.conv2d / F.conv3d call sites of the reference hot path
// (networks/volumetric_avatar/utils.py:661-788 ResBlock, :894-915 Conv2d_ws/Conv3d_ws,
//  decoder.py:77-81,349-356, local_encoder.py:104-108, warp_generator_resnet.py:99-106).
//
// GEMM view:   D[pixel][cout] = sum_{tap, cin} A[pixel + tap][cin] * W[tap][cout][cin]
//   M = 128 output pixels of one (td x th x tw) box of one sample,
//   N = BN output channels, K = taps * Cin walked in chunks of KC channels.
// im2col-free: for every (tap, k-chunk) the A tile is ONE TMA box load of the channels-last activation
//   tensor at the tap-shifted coordinate; out-of-bounds rows/columns are zero-filled by TMA, which is
//   exactly the conv zero padding.  The box lands in shared memory as [pixel][KC] rows with the
//   128B/64B hardware swizzle = the canonical K-major UMMA operand layout.
// Precision: fp32 activations/weights are pre-split into bf16 planes (hi, lo[, lo2]); each K step issues
//   hi*hi + hi*lo + lo*hi (3 x tcgen05.mma kind::f16; 6 products with three planes) into an fp32 TMEM accumulator.
//   The tensor core ACCUMULATES WITH TRUNCATION (tools/accum_probe.py), so an accumulator only takes a short chunk
//   of MMAs (24 / 48); the epilogue warps promote every chunk to fp32 registers (round-to-nearest adds) through a
//   4-deep TMEM chunk ring.  Plain bf16/tf32 operands, or one long accumulation chain, do not hold the 1e-3 budget.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (both issue behind elect.sync),
//   8 accumulator warps, two per TMEM lane quadrant, each owning half of the tile's columns (chunk promotion during
//   the main loop) and running the tile's final phase: EPI = 0 straight from registers to global memory, EPI = 1 through a
//   swizzled shared-memory staging tile and TMA tensor stores / residual loads.  See the template comment at the kernel.
// Main loop: stride-1 convolutions load ONE halo A tile per (kernel column, k-chunk) and serve its kh tap rows through
//   descriptor row offsets (A ring + B ring, "row reuse"); other layers load an {A, B} stage per (tap, k-chunk).  On the
//   instrumented build (tools/conv_timeline.py) the main loop of the dominant layer runs at the tensor pipe's pace for the
//   clock the power cap allows (72 k-steps x 12 MMAs in 35.6 us); what is left of a launch is prologue (2.8 us) and the
//   last tile's final phase.
// CG = 2: CTA pairs (2-CTA clusters) issue cta_group::2 M = 256 MMAs; each CTA stages its own 128 pixels and half of
//   the weight tile, which halves the weight ingest and brings the shared-memory operand reads per MMA under the
//   tensor floor (measured in tools/mma_probe.cu).
// Persistent CTAs walk tiles round-robin; the chunk ring lets the MMA of the next tile run ahead of the epilogue.
#include <type_traits>

#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace emo {

static constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
static constexpr int kEpiThreads = 256;
static constexpr int kMaxStages = 8;
static constexpr int kMaxAStages = 4;  // row-reuse mode: depth cap of the A (halo tile) ring
static constexpr int kTileM = 128;
static constexpr int kMaxBN = 160;  // N tile cap: each epilogue thread keeps half a row of accumulators in registers
static constexpr int kAccBufs = 4;  // max depth of the TMEM accumulation-chunk ring (512 columns / BN, at most 4)

struct ConvKParams {
  int N, Dout, Hout, Wout, Cout;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int kchunks;
  int tw, th, td;
  int tiles_w, tiles_h, tiles_d;
  int m_tiles, n_tiles;
  int BN;
  int stages;
  int flush;  // k-steps per TMEM accumulation chunk
  int nbuf;   // depth of the TMEM chunk ring = min(4, 512 / BN)
  int ksplit;   // split-K factor: work item = (tile, K part); part `k` writes its partial sums to ws + k * ws_part_elems
  float* ws;    // split-K fp32 workspace [ksplit][N][Dout][Hout][Wout][Cout] (no initial state needed)
  long long ws_part_elems;
  int dbg;    // EMO_CONV_DEBUG builds only (tools/conv_bound_probe.py): 1 skip TMA loads, 2 skip MMAs, 4 skip the tile epilogue's
              // global traffic, 8 skip the TMEM chunk reads, 16 / 32 / 64 skip the residual reads / output stores / statistics.  Results are garbage; only the timing is of interest.
  int cg;     // 1, or 2: CTA pairs issue cta_group::2 MMAs (M = 256: two pixel tiles; the weight tile is split across the
              // pair's shared memories, so each SM ingests A 128 x KC + B (BN/2) x KC per k-step instead of A + B BN x KC)
  int cs;     // cluster size (1, 2, 4): CTAs of a cluster take consecutive pixel tiles of the same channel tile and
              // share the weight tile through TMA multicast (each CTA loads 1/cs of it for everybody)
  const float* bias;
  const float* residual;
  int res_shift;
  int rD, rH, rW;
  int act;
  const float* post_add;
  float* out;
  int out_nchw;
  double* stats;
  int G, cpg;
  // sub-pixel mode (conv_igemm_ps_kernel): 3x3 conv over the nearest-x2 upsampling of the input, evaluated on the LOW-resolution
  // planes.  Dout/Hout/Wout above then describe the low-resolution pixel grid the tiles walk; the output tensor is
  // [N][Dout][oH = 2 Hout][oW = 2 Wout][Cout].  The N-tile index carries the output phase: nt = phase * ntc + channel tile,
  // phase = 2 * (row parity) + (column parity); each phase has its own four 2x2 taps (weights [phase * 4 + tap]).
  int oH, oW;
  int ntc;  // channel tiles per phase (= Cout_pad / BN)
  float out_scale;  // conv_igemm_f16_kernel: 1 / (activation plane scale * weight plane scale)
  int res_tma;      // EPI = 1: 1 = same-resolution residual TMA-loaded into the staging tile, 2 = half-resolution residual
                    // (res_shift == 1) TMA-loaded into a quarter-size tile next to it (tm.res); 0 = read from global by the warps
  // row-reuse mode (stride-1 convolutions with kh > 1 on tiles of one depth slice): the A operand of a (kernel column,
  // k-chunk) group is ONE halo tile of a_rows = tw * (th + kh - 1) pixels serving all kh tap rows; A ring of `sa` halo tiles,
  // B ring of `sb` weight tiles (see the kernel's producer role)
  int yreuse, sa, sb, a_rows;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
// cta_group::2 variants (pair = even/odd CTA of a 2-CTA cluster; the even CTA is the leader).  The barrier operand has
// the peer bit cleared, i.e. it names the LEADER's barrier at the same shared-memory offset (cute Sm100MmaPeerBitMask).
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_5d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA tensor store of one staged panel (shared -> global through the async proxy) and its bulk-group bookkeeping
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// One elected lane of a converged warp.  Code guarded by `lane == 0` makes ptxas wrap every TMA / MMA issue in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall (the descriptors live in uniform registers): measured 91 clk per MMA issue
// against a 64 clk tensor floor (tools/mma_probe.cu).  Behind elect.sync the operands are provably uniform.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {  // arrives on the barrier of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
template <int KC>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  constexpr uint64_t row_bytes = KC * 2;          // 128 or 64
  constexpr uint64_t sbo = (8 * row_bytes) >> 4;  // 8-row swizzle atom pitch
  constexpr uint64_t layout = (KC == 64) ? 2ull : 4ull;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Tile statistics for the fused GroupNorm: `acc` holds this thread's row (pixel) of finished outputs, kMaxBN/2 columns.
// R adjacent columns (R | channels per group, R | 16) are summed inside the thread; the resulting 32/R values per
// 16-column block (sums, then sums of squares) are packed 32 to a pass and transpose-reduced over the warp's 32 rows
// (16+8+4+2+1 shuffles); lane l ends up with the warp total of value l and adds it to the CTA's column accumulators
// (only the first column of every R-group receives data; the per-group pass that follows sums whole groups).
template <int R>
__device__ __forceinline__ void tile_stats(const float (&acc)[kMaxBN / 2], int ncols, int lane, float* cs, float* cq) {
  constexpr int kBlocks = kMaxBN / 32;           // 16-column blocks a thread can own
  constexpr int VPB = 32 / R;                    // values per block
  constexpr int GPB = 16 / R;                    // column groups per block
  constexpr int BPP = R < kBlocks ? R : kBlocks; // blocks per pass
#pragma unroll
  for (int b0 = 0; b0 < kBlocks; b0 += BPP) {
    if (b0 * 16 < ncols) {
      float val[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int bl = i / VPB, kind = (i % VPB) / GPB, gi = i % GPB;
        float a = 0.f;
        if (bl < BPP && b0 + bl < kBlocks) {
#pragma unroll
          for (int t = 0; t < R; ++t) {
            const float x = acc[(b0 + bl) * 16 + gi * R + t];
            a += kind ? x * x : x;
          }
        }
        val[i] = a;
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
          const float send = upper ? val[j] : val[j + off];
          const float keep = upper ? val[j + off] : val[j];
          val[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      const int bl = lane / VPB, kind = (lane % VPB) / GPB, gi = lane % GPB;
      const int col = (b0 + bl) * 16 + gi * R;
      if (bl < BPP && col < ncols) (kind ? cq : cs)[col] = val[0];  // this warp is the only writer of (quadrant, column)
    }
  }
}

// Second half of the tile statistics: the per-quadrant column slots (filled by tile_stats, then a named barrier) are summed in
// a fixed order - quadrant 0..3, then the columns of a GroupNorm group - and added to the fp64 global sums.  (Slots of columns
// that never receive data - all but the first of every R-group - stay zero from the kernel's start.)
__device__ __forceinline__ void tile_group_stats(const ConvKParams& p, const float* cs, const float* cq, int n, int n0, int BN, int et, int nthreads) {
  const int cpg = p.cpg;
  if (BN % cpg == 0 && (n0 % cpg) == 0) {
    const int ng = BN / cpg;
    // tile_stats<R> leaves data only in the first column of every R-group (R = largest power of two <= 16 dividing cpg); the other
    // slots hold zeros from the kernel's start, so skipping them changes neither the sums nor their order
    const int rstep = (cpg & -cpg) < 16 ? (cpg & -cpg) : 16;
    for (int g = et; g < ng; g += nthreads) {
      float a = 0.f, b = 0.f;
      for (int q = 0; q < 4; ++q)
        for (int j = 0; j < cpg; j += rstep) { a += cs[q * kMaxBN + g * cpg + j]; b += cq[q * kMaxBN + g * cpg + j]; }
      const int gi = (n0 / cpg) + g;
      if (gi < p.G) {
        atomicAdd(&p.stats[((long long)n * p.G + gi) * 2], (double)a);
        atomicAdd(&p.stats[((long long)n * p.G + gi) * 2 + 1], (double)b);
      }
    }
  } else {
    for (int j = et; j < BN; j += nthreads) {
      const int c = n0 + j;
      if (c < p.Cout) {
        const int gi = c / cpg;
        const float a = (cs[j] + cs[kMaxBN + j]) + (cs[2 * kMaxBN + j] + cs[3 * kMaxBN + j]);
        const float b = (cq[j] + cq[kMaxBN + j]) + (cq[2 * kMaxBN + j] + cq[3 * kMaxBN + j]);
        atomicAdd(&p.stats[((long long)n * p.G + gi) * 2], (double)a);
        atomicAdd(&p.stats[((long long)n * p.G + gi) * 2 + 1], (double)b);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// Instrumented build only (-DEMO_CONV_DEBUG, tools/conv_timeline.py): per-CTA time stamps (%globaltimer, ns) of the phases of
// the LAST stamped launch: 0 entry, 1 prologue done, 2 / 3 producer first / last issue, 4 first operands landed, 5 last MMA
// issued, 6 last accumulation chunk consumed (per tile: last tile wins), 7 final-phase stores issued, 8 statistics done,
// 9 teardown barrier passed, 10 TMEM freed.
#ifdef EMO_CONV_DEBUG
__device__ unsigned long long g_conv_stamps[160 * 16];
__device__ __forceinline__ unsigned long long emo_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define EMO_STAMP(k, cond) do { if ((p.dbg & 256) && (cond) && blockIdx.x < 160) g_conv_stamps[blockIdx.x * 16 + (k)] = emo_gtime(); } while (0)
#define EMO_STAMP_ONCE(k, cond) EMO_STAMP(k, cond)
#else
#define EMO_STAMP(k, cond) do { } while (0)
#define EMO_STAMP_ONCE(k, cond) do { } while (0)
#endif

struct TMaps {
  CUtensorMap a[3];  // activation planes: hi, lo, lo2
  CUtensorMap b[3];  // weight planes
  CUtensorMap out;   // EPI = 1: fp32 output, 32-channel panels of one pixel box (SWIZZLE_128B)
  CUtensorMap res;   // EPI = 1: residual, same panels (half-size boxes for a half-resolution residual)
};

#define EMO_CONV_PS 0
#define EMO_CONV_F16 0
#define EMO_CONV_KERNEL_NAME conv_igemm_kernel
#include "conv_igemm_kernel.inc"
#undef EMO_CONV_PS
#undef EMO_CONV_KERNEL_NAME
#define EMO_CONV_PS 1
#define EMO_CONV_KERNEL_NAME conv_igemm_ps_kernel
#include "conv_igemm_kernel.inc"
#undef EMO_CONV_PS
#undef EMO_CONV_F16
#undef EMO_CONV_KERNEL_NAME
#define EMO_CONV_PS 0
#define EMO_CONV_F16 1
#define EMO_CONV_KERNEL_NAME conv_igemm_f16_kernel
#include "conv_igemm_kernel.inc"
#undef EMO_CONV_PS
#undef EMO_CONV_F16
#undef EMO_CONV_KERNEL_NAME

// split-K finalize: out = act(sum_parts ws[part] + bias + residual) + post_add, statistics.  The parts are summed in part
// order (fixed), so the output is reproducible run to run.
struct FinParams {
  const float* ws;
  int ksplit;
  long long part_elems;
  float* out;
  int N, C;
  long long S;  // spatial positions per sample
  const float* bias;
  const float* residual;
  const float* post_add;
  int act, out_nchw;
  double* stats;
  int G;
};

__global__ void __launch_bounds__(256) splitk_finalize_kernel(const FinParams f) {
  extern __shared__ double sstat[];  // fp64: sums of fp32 values are exact -> independent of the order of the atomics
  const int n = blockIdx.y;
  if (f.stats) {
    for (int i = threadIdx.x; i < 2 * f.G; i += blockDim.x) sstat[i] = 0.0;
    __syncthreads();
  }
  const long long per_n = f.S * f.C;
  const int cpg = f.stats ? f.C / f.G : 1;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < per_n; t += (long long)gridDim.x * blockDim.x) {
    const long long i = (long long)n * per_n + t;
    const int c = (int)(t % f.C);
    const long long sp = t / f.C;
    float v = 0.f;
#pragma unroll 8
    for (int k = 0; k < f.ksplit; ++k) v += __ldcg(f.ws + (long long)k * f.part_elems + i);  // part order: reproducible
    if (f.bias) v += __ldg(f.bias + c);
    if (f.residual) v += __ldg(f.residual + i);
    v = act_apply(v, f.act);
    if (f.post_add) v += __ldg(f.post_add + t);
    if (f.out_nchw) f.out[((long long)n * f.C + c) * f.S + sp] = v;
    else f.out[i] = v;
    if (f.stats) {
      atomicAdd(&sstat[c / cpg], (double)v);
      atomicAdd(&sstat[f.G + c / cpg], (double)(v * v));
    }
  }
  if (f.stats) {
    __syncthreads();
    for (int g = threadIdx.x; g < f.G; g += blockDim.x) {
      atomicAdd(&f.stats[((long long)n * f.G + g) * 2], sstat[g]);
      atomicAdd(&f.stats[((long long)n * f.G + g) * 2 + 1], sstat[f.G + g]);
    }
  }
}

// Fused finalize + normalisation + activation + plane split for SMALL split-K layers (emo_conv_desc.post): one 8-CTA
// cluster per sample does what splitk_finalize_kernel + gn statistics + emo_apply do in three launches for the ResNet
// tails and the first warp-generator blocks (<= 64 Ki elements per sample, 4x4 .. 32x32 maps), whose time is launch
// latency, not work.  Thread -> elements e = rank * 512 + tid + k * 4096: C divides 512, so a thread owns ONE channel.
// GroupNorm statistics: per-thread sums in element order -> per-CTA group sums in thread order -> cluster totals in rank
// order over distributed shared memory: fixed orders throughout, bit-reproducible.
static constexpr int kPostCluster = 8;
static constexpr int kPostThreads = 512;
static constexpr int kPostMaxV = 16;  // elements per thread: <= 65536 elements per sample
struct PostParams {
  const float* ws;
  int ksplit;
  long long part_elems;
  int N, C;
  long long S;
  const float* bias;      // the convolution's own bias / same-resolution residual / activation (before the norm)
  const float* residual;
  int conv_act;
  emo_apply_desc ap;      // the post-op: x ignored (= the convolution's output), up == 1
};

__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ double ld_dsmem_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPostThreads, 1) splitk_post_kernel(const PostParams f) {
  __shared__ float s_sum[kPostThreads], s_sq[kPostThreads];
  __shared__ double s_part[2 * 64];  // this CTA's (sum, sum of squares) per group
  __shared__ double s_tot[2 * 64];   // cluster totals
  const emo_apply_desc& a = f.ap;
  const int tid = threadIdx.x;
  const uint32_t rank = (uint32_t)blockIdx.x % kPostCluster;
  const int n = (int)blockIdx.x / kPostCluster;
  const long long per_n = f.S * f.C;
  const int c = tid % f.C;  // this thread's channel (512 % C == 0 and 4096 % C == 0)
  float v[kPostMaxV];
  float ts = 0.f, tq = 0.f;
  const long long e0 = (long long)rank * kPostThreads + tid;               // this thread's elements: e0 + k * 4096
  const long long base = (long long)n * per_n;
#pragma unroll
  for (int k = 0; k < kPostMaxV; ++k) v[k] = 0.f;
  // K parts summed in part order; four parts x all of the thread's elements are loaded before they are added, so up to 64
  // independent L2 loads are in flight per thread instead of one (ncu, round 2: 26 us per launch with the serial loop)
  for (int part = 0; part < f.ksplit; part += 4) {
    float t[4][kPostMaxV];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float* w = f.ws + (long long)(part + q) * f.part_elems + base;
#pragma unroll
      for (int k = 0; k < kPostMaxV; ++k) {
        const long long e = e0 + (long long)k * (kPostCluster * kPostThreads);
        t[q][k] = (part + q < f.ksplit && e < per_n) ? __ldcg(w + e) : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < kPostMaxV; ++k) v[k] += t[q][k];
  }
#pragma unroll
  for (int k = 0; k < kPostMaxV; ++k) {
    const long long e = e0 + (long long)k * (kPostCluster * kPostThreads);
    if (e < per_n) {
      const long long i = base + e;
      float x = v[k];
      if (f.bias) x += __ldg(f.bias + c);
      if (f.residual) x += __ldg(f.residual + i);
      x = act_apply(x, f.conv_act);
      v[k] = x;
      ts += x; tq = fmaf(x, x, tq);
    } else {
      v[k] = 0.f;
    }
  }
  float A = 1.f, B = 0.f;
  if (a.stats) {
    // GroupNorm of the convolution's output
    const int G = a.G, cpg = f.C / G;
    s_sum[tid] = ts; s_sq[tid] = tq;
    __syncthreads();
    if (tid < G) {
      double sg = 0.0, qg = 0.0;
      for (int m = 0; m < kPostThreads / f.C; ++m)
        for (int cc = tid * cpg; cc < (tid + 1) * cpg; ++cc) { sg += (double)s_sum[cc + m * f.C]; qg += (double)s_sq[cc + m * f.C]; }
      s_part[tid] = sg; s_part[64 + tid] = qg;
    }
    cluster_sync_all();  // every CTA's group sums are in its shared memory
    if (tid < G) {
      double sg = 0.0, qg = 0.0;
      for (uint32_t r = 0; r < (uint32_t)kPostCluster; ++r) {
        sg += ld_dsmem_f64(dsmem_addr(&s_part[tid], r));
        qg += ld_dsmem_f64(dsmem_addr(&s_part[64 + tid], r));
      }
      s_tot[tid] = sg; s_tot[64 + tid] = qg;
      if (rank == 0) {  // publish the statistics too (same layout as every other producer)
        double* st = const_cast<double*>(a.stats);  // == the convolution's `stats` output (checked by the host)
        st[((long long)n * G + tid) * 2] = sg;
        st[((long long)n * G + tid) * 2 + 1] = qg;
      }
    }
    cluster_sync_all();  // nobody retires (or overwrites s_part) while a peer may still read it
    const int g = c / cpg;
    const double mean = s_tot[g] / a.count;
    double var = s_tot[64 + g] / a.count - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    float gam = a.gamma ? a.gamma[c] : 1.f, bet = a.beta ? a.beta[c] : 0.f;
    if (a.ada_w) {
      const float aw = a.ada_w[(long long)n * f.C + c], ab = a.ada_b[(long long)n * f.C + c];
      bet = bet * aw + ab;
      gam = gam * aw;
    }
    A = rstd * gam;
    B = bet - (float)mean * A;
  } else if (a.A) {
    A = a.A[(a.ab_per_sample ? (long long)n * f.C : 0) + c];
    B = a.B[(a.ab_per_sample ? (long long)n * f.C : 0) + c];
  }
  float A2 = 1.f, B2 = 0.f;
  if (a.A2) { A2 = a.A2[c]; B2 = a.B2[c]; }
#pragma unroll
  for (int k = 0; k < kPostMaxV; ++k) {
    const long long e = (long long)rank * kPostThreads + tid + (long long)k * (kPostCluster * kPostThreads);
    if (e < per_n) {
      const long long i = (long long)n * per_n + e;
      float y = fmaf(v[k], A, B);
      if (a.res) y += fmaf(__ldg(a.res + i), A2, B2);
      y = act_apply(y, a.act);
      if (a.out) a.out[i] = y;
      if (a.out_hi) {
        if (a.plane_fp16) {
          __half h, l;
          split_f16(y * a.plane_scale, h, l);
          ((__half*)a.out_hi)[i] = h; ((__half*)a.out_lo)[i] = l;
        } else if (a.out_lo2) {
          __nv_bfloat16 h, l, l2;
          split_bf16x3(y, h, l, l2);
          ((__nv_bfloat16*)a.out_hi)[i] = h; ((__nv_bfloat16*)a.out_lo)[i] = l; ((__nv_bfloat16*)a.out_lo2)[i] = l2;
        } else {
          __nv_bfloat16 h, l;
          split_bf16(y, h, l);
          ((__nv_bfloat16*)a.out_hi)[i] = h; ((__nv_bfloat16*)a.out_lo)[i] = l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = (PFN_cuTensorMapEncodeTiled)p;
  }
  return fn;
}

static int pick_box(int dim, int want) {
  int b = want;
  while (b > dim) b >>= 1;
  return b < 1 ? 1 : b;
}

}  // namespace emo

using namespace emo;

#ifdef EMO_CONV_DEBUG
extern "C" int emo_debug_conv_stamps(unsigned long long* host160x16) {
  return cudaMemcpyFromSymbol(host160x16, g_conv_stamps, sizeof(unsigned long long) * 160 * 16) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int emo_conv_igemm(const emo_conv_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EMO_REQUIRE(d && d->a_hi && d->a_lo && d->w_hi && d->w_lo && d->out, "emo_conv_igemm: null pointer");
  EMO_REQUIRE(d->Cin % 32 == 0, "emo_conv_igemm: Cin=%d must be a multiple of 32 (use emo_conv_direct)", d->Cin);
  EMO_REQUIRE(d->Cout_pad % 16 == 0 && d->Cout <= d->Cout_pad, "emo_conv_igemm: Cout_pad=%d must be a multiple of 16 >= Cout", d->Cout_pad);
  EMO_REQUIRE(d->sd >= 1 && d->sh >= 1 && d->sw >= 1 && d->sd <= 2 && d->sh <= 2 && d->sw <= 2, "emo_conv_igemm: stride must be 1 or 2");
  EMO_REQUIRE(((uintptr_t)d->a_hi % 16) == 0 && ((uintptr_t)d->a_lo % 16) == 0 && ((uintptr_t)d->w_hi % 16) == 0 &&
                  ((uintptr_t)d->w_lo % 16) == 0 && ((uintptr_t)d->out % 16) == 0,
              "emo_conv_igemm: pointers must be 16-byte aligned");
  if (d->stats) EMO_REQUIRE(d->G > 0 && d->Cout % d->G == 0, "emo_conv_igemm: Cout=%d not divisible by G=%d", d->Cout, d->G);
  if (d->post) {
    EMO_REQUIRE(!d->out_nchw && d->post->N == d->N && d->post->C == d->Cout && d->post->D == d->Dout && d->post->H == d->Hout && d->post->W == d->Wout,
                "emo_conv_igemm: post-op shape must be the convolution's channels-last output shape");
    EMO_REQUIRE(!d->post->stats || d->post->stats == d->stats, "emo_conv_igemm: a GroupNorm post-op normalises with the convolution's own statistics (post->stats == stats)");
  }
  const bool vec_ok = (d->Cout % 4 == 0);
  EMO_REQUIRE(vec_ok || d->Cout < 16, "emo_conv_igemm: Cout=%d must be a multiple of 4 (or < 16)", d->Cout);

  PFN_cuTensorMapEncodeTiled encode = get_encode();
  if (!encode) { set_error("emo_conv_igemm: cuTensorMapEncodeTiled entry point unavailable"); return EMO_ERR_CUDA; }

  const int NP = d->a_lo2 ? 3 : 2;
  EMO_REQUIRE((d->a_lo2 == nullptr) == (d->w_lo2 == nullptr), "emo_conv_igemm: a_lo2 and w_lo2 must be given together");
  // sub-pixel mode: the descriptor describes conv3x3(pad 1)(nearest_x2(input)) with the input given at LOW resolution and
  // the weights pre-folded to [4 phases][2x2 taps][Cout_pad][Cin] (emoportraits_b200/ops.py: pack_upconv_weight)
  const int ps = d->upconv ? 1 : 0;
  if (ps) {
    EMO_REQUIRE(d->kd == 1 && d->kh == 3 && d->kw == 3 && d->sd == 1 && d->sh == 1 && d->sw == 1 && d->pd == 0 && d->ph == 1 && d->pw == 1,
                "emo_conv_igemm: upconv needs a 1x3x3 stride-1 pad-1 convolution");
    EMO_REQUIRE(d->Din == 1 && d->Dout == 1 && d->Hout == 2 * d->Hin && d->Wout == 2 * d->Win,
                "emo_conv_igemm: upconv output must be (2 Hin, 2 Win) (got %d x %d from %d x %d)", d->Hout, d->Wout, d->Hin, d->Win);
    EMO_REQUIRE(NP == 2 && d->Cin % 64 == 0 && d->Cout == d->Cout_pad && !d->out_nchw,
                "emo_conv_igemm: upconv needs two-plane operands, Cin %% 64 == 0, Cout %% 16 == 0, channels-last output");
  }
  const int f16 = d->operand_fp16 ? 1 : 0;
  if (f16) EMO_REQUIRE(NP == 2 && !ps && d->out_scale > 0.f, "emo_conv_igemm: fp16 operands need two planes, no upconv, out_scale > 0");
  const int gH = ps ? d->Hin : d->Hout, gW = ps ? d->Win : d->Wout;  // the pixel grid the tiles walk
  const int taps_k = ps ? 4 : d->kd * d->kh * d->kw;                  // taps in one tile's K loop
  const int nt_mult = ps ? 4 : 1;                                     // N tiles per channel tile (one per output phase)
  int KC = (d->Cin % 64 == 0) ? 64 : 32;  // three-plane tiles fall back to 32 below when 64 leaves < 3 pipeline stages
  // N tile: largest multiple of 16 that divides Cout_pad and fits the register-resident accumulator row
  int BN = 0;
  const int bn_step_ok = d->upconv ? 32 : 16;  // sub-pixel mode always runs in pair mode, which needs BN % 32 == 0
  for (int cand = kMaxBN; cand >= 16; cand -= 16)
    if (d->Cout_pad % cand == 0 && cand % bn_step_ok == 0) { BN = cand; break; }
  EMO_REQUIRE(BN > 0, "emo_conv_igemm: no N tile for Cout_pad=%d", d->Cout_pad);
  // prefer 128-wide tiles when that fills the machine better (more tiles than SMs matters more than tile width)
  int sm_count = 148, cur_dev = 0;
  {
    cudaGetDevice(&cur_dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cur_dev);
  }

  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Dout = d->Dout; p.Hout = gH; p.Wout = gW; p.Cout = d->Cout;
  p.oH = d->Hout; p.oW = d->Wout;
  p.kd = d->kd; p.kh = ps ? 2 : d->kh; p.kw = ps ? 2 : d->kw; p.sd = d->sd; p.sh = d->sh; p.sw = d->sw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  // pixel box: 128 pixels, widest along W first
  p.tw = pick_box(gW, 16);
  p.th = pick_box(gH, kTileM / p.tw > 0 ? kTileM / p.tw : 1);
  p.td = pick_box(d->Dout, kTileM / (p.tw * p.th) > 0 ? kTileM / (p.tw * p.th) : 1);
  if (p.tw * p.th * p.td < kTileM && p.tw < gW) {  // shallow/short tensor: widen along W
    int tw = p.tw;
    while (tw * 2 <= gW && tw * 2 * p.th * p.td <= kTileM && tw * 2 * d->sw <= 256) tw *= 2;
    p.tw = tw;
  }
  EMO_REQUIRE(p.tw * d->sw <= 256 && p.th * d->sh <= 256 && p.td * d->sd <= 256, "emo_conv_igemm: TMA box too large");
  p.tiles_w = cdiv(gW, p.tw); p.tiles_h = cdiv(gH, p.th); p.tiles_d = cdiv(d->Dout, p.td);
  p.m_tiles = d->N * p.tiles_d * p.tiles_h * p.tiles_w;
  // small-M layers (ResNet tails, the first warp-generator blocks): the serial K loop of a handful of CTAs is bound by the
  // TMA->MMA->commit round trip (~2.4 us per pipeline turn), not by work.  With a workspace: split K over CTAs
  // (partials red.add'ed in fp32, finalize pass).  Without: narrow the N tile until the tile count fills the machine.
  const int ksteps_total = taps_k * (d->Cin / ((d->Cin % 64 == 0) ? 64 : 32));
  int ksplit = 1;
  {
    const long long tiles = (long long)p.m_tiles * (d->Cout_pad / BN) * nt_mult;
    const long long out_elems = (long long)d->N * d->Dout * d->Hout * d->Wout * d->Cout;
    if (!ps && d->splitk_ws && 2 * out_elems <= d->splitk_ws_elems && d->res_shift == 0 && tiles * 2 <= sm_count) {
      long long parts = (2ll * sm_count) / tiles;
      if (parts > ksteps_total / 4) parts = ksteps_total / 4;
      if (parts * out_elems > d->splitk_ws_elems) parts = d->splitk_ws_elems / out_elems;  // one workspace slice per part
      if (parts >= 2) ksplit = (int)parts;
    }
  }
  if (ksplit == 1 && (long long)p.m_tiles * (d->Cout_pad / BN) * nt_mult < sm_count / 2) {
    for (int cand = BN; cand >= 16; cand -= 16) {
      if (d->Cout_pad % cand || cand % bn_step_ok) continue;
      BN = cand;
      if ((long long)p.m_tiles * (d->Cout_pad / cand) * nt_mult >= sm_count) break;
    }
  }
  p.BN = BN;
  p.ntc = d->Cout_pad / BN;
  p.n_tiles = p.ntc * nt_mult;
  p.ksplit = ksplit;
  p.ws = d->splitk_ws;
  p.ws_part_elems = (long long)d->N * d->Dout * d->Hout * d->Wout * d->Cout;
  p.dbg = 0;
#ifdef EMO_CONV_DEBUG
  { const char* e = getenv("EMO_CONV_DBG"); p.dbg = e ? atoi(e) : 0; }
#endif
  {
    // pair mode (cta_group::2, M = 256): an SS-mode 128 x 128 x 16 MMA of a single CTA reads 8 KB of operands from shared
    // memory and takes 82 clk against the 64 clk tensor floor (tools/mma_probe.cu: ~100 B/clk of operand bandwidth); in a pair
    // each CTA supplies half of the weight tile (6 KB per MMA) and the floor is reached.
    int cg_env = 1;
#ifdef EMO_CONV_DEBUG
    { const char* e = getenv("EMO_CONV_CG2"); if (e) cg_env = atoi(e); }  // instrumented build: single-CTA MMAs for the probes
#endif
    p.cg = (cg_env == 1 && ksplit == 1 && (p.m_tiles % 2) == 0 && BN % 32 == 0) ? 2 : 1;
  }
  {
    // cluster size (single-CTA MMAs only): weight-tile multicast across consecutive pixel tiles.  Measured round 1: no gain at 2
    // and a loss at 4 (multicast does not lower the per-SM fill; pair mode above does): off in the product, reachable in the
    // instrumented build only
    int forced = 0;
#ifdef EMO_CONV_DEBUG
    { const char* e = getenv("EMO_CONV_CLUSTER"); if (e) forced = atoi(e); }
#endif
    int cs = (forced > 0 && ksplit == 1 && p.cg == 1) ? forced : 1;
    while (cs > 1 && (p.m_tiles % cs != 0 || (BN / cs) % 8 != 0 || BN % cs != 0 ||
                      (long long)p.m_tiles * p.n_tiles < 2ll * cs)) cs >>= 1;
    p.cs = cs;
  }
  p.bias = d->bias; p.residual = d->residual; p.res_shift = d->res_shift; p.act = d->act;
  p.rD = d->Dout; p.rH = d->Hout >> d->res_shift; p.rW = d->Wout >> d->res_shift;
  p.post_add = d->post_add; p.out = d->out; p.out_nchw = d->out_nchw;
  p.stats = d->stats; p.G = d->G; p.cpg = d->G > 0 ? d->Cout / d->G : 1;
  p.out_scale = f16 ? d->out_scale : 1.f;
  EMO_REQUIRE(!d->residual || ((d->Hout % (1 << d->res_shift)) == 0 && (d->Wout % (1 << d->res_shift)) == 0),
              "emo_conv_igemm: residual shift does not divide the output size");
  EMO_REQUIRE(!ps || p.cg == 2, "emo_conv_igemm: upconv needs an even number of pixel tiles and BN %% 32 == 0 (pair mode)");

  const size_t tail_bytes = (2 * kMaxStages + 2 * kAccBufs + 2 + 2 * kMaxAStages) * 8 + 16 + 8 * kMaxBN * sizeof(float);
  const size_t smem_limit = 227 * 1024;
  // Final phase of a tile (template parameter EPI, see the kernel's header comment).  TMA epilogue when the layer qualifies
  // (pair mode, full K loop, channels-last output in whole 32-channel panels, no post-add) and pays: per-layer CUDA-graph timings
  // of every shape of the driver frame (tools/conv_layer_bench.py, profiles/conv_layers_r2.md; re-measured after the staging
  // loop was rewritten, profiles/conv_layers_r2b.md) put it ahead for N tiles of >= 96 channels, with or without a residual,
  // except the single-wave layer with a half-resolution residual (128^2 320->320: 425 vs 450 TFLOP/s); narrow-N 3-D layers stay
  // with the in-warp form.
  int epi = 0;
  int want_epi = -1;
#ifdef EMO_CONV_DEBUG
  { const char* e = getenv("EMO_CONV_EPI"); if (e) want_epi = atoi(e); }  // instrumented build: force 0 / 1 for the A/B tools
#endif
  {
    const int want = want_epi;
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    const bool eligible = p.cg == 2 && ksplit == 1 && !d->out_nchw && d->Cout == d->Cout_pad && d->Cout % 32 == 0 && BN % 32 == 0 && !d->post_add &&
                          ((uintptr_t)d->out % 16) == 0 && (!d->residual || ((uintptr_t)d->residual % 16) == 0) && (!ps || d->N * (long long)gH < (1ll << 31));
    const bool res_ok = !d->residual || d->res_shift == 0 || (d->res_shift == 1 && !ps && p.td == 1 && p.tw % 2 == 0 && p.th % 2 == 0 && p.tw * p.th == kTileM);
    const bool pays = BN >= 96 && (!d->residual || d->res_shift == 0 || tiles >= 2ll * sm_count);
    if (eligible && res_ok && (want < 0 ? pays : want == 1)) epi = 1;
  }
  p.res_tma = (epi == 1 && d->residual && !ps) ? (d->res_shift == 0 ? 1 : 2) : 0;
  size_t staging = 0, stage_bytes = 0;
  int stages = 0;
  const int KC0 = KC;
  // row reuse needs: more than one tap row, unit stride along H, tiles inside one depth slice (a row shift is then one uniform
  // offset of the whole tile), whole swizzle atoms per image row of the tile, the full K loop in one CTA, no weight multicast
  const int kh_eff = ps ? 2 : d->kh;
  int yr_env = 1;
#ifdef EMO_CONV_DEBUG
  { const char* e = getenv("EMO_CONV_YREUSE"); if (e) yr_env = atoi(e); }  // instrumented build: per-tap form for the A/B tools
#endif
  bool want_yreuse = yr_env != 0 && kh_eff > 1 && d->sh == 1 && d->sw == 1 && d->sd == 1 && p.td == 1 && p.tw >= 8 && ksplit == 1 && p.cs == 1 && p.th + kh_eff - 1 <= 256;
  size_t ring_bytes = 0;
  for (;;) {
    staging = epi ? (size_t)(kTileM + (p.res_tma == 2 ? kTileM / 4 : 0)) * BN * sizeof(float) : 0;
    const size_t avail = smem_limit - tail_bytes - 1024 - staging;
    KC = KC0;
    if (KC == 64 && avail / ((size_t)NP * (kTileM + BN / p.cg) * 64 * 2) < 3) KC = 32;
    stage_bytes = (size_t)NP * (kTileM + BN / p.cg) * KC * 2;
    stages = (int)(avail / stage_bytes);
    if (epi && (stages < 3 || (ps && KC != 64))) { epi = 0; p.res_tma = 0; continue; }  // (the sub-pixel kernel is built for KC = 64)
    p.yreuse = 0;
    if (epi && want_yreuse && want_epi < 0 && KC0 == 64 &&
        2 * (size_t)NP * p.tw * (p.th + kh_eff - 1) * 64 * 2 + 3 * (size_t)NP * (BN / p.cg) * 64 * 2 > avail) {
      // the staging tile(s) would push the row-reuse rings down to 32-channel k-steps (N tile 160 with a half-resolution
      // residual: measured 84 us against 74 us in-warp): the in-warp final phase keeps the full-width rings
      epi = 0; p.res_tma = 0; continue;
    }
    if (want_yreuse) {
      // A ring of 2-3 halo tiles, the rest of the operand area as B ring (>= 3 tiles: one group's kh taps in flight)
      for (int KCy = KC0; KCy >= 32 && !p.yreuse; KCy >>= 1) {
        if (ps && KCy != 64) break;
        const size_t a_tile = (size_t)NP * p.tw * (p.th + kh_eff - 1) * KCy * 2, b_tile = (size_t)NP * (BN / p.cg) * KCy * 2;
        if (2 * a_tile + 3 * b_tile > avail || (p.tw * KCy * 2) % 1024 != 0) continue;
        const int sa = (3 * a_tile + 4 * b_tile <= avail) ? 3 : 2;
        int sb = (int)((avail - sa * a_tile) / b_tile);
        if (sb > kMaxStages) sb = kMaxStages;
        p.yreuse = 1; p.sa = sa; p.sb = sb; p.a_rows = p.tw * (p.th + kh_eff - 1);
        KC = KCy;
        ring_bytes = sa * a_tile + sb * b_tile;
        stages = sb;
      }
    }
    if (!p.yreuse) ring_bytes = (size_t)(stages > kMaxStages ? kMaxStages : stages) * stage_bytes;
    break;
  }
  p.kchunks = d->Cin / KC;
  if (stages > kMaxStages) stages = kMaxStages;
#ifdef EMO_CONV_DEBUG
  { const char* e = getenv("EMO_CONV_MAXSTAGES"); if (e && atoi(e) >= 2 && stages > atoi(e)) stages = atoi(e); }
#endif
  EMO_REQUIRE(stages >= 2, "emo_conv_igemm: tile does not fit shared memory (BN=%d KC=%d)", BN, KC);
  p.stages = stages;
  p.nbuf = 512 / BN > kAccBufs ? kAccBufs : 512 / BN;
  {
    // ~24 MMAs per accumulation chunk keeps the truncation bias of the tensor-core accumulator near 1e-6 relative
    const int mmas_per_kstep = (KC / 16) * (NP == 3 ? 6 : 3);
    const int target = d->acc_chunk_mmas > 0 ? d->acc_chunk_mmas : (NP == 3 ? 24 : 48);
    p.flush = target / mmas_per_kstep < 1 ? 1 : target / mmas_per_kstep;
  }
  const size_t smem_bytes = ring_bytes + tail_bytes + staging + 1024;

  // ---- tensor maps ----
  TMaps tm;
  memset(&tm, 0, sizeof(tm));
  {
    cuuint64_t gdim[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->Din, (cuuint64_t)d->N};
    cuuint64_t gstr[4] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Win * d->Cin * 2, (cuuint64_t)d->Hin * d->Win * d->Cin * 2,
                          (cuuint64_t)d->Din * d->Hin * d->Win * d->Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)(p.tw * d->sw), (cuuint32_t)(p.yreuse ? p.th + kh_eff - 1 : p.th * d->sh), (cuuint32_t)(p.td * d->sd), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)d->sw, (cuuint32_t)d->sh, (cuuint32_t)d->sd, 1};
    const CUtensorMapSwizzle sw = (KC == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const int taps = ps ? 16 : d->kd * d->kh * d->kw;  // sub-pixel mode: [4 phases][2x2 taps]
    cuuint64_t wdim[3] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Cout_pad, (cuuint64_t)taps};
    cuuint64_t wstr[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cout_pad * d->Cin * 2};
    cuuint32_t wbox[3] = {(cuuint32_t)KC, (cuuint32_t)(BN / p.cs / p.cg), 1};
    cuuint32_t wes[3] = {1, 1, 1};
    const void* ap[3] = {d->a_hi, d->a_lo, d->a_lo2};
    const void* wp[3] = {d->w_hi, d->w_lo, d->w_lo2};
    const CUtensorMapDataType plane_type = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    for (int pl = 0; pl < NP; ++pl) {
      EMO_REQUIRE(((uintptr_t)ap[pl] % 16) == 0 && ((uintptr_t)wp[pl] % 16) == 0, "emo_conv_igemm: planes must be 16-byte aligned");
      CUresult r1 = encode(&tm.a[pl], plane_type, 5, (void*)ap[pl], gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CUresult r2 = encode(&tm.b[pl], plane_type, 3, (void*)wp[pl], wdim, wstr, wbox, wes,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
        set_error("emo_conv_igemm: cuTensorMapEncodeTiled failed: A %d W %d (Cin=%d W=%d H=%d D=%d N=%d box=%d,%d,%d BN=%d)", (int)r1,
                  (int)r2, d->Cin, d->Win, d->Hin, d->Din, d->N, p.tw, p.th, p.td, BN);
        return EMO_ERR_CUDA;
      }
    }
  }

  if (epi == 1) {
    // output (and same-resolution residual) as 32-channel panels of the tile's pixel box; SWIZZLE_128B = the staging tile's
    // chunk order.  Sub-pixel mode: the [2 H][2 W] output seen as (C, column parity, W, row parity, N * H): one phase's pixels
    // of a low-resolution box are again a box.
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    const cuuint64_t C4 = (cuuint64_t)d->Cout * 4;
    if (ps) {
      gdim[0] = (cuuint64_t)d->Cout; gdim[1] = 2; gdim[2] = (cuuint64_t)gW; gdim[3] = 2; gdim[4] = (cuuint64_t)d->N * gH;
      gstr[0] = C4; gstr[1] = 2 * C4; gstr[2] = (cuuint64_t)d->Wout * C4; gstr[3] = 2 * (cuuint64_t)d->Wout * C4;
      box[0] = 32; box[1] = 1; box[2] = (cuuint32_t)p.tw; box[3] = 1; box[4] = (cuuint32_t)p.th;
    } else {
      gdim[0] = (cuuint64_t)d->Cout; gdim[1] = (cuuint64_t)d->Wout; gdim[2] = (cuuint64_t)d->Hout; gdim[3] = (cuuint64_t)d->Dout; gdim[4] = (cuuint64_t)d->N;
      gstr[0] = C4; gstr[1] = (cuuint64_t)d->Wout * C4; gstr[2] = (cuuint64_t)d->Hout * d->Wout * C4; gstr[3] = (cuuint64_t)d->Dout * d->Hout * d->Wout * C4;
      box[0] = 32; box[1] = (cuuint32_t)p.tw; box[2] = (cuuint32_t)p.th; box[3] = (cuuint32_t)p.td; box[4] = 1;
    }
    CUresult r1 = encode(&tm.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)d->out, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = CUDA_SUCCESS;
    if (p.res_tma == 2) {  // half-resolution residual [N][D][H/2][W/2][C], boxes of (tw/2) x (th/2) pixels
      gdim[1] = (cuuint64_t)(d->Wout >> 1); gdim[2] = (cuuint64_t)(d->Hout >> 1);
      gstr[1] = gdim[1] * C4; gstr[2] = gdim[2] * gstr[1]; gstr[3] = (cuuint64_t)d->Dout * gstr[2];
      box[1] = (cuuint32_t)(p.tw >> 1); box[2] = (cuuint32_t)(p.th >> 1);
    }
    if (p.res_tma)
      r2 = encode(&tm.res, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)d->residual, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      set_error("emo_conv_igemm: cuTensorMapEncodeTiled failed for the output / residual map: %d %d (Cout=%d W=%d H=%d D=%d N=%d box=%d,%d,%d)",
                (int)r1, (int)r2, d->Cout, d->Wout, d->Hout, d->Dout, d->N, p.tw, p.th, p.td);
      return EMO_ERR_CUDA;
    }
  }

  const int total_tiles = p.m_tiles * p.n_tiles * p.ksplit;
  int grid = total_tiles < sm_count ? total_tiles : sm_count;
  const int csz = p.cg == 2 ? 2 : p.cs;
  grid = (grid / csz) * csz;  // whole clusters only (total_tiles % csz == 0 by construction)
  cudaError_t e;
#define EMO_LAUNCH_CONV(KC_, NP_, CG_, EPI_) EMO_LAUNCH_CONV5(conv_igemm_kernel, KC_, NP_, CG_, EPI_)
#define EMO_LAUNCH_CONV5(KERNEL_, KC_, NP_, CG_, EPI_)                                                                                        \
  do {                                                                                                                    \
    static bool attr_set_dev[64] = {false}; /* the opt-in is per function AND per device (227 KB covers every configuration) */ \
    bool& attr_set = attr_set_dev[cur_dev & 63];                                                                          \
    if (!attr_set) {                                                                                                      \
      e = cudaFuncSetAttribute(KERNEL_<KC_, NP_, CG_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);      \
      if (e != cudaSuccess) { set_error("emo_conv_igemm: smem attribute: %s", cudaGetErrorString(e)); return EMO_ERR_CUDA; } \
      attr_set = true;                                                                                                    \
    }                                                                                                                     \
    cudaLaunchConfig_t cfg;                                                                                               \
    memset(&cfg, 0, sizeof(cfg));                                                                                         \
    cfg.gridDim = dim3((unsigned)grid);                                                                                   \
    cfg.blockDim = dim3(kThreads);                                                                                        \
    cfg.dynamicSmemBytes = smem_bytes;                                                                                    \
    cfg.stream = stream;                                                                                                  \
    cudaLaunchAttribute attr[1];                                                                                          \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                                     \
    attr[0].val.clusterDim.x = (unsigned)csz;                                                                             \
    attr[0].val.clusterDim.y = 1;                                                                                         \
    attr[0].val.clusterDim.z = 1;                                                                                         \
    cfg.attrs = attr;                                                                                                     \
    cfg.numAttrs = 1;                                                                                                     \
    e = cudaLaunchKernelEx(&cfg, KERNEL_<KC_, NP_, CG_, EPI_>, tm, p);                                                      \
    if (e != cudaSuccess) { set_error("emo_conv_igemm: launch: %s", cudaGetErrorString(e)); return EMO_ERR_CUDA; }         \
  } while (0)
  if (f16) {
    if (epi) {
      if (KC == 64) EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 64, 2, 2, 1);
      else EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 32, 2, 2, 1);
    } else if (p.cg == 2) {
      if (KC == 64) EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 64, 2, 2, 0);
      else EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 32, 2, 2, 0);
    } else {
      if (KC == 64) EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 64, 2, 1, 0);
      else EMO_LAUNCH_CONV5(conv_igemm_f16_kernel, 32, 2, 1, 0);
    }
  } else if (ps) {
    EMO_REQUIRE(KC == 64, "emo_conv_igemm: upconv tile does not fit with KC = 64");
    if (epi) EMO_LAUNCH_CONV5(conv_igemm_ps_kernel, 64, 2, 2, 1);
    else EMO_LAUNCH_CONV5(conv_igemm_ps_kernel, 64, 2, 2, 0);
  } else if (epi) {
    if (NP == 3 && KC == 64) EMO_LAUNCH_CONV(64, 3, 2, 1);
    else if (NP == 3) EMO_LAUNCH_CONV(32, 3, 2, 1);
    else if (KC == 64) EMO_LAUNCH_CONV(64, 2, 2, 1);
    else EMO_LAUNCH_CONV(32, 2, 2, 1);
  } else if (p.cg == 2) {
    if (NP == 3 && KC == 64) EMO_LAUNCH_CONV(64, 3, 2, 0);
    else if (NP == 3) EMO_LAUNCH_CONV(32, 3, 2, 0);
    else if (KC == 64) EMO_LAUNCH_CONV(64, 2, 2, 0);
    else EMO_LAUNCH_CONV(32, 2, 2, 0);
  } else {
    if (NP == 3 && KC == 64) EMO_LAUNCH_CONV(64, 3, 1, 0);
    else if (NP == 3) EMO_LAUNCH_CONV(32, 3, 1, 0);
    else if (KC == 64) EMO_LAUNCH_CONV(64, 2, 1, 0);
    else EMO_LAUNCH_CONV(32, 2, 1, 0);
  }
#undef EMO_LAUNCH_CONV
#undef EMO_LAUNCH_CONV5
  const long long per_n = (long long)d->Dout * d->Hout * d->Wout * d->Cout;
  const emo_apply_desc* post = d->post;
  if (ksplit > 1 && post && post->up == 1 && !d->out_nchw && !d->post_add && per_n <= (long long)kPostCluster * kPostThreads * kPostMaxV &&
      d->Cout <= kPostThreads && kPostThreads % d->Cout == 0 && (!post->stats || (post->G <= 64 && d->Cout % post->G == 0))) {
    // small split-K layer with a post-op: one cluster per sample finishes the convolution, normalises, activates and writes the
    // next convolution's operand planes (splitk_post_kernel)
    PostParams f;
    memset(&f, 0, sizeof(f));
    f.ws = d->splitk_ws; f.ksplit = ksplit; f.part_elems = p.ws_part_elems;
    f.N = d->N; f.C = d->Cout; f.S = (long long)d->Dout * d->Hout * d->Wout;
    f.bias = d->bias; f.residual = d->residual; f.conv_act = d->act;
    f.ap = *post;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(kPostCluster * d->N));
    cfg.blockDim = dim3(kPostThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPostCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, splitk_post_kernel, f);
    if (e != cudaSuccess) { set_error("emo_conv_igemm: post launch: %s", cudaGetErrorString(e)); return EMO_ERR_CUDA; }
    return check_launch("emo_conv_igemm");
  }
  if (ksplit > 1) {
    FinParams f;
    f.ws = d->splitk_ws; f.ksplit = ksplit; f.part_elems = p.ws_part_elems; f.out = d->out; f.N = d->N; f.C = d->Cout;
    f.S = (long long)d->Dout * d->Hout * d->Wout;
    f.bias = d->bias; f.residual = d->residual; f.post_add = d->post_add; f.act = d->act; f.out_nchw = d->out_nchw;
    f.stats = d->stats; f.G = d->G;
    long long bx = cdivll(f.S * f.C, 256 * 2);
    if (bx > 148 * 4) bx = 148 * 4;
    if (bx < 1) bx = 1;
    dim3 fg((unsigned)bx, (unsigned)d->N);
    launch_kernel(splitk_finalize_kernel, fg, 256, d->stats ? 2 * d->G * sizeof(double) : 0, stream, f);
  }
  if (post) {
    // every other layer: the post-op is the ordinary elementwise pass over the convolution's output
    int rc = check_launch("emo_conv_igemm");
    if (rc) return rc;
    emo_apply_desc ap = *post;
    ap.x = d->out;
    return emo_apply(&ap, stream_);
  }
  return check_launch("emo_conv_igemm");
}
