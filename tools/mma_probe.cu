// Microbenchmark (GPU box): how fast does ONE thread feed tcgen05.mma 128 x BN x 16 (kind::f16, bf16, both operands in shared
// memory, SWIZZLE_128B) in the conv kernel's issue pattern, and which part of the pattern costs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
// mode bits: 1 commit per k-step, 2 producer<->issuer mbarrier handshake per k-step (3 stages), 4 one MMA per 16-wide k
// slice instead of three, 8 the three MMAs of a slice go to different accumulators, 16 a second warp streams global memory
// into the stage buffers with st.shared while the MMAs run, 64 A operand from TMEM: each 16-wide k slice of both A planes is
// copied smem -> TMEM once (tcgen05.cp 128x256b) and the three MMAs read A from there (cta_group::1 only; next-round experiment).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
               "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t gtimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int CG, bool ELECT>
__global__ void __launch_bounds__(128) probe(int mode, int BN, int ksteps, int S, const uint4* __restrict__ src, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = 128 * 64 * 2, b_bytes = (uint32_t)(BN / CG) * 64 * 2, stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = (uint64_t*)(smem + S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* done = bars + 8;
  uint32_t* slot = (uint32_t*)(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  for (uint32_t i = threadIdx.x; i < S * stage_bytes / 4; i += blockDim.x) {
    uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
    h ^= h >> 15;
    ((uint32_t*)smem)[i] = (h & 0x3FFF3FFFu) | 0x30003000u;  // bf16 pairs of moderate magnitude, random mantissas
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const uint32_t tmem_base = *slot;
  const int nbuf = (mode & 64) ? (448 / BN > 4 ? 4 : 448 / BN) : (512 / BN > 4 ? 4 : 512 / BN);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  const bool hs = mode & 2;

  if (warp == 0 && lane == 0 && hs) {
    int stage = 0;
    uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive(&full[stage]);
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && crank == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t t0 = clock64(), g0 = gtimer();
    for (int ks = 0; ks < ksteps; ++ks) {
      if (hs) {
        mbar_wait(&full[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (ELECT ? elect_one() : (lane == 0)) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint64_t dA0 = kdesc(sa), dA1 = kdesc(sa + a_bytes), dB0 = kdesc(sa + 2 * a_bytes), dB1 = kdesc(sa + 2 * a_bytes + b_bytes);
        const int buf = (ks / 4) % nbuf;
        const uint32_t d0 = tmem_base + (uint32_t)(buf * BN);
        const uint32_t d1 = (mode & 8) ? tmem_base + (uint32_t)(((buf + 1) % nbuf) * BN) : d0;
        const uint32_t d2 = (mode & 8) ? tmem_base + (uint32_t)(((buf + 2) % nbuf) * BN) : d0;
        if (CG == 1 && (mode & 64)) {
          // A from TMEM: columns 448.. hold two double-buffered {A_hi, A_lo} k slices of 8 columns each
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const uint32_t acc0 = ((ks % 4) == 0 && kk == 0) ? 0u : 1u;
            const uint32_t ta = tmem_base + 448u + (uint32_t)((kk & 1) * 16);
            utccp_128x256b(ta, dA0 + adv);
            utccp_128x256b(ta + 8, dA1 + adv);
            umma_ts(d0, ta + 8, dB0 + adv, idesc, acc0);
            umma_ts(d0, ta, dB1 + adv, idesc, 1);
            umma_ts(d0, ta, dB0 + adv, idesc, 1);
          }
        } else
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);
          const uint32_t acc0 = ((ks % 4) == 0 && kk == 0) ? 0u : 1u;
          umma<CG>(d0, dA1 + adv, dB0 + adv, idesc, acc0);
          if (!(mode & 4)) {
            umma<CG>(d1, dA0 + adv, dB1 + adv, idesc, 1);
            umma<CG>(d2, dA0 + adv, dB0 + adv, idesc, 1);
          }
        }
        if (mode & 3) commit<CG>(&empty[stage]);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) commit<CG>(done);
    mbar_wait(done, 0);
    const uint64_t t1 = clock64(), g1 = gtimer();
    if (lane == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = g1 - g0;
    }
  } else if (warp == 2 && (mode & 16)) {
    // generic-proxy traffic into the stage buffers (stands in for the TMA fill: 64 KB per k-step)
    const uint4* s4 = src + (size_t)blockIdx.x * 65536;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint4* dst = (uint4*)(smem + (size_t)(ks % S) * stage_bytes);
      for (int i = lane; i < (int)(stage_bytes / 16); i += 32) {
        uint4 v = __ldg(s4 + ((ks * 4096 + i) & 65535));
        v.x = (v.x & 0x3FFF3FFFu) | 0x30003000u; v.y = (v.y & 0x3FFF3FFFu) | 0x30003000u;
        v.z = (v.z & 0x3FFF3FFFu) | 0x30003000u; v.w = (v.w & 0x3FFF3FFFu) | 0x30003000u;
        dst[i] = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int CG, bool ELECT>
static void run(int mode, int BN, int ksteps, int sms, const uint4* src, unsigned long long* dout) {
  const size_t stage = 2 * 128 * 64 * 2 + 2 * (size_t)(BN / CG) * 64 * 2;
  const int S = (int)(220 * 1024 / stage) > 3 ? 3 : (int)(220 * 1024 / stage);
  const size_t smem = S * stage + 1024 + 256;
  cudaFuncSetAttribute(probe<CG, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((sms / CG) * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG, ELECT>, mode, BN, ksteps, S, src, dout);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("mode %d BN %d: %s / %s\n", mode, BN, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[4];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas = (double)ksteps * ((mode & 4) ? 4 : 12);
  const double flops = mmas * 2.0 * 128 * CG * BN * 16 * (cfg.gridDim.x / CG);
  printf("%s cg %d mode %2d BN %3d ksteps %d: %.1f clk/MMA  %.1f ns/MMA  (SM clock %.0f MHz)  kernel %.3f ms  %.0f TFLOP/s chip\n", ELECT ? "elect" : "lane0", CG, mode, BN,
         ksteps, h[0] / mmas, h[1] / mmas, 1e3 * h[0] / (double)h[1], best, flops / (best * 1e-3) / 1e12);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* dout;
  cudaMalloc(&dout, 148 * 2 * sizeof(unsigned long long));
  uint4* src;
  cudaMalloc(&src, (size_t)148 * 65536 * 16);
  cudaMemset(src, 0x5a, (size_t)148 * 65536 * 16);
  const int ks = 20000;
  for (int BN : {128, 256, 64, 160})
    for (int mode : {0, 1, 3, 4, 8}) { run<1, false>(mode, BN, ks, sms, src, dout); run<1, true>(mode, BN, ks, sms, src, dout); }
  for (int BN : {32, 64, 128, 160})
    for (int mode : {0, 64, 65, 67}) run<1, true>(mode, BN, ks, sms, src, dout);  // A from shared memory vs A from TMEM
  for (int BN : {128, 256})
    for (int mode : {0, 1}) { run<2, false>(mode, BN, ks, sms, src, dout); run<2, true>(mode, BN, ks, sms, src, dout); }
  return 0;
}
