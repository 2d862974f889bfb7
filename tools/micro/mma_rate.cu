// Micro-benchmark: back-to-back tcgen05.mma issue rate on B200 as a function of cta_group, N, operand layout and
// accumulator reuse.  Operands are whatever is in shared memory (only the timing matters).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_rate tools/micro/mma_rate.cu
//   run  : tools/micro/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../mcvd_b200/csrc/umma_ptx.cuh"
using namespace mcvd::ptx;

struct Cfg {
  int cg;        // 1 | 2
  int N;
  int swz;       // 0: no-swizzle K-major (LBO = rows*16), 1: 128-byte swizzle K-major
  int nacc;      // accumulators cycled through
  int pattern;   // 0: one (A,B) pair repeated; 1: hi/lo triple (A_hi,B_hi),(A_lo,B_hi),(A_hi,B_lo); 2: triple + 9 shifted taps
  int iters;
};

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) k_rate(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;       // target of the per-tap commits (never waited on)
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { if (CG == 2) tmem_alloc2(smem_u32(&slot), 512); else tmem_alloc(smem_u32(&slot), 512); }
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && rank == 0) {
    if (elect_one()) {
      const int M = CG == 2 ? 256 : 128;
      const uint32_t idesc = make_idesc_f16(M, c.N);
      const int HP = 266;
      const int nb = CG == 2 ? c.N / 2 : c.N;                // B rows in this CTA
      const uint32_t a0 = smem_u32(smem), b0 = a0 + 100 * 1024;
      uint64_t ah, al, bh, bl;
      if (c.swz == 0) {
        ah = make_desc_sw(a0 + 67 * 16, HP * 16, 128, 0);
        al = make_desc_sw(a0 + 2 * HP * 16 * 2 + 67 * 16, HP * 16, 128, 0);
        bh = make_desc_sw(b0, nb * 16, 128, 0);
        bl = make_desc_sw(b0 + 2 * nb * 16, nb * 16, 128, 0);
      } else {                                               // SW128 K-major: rows of 128 B, 8-row groups 1024 B apart
        ah = make_desc_sw(a0, 16, 1024, 2);
        al = make_desc_sw(a0 + 32 * 1024, 16, 1024, 2);
        bh = make_desc_sw(b0, 16, 1024, 2);
        bl = make_desc_sw(b0 + 32 * 1024, 16, 1024, 2);
      }
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db) {
        if (CG == 2) umma2_f16(d, da, db, idesc, 1u); else umma_f16(d, da, db, idesc, 1u);
      };
      const uint32_t d0 = tm, d1 = tm + (uint32_t)((c.nacc > 1) ? c.N : 0);
      const long long t0 = clock64();
      if (c.pattern == 0) {                       // one (A, B) pair, 18 MMAs per trip
        for (int it = 0; it < c.iters; it += 18) {
#pragma unroll
          for (int u = 0; u < 9; ++u) { mma(d0, ah, bh); mma(d1, ah, bh); }
        }
      } else if (c.pattern == 1) {                // hi/lo triple, fixed addresses
        for (int it = 0; it < c.iters; it += 6) {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            mma(d0, ah, bh); mma(d0, al, bh); mma(d0, ah, bl);
            mma(d1, ah, bh); mma(d1, al, bh); mma(d1, ah, bl);
          }
        }
      } else if (c.pattern == 6 || c.pattern == 7) {
        // the conv kernel's exact operand geometry (KB = 16, HP = 260, two tiles per unit, 16 weight stages of
        // 32*N bytes, 4 accumulators), no barriers: pattern 6 with the per-tap commit, 7 without
        const int HPk = 260;
        const uint32_t a_plane16 = 2 * HPk, a_stage16 = 2 * a_plane16, b_stage16 = (32u * c.N) >> 4, b_lo16 = (16u * c.N) >> 4;
        const uint64_t ap = make_desc_sw(a0 + 66 * 16, HPk * 16, 128, 0), bp = make_desc_sw(b0, nb * 16, 128, 0);
        const uint32_t cb = smem_u32(&bar2);
        int bst = 0, ast = 0, unit = 0;
        for (int it = 0; it < c.iters; it += 54) {          // one K-block: 9 taps x 2 tiles x 3 MMAs
          const uint32_t dset = tm + (uint32_t)((unit & 1) * 2 * c.N);
          int dy = -1, dx = -1;
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t sh = (uint64_t)(int64_t)(dy * 65 + dx);
            const uint64_t bh = bp + (uint64_t)(bst * b_stage16), bl = bh + b_lo16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t ah_ = ap + (uint64_t)(((ast + j) & 3) * a_stage16) + sh, al_ = ah_ + a_plane16;
              const uint32_t d = dset + (uint32_t)(j * c.N);
              mma(d, ah_, bh); mma(d, al_, bh); mma(d, ah_, bl);
            }
            if (c.pattern == 6) umma2_commit_mc(cb);
            if (++bst == 16) bst = 0;
            if (++dx == 2) { dx = -1; ++dy; }
          }
          ast = (ast + 2) & 3;
          if ((it / 54) % 6 == 5) ++unit;
        }
      } else if (c.pattern >= 3) {                // 6 MMAs + one commit per "tap" (the conv kernel's stage release)
        const uint32_t cb = smem_u32(&bar2);
        const uint32_t cb_remote = CG == 2 ? mapa_u32(cb, 1) : cb;
        for (int it = 0; it < c.iters; it += 2) {
          mma(d0, ah, bh); mma(d0, al, bh); mma(d0, ah, bl);
          mma(d0, ah + 1, bh); mma(d0, al + 1, bh); mma(d0, ah + 1, bl);
          if (CG == 1) umma_commit(cb);
          else if (c.pattern == 3) umma2_commit_mc(cb);                       // multicast to both CTAs
          else if (c.pattern == 4)                                            // leader's barrier only
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cb) : "memory");
          else if (c.pattern == 5) {                                          // two unicast commits
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cb) : "memory");
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cb_remote) : "memory");
          }
        }
      } else {                                    // hi/lo triple over nine shifted views (3x3 taps)
        for (int it = 0; it < c.iters; it += 9) {
#pragma unroll
          for (int u = 0; u < 9; ++u) {
            const uint64_t sh = (uint64_t)((u / 3) * 65 + (u % 3));
            const uint32_t d = (u & 1) ? d1 : d0;
            mma(d, ah + sh, bh); mma(d, al + sh, bh); mma(d, ah + sh, bl);
          }
        }
      }
      if (CG == 2) umma2_commit_mc(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      const long long t1 = clock64();
      if (out) out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) { tc_fence_after(); if (CG == 2) tmem_dealloc2(tm, 512); else tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ns[] = {16, 32, 64, 96, 128, 144, 192, 256};
  printf("cycles per tcgen05.mma (K=16, fp16), mean over issuing CTAs; ideal tensor time = M*N/ (128*2) per SM = N/2 (M=128 per SM)\n");
  for (int swz = 0; swz < 2; ++swz)
    for (int cg = 1; cg <= 2; ++cg)
      for (int pattern = (swz ? 8 : 5); pattern < 8; ++pattern)
        for (int nacc = 1; nacc <= (pattern < 3 ? 2 : 1); ++nacc)
          for (int N : Ns) {
            if (nacc * N > 512) continue;
            if (cg == 1 && pattern > 3) continue;
            if (pattern >= 6 && 4 * N > 512) continue;
            Cfg c{cg, N, swz, nacc, pattern, 1782};
            const int mmas = c.iters * (pattern == 0 ? 1 : 3);
            cudaMemset(out, 0, 148 * sizeof(long long));
            for (int rep = 0; rep < 2; ++rep) {
              if (cg == 1) k_rate<1><<<148, 128, 200 * 1024>>>(c, out);
              else {
                cudaLaunchConfig_t lc = {};
                lc.gridDim = dim3(148); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 200 * 1024;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                cudaLaunchKernelEx(&lc, k_rate<2>, c, out);
              }
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s (swz %d cg %d N %d)\n", cudaGetErrorString(e), swz, cg, N); return 1; }
            long long h[148];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; int n = 0;
            for (int i = 0; i < 148; ++i) if (h[i] > 0) { s += (double)h[i]; ++n; }
            printf("swz %d cg %d pattern %d nacc %d N %3d : %7.1f cycles/mma  (%d issuing CTAs)  per-SM rows*N/cycle = %.1f\n", swz, cg, pattern,
                   nacc, N, s / n / mmas, n, 128.0 * N / (s / n / mmas));
          }
  return 0;
}
