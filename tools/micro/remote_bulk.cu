// Micro test: may a cp.async.bulk (global -> shared::cluster) write into the PEER CTA's shared memory while
// signalling complete_tx on an mbarrier of the ISSUING CTA?  (Would let the leader of a CTA pair load both halves
// of a weight tile and wait on one local barrier.)  Also times local vs remote-destination copies.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/remote_bulk tools/micro/remote_bulk.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../mcvd_b200/csrc/umma_ptx.cuh"
using namespace mcvd::ptx;

__global__ void __cluster_dims__(2, 1, 1) k(const uint32_t* src, int bytes, long long* out, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xdeadbeefu;
  __syncthreads();
  cluster_sync_all();
  long long t = 0;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t dst = mode == 0 ? smem_u32(smem) : mapa_u32(smem_u32(smem), 1);
    const long long t0 = clock64();
    for (int it = 0; it < 16; ++it) {
      mbar_arrive_expect_tx(smem_u32(&bar), bytes);
      bulk_g2s(dst, src, bytes, smem_u32(&bar));
      mbar_wait(smem_u32(&bar), it & 1);
    }
    t = clock64() - t0;
  }
  __syncthreads();
  cluster_sync_all();
  // verify in the CTA that should have received the data
  const uint32_t want_rank = mode == 0 ? 0 : 1;
  int bad = 0;
  if (rank == want_rank)
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) bad += reinterpret_cast<uint32_t*>(smem)[i] != src[i];
  bad = __syncthreads_count(bad);
  if (threadIdx.x == 0) { if (rank == 0) out[blockIdx.x / 2 * 4 + 0] = t; if (rank == want_rank) out[blockIdx.x / 2 * 4 + 1] = bad; }
}

int main() {
  const int bytes = 12288;
  uint32_t* src; long long* out;
  cudaMalloc(&src, bytes); cudaMalloc(&out, 74 * 4 * sizeof(long long));
  uint32_t h[bytes / 4];
  for (int i = 0; i < bytes / 4; ++i) h[i] = i * 2654435761u;
  cudaMemcpy(src, h, bytes, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(out, 0xff, 74 * 4 * sizeof(long long));
    k<<<148, 128, bytes>>>(src, bytes, out, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long r[74 * 4];
    cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
    double t = 0; long long bad = 0;
    for (int c = 0; c < 74; ++c) { t += (double)r[c * 4]; bad += r[c * 4 + 1]; }
    printf("mode %d (%s destination, local mbarrier): %s, mismatching words %lld, %.0f cycles per 12 KB copy round trip\n", mode,
           mode ? "PEER" : "local", cudaGetErrorString(e), bad, t / 74 / 16);
  }
  return 0;
}
