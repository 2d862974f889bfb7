// Final 3x3 convolution of the score network: nf -> C*F (5..15) output channels, with the last
// GroupNorm(affine)+SiLU (reference ncsnpp_more.py:375-379, layerspp.py:539-549) fused into the
// input fetch so the normalised activation is never written to HBM.  N is far too small for a GEMM
// tile: one thread owns one output pixel and all (<= 16) output channels; weights sit in shared
// memory and are read as broadcasts.
#include "mcvd_common.cuh"

namespace mcvd {

namespace {

template <int NP>
__global__ void __launch_bounds__(128) k_conv_smalln(const float* __restrict__ src, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const float4* __restrict__ tab,
                                                     float* __restrict__ dst, int B, int H, int W, int Cin, int Cout,
                                                     int act) {
  extern __shared__ __align__(16) float ws[];  // [9][Cin][NP]
  for (int i = threadIdx.x; i < 9 * Cin * NP; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  long long M = (long long)B * H * W;
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int x = (int)(m % W), y = (int)((m / W) % H);
  int b = (int)(m / ((long long)W * H));
  float acc[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = 0.f;
  const float4* tb = tab ? tab + (long long)b * Cin : nullptr;
  for (int tap = 0; tap < 9; ++tap) {
    int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* sp = src + (((long long)b * H + yy) * W + xx) * Cin;
    const float* wp = ws + tap * Cin * NP;
    for (int c = 0; c < Cin; c += 4) {
      float4 v4 = *reinterpret_cast<const float4*>(sp + c);
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = v[k];
        if (tb) {
          float4 t = tb[c + k];
          a = ((a - t.x) * t.y) * t.z + t.w;
          if (act) a = silu_f(a);
        }
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
          float4 wv = *reinterpret_cast<const float4*>(wp + (c + k) * NP + j);
          acc[j + 0] = fmaf(a, wv.x, acc[j + 0]);
          acc[j + 1] = fmaf(a, wv.y, acc[j + 1]);
          acc[j + 2] = fmaf(a, wv.z, acc[j + 2]);
          acc[j + 3] = fmaf(a, wv.w, acc[j + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NP; ++j)
    if (j < Cout) dst[m * Cout + j] = acc[j] + (bias ? bias[j] : 0.f);
}

template <int NP>
int launch_np(const McvdOp& op, cudaStream_t s) {
  size_t smem = (size_t)9 * op.C0 * NP * sizeof(float);
  MCVD_CHECK(smem <= 227 * 1024, "CONV_SMALLN: weights (%zu B) exceed shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(k_conv_smalln<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  MCVD_CHECK(e == cudaSuccess, "CONV_SMALLN: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  long long M = (long long)op.B * op.H * op.W;
  k_conv_smalln<NP><<<(unsigned)((M + 127) / 128), 128, smem, s>>>(
      (const float*)op.src0, (const float*)op.w, (const float*)op.bias, (const float4*)op.aux0, (float*)op.dst, op.B,
      op.H, op.W, op.C0, op.Cout, (op.flags & MCVD_F_ACT_OUT) ? 1 : 0);
  MCVD_CUDA_LAUNCH_CHECK("conv_smalln");
  return 0;
}

}  // namespace

int launch_conv_smalln(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.w && op.dst, "CONV_SMALLN: null pointer");
  MCVD_CHECK(op.C0 % 4 == 0 && op.C1 == 0, "CONV_SMALLN: Cin must be a multiple of 4, single source");
  MCVD_CHECK(op.Cout <= 16 && op.i1 >= op.Cout, "CONV_SMALLN: Cout %d > 16", op.Cout);
  switch (op.i1) {  // padded Cout of the packed weights
    case 4: return launch_np<4>(op, s);
    case 8: return launch_np<8>(op, s);
    case 12: return launch_np<12>(op, s);
    case 16: return launch_np<16>(op, s);
    default: break;
  }
  set_error("CONV_SMALLN: padded Cout %d unsupported", op.i1);
  return -1;
}

}  // namespace mcvd
