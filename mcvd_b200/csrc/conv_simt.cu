// Direct convolution (3x3 pad 1 / 1x1) as an implicit GEMM on the CUDA cores, fp32 FFMA, NHWC.
// This is the generic path: exact fp32, any channel count, used for the skinny layers (first conv
// K = 9*C*(F+Fc), SPADE cond convs) and as the in-GPU cross-check of the tensor-core kernel.
//
//   M = B*H*W output pixels, N = Cout, K = taps * (C0 + C1)     (src0 | src1 = virtual channel concat)
//   dst = f0 * (A*W + bias + residual)          [optional SiLU]
//
// Tile 128 (pixels) x 64 (couts) x 16 (k) per 256-thread CTA, 8x4 register micro-tile, register
// prefetch of the next k-slab while the current one is consumed from shared memory.
#include "mcvd_common.cuh"

namespace mcvd {

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int AS_LD = BM + 4;  // 132 floats: 16-byte aligned rows

struct ConvArgs {
  const float* s0;
  const float* s1;
  const float* w;     // [taps][Cin][CoutP]
  const float* bias;  // [Cout] or null
  const float* res;   // [M][Cout] or null
  float* dst;         // [M][Cout]
  int B, H, W, C0, C1, Cout, CoutP, ks, flags;
  float scale;
};

__device__ __forceinline__ float4 ld_a4(const ConvArgs& a, long long pix, int ch, int Cin, bool vec) {
  // 4 consecutive input channels ch..ch+3 of pixel `pix` (zero beyond Cin)
  if (vec) {
    if (ch >= Cin) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < a.C0) return *reinterpret_cast<const float4*>(a.s0 + pix * a.C0 + ch);
    return *reinterpret_cast<const float4*>(a.s1 + pix * a.C1 + (ch - a.C0));
  }
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = ch + i;
    float v = 0.f;
    if (c < a.C0) v = a.s0[pix * a.C0 + c];
    else if (c < Cin) v = a.s1[pix * a.C1 + (c - a.C0)];
    r[i] = v;
  }
  return make_float4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(256) k_conv_simt(ConvArgs a) {
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int Cin = a.C0 + a.C1;
  const long long M = (long long)a.B * a.H * a.W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const bool vec = (a.C0 % 4 == 0) && (a.C1 % 4 == 0);
  const int kchunks = cdiv(Cin, BK);
  const int taps = a.ks * a.ks;
  const int total = taps * kchunks;
  const int padk = a.ks / 2;

  // A loader: two float4 per thread: (pixel, kq) = (idx / 4, idx % 4), idx = tid, tid + 256
  int apix[2], akq[2], ay[2], ax[2];
  long long ab[2];
  bool avalid[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int idx = tid + r * 256;
    apix[r] = idx >> 2;
    akq[r] = idx & 3;
    long long m = m0 + apix[r];
    avalid[r] = m < M;
    long long mm = avalid[r] ? m : 0;
    ax[r] = (int)(mm % a.W);
    ay[r] = (int)((mm / a.W) % a.H);
    ab[r] = mm / ((long long)a.W * a.H);
  }
  // B loader: one float4 per thread: (k, n4) = (tid / 16, tid % 16)
  const int bk = tid >> 4, bn4 = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;

  auto fetch = [&](int it) {
    int tap = it / kchunks, kc = it % kchunks;
    int dy = tap / a.ks - padk, dx = tap % a.ks - padk;
    int c0 = kc * BK;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int yy = ay[r] + dy, xx = ax[r] + dx;
      if (avalid[r] && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
        long long pix = (ab[r] * a.H + yy) * a.W + xx;
        ra[r] = ld_a4(a, pix, c0 + akq[r] * 4, Cin, vec);
      } else {
        ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    int kk = c0 + bk, nn = n0 + bn4 * 4;
    if (kk < Cin && nn < a.CoutP)
      rb = *reinterpret_cast<const float4*>(a.w + ((long long)tap * Cin + kk) * a.CoutP + nn);
    else
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int k = akq[r] * 4;
      As[buf][k + 0][apix[r]] = ra[r].x;
      As[buf][k + 1][apix[r]] = ra[r].y;
      As[buf][k + 2][apix[r]] = ra[r].z;
      As[buf][k + 3][apix[r]] = ra[r].w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][bk][bn4 * 4]) = rb;
  };

  fetch(0);
  stash(0);
  __syncthreads();

  for (int it = 0; it < total; ++it) {
    int buf = it & 1;
    if (it + 1 < total) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < total) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const int nn = n0 + tx * 4;
  if (nn >= a.Cout) return;
  const bool full4 = (nn + 3 < a.Cout) && (a.Cout % 4 == 0);
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nn + j < a.Cout) bv[j] = a.bias[nn + j];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bv[j];
    if (full4) {
      if (a.res) {
        float4 r = *reinterpret_cast<const float4*>(a.res + m * a.Cout + nn);
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] *= a.scale;
        if (a.flags & MCVD_F_ACT_OUT) v[j] = silu_f(v[j]);
      }
      *reinterpret_cast<float4*>(a.dst + m * a.Cout + nn) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nn + j < a.Cout) {
          float o = v[j];
          if (a.res) o += a.res[m * a.Cout + nn + j];
          o *= a.scale;
          if (a.flags & MCVD_F_ACT_OUT) o = silu_f(o);
          a.dst[m * a.Cout + nn + j] = o;
        }
      }
    }
  }
}

}  // namespace

int launch_conv_simt(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.w && op.dst && (op.C1 == 0 || op.src1), "CONV_SIMT: null pointer");
  MCVD_CHECK(op.i0 == 1 || op.i0 == 3, "CONV_SIMT: kernel size %d unsupported", op.i0);
  MCVD_CHECK(op.i1 >= op.Cout && op.i1 % 4 == 0, "CONV_SIMT: padded Cout %d invalid for Cout %d", op.i1, op.Cout);
  ConvArgs a;
  a.s0 = (const float*)op.src0; a.s1 = (const float*)op.src1; a.w = (const float*)op.w;
  a.bias = (const float*)op.bias; a.res = (const float*)op.aux0; a.dst = (float*)op.dst;
  a.B = op.B; a.H = op.H; a.W = op.W; a.C0 = op.C0; a.C1 = op.C1; a.Cout = op.Cout; a.CoutP = op.i1;
  a.ks = op.i0; a.flags = op.flags; a.scale = op.f0;
  long long M = (long long)op.B * op.H * op.W;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)cdiv(op.Cout, BN));
  k_conv_simt<<<grid, 256, 0, s>>>(a);
  MCVD_CUDA_LAUNCH_CHECK("conv_simt");
  return 0;
}

}  // namespace mcvd
