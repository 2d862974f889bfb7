// Multi-head self-attention over the H*W tokens of a sample on the tcgen05 tensor cores, flash style
// (reference AttnBlockpp.forward, models/better/layerspp.py:239-245: w = softmax_j(q_i.k_j * Ch^-0.5),
// h_i = sum_j w_ij v_j; the reference materialises w as [B*heads, HW, HW] fp32).
//
//   qkv [B, T, 3C] fp32 (q | k | v along channels; head h = channels [h*d, (h+1)*d)),  out [B, T, C]
//   grid = (ceil(T / 128), heads, B); one CTA owns 128 query rows and walks the keys in tiles of KT.
//
// Per key tile:   S = Q K^T   (tcgen05.mma, M=128, N=KT, K=d)          -> TMEM columns [0, KT)
//                 P = exp(S*scale - m)  by the softmax warps (query row = TMEM lane: no shuffles),
//                     written to smem as the A operand of the next MMA
//                 O_tile = P V (M=128, N=d, K=KT)                       -> TMEM columns [KT, KT+d)
//                 o = o * exp(m_old - m_new) + O_tile   in registers
// fp32 parity: q, k, v and p are split into fp16 hi + lo and every product is 3 MMAs
// (hi*hi + lo*hi + hi*lo), fp32 accumulation -- same scheme as conv_umma.cu.
//
// Operand layouts (canonical K-major, no swizzle): [k-chunk of 8 halfs][row][16 B], LBO = rows*16,
// SBO = 128.   V is staged transposed (rows = channels, k = keys) so it is K-major too.
//
// Two launches per op:
//   k_attn_presplit  converts q, k, v ONCE into fp16 hi/lo "operand images" in a scratch buffer, one image per
//                    (sample, head, tile), byte-for-byte what the MMA wants in shared memory.  (Converting
//                    inside the attention kernel re-did the K/V tiles for each of the T/128 query tiles
//                    and left it bound by the latency of the staging loops: 400 us at 32x32 in cfg2.)
//   k_attention_umma one thread streams the images in with cp.async.bulk (mbarrier complete_tx).
// Warps: 0-7 softmax / output: query row r <-> TMEM lane r is shared by TWO threads (warp w and w+4 both own
//        lane group w%4, the only TMEM lanes either may touch); each takes half of the key columns of S and half of
//        the channels of O, and they meet through shared memory for the row maximum and the final row sum.  (One
//        thread per row left a single warp per SMSP running a dependent MUFU / tcgen05.ld chain: 29 % issue slots.)
//        8 image loader, 9 TMEM allocation + MMA issue.
#include "mcvd_common.cuh"
#include "umma_ptx.cuh"

namespace mcvd {

namespace {

using namespace ptx;

constexpr int QT = 128;
constexpr int ATT_THREADS = 320;       // 8 softmax warps + image loader + MMA issuer
constexpr int SOFTMAX_THREADS = 256;
constexpr int SPLIT_THREADS = 256;

struct AttnArgs {
  const float* qkv;
  float* out;
  uint8_t* img;                 // operand images: [Q tiles | K tiles | V tiles], see image_offsets()
  long long k_off, v_off;       // byte offsets of the K and V image arrays
  int T, C, KT, nkt, nqt, tmem_cols;
  float scale;
};

// fp32 x8 -> fp16 hi (16 B) + lo (16 B)
__device__ __forceinline__ void split8(const float v[8], uint4& hv, uint4& lv) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
  hv = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// stage `rows` token rows x D channels (channel-contiguous in global) as a K-major operand
//   dst[(c8 * rows + r) * 16]  <-  src[(row0 + r) * stride + c8 * 8 .. +8)
template <int D>
__device__ __forceinline__ void stage_rows(uint8_t* hi, uint8_t* lo, const float* src, long long stride, int rows,
                                           int valid_rows, int tid, int nthreads) {
  constexpr int U = 6;                                   // units in flight per thread (12 x LDG.128)
  const int units = (D / 8) * rows;
  for (int u0 = tid; u0 < units; u0 += nthreads * U) {
    float4 va[U], vb[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int u = u0 + i * nthreads;
      const int c8 = u / rows, r = u - c8 * rows;
      va[i] = vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < units && r < valid_rows) {
        const float4* p = reinterpret_cast<const float4*>(src + (long long)r * stride + c8 * 8);
        va[i] = __ldg(p);
        vb[i] = __ldg(p + 1);
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int u = u0 + i * nthreads;
      if (u < units) {
        const int c8 = u / rows, r = u - c8 * rows;
        const float v[8] = {va[i].x, va[i].y, va[i].z, va[i].w, vb[i].x, vb[i].y, vb[i].z, vb[i].w};
        uint4 hv, lv;
        split8(v, hv, lv);
        const size_t off = ((size_t)c8 * rows + r) * 16;
        *reinterpret_cast<uint4*>(hi + off) = hv;
        *reinterpret_cast<uint4*>(lo + off) = lv;
      }
    }
  }
}

// V^T image: rows = channels, k = keys.  unit = (key chunk kc of 8 keys, 4 channels): 8 x LDG.128 (one per
// key, 4 channels each), transposed in registers into 4 rows of 8 keys; 2 units (16 loads) in flight.
template <int D>
__device__ __forceinline__ void stage_v(uint8_t* vh, uint8_t* vl, const float* vsrc, long long stride, int KT, int tid,
                                        int nthreads) {
  constexpr int C4 = D / 4;
  const int vunits = (KT / 8) * C4;
  for (int u0 = tid; u0 < vunits; u0 += 2 * nthreads) {
    float4 ld[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int u = u0 + i * nthreads;
      const int kc = u / C4, c4 = u - kc * C4;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        ld[i][e] = (u < vunits) ? __ldg(reinterpret_cast<const float4*>(vsrc + (long long)(kc * 8 + e) * stride) + c4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int u = u0 + i * nthreads;
      if (u < vunits) {
        const int kc = u / C4, c4 = u - kc * C4;
        const float r0[8] = {ld[i][0].x, ld[i][1].x, ld[i][2].x, ld[i][3].x, ld[i][4].x, ld[i][5].x, ld[i][6].x, ld[i][7].x};
        const float r1[8] = {ld[i][0].y, ld[i][1].y, ld[i][2].y, ld[i][3].y, ld[i][4].y, ld[i][5].y, ld[i][6].y, ld[i][7].y};
        const float r2[8] = {ld[i][0].z, ld[i][1].z, ld[i][2].z, ld[i][3].z, ld[i][4].z, ld[i][5].z, ld[i][6].z, ld[i][7].z};
        const float r3[8] = {ld[i][0].w, ld[i][1].w, ld[i][2].w, ld[i][3].w, ld[i][4].w, ld[i][5].w, ld[i][6].w, ld[i][7].w};
        uint4 hv, lv;
        const size_t off = ((size_t)kc * D + c4 * 4) * 16;
        split8(r0, hv, lv); *reinterpret_cast<uint4*>(vh + off) = hv;      *reinterpret_cast<uint4*>(vl + off) = lv;
        split8(r1, hv, lv); *reinterpret_cast<uint4*>(vh + off + 16) = hv; *reinterpret_cast<uint4*>(vl + off + 16) = lv;
        split8(r2, hv, lv); *reinterpret_cast<uint4*>(vh + off + 32) = hv; *reinterpret_cast<uint4*>(vl + off + 32) = lv;
        split8(r3, hv, lv); *reinterpret_cast<uint4*>(vh + off + 48) = hv; *reinterpret_cast<uint4*>(vl + off + 48) = lv;
      }
    }
  }
}

// image sizes in bytes (hi plane + lo plane)
template <int D> __host__ __device__ constexpr long long q_image_bytes() { return 2LL * (D / 8) * QT * 16; }
template <int D> __host__ __device__ inline long long kv_image_bytes(int KT) { return 2LL * (D / 8) * KT * 16; }

// Pre-pass: grid (nqt + 2 nkt, heads, B); block x builds one Q, K or V image of (sample, head).
template <int D>
__global__ void __launch_bounds__(SPLIT_THREADS) k_attn_presplit(const AttnArgs a) {
  const int x = blockIdx.x, h = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
  const int C3 = 3 * a.C, KT = a.KT;
  const float* base = a.qkv + (long long)b * a.T * C3 + h * D;
  const long long bh = (long long)b * heads + h;
  if (x < a.nqt) {
    uint8_t* dst = a.img + (bh * a.nqt + x) * q_image_bytes<D>();
    const int q0 = x * QT;
    stage_rows<D>(dst, dst + q_image_bytes<D>() / 2, base + (long long)q0 * C3, C3, QT, min(QT, a.T - q0), threadIdx.x,
                  SPLIT_THREADS);
  } else if (x < a.nqt + a.nkt) {
    const int kt = x - a.nqt;
    uint8_t* dst = a.img + a.k_off + (bh * a.nkt + kt) * kv_image_bytes<D>(KT);
    stage_rows<D>(dst, dst + kv_image_bytes<D>(KT) / 2, base + a.C + (long long)kt * KT * C3, C3, KT, KT, threadIdx.x,
                  SPLIT_THREADS);
  } else {
    const int kt = x - a.nqt - a.nkt;
    uint8_t* dst = a.img + a.v_off + (bh * a.nkt + kt) * kv_image_bytes<D>(KT);
    stage_v<D>(dst, dst + kv_image_bytes<D>(KT) / 2, base + 2 * a.C + (long long)kt * KT * C3, C3, KT, threadIdx.x,
               SPLIT_THREADS);
  }
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1) k_attention_umma(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int KT = a.KT;
  const uint32_t q_half = (D / 8) * QT * 16, k_half = (D / 8) * KT * 16, v_half = (KT / 8) * D * 16,
                 p_half = (KT / 8) * QT * 16;
  uint8_t* qh = smem_raw;           uint8_t* ql = qh + q_half;
  uint8_t* kh = ql + q_half;        uint8_t* kl = kh + k_half;
  uint8_t* vh = kl + k_half;        uint8_t* vl = vh + v_half;
  uint8_t* ph = vl + v_half;        uint8_t* pl = ph + p_half;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pl + p_half);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t K_FULL = bar0, V_FULL = bar0 + 8, P_FULL = bar0 + 16, S_FULL = bar0 + 24, O_FULL = bar0 + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* xmax = reinterpret_cast<float*>(bars + 8);        // [tile parity][half][QT] partial row maxima
  float* xsum = xmax + 4 * QT;                             // [half][QT] partial row sums

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const long long bh = (long long)b * gridDim.y + h;

  if (tid == 0) {
    mbar_init(K_FULL, 1); mbar_init(V_FULL, 1); mbar_init(P_FULL, SOFTMAX_THREADS);
    mbar_init(S_FULL, 1); mbar_init(O_FULL, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + (uint32_t)KT;

  if (warp < 8) {
    // ================= softmax + output accumulation: two threads per query row =================
    const int lg = warp & 3, hf = warp >> 2;                // TMEM lane group, column / channel half
    const int r = lg * 32 + (tid & 31);
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    constexpr int D0 = ((D / 2 + 15) / 16) * 16;            // channels of half 0 (half 1 takes the rest)
    const int oc0 = hf ? D0 : 0, n_o = hf ? D - D0 : D0;
    const int half_cols = KT / 2;                           // 64 or 32 key columns per thread and tile
    const int col0 = hf * half_cols;
    float o[D0];
#pragma unroll
    for (int i = 0; i < D0; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;                           // l: this thread's share of the row sum
    for (int j = 0; j < a.nkt; ++j) {
      mbar_wait(S_FULL, j & 1);
      tc_fence_after();
      // this thread's columns of S(j): read once, kept in registers across the max exchange
      uint32_t rr[64];                                      // half_cols = 16 (head dim 192), 32 or 64 of them are live
      tmem_ld16(tS + lane_off + col0, rr);
      if (half_cols >= 32) tmem_ld16(tS + lane_off + col0 + 16, rr + 16);
      if (half_cols == 64) {
        tmem_ld16(tS + lane_off + col0 + 32, rr + 32);
        tmem_ld16(tS + lane_off + col0 + 48, rr + 48);
      }
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(rr[e]));
      if (half_cols >= 32) {
#pragma unroll
        for (int e = 16; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(rr[e]));
      }
      if (half_cols == 64) {
#pragma unroll
        for (int e = 32; e < 64; ++e) mx = fmaxf(mx, __uint_as_float(rr[e]));
      }
      float* xm = xmax + (j & 1) * 2 * QT;                  // double-buffered by tile parity
      xm[hf * QT + r] = mx;
      named_bar_sync(1, SOFTMAX_THREADS);
      mx = fmaxf(mx, xm[(hf ^ 1) * QT + r]);
      // scale > 0, so max(s)*scale == max(s*scale)
      const float mnew = fmaxf(m, mx * a.scale);
      const float corr = __expf(m - mnew);
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (g * 8 < half_cols) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = __expf(fmaf(__uint_as_float(rr[g * 8 + e]), a.scale, -mnew));
            sum += p[e];
          }
          uint4 hv, lv;
          split8(p, hv, lv);
          const size_t off = ((size_t)(col0 / 8 + g) * QT + r) * 16;
          *reinterpret_cast<uint4*>(ph + off) = hv;
          *reinterpret_cast<uint4*>(pl + off) = lv;
        }
      }
      l = l * corr + sum;
      m = mnew;
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(P_FULL);
      mbar_wait(O_FULL, j & 1);
      tc_fence_after();
      {
        uint32_t ro[D0];
#pragma unroll
        for (int c = 0; c < D0 / 16; ++c)
          if (c * 16 < n_o) tmem_ld16(tO + lane_off + oc0 + c * 16, ro + c * 16);     // warp-uniform predicate
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < D0; ++e)
          if (e < n_o) o[e] = fmaf(o[e], corr, __uint_as_float(ro[e]));
      }
    }
    tc_fence_before();
    // row sum = the two halves' shares (same running maximum, so they simply add)
    xsum[hf * QT + r] = l;
    named_bar_sync(1, SOFTMAX_THREADS);
    if (q0 + r < a.T) {
      const float inv = 1.0f / (xsum[r] + xsum[QT + r]);
      float* op = a.out + ((long long)b * a.T + q0 + r) * a.C + h * D + oc0;
#pragma unroll
      for (int c = 0; c < D0; c += 4)
        if (c < n_o)
          *reinterpret_cast<float4*>(op + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
    }
  } else if (warp == 8) {
    // ================= image loader: one thread, cp.async.bulk global -> smem =================
    if (elect_one()) {
      const uint32_t qbytes = (uint32_t)q_image_bytes<D>(), kvbytes = (uint32_t)kv_image_bytes<D>(KT);
      const uint8_t* qimg = a.img + (bh * a.nqt + blockIdx.x) * q_image_bytes<D>();
      const uint8_t* kimg = a.img + a.k_off + bh * a.nkt * kv_image_bytes<D>(KT);
      const uint8_t* vimg = a.img + a.v_off + bh * a.nkt * kv_image_bytes<D>(KT);
      for (int j = 0; j < a.nkt; ++j) {
        if (j > 0) mbar_wait(S_FULL, (j - 1) & 1);          // S_{j-1} done: K buffer free
        // the Q image rides on the first K phase (q | k buffers are adjacent but filled by two copies)
        mbar_arrive_expect_tx(K_FULL, kvbytes + (j == 0 ? qbytes : 0u));
        if (j == 0) bulk_g2s(smem_u32(qh), qimg, qbytes, K_FULL);
        bulk_g2s(smem_u32(kh), kimg + (long long)j * kvbytes, kvbytes, K_FULL);
        if (j > 0) mbar_wait(O_FULL, (j - 1) & 1);          // PV_{j-1} done: V buffer free
        mbar_arrive_expect_tx(V_FULL, kvbytes);
        bulk_g2s(smem_u32(vh), vimg + (long long)j * kvbytes, kvbytes, V_FULL);
      }
    }
  } else if (elect_one()) {
    // ================= MMA issuer (one elected lane of warp 9) =================
    const uint32_t idesc_s = make_idesc_f16(QT, KT), idesc_o = make_idesc_f16(QT, D);
    const uint32_t q_lbo = QT * 16, k_lbo = (uint32_t)KT * 16, v_lbo = D * 16, p_lbo = QT * 16;
    const uint32_t sqh = smem_u32(qh), sql = smem_u32(ql), skh = smem_u32(kh), skl = smem_u32(kl);
    const uint32_t svh = smem_u32(vh), svl = smem_u32(vl), sph = smem_u32(ph), spl = smem_u32(pl);
    for (int j = 0; j < a.nkt; ++j) {
      mbar_wait(K_FULL, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint64_t dqh = make_desc(sqh + 2 * ks * q_lbo, q_lbo, 128), dql = make_desc(sql + 2 * ks * q_lbo, q_lbo, 128);
        const uint64_t dkh = make_desc(skh + 2 * ks * k_lbo, k_lbo, 128), dkl = make_desc(skl + 2 * ks * k_lbo, k_lbo, 128);
        umma_f16(tS, dqh, dkh, idesc_s, ks > 0 ? 1u : 0u);
        umma_f16(tS, dql, dkh, idesc_s, 1u);
        umma_f16(tS, dqh, dkl, idesc_s, 1u);
      }
      umma_commit(S_FULL);
      mbar_wait(V_FULL, j & 1);
      mbar_wait(P_FULL, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ks = 0; ks < KT / 16; ++ks) {
        const uint64_t dph = make_desc(sph + 2 * ks * p_lbo, p_lbo, 128), dpl = make_desc(spl + 2 * ks * p_lbo, p_lbo, 128);
        const uint64_t dvh = make_desc(svh + 2 * ks * v_lbo, v_lbo, 128), dvl = make_desc(svl + 2 * ks * v_lbo, v_lbo, 128);
        umma_f16(tO, dph, dvh, idesc_o, ks > 0 ? 1u : 0u);
        umma_f16(tO, dpl, dvh, idesc_o, 1u);
        umma_f16(tO, dph, dvl, idesc_o, 1u);
      }
      umma_commit(O_FULL);
    }
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

template <int D>
int launch_d(const McvdOp& op, cudaStream_t s) {
  AttnArgs a;
  a.qkv = (const float*)op.src0; a.out = (float*)op.dst;
  a.T = op.H * op.W; a.C = op.C0; a.scale = op.f0;
  // key tile: as large as the operand images allow in 227 KB (Q is resident: 4*D*128 bytes; K, V^T, P scale with KT)
  a.KT = (D <= 96) ? 128 : (D <= 128 ? 64 : 32);
  if (a.T < a.KT) a.KT = a.T;
  // two softmax threads per query row split the key columns of a tile: 16, 32 or 64 columns each
  MCVD_CHECK((a.KT == 32 || a.KT == 64 || a.KT == 128) && a.T % a.KT == 0,
             "ATTENTION_UMMA: %d tokens not tileable by the key tile %d", a.T, a.KT);
  a.nkt = a.T / a.KT;
  int cols = a.KT + D, p2 = 32;
  while (p2 < cols) p2 <<= 1;
  a.tmem_cols = p2;
  a.nqt = cdiv(a.T, QT);
  a.img = (uint8_t*)op.dst2;
  const long long bh = (long long)op.B * op.i0;
  a.k_off = bh * a.nqt * q_image_bytes<D>();
  a.v_off = a.k_off + bh * a.nkt * kv_image_bytes<D>(a.KT);
  MCVD_CHECK(op.dst2, "ATTENTION_UMMA: dst2 (operand-image scratch, mcvd_attention_scratch_bytes) is NULL");
  MCVD_CHECK((reinterpret_cast<uintptr_t>(op.dst2) & 15) == 0, "ATTENTION_UMMA: scratch must be 16-byte aligned");
  const size_t smem = 2 * ((size_t)(D / 8) * QT * 16 + (size_t)(D / 8) * a.KT * 16 + (size_t)(a.KT / 8) * D * 16 +
                           (size_t)(a.KT / 8) * QT * 16) + 64 + 6 * QT * sizeof(float);
  MCVD_CHECK(smem <= 227 * 1024, "ATTENTION_UMMA: %zu B of shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(k_attention_umma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  MCVD_CHECK(e == cudaSuccess, "ATTENTION_UMMA: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  dim3 sgrid(a.nqt + 2 * a.nkt, op.i0, op.B);
  k_attn_presplit<D><<<sgrid, SPLIT_THREADS, 0, s>>>(a);
  MCVD_CUDA_LAUNCH_CHECK("attention presplit");
  dim3 grid(a.nqt, op.i0, op.B);
  k_attention_umma<D><<<grid, ATT_THREADS, smem, s>>>(a);
  MCVD_CUDA_LAUNCH_CHECK("attention_umma");
  return 0;
}

}  // namespace

// bytes of operand-image scratch an ATTENTION_UMMA op needs: Q padded to whole 128-row tiles, K and V
// exactly T rows; fp16 hi + lo = 4 bytes per element
long long attention_umma_scratch_bytes(int B, int T, int C) {
  if (B <= 0 || T <= 0 || C <= 0) return 0;
  return 4LL * B * C * ((long long)cdiv(T, QT) * QT + 2LL * T);
}

int launch_attention_umma(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.dst, "ATTENTION_UMMA: null pointer");
  MCVD_CHECK(op.i0 * op.i1 == op.C0, "ATTENTION_UMMA: heads %d x dim %d != channels %d", op.i0, op.i1, op.C0);
  switch (op.i1) {
    case 32: return launch_d<32>(op, s);
    case 48: return launch_d<48>(op, s);
    case 64: return launch_d<64>(op, s);
    case 96: return launch_d<96>(op, s);
    case 128: return launch_d<128>(op, s);
    case 192: return launch_d<192>(op, s);          // cfg4 (bair_big, n_head_channels = 192): key tile 32
    default: break;
  }
  set_error("ATTENTION_UMMA: head dim %d unsupported (32/48/64/96/128/192)", op.i1);
  return -1;
}

}  // namespace mcvd

extern "C" long long mcvd_attention_scratch_bytes(int B, int T, int C) {
  return mcvd::attention_umma_scratch_bytes(B, T, C);
}
