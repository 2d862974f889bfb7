// 3x3 (pad 1) / 1x1 convolution on the Blackwell 5th-gen tensor cores (tcgen05.mma, accumulators in
// TMEM), with the GroupNorm / FiLM / SiLU input transform fused into the shared-memory staging.
//
// Reference semantics: nn.Conv2d 3x3 / 1x1 and NIN of models/better/layers.py:89-113,541-544 applied
// to get_act_norm's output (layerspp.py:518-549), i.e. Conv_0 / Conv_1 / Conv_2 / NIN_* of
// ResnetBlockBigGANppGN (layerspp.py:595-624) and AttnBlockpp (:230-249).
//
// fp32 parity on fp16 tensor cores: both operands are split  v = hi + lo  (hi = fp16(v),
// lo = fp16(v - hi)) and three MMAs  hi*hi + lo*hi + hi*lo  accumulate in fp32 (TMEM); the dropped
// lo*lo term and the second-level rounding are ~2^-22 relative.  Weights are pre-scaled by a power of
// two (undone in the epilogue) so their lo parts stay in the fp16 normal range.
//
// "Padded-flat" implicit GEMM.  The batch is viewed as one flat array of positions
//     q = b*(H+1)*(W+1) + r*(W+1) + c,   r in [0,H], c in [0,W],   pixel (y,x) = (r-1, c-1)
// where row r = 0 and column c = 0 are zero padding shared between neighbouring rows / images.  In
// this indexing every 3x3 tap is a CONSTANT flat offset  dy*(W+1) + dx,  so a CTA stages ONE halo
// slab [tile + 2*(W+1) + 2 positions] x [32 channels] of the transformed input in shared memory per
// K-block and all nine taps are shifted views of it: the A-operand descriptor of tap (dy,dx) is the
// slab descriptor advanced by (dy*(W+1)+dx) * 16 bytes.  That needs a layout whose M-stride is 16
// bytes, which is exactly the canonical no-swizzle K-major core-matrix layout
//     [k-chunk of 8 halfs][position][8 halfs = 16 B]      LBO = positions*16 B,  SBO = 128 B.
// Outputs at padding positions are computed and discarded (2..20 % of the rows).
//
// Warp roles (320 threads, one CTA per SM):
//   warps 0-7  producers: fp32 NHWC global -> normalise/FiLM/SiLU -> fp16 hi/lo -> smem slab
//              (generic-proxy stores + fence.proxy.async), then the epilogue (TMEM -> regs -> global)
//   warp  8    weight loader: cp.async.bulk (TMA 1-D) of pre-packed fp16 hi/lo smem images
//   warp  9    TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
#include <cuda_fp16.h>

#include "mcvd_common.cuh"

namespace mcvd {

namespace {

constexpr int NPROD = 256;      // producer / epilogue threads (8 warps)
constexpr int NTHREADS = 320;   // + loader warp + MMA warp
constexpr int MT = 128;         // rows per accumulator (UMMA M)

struct UmmaArgs {
  const float* s0;
  const float* s1;
  const __half* wpk;     // packed weights (see k_pack_weights)
  const float* bias;
  const float* res;
  const float4* tab;     // norm table [B][Cin] (mean, rstd, G, S) or null
  float* dst;
  int B, H, W, C0, C1, Cout;
  int ks;                // 1 or 3
  int Wp, Pimg;          // padded row pitch, positions per image
  long long Qtot;        // total flat positions
  int NT, NACC, KB;      // n tile, accumulators per CTA, channels per K-block (16|32)
  int HP;                // halo slab positions (multiple of 8)
  int halo0;             // slab index of the tile's first output position
  int nKB;               // K-blocks
  int NB;                // weight ring stages
  int tmem_cols;
  int act_in, act_out;
  float wscale, oscale;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (K-direction core-matrix stride) | [32,46) SBO>>4 (8-row group
//   stride) | [46,48) version = 1 | [61,64) layout = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// position decode: flat q -> pixel index (b*H + y)*W + x, or -1 for padding / out of range
__device__ __forceinline__ long long decode_pos(const UmmaArgs& a, long long q, int& b_out) {
  if (q < 0 || q >= a.Qtot) return -1;
  int b = (int)(q / a.Pimg);
  int r = (int)(q - (long long)b * a.Pimg);
  int rr = r / a.Wp, cc = r - rr * a.Wp;
  b_out = b;
  if (a.ks == 3) {
    if (rr == 0 || cc == 0) return -1;
    return ((long long)b * a.H + (rr - 1)) * a.W + (cc - 1);
  }
  return ((long long)b * a.H + rr) * a.W + cc;
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// ---- the kernel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) k_conv_umma(const UmmaArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // carve-up: [A stage 0][A stage 1][B ring][pinfo int2[HP]][barriers][tmem slot]
  const int chunks = a.KB / 8;
  const uint32_t a_half_bytes = (uint32_t)chunks * a.HP * 16;       // one of hi / lo
  const uint32_t a_stage_bytes = 2 * a_half_bytes;
  const uint32_t b_step_bytes = 64u * a.NT;                          // one k16 step: hi (2 chunks) + lo
  const uint32_t b_stage_bytes = (uint32_t)(a.KB / 16) * b_step_bytes;
  uint8_t* a_base = smem_raw;
  uint8_t* b_base = a_base + 2 * a_stage_bytes;
  int2* pinfo = reinterpret_cast<int2*>(b_base + (size_t)a.NB * b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(pinfo) + (size_t)a.HP * sizeof(int2));
  // bars: a_full[2], a_empty[2], b_full[NB], b_empty[NB], acc_full
  uint32_t bar0 = smem_u32(bars);
  auto A_FULL = [&](int i) { return bar0 + 8u * i; };
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (2 + i); };
  auto B_FULL = [&](int i) { return bar0 + 8u * (4 + i); };
  auto B_EMPTY = [&](int i) { return bar0 + 8u * (4 + a.NB + i); };
  const uint32_t ACC_FULL = bar0 + 8u * (4 + 2 * a.NB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + (5 + 2 * a.NB));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int taps = a.ks * a.ks;
  const int MTOT = MT * a.NACC;
  const long long p0 = (long long)blockIdx.x * MTOT;   // first output position of this tile
  const int n0 = blockIdx.y * a.NT;

  if (tid == 0) {
    mbar_init(A_FULL(0), NPROD); mbar_init(A_FULL(1), NPROD);
    mbar_init(A_EMPTY(0), 1); mbar_init(A_EMPTY(1), 1);
    for (int i = 0; i < a.NB; ++i) { mbar_init(B_FULL(i), 1); mbar_init(B_EMPTY(i), 1); }
    mbar_init(ACC_FULL, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  // position table for the halo slab
  for (int h = tid; h < a.HP; h += NTHREADS) {
    int b = 0;
    long long pix = decode_pos(a, p0 - a.halo0 + h, b);
    pinfo[h] = make_int2((int)pix, b);   // pixel index < 2^31 is checked on the host
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =========================== producers ===========================
    const int Cin = a.C0 + a.C1;
    const int units = chunks * a.HP;
    for (int kb = 0; kb < a.nKB; ++kb) {
      const int st = kb & 1;
      mbar_wait(A_EMPTY(st), ((kb >> 1) & 1) ^ 1);
      const int c0 = kb * a.KB;
      const float* src;
      int cs, cc0;
      if (c0 < a.C0) { src = a.s0; cs = a.C0; cc0 = c0; } else { src = a.s1; cs = a.C1; cc0 = c0 - a.C0; }
      uint8_t* hi_base = a_base + (size_t)st * a_stage_bytes;
      uint8_t* lo_base = hi_base + a_half_bytes;
      for (int u = tid; u < units; u += NPROD) {
        const int ch = u / a.HP, h = u - ch * a.HP;
        const int2 pi = pinfo[h];
        uint4 hv = make_uint4(0u, 0u, 0u, 0u), lv = hv;
        if (pi.x >= 0) {
          const float* sp = src + (long long)pi.x * cs + cc0 + ch * 8;
          float4 v0 = *reinterpret_cast<const float4*>(sp);
          float4 v1 = *reinterpret_cast<const float4*>(sp + 4);
          float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          if (a.tab) {
            const float4* tb = a.tab + (long long)pi.y * Cin + c0 + ch * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float4 t = __ldg(tb + e);
              float n = ((v[e] - t.x) * t.y) * t.z + t.w;
              if (a.act_in) n = silu_f(n);
              v[e] = n;
            }
          }
          __half hh[8], ll[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            hh[e] = __float2half_rn(v[e]);
            ll[e] = __float2half_rn(v[e] - __half2float(hh[e]));
          }
          hv = make_uint4(pack_half2(hh[0], hh[1]), pack_half2(hh[2], hh[3]), pack_half2(hh[4], hh[5]),
                          pack_half2(hh[6], hh[7]));
          lv = make_uint4(pack_half2(ll[0], ll[1]), pack_half2(ll[2], ll[3]), pack_half2(ll[4], ll[5]),
                          pack_half2(ll[6], ll[7]));
        }
        const size_t off = ((size_t)ch * a.HP + h) * 16;
        *reinterpret_cast<uint4*>(hi_base + off) = hv;
        *reinterpret_cast<uint4*>(lo_base + off) = lv;
      }
      fence_proxy_async();          // make the generic-proxy stores visible to the tensor-core proxy
      mbar_arrive(A_FULL(st));
    }

    // =========================== epilogue ===========================
    mbar_wait(ACC_FULL, 0);
    tc_fence_after();
    const int lq = warp & 3, grp = warp >> 2;
    const int nchunk = a.NT / 16;
    for (int acc = 0; acc < a.NACC; ++acc) {
      const long long q = p0 + (long long)acc * MT + lq * 32 + lane;
      int b = 0;
      const long long pix = decode_pos(a, q, b);
      const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * a.NT);
      for (int j = grp; j < nchunk; j += 2) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)(j * 16), r);
        tmem_ld_wait();
        if (pix >= 0) {
          const int n = n0 + j * 16;
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(r[e]) * a.wscale;
          if (a.bias) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] += __ldg(a.bias + n + e);
          }
          float* dp = a.dst + pix * a.Cout + n;
          if (a.res) {
            const float* rp = a.res + pix * a.Cout + n;
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float4 rv = *reinterpret_cast<const float4*>(rp + e);
              o[e] += rv.x; o[e + 1] += rv.y; o[e + 2] += rv.z; o[e + 3] += rv.w;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            o[e] *= a.oscale;
            if (a.act_out) o[e] = silu_f(o[e]);
          }
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(dp + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // =========================== weight loader ===========================
    if (lane == 0) {
      const int total = a.nKB * taps;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpk) +
                            (size_t)blockIdx.y * (size_t)a.nKB * taps * b_stage_bytes;
      for (int i = 0; i < total; ++i) {
        const int st = i % a.NB;
        mbar_wait(B_EMPTY(st), ((i / a.NB) & 1) ^ 1);
        mbar_arrive_expect_tx(B_FULL(st), b_stage_bytes);
        bulk_g2s(smem_u32(b_base + (size_t)st * b_stage_bytes), wsrc + (size_t)i * b_stage_bytes, b_stage_bytes,
                 B_FULL(st));
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format F16 = 0,
      // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)a.HP * 16, b_lbo = (uint32_t)a.NT * 16;
      const int ksteps = a.KB / 16;
      int bi = 0;
      for (int kb = 0; kb < a.nKB; ++kb) {
        const int st = kb & 1;
        mbar_wait(A_FULL(st), (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(a_base + (size_t)st * a_stage_bytes);
        const uint32_t a_lo = a_hi + a_half_bytes;
        for (int tap = 0; tap < taps; ++tap, ++bi) {
          const int bst = bi % a.NB;
          mbar_wait(B_FULL(bst), (bi / a.NB) & 1);
          tc_fence_after();
          const int shift = (a.ks == 3) ? ((tap / 3 - 1) * a.Wp + (tap % 3 - 1)) : 0;
          const uint32_t b_stage = smem_u32(b_base + (size_t)bst * b_stage_bytes);
          for (int s = 0; s < ksteps; ++s) {
            const uint32_t b_hi = b_stage + (uint32_t)s * b_step_bytes;
            const uint32_t b_lo = b_hi + 32u * a.NT;
            const uint64_t dbh = make_desc(b_hi, b_lbo, 128);
            const uint64_t dbl = make_desc(b_lo, b_lbo, 128);
            for (int acc = 0; acc < a.NACC; ++acc) {
              const uint32_t row_off = (uint32_t)((a.halo0 + shift + acc * MT) * 16) + (uint32_t)(s * 2) * a_lbo;
              const uint64_t dah = make_desc(a_hi + row_off, a_lbo, 128);
              const uint64_t dal = make_desc(a_lo + row_off, a_lbo, 128);
              const uint32_t d = tmem_base + (uint32_t)(acc * a.NT);
              const uint32_t first = (kb == 0 && tap == 0 && s == 0) ? 0u : 1u;
              umma_f16(d, dah, dbh, idesc, first);
              umma_f16(d, dal, dbh, idesc, 1u);
              umma_f16(d, dah, dbl, idesc, 1u);
            }
          }
          umma_commit(B_EMPTY(bst));        // weights of this stage consumed
        }
        umma_commit(A_EMPTY(st));           // slab of this K-block consumed
      }
      umma_commit(ACC_FULL);
    }
  }

  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ---- weight packing ---------------------------------------------------------------------------------
// in : w_taps fp32 [taps][Cin][Cout]
// out: fp16, for nt, kb, tap, s(k16):  hi[2 chunks][NT][8]  then  lo[2 chunks][NT][8]
__global__ void k_pack_weights(const float* __restrict__ w, __half* __restrict__ out, int taps, int Cin, int Cout,
                               int NT, int KB, float scale) {
  const int ksteps = KB / 16, nKB = Cin / KB, nNT = Cout / NT;
  const long long total = (long long)nNT * nKB * taps * ksteps * 2 * NT * 8;  // (chunk j, n, e) per step
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i % 8);
    long long t = i / 8;
    int n = (int)(t % NT); t /= NT;
    int j = (int)(t % 2); t /= 2;
    int s = (int)(t % ksteps); t /= ksteps;
    int tap = (int)(t % taps); t /= taps;
    int kb = (int)(t % nKB); t /= nKB;
    int nt = (int)t;
    int c = kb * KB + s * 16 + j * 8 + e;
    float v = w[((long long)tap * Cin + c) * Cout + nt * NT + n] * scale;
    __half h = __float2half_rn(v);
    __half l = __float2half_rn(v - __half2float(h));
    long long step = (((long long)nt * nKB + kb) * taps + tap) * ksteps + s;
    long long base = step * (4LL * NT * 8);      // halfs per step: hi 2*NT*8 + lo 2*NT*8
    long long o = (long long)(j * NT + n) * 8 + e;
    out[base + o] = h;
    out[base + 2LL * NT * 8 + o] = l;
  }
}

int pick_kb(int C0, int C1) {
  if (C0 % 32 == 0 && C1 % 32 == 0) return 32;
  if (C0 % 16 == 0 && C1 % 16 == 0) return 16;
  return 0;
}

}  // namespace

int launch_conv_umma(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.w && op.dst && (op.C1 == 0 || op.src1), "CONV_UMMA: null pointer");
  MCVD_CHECK(op.i0 == 1 || op.i0 == 3, "CONV_UMMA: kernel size %d unsupported", op.i0);
  UmmaArgs a;
  a.s0 = (const float*)op.src0; a.s1 = (const float*)op.src1; a.wpk = (const __half*)op.w;
  a.bias = (const float*)op.bias; a.res = (const float*)op.aux0; a.tab = (const float4*)op.aux1;
  a.dst = (float*)op.dst;
  a.B = op.B; a.H = op.H; a.W = op.W; a.C0 = op.C0; a.C1 = op.C1; a.Cout = op.Cout; a.ks = op.i0;
  a.NT = op.i1; a.NACC = op.i2;
  a.KB = pick_kb(op.C0, op.C1);
  MCVD_CHECK(a.KB != 0, "CONV_UMMA: input channels (%d,%d) must be multiples of 16", op.C0, op.C1);
  MCVD_CHECK(a.NT >= 16 && a.NT <= 256 && a.NT % 16 == 0 && op.Cout % a.NT == 0,
             "CONV_UMMA: n tile %d invalid for Cout %d", a.NT, op.Cout);
  MCVD_CHECK(a.NACC == 1 || a.NACC == 2, "CONV_UMMA: accumulators %d", a.NACC);
  if (a.ks == 3) { a.Wp = op.W + 1; a.Pimg = (op.H + 1) * (op.W + 1); }
  else { a.Wp = op.W; a.Pimg = op.H * op.W; }
  a.Qtot = (long long)op.B * a.Pimg;
  MCVD_CHECK((long long)op.B * op.H * op.W < (1LL << 31), "CONV_UMMA: too many pixels");
  const int MTOT = MT * a.NACC;
  a.halo0 = (a.ks == 3) ? a.Wp + 1 : 0;
  a.HP = (MTOT + 2 * a.halo0 + 7) & ~7;
  a.nKB = (op.C0 + op.C1) / a.KB;
  a.act_in = (op.flags & MCVD_F_ACT_IN) ? 1 : 0;
  a.act_out = (op.flags & MCVD_F_ACT_OUT) ? 1 : 0;
  a.wscale = op.f1; a.oscale = op.f0;
  int cols = a.NACC * a.NT, p2 = 32;
  while (p2 < cols) p2 <<= 1;
  MCVD_CHECK(p2 <= 512, "CONV_UMMA: %d TMEM columns", cols);
  a.tmem_cols = p2;
  const size_t a_stage = (size_t)2 * (a.KB / 8) * a.HP * 16;
  const size_t b_stage = (size_t)(a.KB / 16) * 64 * a.NT;
  const size_t fixed = 2 * a_stage + (size_t)a.HP * 8 + 8 * 64 + 16;
  const size_t limit = 227 * 1024;
  MCVD_CHECK(fixed + 2 * b_stage <= limit, "CONV_UMMA: tile does not fit shared memory (W=%d)", op.W);
  int NB = (int)((limit - fixed) / b_stage);
  if (NB > 8) NB = 8;
  a.NB = NB;
  const size_t smem = 2 * a_stage + (size_t)NB * b_stage + (size_t)a.HP * 8 + (size_t)(5 + 2 * NB) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(k_conv_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
  MCVD_CHECK(e == cudaSuccess, "CONV_UMMA: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  long long tiles = (a.Qtot + MTOT - 1) / MTOT;
  dim3 grid((unsigned)tiles, (unsigned)(op.Cout / a.NT));
  k_conv_umma<<<grid, NTHREADS, smem, s>>>(a);
  MCVD_CUDA_LAUNCH_CHECK("conv_umma");
  return 0;
}

}  // namespace mcvd

extern "C" int mcvd_umma_kblock(int C0, int C1) { return mcvd::pick_kb(C0, C1); }

extern "C" long long mcvd_umma_pack_weights(const float* w_taps, int taps, int Cin, int Cout, int n_tile, int KB,
                                            void* out, int scale_log2, void* stream) {
  if ((KB != 16 && KB != 32) || Cin % KB || n_tile < 16 || n_tile % 16 || Cout % n_tile) {
    mcvd::set_error("umma_pack_weights: Cin %d / Cout %d / n_tile %d unsupported", Cin, Cout, n_tile);
    return -1;
  }
  long long halfs = (long long)taps * Cin * Cout * 2;  // hi + lo
  long long bytes = halfs * 2;
  if (!out) return bytes;
  if (!w_taps) {
    mcvd::set_error("umma_pack_weights: null input");
    return -1;
  }
  float scale = ldexpf(1.0f, scale_log2);
  long long total = (long long)taps * Cin * Cout;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  mcvd::k_pack_weights<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_taps, (__half*)out, taps, Cin, Cout,
                                                                           n_tile, KB, scale);
  if (cudaGetLastError() != cudaSuccess) {
    mcvd::set_error("umma_pack_weights: launch failed");
    return -2;
  }
  return bytes;
}
