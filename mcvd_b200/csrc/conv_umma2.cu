// 3x3 (pad 1) / 1x1 convolution on the Blackwell tensor cores, CTA-pair edition: tcgen05.mma.cta_group::2
// (M = 256 across the two SMs of a TPC), fp32 accumulators in TMEM, GroupNorm / FiLM / SiLU input transform
// fused into the shared-memory staging, GroupNorm partial sums of the OUTPUT fused into the epilogue.
//
// Reference semantics: nn.Conv2d 3x3 / 1x1 and NIN of models/better/layers.py:89-113,541-544 applied to
// get_act_norm's output (layerspp.py:518-549), i.e. Conv_0 / Conv_1 / Conv_2 / NIN_* of
// ResnetBlockBigGANppGN (layerspp.py:595-624) and AttnBlockpp (:230-249); the epilogue statistics are the
// sums nn.GroupNorm (layerspp.py:474-477) of the NEXT act-norm needs.
//
// fp32 parity on fp16 tensor cores: v = hi + lo (hi = fp16(v), lo = fp16(v - hi)) for both operands, three
// MMAs hi*hi + lo*hi + hi*lo, fp32 accumulation.
//
// "Padded-flat" implicit GEMM (as in round 1): positions q = b*Pimg + r*(W+1) + c with row 0 / column 0 of
// every image being shared zero padding, so a 3x3 tap is a constant flat offset and all nine taps are shifted
// views (descriptor start + offset*16 B) of ONE halo slab  [k-chunk of 8 halfs][position][16 B]  per K-block.
//
// What changed against conv_umma.cu (round 1), and why (profiles/r2_timing_v1_round1_kernel.txt):
//   * CTA pairs.  A unit of work is 256 positions x NT output channels.  CTA r of the pair stages the slab of
//     ITS 128 positions and loads ITS half (NT/2 columns) of the weight tile; one tcgen05.mma.cta_group::2
//     issued by the leader covers both.  Per instruction each SM now reads 4 KB of A + NT*16 B of B from its
//     shared memory instead of 4 KB + NT*32 B for half the work: the A-operand read floor of round 1
//     (64 + N/4 cycles per M128 instruction) becomes max(N/2, 64 + N/8) per M256, and the weight traffic
//     from L2 per SM halves.
//   * Producers: 9 warps, unit = (slab row, 16-byte piece) so global reads are whole 128-byte lines and
//     shared-memory accesses are conflict-free; the fp32 rows arrive through a cp.async ring (R stages,
//     prefetch distance R-1 K-blocks, zero-fill for padding rows) instead of registers, every thread reads
//     back only what it copied itself (no producer-side barriers), and the norm table comes from L1.
//     Round 1 ran the 1x1 convs with ONE producer warp per SM sub-partition (2.5k cycles per K-block).
//   * The fused 1x1 shortcut segment stages only the 128 centre rows (no halo).
//   * Epilogue: optional per-(image, channel) sum / sum of squares of the stored output, accumulated
//     EXACTLY in 64-bit fixed point (x * 2^16 rounded to an integer), so the result does not depend on how
//     the rows of an image fall into tiles, warps or GPUs (bit-exact clip sharding is preserved) -- this
//     replaces the k_gn_partial pass (one extra read of every activation).
//
// Warp roles (480 threads, one CTA per SM, clusters of 2):
//   warps 0-8   producers   warp 9 weight loader (cp.async.bulk)   warp 10 MMA issuer (leader) / barrier
//   forwarder (peer)        warps 11-14 epilogue (TMEM lane quadrants 3,0,1,2)
#include <cuda_fp16.h>

#include "mcvd_common.cuh"
#include "umma_ptx.cuh"

namespace mcvd {

namespace {

using namespace ptx;

constexpr int NPROD = 288;
constexpr int W_LOAD = 9;
constexpr int W_MMA = 10;
constexpr int W_EPI = 11;
constexpr int NTHREADS = 480;
constexpr int MT = 128;                 // positions per CTA and unit
constexpr float STAT_SCALE = 65536.0f;  // fixed-point scale of the epilogue statistics
constexpr int STAT_CLAMP = 1 << 28;

struct C2Args {
  const float* s0;
  const float* s1;
  const float* s2;
  const float* s3;
  int C0, C1, C2, C3;
  int nKB0, nKB;
  const uint8_t* wpk;
  const float* bias;
  const float* res;
  const float* tab3;            // [B][3][C0+C1]: mean | rstd*G | S   (null: raw input)
  float* dst;
  unsigned long long* stats;    // [tiles128][NJ][2][Cout] fixed-point partial sums, or null
  long long* dbg;
  int B, H, W, Cout, ks, Wp, Pimg, HW;
  long long Qtot;
  int NT, KB, HP, halo0, NB, SA, R, tiles_n, nunits, tmem_cols, NJ;
  int act_in, act_out, split;
  int dbgf;                     // bring-up switches (tools/conv2_check.py): 1 producers skip copies + transform, 2 epilogue skips
                                // its loads / stores, 4 weight stages are not reloaded, 8 leader ignores the peer's stage barriers, 16 no slab
                                // pipeline at all (producers idle, MMA does not wait for slabs), 32 no weight pipeline -- garbage results, timing only
  float wscale, oscale;
  uint32_t off_img, off_raw, off_b, off_pad, off_row, off_stat, off_bias, off_bar;
  uint32_t a_plane, raw_stage, b_stage;
};

// flat position -> pixel index (b*H + y)*W + x or -1 (padding / out of range); b_out = image
__device__ __forceinline__ int decode_pos(const C2Args& a, long long q, int& b_out) {
  b_out = 0;
  if (q < 0 || q >= a.Qtot) return -1;
  const int b = (int)(q / a.Pimg);
  const int r = (int)(q - (long long)b * a.Pimg);
  const int rr = r / a.Wp, cc = r - rr * a.Wp;
  b_out = b;
  if (a.ks == 3) {
    if (rr == 0 || cc == 0 || rr > a.H) return -1;
    return (b * a.H + (rr - 1)) * a.W + (cc - 1);
  }
  if (rr >= a.H) return -1;
  return (b * a.H + rr) * a.W + cc;
}

// fp32 pair -> fp16 hi pair + fp16 lo pair (x in the low half).  Both conversions saturate to +-65504
// (F2FP.SATFINITE), so an activation beyond the fp16 range degrades to a coarser finite value instead of
// inf - inf = NaN (round 1's split2)
__device__ __forceinline__ void split2_sat(float x, float y, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(y - hf.y), "f"(x - hf.x));
}

__device__ __forceinline__ float4 ld_nc_na(const float* p) {      // read-only, do not allocate in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

template <int PEND>
__device__ __forceinline__ void wait_copies() { cp_async_wait<PEND>(); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) k_conv_umma2(const __grid_constant__ C2Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + a.off_bar;
  const int SA = a.SA, NB = a.NB;
  auto A_FULL = [&](int i) { return bar0 + 8u * i; };
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (SA + i); };
  auto PA_FULL = [&](int i) { return bar0 + 8u * (2 * SA + i); };
  auto B_FULL = [&](int i) { return bar0 + 8u * (3 * SA + i); };
  auto B_EMPTY = [&](int i) { return bar0 + 8u * (3 * SA + NB + i); };
  auto PB_FULL = [&](int i) { return bar0 + 8u * (3 * SA + 2 * NB + i); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (3 * SA + 3 * NB + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (3 * SA + 3 * NB + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.off_bar + 8 * (3 * SA + 3 * NB + 4));
  const int taps = a.ks * a.ks;
  const int per_unit = a.nKB0 * taps + (a.nKB - a.nKB0);       // weight stages per unit

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(A_FULL(i), NPROD / 32); mbar_init(A_EMPTY(i), 1); mbar_init(PA_FULL(i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(B_FULL(i), 1); mbar_init(B_EMPTY(i), 1); mbar_init(PB_FULL(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 256); }
    fence_barrier_init();
  }
  if (a.stats) {
    unsigned long long* st = reinterpret_cast<unsigned long long*>(smem + a.off_stat);
    for (int i = tid; i < a.NJ * 2 * a.NT; i += NTHREADS) st[i] = 0ull;
  }
  if (warp == W_MMA) tmem_alloc2(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;
  const long long t_begin = dbg ? clock64() : 0;
#define DBG_T(var) long long var = dbg ? clock64() : 0
#define DBG_ADD(slot, since, cond) do { if (dbg && (cond)) { long long t__ = clock64(); dbg[slot] += t__ - since; since = t__; } } while (0)

  const int my_units = (a.nunits - cid + ncl - 1) / ncl;        // units cid, cid + ncl, ...

  if (warp < W_LOAD) {
    // =============================== producers ===============================
    const int pj_shift = (a.KB == 32) ? 3 : 2;                 // 16-byte pieces per slab row: 8 or 4
    const int PJ = 1 << pj_shift;
    const int j = tid & (PJ - 1);
    const int hrow = tid >> pj_shift;
    const int RPP = NPROD >> pj_shift;                         // rows per pass: 36 or 72
    const int Cin = a.C0 + a.C1;
    const uint32_t raw0 = sbase + a.off_raw, img0 = sbase + a.off_img;
    int2* rowinfo = reinterpret_cast<int2*>(smem + a.off_row);   // [2][HP] (pix, image)
    const int total = my_units * a.nKB;
    const int Rm1 = a.R - 1;

    auto src_of = [&](int kb, const float*& src, int& cs, int& cc0) {
      if (kb < a.nKB0) {
        const int c0 = kb * a.KB;
        if (c0 < a.C0) { src = a.s0; cs = a.C0; cc0 = c0; } else { src = a.s1; cs = a.C1; cc0 = c0 - a.C0; }
      } else {
        const int c0 = (kb - a.nKB0) * a.KB;
        if (c0 < a.C2) { src = a.s2; cs = a.C2; cc0 = c0; } else { src = a.s3; cs = a.C3; cc0 = c0 - a.C2; }
      }
    };

    int i_unit = 0, i_kb = 0;        // issue stream position
    int t_unit = 0, t_kb = 0;        // transform stream position
    for (int it = 0; it < ((a.dbgf & 16) ? 0 : total + Rm1); ++it) {
      // ---- issue the copies of job `it` (K-block i_kb of unit i_unit) into raw stage it % R ----
      if (it < total) {
        const int par = i_unit & 1;
        if (i_kb == 0) {
          // slab row -> pixel table of this unit (the transform stream may still read the other buffer)
          named_bar_sync(1, NPROD);
          const int u = cid + i_unit * ncl;
          const long long p0 = (long long)(u / a.tiles_n) * (2 * MT) + (long long)rank * MT - a.halo0;
          for (int h = tid; h < a.HP; h += NPROD) {
            int b;
            const int pix = decode_pos(a, p0 + h, b);
            rowinfo[par * a.HP + h] = make_int2(pix, b);
          }
          named_bar_sync(1, NPROD);
        }
        const float* src; int cs, cc0;
        src_of(i_kb, src, cs, cc0);
        // slab row h always belongs to thread group h % RPP (whatever the segment), so a raw-ring slot is
        // only ever touched by one thread and needs no barrier; the 1x1 shortcut segment stages the centre rows only
        const bool seg1 = i_kb >= a.nKB0;
        const int hlo = seg1 ? a.halo0 : 0, hhi = seg1 ? a.halo0 + MT : a.HP;
        const uint32_t rst = raw0 + (uint32_t)(it % a.R) * a.raw_stage;
        const float* sj = src + cc0 + j * 4;
        for (int h = hrow; h < hhi && !(a.dbgf & 1); h += RPP) {
          if (h < hlo) continue;
          const int pix = rowinfo[par * a.HP + h].x;
          const float* p = pix >= 0 ? sj + (long long)pix * cs : src;
          cp_async16(rst + (uint32_t)((h << pj_shift) + j) * 16u, p, pix >= 0 ? 16u : 0u);
        }
        if (++i_kb == a.nKB) { i_kb = 0; ++i_unit; }
      }
      cp_async_commit();
      // ---- transform job it - (R-1) ----
      const int gt = it - Rm1;
      if (gt < 0) continue;
      if (Rm1 == 1) wait_copies<1>(); else if (Rm1 == 2) wait_copies<2>(); else wait_copies<3>();
      const int st = gt % SA;
      DBG_T(tp);
      // one lane per warp polls (32 lanes spinning on the same mbarrier steal shared-memory cycles from the
      // tensor core's operand reads); __syncwarp orders the other lanes behind lane 0's acquire
      if (lane == 0) mbar_wait(A_EMPTY(st), ((gt / SA) & 1) ^ 1);
      __syncwarp();
      DBG_ADD(1, tp, tid == 0);
      const int par = t_unit & 1;
      const bool seg1 = t_kb >= a.nKB0;
      const bool norm = a.tab3 != nullptr && !seg1;
      const int hlo = seg1 ? a.halo0 : 0, hhi = seg1 ? a.halo0 + MT : a.HP;
      const uint32_t rst = raw0 + (uint32_t)(gt % a.R) * a.raw_stage;
      const uint32_t hi_base = img0 + (uint32_t)st * 2u * a.a_plane, lo_base = hi_base + a.a_plane;
      const uint32_t img_off = (uint32_t)(j >> 1) * (uint32_t)a.HP * 16u + (uint32_t)(j & 1) * 8u;
      const float* tabc = a.tab3 + t_kb * a.KB + j * 4;
      // (mean, rstd*G, S) of this thread's 4 channels for the image of the current row: a slab touches one or two
      // images on the large maps, so the three L1 loads are paid once per K-block, not once per row
      float4 tm = make_float4(0.f, 0.f, 0.f, 0.f), tg = tm, ts = tm;
      int tb_img = -1;
#pragma unroll 4
      for (int h = hrow; h < hhi && !(a.dbgf & 1); h += RPP) {
        if (h < hlo) continue;
        const int2 info = rowinfo[par * a.HP + h];
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                     : "r"(rst + (uint32_t)((h << pj_shift) + j) * 16u));
        uint32_t h0 = 0u, h1 = 0u, l0 = 0u, l1 = 0u;
        if (info.x >= 0) {
          float v[4] = {x.x, x.y, x.z, x.w};
          if (norm) {
            if (info.y != tb_img) {
              const float* tb = tabc + (long long)info.y * 3 * Cin;
              tm = __ldg(reinterpret_cast<const float4*>(tb));
              tg = __ldg(reinterpret_cast<const float4*>(tb + Cin));
              ts = __ldg(reinterpret_cast<const float4*>(tb + 2 * Cin));
              tb_img = info.y;
            }
            v[0] = fmaf(v[0] - tm.x, tg.x, ts.x); v[1] = fmaf(v[1] - tm.y, tg.y, ts.y);
            v[2] = fmaf(v[2] - tm.z, tg.z, ts.z); v[3] = fmaf(v[3] - tm.w, tg.w, ts.w);
            if (a.act_in) silu_fast4(v);
          }
          split2_sat(v[0], v[1], h0, l0);
          split2_sat(v[2], v[3], h1, l1);
        }
        const uint32_t off = img_off + (uint32_t)h * 16u;
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(hi_base + off), "r"(h0), "r"(h1) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(lo_base + off), "r"(l0), "r"(l1) : "memory");
      }
      DBG_ADD(3, tp, tid == 0);
      fence_proxy_async();                // generic-proxy smem stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(A_FULL(st));          // one arrival per producer warp
      DBG_ADD(4, tp, tid == 0);
      if (++t_kb == a.nKB) { t_kb = 0; ++t_unit; }
    }
  } else if (warp == W_LOAD) {
    // =============================== weight loader ===============================
    if (elect_one()) {
      const uint32_t b0 = sbase + a.off_b;
      int st = 0, ph = 1;
      for (int iu = 0; iu < ((a.dbgf & 32) ? 0 : my_units); ++iu) {
        const int u = cid + iu * ncl;
        const uint8_t* wsrc = a.wpk + ((size_t)(u % a.tiles_n) * per_unit * 2 + rank) * a.b_stage;
        for (int i = 0; i < per_unit; ++i) {
          mbar_wait(B_EMPTY(st), ph);
          if (a.dbgf & 4) { mbar_arrive(B_FULL(st)); if (++st == NB) { st = 0; ph ^= 1; } continue; }
          mbar_arrive_expect_tx(B_FULL(st), a.b_stage);
          bulk_g2s(b0 + (uint32_t)st * a.b_stage, wsrc + (size_t)i * 2 * a.b_stage, a.b_stage, B_FULL(st));
          if (++st == NB) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    if (rank == 0) {
      // =============================== MMA issuer (leader CTA) ===============================
      if (elect_one()) {
        const uint32_t idesc = make_idesc_f16(2 * MT, a.NT);
        const uint32_t a_lbo16 = (uint32_t)a.HP, b_lbo16 = (uint32_t)(a.NT >> 1);
        const uint64_t a_proto = make_desc(0, a_lbo16 * 16, 128), b_proto = make_desc(0, b_lbo16 * 16, 128);
        const int ksteps = a.KB / 16;
        const uint32_t a_plane16 = a.a_plane >> 4, b_step16 = (32u * a.NT) >> 4, b_lo16 = (16u * a.NT) >> 4;
        const uint32_t a0_16 = (sbase + a.off_img) >> 4, b0_16 = (sbase + a.off_b) >> 4;
        const uint32_t a_stage16 = 2 * a_plane16, b_stage16 = a.b_stage >> 4;
        const bool w_lo = (a.split & 2) != 0, a_lo = (a.split & 1) != 0;
        int bst = 0, bph = 0, g = 0;
        for (int iu = 0; iu < my_units; ++iu) {
          const int set = iu & 1;
          DBG_T(tm);
          mbar_wait_cluster(ACC_EMPTY(set), ((iu >> 1) & 1) ^ 1);    // both epilogues drained this set
          DBG_ADD(5, tm, true);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(set * a.NT);
          uint32_t accum = 0;
          for (int kb = 0; kb < a.nKB; ++kb, ++g) {
            const int st = g % SA;
            const uint32_t aph = (uint32_t)((g / SA) & 1);
            DBG_ADD(8, tm, true);
            if (!(a.dbgf & 16)) {
              mbar_wait(A_FULL(st), aph);
              if (!(a.dbgf & 8)) mbar_wait_cluster(PA_FULL(st), aph);
            }
            DBG_ADD(6, tm, true);
            // no tcgen05.fence here: the slab was written through the generic proxy and published with
            // fence.proxy.async + mbarrier release/acquire; a tcgen05.fence::after_thread_sync per stage drained
            // the MMA pipeline (~500 cycles per weight stage in round 1 and in the first version of this kernel)
            const uint32_t a_hi16 = a0_16 + (uint32_t)st * a_stage16 + (uint32_t)a.halo0;
            const bool main = kb < a.nKB0;
            const int ntap = main ? taps : 1;
            for (int tap = 0; tap < ntap; ++tap) {
              DBG_ADD(8, tm, true);
              if (!(a.dbgf & 32)) {
                mbar_wait(B_FULL(bst), bph);
                if (!(a.dbgf & 8)) mbar_wait_cluster(PB_FULL(bst), bph);
              }
              DBG_ADD(7, tm, true);
              const int shift = (a.ks == 3 && main) ? ((tap / 3 - 1) * a.Wp + (tap % 3 - 1)) : 0;
              const uint32_t a_tap16 = a_hi16 + (uint32_t)shift;
              const uint32_t b_tap16 = b0_16 + (uint32_t)bst * b_stage16;
              for (int s = 0; s < ksteps; ++s) {
                const uint64_t dbh = desc_add(b_proto, b_tap16 + (uint32_t)s * b_step16);
                const uint64_t dbl = desc_add(dbh, b_lo16);
                const uint64_t dah = desc_add(a_proto, a_tap16 + (uint32_t)(2 * s) * a_lbo16);
                const uint64_t dal = desc_add(dah, a_plane16);
                umma2_f16(d, dah, dbh, idesc, accum);
                if (a_lo) umma2_f16(d, dal, dbh, idesc, 1u);
                if (w_lo) umma2_f16(d, dah, dbl, idesc, 1u);
                accum = 1u;
              }
              umma2_commit_mc(B_EMPTY(bst));       // weight stage consumed (both CTAs)
              if (++bst == NB) { bst = 0; bph ^= 1; }
            }
            umma2_commit_mc(A_EMPTY(st));          // slab stage consumed (both CTAs)
          }
          umma2_commit_mc(ACC_FULL(set));
          DBG_ADD(8, tm, true);
        }
      }
    } else {
      // ===================== peer CTA: forward "stage filled" to the leader's barriers =====================
      if (elect_one()) {
        int bst = 0, bph = 0, g = 0;
        for (int iu = 0; iu < my_units; ++iu) {
          for (int kb = 0; kb < a.nKB; ++kb, ++g) {
            const int st = g % SA;
            if (!(a.dbgf & 16)) {
              mbar_wait(A_FULL(st), (uint32_t)((g / SA) & 1));
              mbar_arrive_remote(mapa_u32(PA_FULL(st), 0));
            }
            const int ntap = (kb < a.nKB0) ? taps : 1;
            for (int tap = 0; tap < ntap; ++tap) {
              if (!(a.dbgf & 32)) {
                mbar_wait(B_FULL(bst), bph);
                mbar_arrive_remote(mapa_u32(PB_FULL(bst), 0));
              }
              if (++bst == NB) { bst = 0; bph ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    // Each of the 4 warps drains its 32 TMEM lanes (rows) in column blocks of 32 (or a 16-wide tail) through an
    // XOR-swizzled 4 KB transpose pad, so global loads / stores cover whole 128-byte lines of dst / res.
    const int lq = warp & 3;
    const int et = tid - W_EPI * 32;
    float4* pad = reinterpret_cast<float4*>(smem + a.off_pad) + (size_t)(warp - W_EPI) * 256;
    float* bias_s = reinterpret_cast<float*>(smem + a.off_bias);
    unsigned long long* stat_s = reinterpret_cast<unsigned long long*>(smem + a.off_stat);
    const int nblk = (a.NT + 31) / 32;
    const float* __restrict__ resp = a.res;
    float* __restrict__ dstp = a.dst;
    const uint32_t acc_empty_leader = mapa_u32(ACC_EMPTY(0), 0);
    for (int iu = 0; iu < my_units; ++iu) {
      const int u = cid + iu * ncl;
      const int set = iu & 1;
      const int mtile = (u / a.tiles_n) * 2 + (int)rank;               // 128-row tile index
      const long long p0 = (long long)mtile * MT;
      const int n0 = (u % a.tiles_n) * a.NT;
      named_bar_sync(2, 128);                                         // previous unit's readers are done
      for (int i = et; i < a.NT; i += 128) bias_s[i] = a.bias ? __ldg(a.bias + n0 + i) : 0.f;
      named_bar_sync(2, 128);
      int myb;
      const int mypix = decode_pos(a, p0 + lq * 32 + lane, myb);
      const int tile_b0 = (int)min((long long)(a.B - 1), p0 / a.Pimg);
      const int myj = mypix >= 0 ? myb - tile_b0 : -1;                 // image slot of this row (stats)
      int px8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) px8[k] = __shfl_sync(0xffffffffu, mypix, k * 4 + (lane >> 3));
      const int q8 = lane & 7;
      unsigned jmask = 0;                                             // image slots present in this warp's rows
      if (a.stats) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (__ballot_sync(0xffffffffu, myj == jj)) jmask |= 1u << jj;
      }
      float4 rnext[8];
      auto res_fetch = [&](int blk) {
        if (!resp) return;
        const int cb = blk * 32;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (px8[k] >= 0) rnext[k] = ld_nc_na(resp + (long long)px8[k] * a.Cout + n0 + cb + q8 * 4);
      };
      if (a.NT >= 32) res_fetch(0);
      DBG_T(te);
      if (lane == 0) mbar_wait(ACC_FULL(set), (uint32_t)((iu >> 1) & 1));
      __syncwarp();
      DBG_ADD(9, te, tid == W_EPI * 32);
      tc_fence_after();
      const uint32_t trow0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(set * a.NT);
      if (a.dbgf & 2) {
        tc_fence_before();
        if (rank == 0) mbar_arrive(ACC_EMPTY(set)); else mbar_arrive_remote(acc_empty_leader + 8u * set);
        continue;
      }
      for (int blk = 0; blk < nblk; ++blk) {
        const int cb = blk * 32;
        const int w = min(32, a.NT - cb);
        uint32_t r[32];
        tmem_ld16(trow0 + (uint32_t)cb, r);
        if (w == 32) tmem_ld16(trow0 + (uint32_t)(cb + 16), r + 16);
        tmem_ld_wait();
        if (blk == nblk - 1) {                       // accumulator fully read: hand the TMEM set back
          tc_fence_before();
          if (rank == 0) mbar_arrive(ACC_EMPTY(set)); else mbar_arrive_remote(acc_empty_leader + 8u * set);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q * 4 < w)
            pad[lane * 8 + (q ^ (lane & 7))] =
                make_float4(__uint_as_float(r[4 * q]) * a.wscale, __uint_as_float(r[4 * q + 1]) * a.wscale,
                            __uint_as_float(r[4 * q + 2]) * a.wscale, __uint_as_float(r[4 * q + 3]) * a.wscale);
        __syncwarp();
        if (w == 32) {
          float4 rcur[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) rcur[k] = rnext[k];
          if ((blk + 1) * 32 + 32 <= a.NT) res_fetch(blk + 1);
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + cb + q8 * 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = k * 4 + (lane >> 3);
            float4 v = pad[row * 8 + (q8 ^ (row & 7))];
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (resp && px8[k] >= 0) { v.x += rcur[k].x; v.y += rcur[k].y; v.z += rcur[k].z; v.w += rcur[k].w; }
            v.x *= a.oscale; v.y *= a.oscale; v.z *= a.oscale; v.w *= a.oscale;
            if (a.act_out) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
            if (px8[k] >= 0) *reinterpret_cast<float4*>(dstp + (long long)px8[k] * a.Cout + n0 + cb + q8 * 4) = v;
            if (a.stats) pad[row * 8 + (q8 ^ (row & 7))] = v;      // same thread re-reads it below
          }
          if (a.stats) {
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
              if (!(jmask & (1u << jj))) continue;
              long long s1[4] = {0, 0, 0, 0};
              unsigned long long s2[4] = {0, 0, 0, 0};
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int row = k * 4 + (lane >> 3);
                if (__shfl_sync(0xffffffffu, myj, row) == jj) {
                  const float4 v = pad[row * 8 + (q8 ^ (row & 7))];
                  const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    int xi = __float2int_rn(f[e] * STAT_SCALE);
                    xi = max(-STAT_CLAMP, min(STAT_CLAMP, xi));
                    s1[e] += xi;
                    s2[e] += (unsigned long long)((long long)xi * (long long)xi);
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8);
                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
              }
              if (lane < 8) {
                unsigned long long* sp = stat_s + (size_t)jj * 2 * a.NT + cb + q8 * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  atomicAdd(sp + e, (unsigned long long)s1[e]);
                  atomicAdd(sp + a.NT + e, s2[e]);
                }
              }
            }
          }
        } else {                                          // 16-wide tail: 4 lanes per row, 8 rows per instruction
          const int q = lane & 3, rsub = lane >> 2;
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + cb + q * 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = k * 8 + rsub;
            const int px = __shfl_sync(0xffffffffu, mypix, row);
            float4 v = pad[row * 8 + (q ^ (row & 7))];
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            const long long off = (long long)px * a.Cout + n0 + cb + q * 4;
            if (resp && px >= 0) {
              const float4 rv = ld_nc_na(resp + off);
              v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
            }
            v.x *= a.oscale; v.y *= a.oscale; v.z *= a.oscale; v.w *= a.oscale;
            if (a.act_out) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
            if (px >= 0) *reinterpret_cast<float4*>(dstp + off) = v;
            if (a.stats) pad[row * 8 + (q ^ (row & 7))] = v;
          }
          if (a.stats) {
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
              if (!(jmask & (1u << jj))) continue;
              long long s1[4] = {0, 0, 0, 0};
              unsigned long long s2[4] = {0, 0, 0, 0};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int row = k * 8 + rsub;
                if (__shfl_sync(0xffffffffu, myj, row) == jj) {
                  const float4 v = pad[row * 8 + (q ^ (row & 7))];
                  const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    int xi = __float2int_rn(f[e] * STAT_SCALE);
                    xi = max(-STAT_CLAMP, min(STAT_CLAMP, xi));
                    s1[e] += xi;
                    s2[e] += (unsigned long long)((long long)xi * (long long)xi);
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                  s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                  s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
                }
              }
              if (lane < 4) {
                unsigned long long* sp = stat_s + (size_t)jj * 2 * a.NT + cb + q * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  atomicAdd(sp + e, (unsigned long long)s1[e]);
                  atomicAdd(sp + a.NT + e, s2[e]);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      if (a.stats) {
        // the four warps' contributions to this 128-row tile -> global [tile][NJ][2][Cout], then re-zero
        named_bar_sync(2, 128);
        unsigned long long* gp = a.stats + (size_t)mtile * a.NJ * 2 * a.Cout;
        for (int i = et; i < a.NJ * 2 * a.NT; i += 128) {
          const int jp = i / a.NT, n = i - jp * a.NT;
          gp[(size_t)jp * a.Cout + n0 + n] = stat_s[i];
          stat_s[i] = 0ull;
        }
      }
      DBG_ADD(10, te, tid == W_EPI * 32);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's shared memory / TMEM / barriers stay valid until both are done
  if (dbg && tid == 0) dbg[0] = clock64() - t_begin;
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ---- weight packing ---------------------------------------------------------------------------------
// in : w_taps fp32 [taps][Cin][Cout]
// out: fp16 stage images; stage (nt, kb, tap) of a unit sits at ((nt*per_unit + stage_off + kb*taps + tap)*2 + rank)
//      * b_stage bytes, b_stage = (KB/16) * 32*NT; inside: k16 step s: hi[2 chunks][NT/2 cols][8] | lo[...],
//      rank r holding output columns nt*NT + r*NT/2 + [0, NT/2)
__global__ void k_pack_weights2(const float* __restrict__ w, __half* __restrict__ out, int taps, int Cin, int Cout,
                                int NT, int KB, float scale, int stage_off, int per_unit) {
  const int ksteps = KB / 16, nKB = Cin / KB, nNT = Cout / NT, NH = NT / 2;
  const long long total = (long long)nNT * nKB * taps * 2 * ksteps * 2 * NH * 8;
  const long long stage_halfs = (long long)ksteps * 16 * NT;        // b_stage / 2 bytes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i % 8);
    long long t = i / 8;
    int n = (int)(t % NH); t /= NH;
    int jc = (int)(t % 2); t /= 2;
    int s = (int)(t % ksteps); t /= ksteps;
    int rk = (int)(t % 2); t /= 2;
    int tap = (int)(t % taps); t /= taps;
    int kb = (int)(t % nKB); t /= nKB;
    int nt = (int)t;
    const int c = kb * KB + s * 16 + jc * 8 + e;
    const float v = w[((long long)tap * Cin + c) * Cout + nt * NT + rk * NH + n] * scale;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const long long stage = ((long long)nt * per_unit + stage_off + kb * taps + tap) * 2 + rk;
    const long long base = stage * stage_halfs + (long long)s * 16 * NT;
    const long long o = (long long)(jc * NH + n) * 8 + e;
    out[base + o] = h;
    out[base + 8LL * NT + o] = l;
  }
}

struct Plan {
  int KB, NT, HP, halo0, SA, R, NB, NJ, Wp, Pimg, tmem_cols;
  uint32_t off_img, off_raw, off_b, off_pad, off_row, off_stat, off_bias, off_bar, a_plane, raw_stage, b_stage;
  size_t smem;
};

int kb_of(int C0, int C1, int C2, int C3) {
  auto ok = [](int c, int m) { return c % m == 0; };
  if (ok(C0, 32) && ok(C1, 32) && ok(C2, 32) && ok(C3, 32)) return 32;
  if (ok(C0, 16) && ok(C1, 16) && ok(C2, 16) && ok(C3, 16)) return 16;
  return 0;
}

// shared-memory plan for one conv; returns false when nothing fits
bool make_plan(int H, int W, int ks, int C0, int C1, int C2, int C3, int NT, int KB, bool stats, Plan& p) {
  p.KB = KB; p.NT = NT;
  if (ks == 3) { p.Wp = W + 1; p.Pimg = (H + 1) * (W + 1); p.halo0 = p.Wp + 1; }
  else { p.Wp = W; p.Pimg = H * W; p.halo0 = 0; }
  int hp = MT + 2 * p.halo0;
  const int want = (KB == 32) ? 2 : 4;                   // HP mod 8 that keeps the image stores conflict-free
  while (hp % 8 != want) ++hp;
  p.HP = hp;
  p.NJ = (MT - 1) / p.Pimg + 2;
  p.a_plane = (uint32_t)(KB / 8) * hp * 16;
  p.raw_stage = (uint32_t)hp * KB * 4;
  p.b_stage = (uint32_t)(KB / 16) * 32 * NT;
  const int nKB0 = (C0 + C1) / KB, nKB = nKB0 + (C2 + C3) / KB;
  const size_t stat_bytes = stats ? (size_t)p.NJ * 2 * NT * 8 : 0;
  const size_t fixed = 4 * 4096 + (size_t)2 * hp * 8 + stat_bytes + 1024 + 1024;
  const size_t limit = 227 * 1024;
  const size_t a_stage = 2 * (size_t)p.a_plane;
  int SA = 2, R = 2, NB = 3;
  if (fixed + SA * a_stage + R * (size_t)p.raw_stage + NB * (size_t)p.b_stage > limit) return false;
  // grow: weights ring first (to 4), then the raw ring / image stages of short K-blocks (1x1), then weights again
  auto fits = [&](int sa, int r, int nb) {
    return fixed + sa * a_stage + r * (size_t)p.raw_stage + nb * (size_t)p.b_stage <= limit;
  };
  while (NB < 4 && fits(SA, R, NB + 1)) ++NB;
  const int r_max = (nKB + 1 < 4) ? nKB + 1 : 4;         // prefetch distance R-1 <= K-blocks per unit
  if (ks == 1) {
    while ((R < r_max || SA < 4) ) {
      bool grew = false;
      if (R < r_max && fits(SA, R + 1, NB)) { ++R; grew = true; }
      if (SA < 4 && fits(SA + 1, R, NB)) { ++SA; grew = true; }
      if (!grew) break;
    }
  }
  while (NB < 8 && fits(SA, R, NB + 1)) ++NB;
  if (ks == 3) {
    if (R < r_max && R < 3 && fits(SA, R + 1, NB)) ++R;
    if (SA < 3 && fits(SA + 1, R, NB)) ++SA;
  }
  if (R > r_max) R = r_max;
  if (R < 2) R = 2;
  p.SA = SA; p.R = R; p.NB = NB;
  size_t off = 0;
  p.off_img = (uint32_t)off; off += SA * a_stage;
  p.off_raw = (uint32_t)off; off += R * (size_t)p.raw_stage;
  off = (off + 127) & ~(size_t)127;
  p.off_b = (uint32_t)off; off += NB * (size_t)p.b_stage;
  off = (off + 127) & ~(size_t)127;
  p.off_pad = (uint32_t)off; off += 4 * 4096;
  p.off_row = (uint32_t)off; off += (size_t)2 * hp * 8;
  off = (off + 15) & ~(size_t)15;
  p.off_stat = (uint32_t)off; off += stat_bytes;
  p.off_bias = (uint32_t)off; off += 1024;
  p.off_bar = (uint32_t)off; off += 8 * (3 * SA + 3 * NB + 4) + 16;
  p.smem = off;
  int cols = 2 * NT, p2 = 32;
  while (p2 < cols) p2 <<= 1;
  p.tmem_cols = p2;
  return off <= limit && p2 <= 512;
}

// K-block size: 32 channels when every source allows it and the slab fits, else 16
int plan_kb(int H, int W, int ks, int C0, int C1, int C2, int C3, int NT, bool stats) {
  int kb = kb_of(C0, C1, C2, C3);
  Plan p;
  while (kb >= 16) {
    if (make_plan(H, W, ks, C0, C1, C2, C3, NT, kb, stats, p)) return kb;
    kb >>= 1;
  }
  return 0;
}

}  // namespace

int launch_conv_umma2(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.w && op.dst && (op.C1 == 0 || op.src1), "CONV_UMMA: null pointer");
  MCVD_CHECK(op.i0 == 1 || op.i0 == 3, "CONV_UMMA: kernel size %d unsupported", op.i0);
  C2Args a;
  memset(&a, 0, sizeof(a));
  a.s0 = (const float*)op.src0; a.s1 = (const float*)op.src1;
  a.s2 = (const float*)op.src2; a.s3 = (const float*)op.src3;
  a.C0 = op.C0; a.C1 = op.C1; a.C2 = op.src2 ? op.C2 : 0; a.C3 = op.src3 ? op.C3 : 0;
  a.wpk = (const uint8_t*)op.w; a.bias = (const float*)op.bias; a.res = (const float*)op.aux0;
  a.tab3 = (const float*)op.aux1; a.dst = (float*)op.dst;
  a.stats = (unsigned long long*)op.dst2;
  a.dbg = (long long*)op.aux2;
  a.B = op.B; a.H = op.H; a.W = op.W; a.Cout = op.Cout; a.ks = op.i0; a.HW = op.H * op.W;
  a.NT = op.i1;
  MCVD_CHECK(a.NT >= 16 && a.NT <= 256 && a.NT % 16 == 0 && op.Cout % a.NT == 0,
             "CONV_UMMA: n tile %d invalid for Cout %d", a.NT, op.Cout);
  const bool stats = a.stats != nullptr;
  const int KB = plan_kb(op.H, op.W, op.i0, a.C0, a.C1, a.C2, a.C3, a.NT, stats);
  MCVD_CHECK(KB != 0, "CONV_UMMA: channels (%d,%d | %d,%d) must be multiples of 16 and the %dx%d slab must fit",
             a.C0, a.C1, a.C2, a.C3, op.H, op.W);
  MCVD_CHECK(op.i2 == 0 || op.i2 == KB, "CONV_UMMA: weights were packed for K-block %d, the plan says %d", op.i2, KB);
  Plan p;
  make_plan(op.H, op.W, op.i0, a.C0, a.C1, a.C2, a.C3, a.NT, KB, stats, p);
  a.KB = KB; a.HP = p.HP; a.halo0 = p.halo0; a.NB = p.NB; a.SA = p.SA; a.R = p.R; a.NJ = p.NJ;
  a.Wp = p.Wp; a.Pimg = p.Pimg; a.tmem_cols = p.tmem_cols;
  a.off_img = p.off_img; a.off_raw = p.off_raw; a.off_b = p.off_b; a.off_pad = p.off_pad; a.off_row = p.off_row;
  a.off_stat = p.off_stat; a.off_bias = p.off_bias; a.off_bar = p.off_bar;
  a.a_plane = p.a_plane; a.raw_stage = p.raw_stage; a.b_stage = p.b_stage;
  a.Qtot = (long long)op.B * a.Pimg;
  MCVD_CHECK((long long)op.B * op.H * op.W < (1LL << 31) && a.Qtot < (1LL << 31), "CONV_UMMA: too many pixels");
  MCVD_CHECK(!stats || a.Pimg >= 64, "CONV_UMMA: epilogue statistics need images of >= 64 positions");
  a.nKB0 = (a.C0 + a.C1) / KB;
  a.nKB = a.nKB0 + (a.C2 + a.C3) / KB;
  a.act_in = (op.flags & MCVD_F_ACT_IN) ? 1 : 0;
  a.act_out = (op.flags & MCVD_F_ACT_OUT) ? 1 : 0;
  a.split = (op.i3 >= 1 && op.i3 <= 3) ? op.i3 : 3;
  if (op.i3 == 4) a.split = 0;                     // single fp16 MMA (experiments only)
  a.wscale = op.f1; a.oscale = op.f0;
  a.dbgf = op.i7;
  a.tiles_n = op.Cout / a.NT;
  const long long pairs_m = (a.Qtot + 2 * MT - 1) / (2 * MT);
  a.nunits = (int)(pairs_m * a.tiles_n);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ncl = sms / 2;
  if (a.nunits < ncl) ncl = a.nunits;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_umma2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MCVD_CHECK(e == cudaSuccess, "CONV_UMMA: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  k_conv_umma2<<<2 * ncl, NTHREADS, p.smem, s>>>(a);
  MCVD_CUDA_LAUNCH_CHECK("conv_umma2");
  return 0;
}

}  // namespace mcvd

extern "C" int mcvd_umma2_plan(int H, int W, int ks, int C0, int C1, int C2, int C3, int n_tile, int stats) {
  return mcvd::plan_kb(H, W, ks, C0, C1, C2, C3, n_tile, stats != 0);
}

// shared-memory plan of a conv (diagnostics / tests): out[0..7] = KB, HP, image stages, raw-ring stages, weight
// stages, image slots per tile, TMEM columns, dynamic shared memory bytes; returns 0, or -1 when nothing fits
extern "C" int mcvd_umma2_plan_info(int H, int W, int ks, int C0, int C1, int C2, int C3, int n_tile, int stats,
                                    int* out) {
  const int kb = mcvd::plan_kb(H, W, ks, C0, C1, C2, C3, n_tile, stats != 0);
  if (!kb || !out) return -1;
  mcvd::Plan p;
  mcvd::make_plan(H, W, ks, C0, C1, C2, C3, n_tile, kb, stats != 0, p);
  out[0] = kb; out[1] = p.HP; out[2] = p.SA; out[3] = p.R; out[4] = p.NB; out[5] = p.NJ; out[6] = p.tmem_cols;
  out[7] = (int)p.smem;
  return 0;
}

extern "C" long long mcvd_umma2_stats_bytes(int B, int H, int W, int ks, int Cout) {
  const long long pimg = ks == 3 ? (long long)(H + 1) * (W + 1) : (long long)H * W;
  const long long tiles = 2 * ((B * pimg + 2 * mcvd::MT - 1) / (2 * mcvd::MT));
  const long long nj = (mcvd::MT - 1) / pimg + 2;
  return tiles * nj * 2 * Cout * 8;
}

extern "C" long long mcvd_umma2_pack_weights(const float* w_taps, int taps, int Cin, int Cout, int n_tile, int KB,
                                             void* out, int scale_log2, int stage_off, int per_unit, void* stream) {
  if ((KB != 16 && KB != 32) || Cin % KB || n_tile < 16 || n_tile % 16 || Cout % n_tile) {
    mcvd::set_error("umma2_pack_weights: Cin %d / Cout %d / n_tile %d / KB %d unsupported", Cin, Cout, n_tile, KB);
    return -1;
  }
  const long long bytes = (long long)taps * Cin * Cout * 4;      // hi + lo fp16
  if (!out) return bytes;
  if (!w_taps) {
    mcvd::set_error("umma2_pack_weights: null input");
    return -1;
  }
  const float scale = ldexpf(1.0f, scale_log2);
  const long long total = (long long)taps * Cin * Cout;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  mcvd::k_pack_weights2<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_taps, (__half*)out, taps, Cin, Cout,
                                                                            n_tile, KB, scale, stage_off, per_unit);
  if (cudaGetLastError() != cudaSuccess) {
    mcvd::set_error("umma2_pack_weights: launch failed");
    return -2;
  }
  return bytes;
}
