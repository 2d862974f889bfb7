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
// What the measurements of round 2 said (profiles/r2_mma_rate_microbench.txt, r2_conv2_isolation.txt) and what
// this kernel does about it:
//   * Back-to-back tcgen05.mma run at the tensor floor (N/2 cycles per 128 rows per SM for N >= 96 in cta_group::2,
//     any operand layout, shifted descriptors included): the round-1 "A-operand read floor" does not exist.  What
//     cost round 1 (and the first version of this file) a factor 2-3 was the pipeline AROUND the MMAs.
//   * Weights.  Re-streaming the weight tile for every 128 positions needs ~20 B/clk per SM = 5.8 TB/s from L2
//     at full tensor speed; measured, the loads alone cost +50 % (L2 delivers these hot lines at ~2.8 TB/s).  So a
//     unit of work is J (= 2..4) position tiles per CTA against one n tile: loop order K-block -> tap -> tile, the
//     J slabs of a K-block are resident together and every weight stage is used J times.  The pair splits the n
//     tile (cta_group::2: each CTA stages NT/2 columns), so per SM the weight bytes per position drop 2J-fold
//     against a single-CTA, single-tile schedule.
//   * No forwarding hops.  Weight stages arrive by TMA (cp.async.bulk.tensor ... .cta_group::2) that signals the
//     LEADER's mbarrier from both CTAs; the peer's producer warps arrive on the leader's slab barrier directly
//     (one remote arrive per warp); tcgen05.commit multicasts "stage free" to both CTAs.
//   * One arrive / one polling lane per warp everywhere (128 remote arrives per tile on the accumulator barrier
//     cost ~8k cycles per tile before).
//   * Producers: 9 warps, unit = (slab row, 16-byte piece): whole 128-byte lines from global, conflict-free shared
//     memory, fp32 rows staged through a cp.async ring (zero-fill for padding), thread-private (no barriers).
//   * Epilogue: 8 warps (two groups of four TMEM-lane quadrants, one position tile each), smem-transposed
//     128-byte-line stores, optional exact integer GroupNorm partial sums of the stored output (fixed point
//     2^-16, order-independent => bit-exact under clip sharding), which replaces the k_gn_partial pass.
//
// Warp roles (640 threads, one CTA per SM, clusters of 2):
//   warps 0-8 producers | 9 weight loader (TMA) | 10-11 MMA issuers (leader CTA only) | 12-19 epilogue
#include <cuda.h>
#include <cuda_fp16.h>

#include "mcvd_common.cuh"
#include "umma_ptx.cuh"

namespace mcvd {

namespace {

using namespace ptx;

constexpr int NPROD = 288;
constexpr int NPW = NPROD / 32;
constexpr int W_LOAD = 9;
constexpr int W_MMA = 10;               // two MMA-issuing warps (10, 11): warp m issues for position tiles j = m, m+2
constexpr int NMMA_W = 2;
constexpr int W_EPI = 12;
constexpr int NEPI_W = 8;
constexpr int NTHREADS = (W_EPI + NEPI_W) * 32;      // 640
constexpr int MT = 128;                 // positions per CTA and tile
constexpr int JMAX = 4;                 // position tiles per unit (accumulators per TMEM set)
constexpr float STAT_SCALE = 65536.0f;  // fixed-point scale of the epilogue statistics

struct C2Args {
  const float* s0;
  const float* s1;
  const float* s2;
  const float* s3;
  int C0, C1, C2, C3;
  int nKB0, nKB;
  const float* bias;
  const float* res;
  const float* tab3;            // [B][3][C0+C1]: mean | rstd*G | S   (null: raw input)
  float* dst;
  unsigned long long* stats;    // [tiles128][NJ][2][Cout] fixed-point partial sums, or null
  long long* dbg;
  int B, H, W, Cout, ks, Wp, Pimg, HW;
  long long Qtot;
  int NT, KB, HP, halo0, NB, SA, R, J, nsets, tiles_n, PT, njobs, tmem_cols, NJ;
  int act_in, act_out, split;
  int dbgf;                     // timing switches (tools/conv2_check.py): 1 producers skip copies + transform, 2 the
                                // epilogue skips loads / stores, 4 weight stages are not reloaded -- garbage results
  float wscale, oscale;
  uint32_t off_img, off_raw, off_b, off_pad, off_row, off_stat, off_bias, off_bar;
  uint32_t a_plane, raw_stage, b_stage, b_rows;      // b_rows = 512-byte rows of the weight tensor map per stage
};

// flat position -> pixel index (b*H + y)*W + x or -1 (padding / out of range); b_out = image of the position
__device__ __forceinline__ int decode_pos(const C2Args& a, long long q, int& b_out) {
  b_out = 0;
  if (q < 0 || q >= a.Qtot) return -1;
  const int b = (int)(q / a.Pimg);
  b_out = b;
  const int r = (int)(q - (long long)b * a.Pimg);
  const int rr = r / a.Wp, cc = r - rr * a.Wp;
  if (a.ks == 3) {
    if (rr == 0 || cc == 0) return -1;
    return (b * a.H + (rr - 1)) * a.W + (cc - 1);
  }
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

// TMA: one weight stage (b_rows x 512 B) -> this CTA's shared memory, complete_tx on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_stage(uint32_t dst, const CUtensorMap* tmap, int row, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(0), "r"(row)
      : "memory");
}

// the units of a cluster: consecutive (n tile, position-pair tile) jobs [i, end), grouped J at a time inside an n tile
struct Units {
  int i, end, PT, J;
  __device__ Units(const C2Args& a, int cid, int ncl) {
    i = (int)((long long)cid * a.njobs / ncl);
    end = (int)((long long)(cid + 1) * a.njobs / ncl);
    PT = a.PT; J = a.J;
  }
  __device__ bool next(int& nt, int& pt0, int& cnt) {
    if (i >= end) return false;
    nt = i / PT; pt0 = i - nt * PT;
    cnt = min(J, min(end - i, PT - pt0));
    i += cnt;
    return true;
  }
};

template <int KSTEPS, bool DBGC>      // k16 steps per K-block: 1 (KB = 16) or 2 (KB = 32); cycle counters compiled in
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
k_conv_umma2(const __grid_constant__ C2Args a, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + a.off_bar;
  const int SA = a.SA, NB = a.NB;
  // barriers (same offsets in both CTAs; the *_FULL ones are only waited on in the leader)
  auto A_FULL = [&](int i) { return bar0 + 8u * i; };                       // 2 * NPW warp arrivals (local + peer)
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (SA + i); };               // tcgen05.commit, multicast
  auto B_FULL = [&](int i) { return bar0 + 8u * (2 * SA + i); };            // expect_tx (leader) + TMA bytes of both CTAs
  auto B_EMPTY = [&](int i) { return bar0 + 8u * (2 * SA + NB + i); };      // one tcgen05.commit (multicast) per MMA warp
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * SA + 2 * NB + i); };             // [set*JMAX + j], commit mc
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * SA + 2 * NB + 2 * JMAX + i); }; // 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.off_bar + 8 * (2 * SA + 2 * NB + 4 * JMAX));
  const int taps = a.ks * a.ks;

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(A_FULL(i), 2 * NPW); mbar_init(A_EMPTY(i), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(B_FULL(i), 1); mbar_init(B_EMPTY(i), NMMA_W); }
    for (int i = 0; i < 2 * JMAX; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 8); }
    fence_barrier_init();
  }
  if (a.stats) {
    unsigned long long* st = reinterpret_cast<unsigned long long*>(smem + a.off_stat);
    for (int i = tid; i < 2 * a.NJ * 2 * a.NT; i += NTHREADS) st[i] = 0ull;
  }
  if (warp == W_MMA) tmem_alloc2(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = (DBGC && a.dbg) ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;
  const long long t_begin = dbg ? clock64() : 0;
#define DBG_T(var) long long var = (DBGC && dbg) ? clock64() : 0
#define DBG_ADD(slot, since, cond) do { if (DBGC && dbg && (cond)) { long long t__ = clock64(); dbg[slot] += t__ - since; since = t__; } } while (0)

  if (warp < W_LOAD) {
    // =============================== producers ===============================
    // job = (unit, K-block kb, tile j): the slab of position tile pt0+j for channels [kb*KB, +KB), in the order the
    // MMA warp consumes them (unit -> kb -> j).  Copies of job g go to raw stage g % R, its fp16 image to image
    // stage g % SA; the copies run R-1 jobs ahead of the transform.
    const int pj_shift = (a.KB == 32) ? 3 : 2;                 // 16-byte pieces per slab row: 8 or 4
    const int PJ = 1 << pj_shift;
    const int j16 = tid & (PJ - 1);
    const int hrow = tid >> pj_shift;
    const int RPP = NPROD >> pj_shift;                         // rows per pass: 36 or 72
    const int Cin = a.C0 + a.C1;
    const uint32_t raw0 = sbase + a.off_raw, img0 = sbase + a.off_img;
    int* rowinfo = reinterpret_cast<int*>(smem + a.off_row);     // [2 (unit parity)][J][HP] pixel | image << 24, or -1
    int* tile_b0 = rowinfo + 2 * a.J * a.HP;                     // [2][J] first image of each slab
    const uint32_t a_full_leader = mapa_u32(A_FULL(0), 0);
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

    int total = 0;
    {
      Units u0(a, cid, ncl);
      int n_, p_, c_;
      while (u0.next(n_, p_, c_)) total += c_ * a.nKB;
    }
    // two cursors over the same job sequence
    Units ui(a, cid, ncl), ut(a, cid, ncl);
    int i_nt = 0, i_pt0 = 0, i_cnt = 0, i_kb = 0, i_j = 0, i_par = 0;
    int t_nt = 0, t_pt0 = 0, t_cnt = 0, t_kb = 0, t_j = 0, t_par = 0;
    ui.next(i_nt, i_pt0, i_cnt);
    ut.next(t_nt, t_pt0, t_cnt);
    bool i_new = true;                                          // the issue cursor entered a new unit
    for (int it = 0; it < total + Rm1; ++it) {
      // ---- issue the copies of job `it` into raw stage it % R ----
      if (it < total) {
        if (i_new) {
          // slab row -> pixel table of the unit's tiles (the transform cursor may still read the other buffer: it is
          // at most R-1 <= nKB jobs, i.e. at most one unit, behind)
          named_bar_sync(1, NPROD);
          for (int x = tid; x < i_cnt * a.HP; x += NPROD) {
            const int jj = x / a.HP, h = x - jj * a.HP;
            const long long p0 = (long long)(i_pt0 + jj) * (2 * MT) + (long long)rank * MT - a.halo0;
            const int b_first = p0 <= 0 ? 0 : (int)min((long long)(a.B - 1), p0 / a.Pimg);   // image of the slab's first row
            int b;
            const int pix = decode_pos(a, p0 + h, b);
            // pixel index (< 2^24, checked by the launcher) | image relative to the slab's first one << 24
            rowinfo[(i_par * a.J + jj) * a.HP + h] = pix < 0 ? -1 : (pix | ((b - b_first) << 24));
            if (h == 0) tile_b0[i_par * a.J + jj] = b_first;
          }
          named_bar_sync(1, NPROD);
          i_new = false;
        }
        const float* src; int cs, cc0;
        src_of(i_kb, src, cs, cc0);
        // slab row h always belongs to thread group h % RPP (whatever the segment), so a raw-ring slot is only ever
        // touched by one thread and needs no barrier; the 1x1 shortcut segment stages the centre rows only
        const bool seg1 = i_kb >= a.nKB0;
        const int hlo = seg1 ? a.halo0 : 0, hhi = seg1 ? a.halo0 + MT : a.HP;
        const uint32_t rst = raw0 + (uint32_t)(it % a.R) * a.raw_stage;
        const float* sj = src + cc0 + j16 * 4;
        const int* ri = rowinfo + (i_par * a.J + i_j) * a.HP;
        if (!(a.dbgf & 1)) {
#pragma unroll 2
          for (int h = hrow; h < hhi; h += RPP) {
            if (h < hlo) continue;
            const int info = ri[h];
            const float* p = info >= 0 ? sj + (long long)(info & 0xffffff) * cs : src;
            cp_async16(rst + (uint32_t)((h << pj_shift) + j16) * 16u, p, info >= 0 ? 16u : 0u);
          }
        }
        if (++i_j == i_cnt) {
          i_j = 0;
          if (++i_kb == a.nKB) { i_kb = 0; ui.next(i_nt, i_pt0, i_cnt); i_par ^= 1; i_new = true; }
        }
      }
      cp_async_commit();
      // ---- transform job it - (R-1) ----
      const int gt = it - Rm1;
      if (gt < 0) continue;
      // norm table of this job's K-block for the slab's first image: requested now, so the L2 round trip of the three
      // loads overlaps the waits below (every job reads other channels, so they always miss L1)
      const bool seg1 = t_kb >= a.nKB0;
      const bool norm = a.tab3 != nullptr && !seg1;
      const float* tabc = a.tab3 + t_kb * a.KB + j16 * 4;
      const int b0_tile = tile_b0[t_par * a.J + t_j];
      float4 tm = make_float4(0.f, 0.f, 0.f, 0.f), tg = tm, ts = tm;
      int tb_img = -1;
      if (norm && !(a.dbgf & 1)) {
        const float* tb = tabc + (long long)b0_tile * 3 * Cin;
        tm = __ldg(reinterpret_cast<const float4*>(tb));
        tg = __ldg(reinterpret_cast<const float4*>(tb + Cin));
        ts = __ldg(reinterpret_cast<const float4*>(tb + 2 * Cin));
        tb_img = b0_tile;
      }
      if (Rm1 == 1) cp_async_wait<1>(); else if (Rm1 == 2) cp_async_wait<2>(); else cp_async_wait<3>();
      const int st = gt % SA;
      DBG_T(tp);
      // one lane per warp polls; __syncwarp orders the other lanes behind lane 0's acquire
      if (lane == 0) { if (a.dbgf & 8) mbar_wait(A_EMPTY(st), ((gt / SA) & 1) ^ 1); else mbar_wait_parked(A_EMPTY(st), ((gt / SA) & 1) ^ 1); }
      __syncwarp();
      DBG_ADD(1, tp, tid == 0);
      const int hlo = seg1 ? a.halo0 : 0, hhi = seg1 ? a.halo0 + MT : a.HP;
      const uint32_t rst = raw0 + (uint32_t)(gt % a.R) * a.raw_stage;
      const uint32_t hi_base = img0 + (uint32_t)st * 2u * a.a_plane, lo_base = hi_base + a.a_plane;
      const uint32_t img_off = (uint32_t)(j16 >> 1) * (uint32_t)a.HP * 16u + (uint32_t)(j16 & 1) * 8u;
      const int* ri = rowinfo + (t_par * a.J + t_j) * a.HP;
      // Two slab rows per trip (independent instruction chains; more unrolling cost more in instruction-cache misses
      // than it gained).  (mean, rstd*G, S) of this thread's 4 channels are cached for the image of the previous row:
      // a slab touches one or two images on the large maps.
      if (!(a.dbgf & 1)) {
#pragma unroll 1
        for (int hb = hrow; hb < hhi; hb += 2 * RPP) {
          int info[2];
          float4 x[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int h = hb + i * RPP;
            info[i] = -2;                                        // -2: no such row for this thread
            if (h < hhi && h >= hlo) {
              info[i] = ri[h];
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                           : "r"(rst + (uint32_t)((h << pj_shift) + j16) * 16u));
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (info[i] == -2) continue;
            const int h = hb + i * RPP;
            uint32_t h0 = 0u, h1 = 0u, l0 = 0u, l1 = 0u;
            if (info[i] >= 0) {
              float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
              if (norm) {
                const int img = b0_tile + (info[i] >> 24);
                if (img != tb_img) {
                  const float* tb = tabc + (long long)img * 3 * Cin;
                  tm = __ldg(reinterpret_cast<const float4*>(tb));
                  tg = __ldg(reinterpret_cast<const float4*>(tb + Cin));
                  ts = __ldg(reinterpret_cast<const float4*>(tb + 2 * Cin));
                  tb_img = img;
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
        }
      }
      DBG_ADD(3, tp, tid == 0);
      fence_proxy_async();                // generic-proxy smem stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) {                    // one arrival per producer warp, on the LEADER's barrier
        if (rank == 0) mbar_arrive(A_FULL(st)); else mbar_arrive_remote(a_full_leader + 8u * st);
      }
      DBG_ADD(4, tp, tid == 0);
      if (++t_j == t_cnt) {
        t_j = 0;
        if (++t_kb == a.nKB) { t_kb = 0; ut.next(t_nt, t_pt0, t_cnt); t_par ^= 1; }
      }
    }
  } else if (warp == W_LOAD) {
    // =============================== weight loader (both CTAs) ===============================
    if (elect_one()) {
      const uint32_t b0 = sbase + a.off_b;
      const uint32_t b_full_leader = mapa_u32(B_FULL(0), 0);
      const int per_unit = a.nKB0 * taps + (a.nKB - a.nKB0);       // weight stages per unit
      Units un(a, cid, ncl);
      int nt, pt0, cnt, st = 0, ph = 1;
      while (un.next(nt, pt0, cnt)) {
        const int row0 = (nt * per_unit * 2 + (int)rank) * (int)a.b_rows;
        for (int i = 0; i < per_unit; ++i) {
          if (a.dbgf & 8) mbar_wait(B_EMPTY(st), ph); else mbar_wait_parked(B_EMPTY(st), ph);
          if (rank == 0) mbar_arrive_expect_tx(B_FULL(st), 2u * a.b_stage);     // both halves report here
          if (a.dbgf & 4) {
            // timing switch: no reload; complete the transaction count by hand
            asm volatile("mbarrier.complete_tx.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(b_full_leader + 8u * st),
                         "r"(a.b_stage) : "memory");
          } else {
            tma_load_stage(b0 + (uint32_t)st * a.b_stage, &wmap, row0 + i * 2 * (int)a.b_rows, b_full_leader + 8u * st);
          }
          if (++st == NB) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp < W_EPI) {
    if (rank == 0) {
      // =============================== MMA issuers (leader CTA) ===============================
      // Two warps, one per position tile parity: warp m issues the MMAs of tiles j = m, m+2 (each accumulator is fed
      // by one warp, in order), waits for its own slabs, and both commit every weight stage (B_EMPTY counts 2).  One
      // warp alone was bound by the latency of its own instruction stream (~130 instructions / ~950 cycles per weight
      // stage at ~7 cycles per dependent instruction, 3x the tensor time of the 6 MMAs: ncu source page in
      // profiles/r2_ncu_conv2_summary.txt).  The whole warp walks the loops (uniform control flow); only tcgen05.mma /
      // tcgen05.commit sit under elect.sync.  Descriptors differ only in their low word (start address field,
      // < 2^14 16-byte units, never carries into the LBO field).  The waits are plain CTA-scope try_waits: an
      // .acquire.cluster wait compiles to TRYWAIT + CCTL.IVALL (an L1 invalidation per stage), and nothing this thread
      // reads afterwards needs it -- the operands are read by each SM's tensor core from its own shared memory,
      // published by fence.proxy.async (slabs) or written by TMA (weights).
      const int m = warp - W_MMA;
      const uint32_t idesc = make_idesc_f16(2 * MT, a.NT);
      const uint64_t a_proto = make_desc(0, (uint32_t)a.HP * 16, 128), b_proto = make_desc(0, (uint32_t)(a.NT >> 1) * 16, 128);
      const uint32_t a_hiw = (uint32_t)(a_proto >> 32), b_hiw = (uint32_t)(b_proto >> 32);
      const uint32_t a_plane16 = a.a_plane >> 4, b_step16 = (32u * a.NT) >> 4, b_lo16 = (16u * a.NT) >> 4;
      const uint32_t a_kstep16 = 2u * (uint32_t)a.HP;
      const uint32_t a_stage16 = 2 * a_plane16, b_stage16 = a.b_stage >> 4;
      const uint32_t a_org = (uint32_t)a_proto + ((sbase + a.off_img) >> 4) + (uint32_t)a.halo0;   // stage 0, centre tap
      const uint32_t b_org = (uint32_t)b_proto + ((sbase + a.off_b) >> 4);
      const bool w_lo = (a.split & 2) != 0, a_lo = (a.split & 1) != 0;
      auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
      Units un(a, cid, ncl);
      int nt, pt0, cnt, bst = 0, ast = 0, iu = 0;
      uint32_t bph = 0, aph = 0;
      uint32_t b_cur = b_org;                               // low descriptor word of weight stage bst
      while (un.next(nt, pt0, cnt)) {
        const int set = iu % a.nsets;
        const uint32_t acc_ph = (uint32_t)((iu / a.nsets) & 1);
        DBG_T(tm);
        for (int j = m; j < cnt; j += NMMA_W) mbar_wait(ACC_EMPTY(set * JMAX + j), acc_ph ^ 1);   // epilogue drained it
        DBG_ADD(5, tm, lane == 0 && m == 0);
        tc_fence_after();
        const uint32_t d_set = tmem_base + (uint32_t)(set * a.J * a.NT);
        for (int kb = 0; kb < a.nKB; ++kb) {
          DBG_ADD(8, tm, lane == 0 && m == 0);
          // my (at most two) slabs of this K-block: tile j lives in stage (ast + j) mod SA
          uint32_t a_w[2] = {0u, 0u}, d_w[2] = {0u, 0u};
          int st_w[2] = {0, 0};
          int nmine = 0;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int j = m + q * NMMA_W;
            if (j < cnt) {
              int st = ast + j;
              uint32_t ph = aph;
              if (st >= SA) { st -= SA; ph ^= 1; }
              mbar_wait(A_FULL(st), ph);
              a_w[q] = a_org + (uint32_t)st * a_stage16;
              d_w[q] = d_set + (uint32_t)(j * a.NT);
              st_w[q] = st;
              nmine = q + 1;
            }
          }
          DBG_ADD(6, tm, lane == 0 && m == 0);
          const bool main = kb < a.nKB0;
          const int ntap = main ? taps : 1;
          // (no unrolling over the taps: the hot code of all roles has to stay inside the instruction cache -- with the
          // taps unrolled the kernel ran at a 70 % i-cache hit rate and "no instruction" was the top stall reason)
          int dy = (ntap == 1) ? 0 : -1, dx = dy;
#pragma unroll 1
          for (int tap = 0; tap < ntap; ++tap) {
            DBG_ADD(8, tm, lane == 0 && m == 0);
            mbar_wait(B_FULL(bst), bph);
            DBG_ADD(7, tm, lane == 0 && m == 0);
            const uint32_t shift = (uint32_t)(dy * a.Wp + dx);
            if (++dx == 2) { dx = -1; ++dy; }
            const uint32_t first = (kb > 0 || tap > 0) ? 1u : 0u;      // 0 only for the first MMA of an accumulator
            if (elect_one()) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                if (q < nmine) {
                  uint32_t al = a_w[q] + shift, bl = b_cur;
#pragma unroll
                  for (int s = 0; s < KSTEPS; ++s) {
                    const uint64_t dah = desc(al, a_hiw), dal = desc(al + a_plane16, a_hiw);
                    const uint64_t dbh = desc(bl, b_hiw), dbl = desc(bl + b_lo16, b_hiw);
                    umma2_f16(d_w[q], dah, dbh, idesc, (s > 0) ? 1u : first);
                    if (a_lo) umma2_f16(d_w[q], dal, dbh, idesc, 1u);
                    if (w_lo) umma2_f16(d_w[q], dah, dbl, idesc, 1u);
                    al += a_kstep16; bl += b_step16;
                  }
                }
              }
              umma2_commit_mc(B_EMPTY(bst));           // this warp is done with the weight stage (both CTAs)
            }
            __syncwarp();
            b_cur += b_stage16;
            if (++bst == NB) { bst = 0; bph ^= 1; b_cur = b_org; }
          }
          if (elect_one()) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
              if (q < nmine) umma2_commit_mc(A_EMPTY(st_w[q]));     // slab stages consumed (both CTAs)
          }
          __syncwarp();
          ast += cnt;
          if (ast >= SA) { ast -= SA; aph ^= 1; }
        }
        if (elect_one())
          for (int j = m; j < cnt; j += NMMA_W) umma2_commit_mc(ACC_FULL(set * JMAX + j));
        __syncwarp();
        DBG_ADD(8, tm, lane == 0 && m == 0);
        ++iu;
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    // Two groups of four warps (TMEM lane quadrants); group e owns position tiles j = e, e+2 of every unit.  A warp
    // drains its 32 TMEM lanes (rows) in column blocks of 32 (or a 16-wide tail) through an XOR-swizzled 4 KB
    // transpose pad, so global loads / stores cover whole 128-byte lines of dst / res.
    const int ew = warp - W_EPI, grp = ew >> 2, lq = warp & 3;
    const int et = tid - (W_EPI + 4 * grp) * 32;                                // 0..127 inside the group
    float4* pad = reinterpret_cast<float4*>(smem + a.off_pad) + (size_t)ew * 256;
    float* bias_s = reinterpret_cast<float*>(smem + a.off_bias) + grp * 256;
    unsigned long long* stat_s = reinterpret_cast<unsigned long long*>(smem + a.off_stat) + (size_t)grp * a.NJ * 2 * a.NT;
    const int nblk = (a.NT + 31) / 32;
    const float* __restrict__ resp = a.res;
    float* __restrict__ dstp = a.dst;
    const uint32_t acc_empty_leader = mapa_u32(ACC_EMPTY(0), 0);
    const int bar_id = 2 + grp;
    Units un(a, cid, ncl);
    int nt, pt0, cnt, iu = 0;
    while (un.next(nt, pt0, cnt)) {
      const int set = iu % a.nsets;
      const uint32_t acc_ph = (uint32_t)((iu / a.nsets) & 1);
      ++iu;
      const int n0 = nt * a.NT;
      for (int j = grp; j < cnt; j += 2) {
        const int mtile = (pt0 + j) * 2 + (int)rank;                          // 128-row tile index
        const long long p0 = (long long)mtile * MT;
        named_bar_sync(bar_id, 128);                                         // previous tile's readers are done
        for (int i = et; i < a.NT; i += 128) bias_s[i] = a.bias ? __ldg(a.bias + n0 + i) : 0.f;
        named_bar_sync(bar_id, 128);
        int myb;
        const int mypix = decode_pos(a, p0 + lq * 32 + lane, myb);
        const int tile_b0 = (int)min((long long)(a.B - 1), p0 / a.Pimg);
        const int myj = mypix >= 0 ? myb - tile_b0 : -1;                       // image slot of this row (stats)
        unsigned jmask = 0;                                                   // image slots present in this warp's rows
        if (a.stats) {
#pragma unroll 1
          for (int jj = 0; jj < 4; ++jj)
            if (__ballot_sync(0xffffffffu, myj == jj)) jmask |= 1u << jj;
        }
        DBG_T(te);
        if (lane == 0) { if (a.dbgf & 8) mbar_wait(ACC_FULL(set * JMAX + j), acc_ph); else mbar_wait_parked(ACC_FULL(set * JMAX + j), acc_ph); }
        __syncwarp();
        DBG_ADD(9, te, et == 0);
        tc_fence_after();
        const uint32_t trow0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((set * a.J + j) * a.NT);
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk) {
          const int cb = blk * 32;
          const int w = min(32, a.NT - cb);
          const bool wide = (w == 32);
          uint32_t r[32];
          tmem_ld16(trow0 + (uint32_t)cb, r);
          if (wide) tmem_ld16(trow0 + (uint32_t)(cb + 16), r + 16);
          tmem_ld_wait();
          if (blk == nblk - 1) {                       // accumulator fully read: hand it back (one arrive per warp)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(ACC_EMPTY(set * JMAX + j));
              else mbar_arrive_remote(acc_empty_leader + 8u * (set * JMAX + j));
            }
          }
          if (a.dbgf & 2) continue;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q * 4 < w)
              pad[lane * 8 + (q ^ (lane & 7))] =
                  make_float4(__uint_as_float(r[4 * q]) * a.wscale, __uint_as_float(r[4 * q + 1]) * a.wscale,
                              __uint_as_float(r[4 * q + 2]) * a.wscale, __uint_as_float(r[4 * q + 3]) * a.wscale);
          __syncwarp();
          // transposed phase: 8 lanes per row (32-wide block) or 4 lanes per row (16-wide tail); two rows per trip.
          // Loops are kept rolled on purpose (instruction-cache footprint).
          const int lpr = wide ? 8 : 4;
          const int qc = lane & (lpr - 1), rsub = wide ? (lane >> 3) : (lane >> 2);
          const int rpi = wide ? 4 : 8;                       // rows per instruction
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + cb + qc * 4);
#pragma unroll 1
          for (int r0 = rsub; r0 < 32; r0 += 2 * rpi) {
            float4 v[2], rv[2];
            int px[2];
            long long off[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int row = r0 + u * rpi;
              px[u] = __shfl_sync(0xffffffffu, mypix, row);
              off[u] = (long long)px[u] * a.Cout + n0 + cb + qc * 4;
              rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (resp && px[u] >= 0) rv[u] = ld_nc_na(resp + off[u]);
              v[u] = pad[row * 8 + (qc ^ (row & 7))];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int row = r0 + u * rpi;
              float4 t = v[u];
              t.x = (t.x + bv.x + rv[u].x) * a.oscale; t.y = (t.y + bv.y + rv[u].y) * a.oscale;
              t.z = (t.z + bv.z + rv[u].z) * a.oscale; t.w = (t.w + bv.w + rv[u].w) * a.oscale;
              if (a.act_out) { t.x = silu_fast(t.x); t.y = silu_fast(t.y); t.z = silu_fast(t.z); t.w = silu_fast(t.w); }
              if (px[u] >= 0) *reinterpret_cast<float4*>(dstp + off[u]) = t;
              if (a.stats) pad[row * 8 + (qc ^ (row & 7))] = t;      // same thread re-reads it below
            }
          }
          if (a.stats) {
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
              if (!(jmask & (1u << jj))) continue;
              long long s1[4] = {0, 0, 0, 0};
              unsigned long long s2[4] = {0, 0, 0, 0};
#pragma unroll 2
              for (int row = rsub; row < 32; row += rpi) {
                if (__shfl_sync(0xffffffffu, myj, row) == jj) {
                  const float4 t = pad[row * 8 + (qc ^ (row & 7))];
                  const float f[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int xi = __float2int_rn(f[e] * STAT_SCALE);       // saturates at +-2^31
                    s1[e] += xi;
                    s2[e] += (unsigned long long)((long long)xi * (long long)xi);
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                for (int o = lpr; o <= 16; o <<= 1) {
                  s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                  s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
                }
              }
              if (lane < lpr) {
                unsigned long long* sp = stat_s + (size_t)jj * 2 * a.NT + cb + qc * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  atomicAdd(sp + e, (unsigned long long)s1[e]);
                  atomicAdd(sp + a.NT + e, s2[e]);
                }
              }
            }
          }
          __syncwarp();
        }
        if (a.stats && !(a.dbgf & 2)) {
          // the four warps' contributions to this 128-row tile -> global [tile][NJ][2][Cout], then re-zero
          named_bar_sync(bar_id, 128);
          unsigned long long* gp = a.stats + (size_t)mtile * a.NJ * 2 * a.Cout;
          for (int i = et; i < a.NJ * 2 * a.NT; i += 128) {
            const int jp = i / a.NT, n = i - jp * a.NT;
            gp[(size_t)jp * a.Cout + n0 + n] = stat_s[i];
            stat_s[i] = 0ull;
          }
        }
        DBG_ADD(10, te, et == 0);
      }
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
  int KB, NT, HP, halo0, SA, R, NB, NJ, J, nsets, Wp, Pimg, tmem_cols;
  uint32_t off_img, off_raw, off_b, off_pad, off_row, off_stat, off_bias, off_bar, a_plane, raw_stage, b_stage;
  size_t smem;
};

int kb_max(int C0, int C1, int C2, int C3) {
  auto ok = [](int c, int m) { return c % m == 0; };
  if (ok(C0, 32) && ok(C1, 32) && ok(C2, 32) && ok(C3, 32)) return 32;
  if (ok(C0, 16) && ok(C1, 16) && ok(C2, 16) && ok(C3, 16)) return 16;
  return 0;
}

// shared-memory plan for one conv; returns false when it does not fit
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
  // tiles per unit: as many accumulators as fit twice into TMEM (double-buffered sets), at most JMAX
  int J = 256 / NT;
  if (J > JMAX) J = JMAX;
  if (J < 1) J = 1;
  p.J = J;
  p.nsets = (2 * J * NT <= 512) ? 2 : 1;
  const int nKB0 = (C0 + C1) / KB, nKB = nKB0 + (C2 + C3) / KB;
  const size_t stat_bytes = stats ? (size_t)2 * p.NJ * 2 * NT * 8 : 0;
  const size_t row_bytes = (size_t)2 * J * hp * 4 + 2 * JMAX * 4;
  const size_t fixed = (size_t)NEPI_W * 4096 + row_bytes + stat_bytes + 2 * 1024 + 1024;
  const size_t limit = 227 * 1024;
  const size_t a_stage = 2 * (size_t)p.a_plane;
  // minimum: the J slabs of a K-block resident + one being filled, 2 raw stages, 4 weight stages
  int SA = J + 1, R = 2, NB = 4;
  auto fits = [&](int sa, int r, int nb) {
    return fixed + sa * a_stage + r * (size_t)p.raw_stage + nb * (size_t)p.b_stage <= limit;
  };
  if (!fits(SA, R, NB)) return false;
  const int r_max = (nKB + 1 < 4) ? nKB + 1 : 4;         // prefetch distance R-1 <= K-blocks (>= jobs) per unit
  // grow in the order that matters: weight ring to 8, slabs to 2J (next K-block fully buffered), raw ring, weights
  while (NB < 8 && fits(SA, R, NB + 1)) ++NB;
  while (SA < 2 * J && fits(SA + 1, R, NB)) ++SA;
  while (R < r_max && R < 3 && fits(SA, R + 1, NB)) ++R;
  while (NB < 16 && fits(SA, R, NB + 1)) ++NB;
  while (R < r_max && fits(SA, R + 1, NB)) ++R;
  if (R > r_max) R = r_max;
  if (R < 2) R = 2;
  p.SA = SA; p.R = R; p.NB = NB;
  size_t off = 0;
  p.off_img = (uint32_t)off; off += SA * a_stage;
  p.off_raw = (uint32_t)off; off += R * (size_t)p.raw_stage;
  off = (off + 127) & ~(size_t)127;
  p.off_b = (uint32_t)off; off += NB * (size_t)p.b_stage;
  off = (off + 127) & ~(size_t)127;
  p.off_pad = (uint32_t)off; off += (size_t)NEPI_W * 4096;
  p.off_row = (uint32_t)off; off += row_bytes;
  off = (off + 15) & ~(size_t)15;
  p.off_stat = (uint32_t)off; off += stat_bytes;
  p.off_bias = (uint32_t)off; off += 2 * 1024;
  p.off_bar = (uint32_t)off; off += 8 * (2 * SA + 2 * NB + 4 * JMAX) + 16;
  p.smem = off;
  int cols = p.nsets * J * NT, p2 = 32;
  while (p2 < cols) p2 <<= 1;
  p.tmem_cols = p2;
  return off <= limit && p2 <= 512;
}

// K-block size: 3x3 convs use 16 channels (the J slabs of a K-block and their successors must be co-resident),
// 1x1 convs 32 when every source allows it
int plan_kb(int H, int W, int ks, int C0, int C1, int C2, int C3, int NT, bool stats) {
  int kb = kb_max(C0, C1, C2, C3);
  if (ks == 3 && kb > 16) kb = 16;
  Plan p;
  while (kb >= 16) {
    if (make_plan(H, W, ks, C0, C1, C2, C3, NT, kb, stats, p)) return kb;
    kb >>= 1;
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

int launch_conv_umma2(const McvdOp& op, cudaStream_t s) {
  MCVD_CHECK(op.src0 && op.w && op.dst && (op.C1 == 0 || op.src1), "CONV_UMMA2: null pointer");
  MCVD_CHECK(op.i0 == 1 || op.i0 == 3, "CONV_UMMA2: kernel size %d unsupported", op.i0);
  C2Args a;
  memset(&a, 0, sizeof(a));
  a.s0 = (const float*)op.src0; a.s1 = (const float*)op.src1;
  a.s2 = (const float*)op.src2; a.s3 = (const float*)op.src3;
  a.C0 = op.C0; a.C1 = op.C1; a.C2 = op.src2 ? op.C2 : 0; a.C3 = op.src3 ? op.C3 : 0;
  a.bias = (const float*)op.bias; a.res = (const float*)op.aux0;
  a.tab3 = (const float*)op.aux1; a.dst = (float*)op.dst;
  a.stats = (unsigned long long*)op.dst2;
  a.dbg = (long long*)op.aux2;
  a.B = op.B; a.H = op.H; a.W = op.W; a.Cout = op.Cout; a.ks = op.i0; a.HW = op.H * op.W;
  a.NT = op.i1;
  MCVD_CHECK(a.NT >= 16 && a.NT <= 256 && a.NT % 16 == 0 && op.Cout % a.NT == 0,
             "CONV_UMMA2: n tile %d invalid for Cout %d", a.NT, op.Cout);
  const bool stats = a.stats != nullptr;
  const int KB = plan_kb(op.H, op.W, op.i0, a.C0, a.C1, a.C2, a.C3, a.NT, stats);
  MCVD_CHECK(KB != 0, "CONV_UMMA2: channels (%d,%d | %d,%d) must be multiples of 16 and the %dx%d slab must fit",
             a.C0, a.C1, a.C2, a.C3, op.H, op.W);
  MCVD_CHECK(op.i2 == 0 || op.i2 == KB, "CONV_UMMA2: weights were packed for K-block %d, the plan says %d", op.i2, KB);
  Plan p;
  make_plan(op.H, op.W, op.i0, a.C0, a.C1, a.C2, a.C3, a.NT, KB, stats, p);
  a.KB = KB; a.HP = p.HP; a.halo0 = p.halo0; a.NB = p.NB; a.SA = p.SA; a.R = p.R; a.NJ = p.NJ; a.J = p.J;
  a.nsets = p.nsets; a.Wp = p.Wp; a.Pimg = p.Pimg; a.tmem_cols = p.tmem_cols;
  a.off_img = p.off_img; a.off_raw = p.off_raw; a.off_b = p.off_b; a.off_pad = p.off_pad; a.off_row = p.off_row;
  a.off_stat = p.off_stat; a.off_bias = p.off_bias; a.off_bar = p.off_bar;
  a.a_plane = p.a_plane; a.raw_stage = p.raw_stage; a.b_stage = p.b_stage; a.b_rows = p.b_stage / 512;
  a.Qtot = (long long)op.B * a.Pimg;
  MCVD_CHECK((long long)op.B * op.H * op.W <= (1LL << 24) && a.Qtot < (1LL << 31),
             "CONV_UMMA2: more than 2^24 pixels per launch (B*H*W = %lld); split the batch", (long long)op.B * op.H * op.W);
  MCVD_CHECK(!stats || a.Pimg >= 64, "CONV_UMMA2: epilogue statistics need images of >= 64 positions");
  a.nKB0 = (a.C0 + a.C1) / KB;
  a.nKB = a.nKB0 + (a.C2 + a.C3) / KB;
  a.act_in = (op.flags & MCVD_F_ACT_IN) ? 1 : 0;
  a.act_out = (op.flags & MCVD_F_ACT_OUT) ? 1 : 0;
  a.split = (op.i3 >= 1 && op.i3 <= 3) ? op.i3 : 3;
  if (op.i3 == 4) a.split = 0;                     // single fp16 MMA (experiments only)
  a.wscale = op.f1; a.oscale = op.f0;
  a.dbgf = op.i7;
  a.tiles_n = op.Cout / a.NT;
  a.PT = (int)((a.Qtot + 2 * MT - 1) / (2 * MT));
  a.njobs = a.PT * a.tiles_n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ncl = sms / 2;
  if (a.njobs < ncl) ncl = a.njobs;
  // weight tensor map: the packed stage images as rows of 512 bytes (64 x u64), one box = one stage of one CTA
  EncodeTiledFn enc = encode_tiled();
  MCVD_CHECK(enc != nullptr, "CONV_UMMA2: cuTensorMapEncodeTiled is not available from this driver");
  const int taps = a.ks * a.ks;
  const long long per_unit = (long long)a.nKB0 * taps + (a.nKB - a.nKB0);
  const long long total_rows = per_unit * 2 * a.tiles_n * a.b_rows;
  CUtensorMap wmap;
  {
    cuuint64_t gdim[2] = {64, (cuuint64_t)total_rows};
    cuuint64_t gstride[1] = {512};
    cuuint32_t box[2] = {64, a.b_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(op.w), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCVD_CHECK(r == CUDA_SUCCESS, "CONV_UMMA2: cuTensorMapEncodeTiled failed (%d) for %lld rows of 512 B, box %u rows",
               (int)r, total_rows, a.b_rows);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_umma2<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_umma2<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_umma2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_umma2<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MCVD_CHECK(e == cudaSuccess, "CONV_UMMA2: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const bool dbgc = a.dbg != nullptr;
  if (KB == 16) {
    if (dbgc) k_conv_umma2<1, true><<<2 * ncl, NTHREADS, p.smem, s>>>(a, wmap);
    else k_conv_umma2<1, false><<<2 * ncl, NTHREADS, p.smem, s>>>(a, wmap);
  } else {
    if (dbgc) k_conv_umma2<2, true><<<2 * ncl, NTHREADS, p.smem, s>>>(a, wmap);
    else k_conv_umma2<2, false><<<2 * ncl, NTHREADS, p.smem, s>>>(a, wmap);
  }
  MCVD_CUDA_LAUNCH_CHECK("conv_umma2");
  return 0;
}

}  // namespace mcvd

extern "C" int mcvd_umma2_plan(int H, int W, int ks, int C0, int C1, int C2, int C3, int n_tile, int stats) {
  return mcvd::plan_kb(H, W, ks, C0, C1, C2, C3, n_tile, stats != 0);
}

// shared-memory plan of a conv (diagnostics / tests): out[0..9] = KB, HP, image stages, raw-ring stages, weight
// stages, image slots per tile, TMEM columns, dynamic shared memory bytes, tiles per unit, TMEM sets; returns 0, or -1
extern "C" int mcvd_umma2_plan_info(int H, int W, int ks, int C0, int C1, int C2, int C3, int n_tile, int stats,
                                    int* out) {
  const int kb = mcvd::plan_kb(H, W, ks, C0, C1, C2, C3, n_tile, stats != 0);
  if (!kb || !out) return -1;
  mcvd::Plan p;
  mcvd::make_plan(H, W, ks, C0, C1, C2, C3, n_tile, kb, stats != 0, p);
  out[0] = kb; out[1] = p.HP; out[2] = p.SA; out[3] = p.R; out[4] = p.NB; out[5] = p.NJ; out[6] = p.tmem_cols;
  out[7] = (int)p.smem; out[8] = p.J; out[9] = p.nsets;
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
