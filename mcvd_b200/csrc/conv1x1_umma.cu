// 1x1 convolution / NIN on tcgen05 with the INPUT tile held stationary in shared memory.
//
// Reference semantics: NIN (models/better/layers.py:541-556) of AttnBlockpp's q/k/v/out projections
// (layerspp.py:230-249) and the 1x1 Conv_2 skip projection of ResnetBlockBigGANppGN (layerspp.py:595-624),
// with the GroupNorm of the attention block folded into the staging exactly as conv_umma.cu does.
//
// Why a second kernel (profiles/r2_launch_metrics_v11_cfg2_b64.txt, profiles/r2_k1_attribution_{1,2,3}_*.txt): the general
// kernel re-stages the transformed input once per n-tile and keeps one K-block of register-staged global loads in
// flight, so a 192->576 projection ran at 8-10 % of the tensor pipe and ~1 TB/s of DRAM traffic.  A 1x1 convolution
// has no halo and no taps, so here
//   * a CTA owns a 128-position x Cin tile whose Cin/32 K-blocks are ALL resident in shared memory (when they fit next
//     to two weight stages: Cin <= 320 for the shapes of this network) and loops over every n-tile of the output
//     against that resident tile: the input is converted once.  Wider inputs stream through a ring of 8 operand
//     stages with one n-tile per work item (the 384-channel 8x8 level);
//   * the raw fp32 [128 x 32] box of a K-block is brought in by TMA (128-byte swizzle) straight into that K-block's
//     operand stage the moment the previous tile's MMAs released it -- up to Cin/32 x 16 KB in flight per SM with no
//     registers involved (register-staged loads capped the first version at ~2 TB/s) -- and is converted IN PLACE:
//     the 128 threads of a producer group read their rows, meet at a named barrier, and write the fp16 hi | lo
//     core-matrix image over it (16 KB either way);
//   * the two producer groups own alternating operand STAGES (even / odd): a stage's barriers are then only ever
//     waited on by one group, in order -- a group that could run a whole phase ahead of another group's stage would
//     alias the mbarrier parity bit (a real deadlock on 32-channel inputs before this rule).
// The weight images, operand split and epilogue arithmetic are those of conv_umma.cu (same packed-weight format), so
// results are bit-identical to the general kernel's (tests/test_gpu_ops.py::test_conv1x1_stationary).
//
// Warp roles (608 threads, one CTA per SM):
//   warps 0-3 / 4-7  producer groups 0 / 1: operand stages st = group, group + 2, ...
//   warp  8          weight loader: cp.async.bulk of the packed fp16 hi/lo stage images (ring of NB stages)
//   warp  9          TMEM allocation + single-thread tcgen05.mma issue
//   warps 10-17      epilogue: two warps per TMEM lane quadrant, alternating 32-column blocks; every lane owns one
//                    output position and writes its columns with 256-bit stores (whole 32-byte sectors) -- measured
//                    20 % faster than the shared-memory transpose of conv_umma.cu
//   warp  18         input loader (TMA)
// The bias (all Cout values, once per CTA) and the norm-table rows of the <= 3 images a tile touches (once per tile)
// live in shared memory: as global loads behind asm stores they serialised into one L2 round trip per 8 columns.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "mcvd_common.cuh"
#include "umma_ptx.cuh"

namespace mcvd {

namespace {

using namespace ptx;

constexpr int K1_GROUP = 128;     // threads per producer group
constexpr int K1_W_LOAD = 8;
constexpr int K1_W_MMA = 9;
constexpr int K1_W_EPI = 10;
constexpr int K1_EPI_WARPS = 8;
constexpr int K1_W_ALOAD = K1_W_EPI + K1_EPI_WARPS;   // input loader (TMA)
constexpr int K1_THREADS = (K1_W_ALOAD + 1) * 32;
constexpr int K1_MT = 128;
constexpr int K1_KB = 32;
constexpr int K1_MAX_KB = 24;      // <= 768 input channels
constexpr int K1_RING = 8;         // operand stages when the tile is not resident
constexpr int K1_TAB_IMGS = 3;     // images a 128-position tile can touch (images of >= 64 positions)
constexpr uint32_t K1_A_STAGE = 2u * 4u * K1_MT * 16u;     // hi | lo, 4 chunks of 8 channels, 16 B per row

// TMA: a [128 positions][32 channels] fp32 box of the NHWC activation -> shared memory (128-byte swizzle: the 16-byte
// chunk j of row r lands at chunk j ^ (r & 7)), complete_tx on `bar`; rows past the end of the tensor arrive as zeros
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* tmap, int c0, int row0, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(row0)
      : "memory");
}

struct K1Args {
  const float* s0;
  const float* s1;
  int C0, C1;
  const __half* wpk;
  const float* bias;
  const float* res;
  const float4* tab;       // [B][Cin] (mean, rstd, G, S) or null
  float* dst;
  long long Qtot;          // positions (B * H * W)
  int Pimg, B, Cout;
  int NT, tiles_n, nKB, NB, nsets, tmem_cols;
  int NA;                  // operand stages: == nKB (tile resident, n-tiles loop over it) or a ring of fewer (one n-tile per item)
  int n_items, n_groups, npg;   // work items = m-tiles x n-groups; n-tiles per group
  int act_in, act_out, split;
  float wscale, oscale;
  int dbgf;                // timing experiments (MCVD_K1_DBGF): 1 = no global stores, 2 = no input loads, 4 = no table
  long long* dbg;          // optional per-CTA cycle attribution (tools/umma_timing.py), DBG instantiation only
};

// DBG: per-role cycle attribution into a.dbg[blockIdx.x * 16 + slot]:
//   0 CTA total | producer thread 0: 1 wait RAW_FULL (input box landed), 2 transform + stores | MMA: 3 wait ACC_EMPTY, 4 wait A_FULL,
//   5 wait B_FULL, 6 issue | epilogue warp 0: 7 wait ACC_FULL, 8 TMEM load wait, 10 bias / residual / global stores |
//   loader: 11 wait B_EMPTY
#define K1_T0(var) unsigned var = DBG ? (unsigned)clock() : 0u
#define K1_ACC(acc, since) do { if (DBG) { unsigned t__ = (unsigned)clock(); acc += t__ - since; since = t__; } } while (0)
template <bool DBG>
__global__ void __launch_bounds__(K1_THREADS, 1) k_conv1x1_umma(const __grid_constant__ K1Args a,
                                                              const __grid_constant__ CUtensorMap map0,
                                                              const __grid_constant__ CUtensorMap map1) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t b_step_bytes = 64u * a.NT;
  const uint32_t b_stage_bytes = 2u * b_step_bytes;
  const int Cin = a.C0 + a.C1;
  uint8_t* a_base = smem_raw;
  uint8_t* b_base = a_base + (size_t)a.NA * K1_A_STAGE;
  float* bias_s = reinterpret_cast<float*>(b_base + (size_t)a.NB * b_stage_bytes);                 // [Cout]
  float* tab_s = bias_s + a.Cout;                            // [2 items][K1_TAB_IMGS][mean | rstd*G | S][Cin] (when a.tab)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab_s + (a.tab ? 2 * K1_TAB_IMGS * 3 * Cin : 0));
  const uint32_t bar0 = smem_u32(bars);
  const int NA = a.NA;
  auto A_FULL = [&](int i) { return bar0 + 8u * i; };
  auto A_EMPTY = [&](int i) { return bar0 + 8u * (NA + i); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * NA + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * NA + 2 + i); };
  auto B_FULL = [&](int i) { return bar0 + 8u * (2 * NA + 4 + i); };
  auto B_EMPTY = [&](int i) { return bar0 + 8u * (2 * NA + 4 + a.NB + i); };
  auto RAW_FULL = [&](int i) { return bar0 + 8u * (2 * NA + 4 + 2 * a.NB + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 63);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(A_FULL(i), K1_GROUP / 32); mbar_init(A_EMPTY(i), 1); mbar_init(RAW_FULL(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), K1_EPI_WARPS); }
    for (int i = 0; i < a.NB; ++i) { mbar_init(B_FULL(i), 1); mbar_init(B_EMPTY(i), 1); }
    fence_barrier_init();
  }
  if (warp == K1_W_MMA) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  for (int i = tid; i < a.Cout; i += K1_THREADS) bias_s[i] = a.bias ? __ldg(a.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = DBG ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;
  const unsigned t_begin = DBG ? (unsigned)clock() : 0u;

  if (warp < 8) {
    // =========================== producers ===========================
    const int grp = tid / K1_GROUP, r = tid % K1_GROUP;
    const bool use_tab = a.tab != nullptr && !(DBG && (a.dbgf & 4));
    int it = 0, g = 0;                                          // g: K-blocks staged so far (all items)
    unsigned d1 = 0, d2 = 0;
    K1_T0(tp);
    for (int w = blockIdx.x; w < a.n_items; w += gridDim.x, ++it) {
      const long long q0 = (long long)(w / a.n_groups) * K1_MT;
      const long long p = q0 + r;
      const bool valid = p < a.Qtot;
      const int b_first = (int)(q0 / a.Pimg);
      const int bi = valid ? (int)(p / a.Pimg) - b_first : 0;            // image slot of this row within the tile
      const float* tb = tab_s + ((size_t)(it & 1) * K1_TAB_IMGS + bi) * 3 * Cin;
      if (use_tab) {
        // (mean, rstd*G, S) of the <= 3 images this tile touches, all channels -> smem; double-buffered by item parity
        float* dstb = tab_s + (size_t)(it & 1) * K1_TAB_IMGS * 3 * Cin;
        for (int i = tid; i < K1_TAB_IMGS * Cin; i += 2 * K1_GROUP) {
          const int bj = i / Cin, c = i - bj * Cin;
          const float4 t = __ldg(a.tab + (long long)min(b_first + bj, a.B - 1) * Cin + c);
          float* o = dstb + (size_t)bj * 3 * Cin + c;
          o[0] = t.x; o[Cin] = t.y * t.z; o[2 * Cin] = t.w;
        }
        named_bar_sync(1, 2 * K1_GROUP);
      }
      for (int kb = 0; kb < a.nKB; ++kb, ++g) {
        const int st = g % NA;
        // even stages belong to group 0, odd ones to group 1: a stage's barriers are then only ever waited on by one
        // group, in order (a group that could run a whole phase ahead of another's stage would alias the parity bit)
        if ((st & 1) != grp) continue;
        K1_ACC(d2, tp);
        mbar_wait(RAW_FULL(st), (g / NA) & 1);                // the raw fp32 box of this K-block has landed in the stage
        K1_ACC(d1, tp);
        uint8_t* hi_base = a_base + (size_t)st * K1_A_STAGE;
        uint8_t* lo_base = hi_base + K1_A_STAGE / 2;
        float4 cur[8];
        {
          const uint8_t* rowp = hi_base + (size_t)r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) cur[j] = *reinterpret_cast<const float4*>(rowp + ((j ^ (r & 7)) << 4));
        }
        named_bar_sync(2 + grp, K1_GROUP);                    // converted IN PLACE: every row is in registers first
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 hv = make_uint4(0u, 0u, 0u, 0u), lv = hv;
          if (valid) {
            float v[8] = {cur[2 * ch].x, cur[2 * ch].y, cur[2 * ch].z, cur[2 * ch].w,
                          cur[2 * ch + 1].x, cur[2 * ch + 1].y, cur[2 * ch + 1].z, cur[2 * ch + 1].w};
            if (use_tab) {
              const float* t0 = tb + kb * K1_KB + ch * 8;
              const float4 m0 = *reinterpret_cast<const float4*>(t0), m1 = *reinterpret_cast<const float4*>(t0 + 4);
              const float4 g0 = *reinterpret_cast<const float4*>(t0 + Cin), g1 = *reinterpret_cast<const float4*>(t0 + Cin + 4);
              const float4 s0 = *reinterpret_cast<const float4*>(t0 + 2 * Cin), s1 = *reinterpret_cast<const float4*>(t0 + 2 * Cin + 4);
              const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e] - mm[e], gg[e], ss[e]);
              if (a.act_in) silu_fast8(v);
            }
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
            hv = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          const size_t off = ((size_t)ch * K1_MT + r) * 16;
          *reinterpret_cast<uint4*>(hi_base + off) = hv;
          *reinterpret_cast<uint4*>(lo_base + off) = lv;
        }
        fence_proxy_async();          // generic-proxy stores -> visible to the tensor-core proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(A_FULL(st));
      }
    }
    if (DBG && tid == 0) { dbg[1] = d1; dbg[2] = d2; }
  } else if (warp == K1_W_LOAD) {
    // =========================== weight loader ===========================
    if (elect_one()) {
      const uint32_t b0 = smem_u32(b_base);
      int st = 0, ph = 1;
      unsigned d11 = 0;
      for (int w = blockIdx.x; w < a.n_items; w += gridDim.x) {
        const int ng = w % a.n_groups;
        const int n_lo = ng * a.npg, n_hi = min(a.tiles_n, n_lo + a.npg);
        for (int n = n_lo; n < n_hi; ++n) {
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)n * a.nKB * b_stage_bytes;
          for (int kb = 0; kb < a.nKB; ++kb) {
            K1_T0(tl);
            mbar_wait(B_EMPTY(st), ph);
            K1_ACC(d11, tl);
            mbar_arrive_expect_tx(B_FULL(st), b_stage_bytes);
            bulk_g2s(b0 + (uint32_t)st * b_stage_bytes, wsrc + (size_t)kb * b_stage_bytes, b_stage_bytes, B_FULL(st));
            if (++st == a.NB) { st = 0; ph ^= 1; }
          }
        }
      }
      if (DBG) dbg[11] = d11;
    }
    __syncwarp();
  } else if (warp == K1_W_MMA) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(K1_MT, a.NT);
      const uint32_t a_lbo16 = K1_MT, b_lbo16 = (uint32_t)a.NT;
      const uint64_t a_proto = make_desc(0, a_lbo16 * 16, 128), b_proto = make_desc(0, b_lbo16 * 16, 128);
      const uint32_t a_half16 = (K1_A_STAGE / 2) >> 4, b_step16 = b_step_bytes >> 4, b_lo16 = 2u * a.NT;
      const uint32_t a0_16 = smem_u32(a_base) >> 4, b0_16 = smem_u32(b_base) >> 4;
      const uint32_t b_stage16 = b_stage_bytes >> 4;
      int bst = 0, bph = 0, it = 0, tc = 0;
      unsigned d3 = 0, d4 = 0, d5 = 0, d6 = 0;
      K1_T0(tm);
      for (int w = blockIdx.x; w < a.n_items; w += gridDim.x, ++it) {
        const int ng = w % a.n_groups;
        const int n_lo = ng * a.npg, n_hi = min(a.tiles_n, n_lo + a.npg);
        for (int n = n_lo; n < n_hi; ++n, ++tc) {
          const int set = tc % a.nsets;
          K1_ACC(d6, tm);
          mbar_wait(ACC_EMPTY(set), ((tc / a.nsets) & 1) ^ 1);
          K1_ACC(d3, tm);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(set * a.NT);
          uint32_t accum = 0;
          for (int kb = 0; kb < a.nKB; ++kb) {
            K1_ACC(d6, tm);
            const int g = it * a.nKB + kb, st = g % NA;
            if (n == n_lo) { mbar_wait(A_FULL(st), (g / NA) & 1); }
            K1_ACC(d4, tm);
            mbar_wait(B_FULL(bst), bph);
            K1_ACC(d5, tm);
            tc_fence_after();
            const uint32_t a_st16 = a0_16 + (uint32_t)st * (K1_A_STAGE >> 4);
            const uint32_t b_st16 = b0_16 + (uint32_t)bst * b_stage16;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const uint64_t dbh = desc_add(b_proto, b_st16 + (uint32_t)s * b_step16);
              const uint64_t dbl = desc_add(dbh, b_lo16);
              const uint64_t dah = desc_add(a_proto, a_st16 + (uint32_t)(2 * s) * a_lbo16);
              const uint64_t dal = desc_add(dah, a_half16);
              umma_f16(d, dah, dbh, idesc, accum);
              if (a.split & 1) umma_f16(d, dal, dbh, idesc, 1u);
              if (a.split & 2) umma_f16(d, dah, dbl, idesc, 1u);
              accum = 1u;
            }
            umma_commit(B_EMPTY(bst));
            if (++bst == a.NB) { bst = 0; bph ^= 1; }
            if (n == n_hi - 1) umma_commit(A_EMPTY(st));       // this K-block's stage may be overwritten
          }
          umma_commit(ACC_FULL(set));
        }
      }
      if (DBG) { dbg[3] = d3; dbg[4] = d4; dbg[5] = d5; dbg[6] = d6; }
    }
    __syncwarp();
  } else if (warp == K1_W_ALOAD) {
    // =========================== input loader ===========================
    // one TMA box per K-block straight into that K-block's operand stage as soon as the MMAs of the previous tile
    // released it: up to nKB x 16 KB of input in flight per SM, no registers involved
    if (elect_one()) {
      const uint32_t a0 = smem_u32(a_base);
      int g = 0;
      for (int w = blockIdx.x; w < a.n_items; w += gridDim.x) {
        const int row0 = (w / a.n_groups) * K1_MT;
        for (int kb = 0; kb < a.nKB; ++kb, ++g) {
          const int st = g % NA;
          mbar_wait(A_EMPTY(st), ((g / NA) & 1) ^ 1);
          if (DBG && (a.dbgf & 2)) { mbar_arrive(RAW_FULL(st)); continue; }
          mbar_arrive_expect_tx(RAW_FULL(st), K1_A_STAGE);
          const int c0 = kb * K1_KB;
          if (c0 < a.C0) tma_load_box(a0 + (uint32_t)st * K1_A_STAGE, &map0, c0, row0, RAW_FULL(st));
          else tma_load_box(a0 + (uint32_t)st * K1_A_STAGE, &map1, c0 - a.C0, row0, RAW_FULL(st));
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    // Row-per-lane drain: lane owns position row0 + lane and writes its 32 (16) columns of each block as 256-bit
    // stores (whole 32-byte sectors; no shared-memory transpose).  The residual of the block is requested before
    // the TMEM load is waited for; the bias comes from shared memory.
    const int ew = warp - K1_W_EPI;
    const int lq = warp & 3;                 // TMEM lane quadrant this warp may read
    const int hsel = ew >> 2;                // parity of the 32-column blocks this warp drains
    const int nblk = (a.NT + 31) / 32;
    const float* __restrict__ resp = a.res;
    float* __restrict__ dstp = a.dst;
    int tc = 0;
    unsigned d7 = 0, d8 = 0, d10 = 0;
    K1_T0(te);
    for (int w = blockIdx.x; w < a.n_items; w += gridDim.x) {
      const int ng = w % a.n_groups;
      const int n_lo = ng * a.npg, n_hi = min(a.tiles_n, n_lo + a.npg);
      const long long p = (long long)(w / a.n_groups) * K1_MT + lq * 32 + lane;
      const bool ok = p < a.Qtot;
      for (int n = n_lo; n < n_hi; ++n, ++tc) {
        const int set = tc % a.nsets;
        const int n0 = n * a.NT;
        const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(set * a.NT);
        float rr[32];
        auto res_fetch = [&](int blk) {
          if (!resp || !ok) return;
          const int cb = blk * 32;
          const float* rp = resp + p * a.Cout + n0 + cb;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8)
            if (cb + c8 * 8 < a.NT) ldg256(rp + c8 * 8, rr + c8 * 8);
        };
        if (resp && ok) {                                    // later blocks' residual: into L2 while the MMAs run
          const float* rp = resp + p * a.Cout + n0;
          for (int blk = hsel + 2; blk < nblk; blk += 2) prefetch_l2(rp + blk * 32);
        }
        if (hsel < nblk) res_fetch(hsel);
        K1_ACC(d10, te);
        mbar_wait(ACC_FULL(set), (tc / a.nsets) & 1);
        K1_ACC(d7, te);
        tc_fence_after();
        for (int blk = hsel; blk < nblk; blk += 2) {
          const int cb = blk * 32;
          const int wd = min(32, a.NT - cb);                // 32 or 16 columns
          uint32_t r[32];
          tmem_ld16(trow + (uint32_t)cb, r);
          if (wd == 32) tmem_ld16(trow + (uint32_t)(cb + 16), r + 16);
          K1_ACC(d10, te);
          tmem_ld_wait();
          K1_ACC(d8, te);
          float* op = dstp + p * a.Cout + n0 + cb;
          const float* bs = bias_s + n0 + cb;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            if (c8 * 8 < wd) {
              const float4 b0 = *reinterpret_cast<const float4*>(bs + c8 * 8), b1 = *reinterpret_cast<const float4*>(bs + c8 * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                v[e] = __uint_as_float(r[c8 * 8 + e]) * a.wscale + bb[e];      // wscale is a power of two: exact product
                if (resp) v[e] += rr[c8 * 8 + e];
                v[e] *= a.oscale;
                if (a.act_out) v[e] = silu_f(v[e]);
              }
              if (ok && !(DBG && (a.dbgf & 1))) stg256(op + c8 * 8, v);
            }
          }
          if (blk + 2 < nblk) res_fetch(blk + 2);
        }
        K1_ACC(d10, te);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ACC_EMPTY(set));        // this warp's share of the accumulator set is drained
      }
    }
    if (DBG && ew == 0 && lane == 0) { dbg[7] = d7; dbg[8] = d8; dbg[9] = 0; dbg[10] = d10; }
  }

  __syncthreads();
  if (DBG && tid == 0) dbg[0] = (unsigned)clock() - t_begin;
  if (warp == K1_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

}  // namespace

namespace {
typedef CUresult (*K1EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
K1EncodeFn k1_encode_fn() {
  static K1EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<K1EncodeFn>(p);
    return (K1EncodeFn) nullptr;
  }();
  return fn;
}
// NHWC activation [rows][C] fp32, box = 128 rows x 32 channels, 128-byte swizzle, zero fill past the end
bool k1_make_map(CUtensorMap* m, const void* base, long long rows, int C) {
  K1EncodeFn enc = k1_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {K1_KB, K1_MT};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// Launches the input-stationary kernel when the op qualifies; returns false (nothing launched) otherwise.
bool try_launch_conv1x1(const McvdOp& op, cudaStream_t s, int& rc) {
  static const int enabled = [] { const char* e = getenv("MCVD_CONV1X1"); return (e && e[0] == '0') ? 0 : 1; }();
  rc = 0;
  if (!enabled || op.i0 != 1 || op.src2 || op.src3 || op.dst2) return false;
  if (op.C0 % K1_KB || op.C1 % K1_KB || (op.C1 && !op.src1)) return false;
  const int Cin = op.C0 + op.C1, nKB = Cin / K1_KB;
  const int NT = op.i1;
  if (nKB < 1 || nKB > K1_MAX_KB || NT < 16 || NT > 256 || NT % 16 || op.Cout % NT) return false;
  if (op.aux1 && op.H * op.W < 64) return false;        // norm-table staging covers <= 3 images per tile
  const size_t b_stage = (size_t)128 * NT;
  const size_t tab_bytes = op.aux1 ? (size_t)2 * K1_TAB_IMGS * 3 * Cin * 4 : 0;
  const size_t limit = 227 * 1024;
  // the whole tile resident (every n-tile reuses the converted input) when it fits next to >= 2 weight stages,
  // else a ring of >= 4 operand stages and one n-tile per work item
  const size_t other = (size_t)op.Cout * 4 + tab_bytes + 1024 + 1024;     // bias, table, barriers, base alignment
  int NA = nKB;
  while (NA > 0 && (size_t)NA * K1_A_STAGE + other + 2 * b_stage > limit) --NA;
  if (NA < nKB && NA > K1_RING) NA = K1_RING;
  if (NA < 4 && NA < nKB) return false;
  const size_t fixed = (size_t)NA * K1_A_STAGE + other;
  K1Args a;
  a.s0 = (const float*)op.src0; a.s1 = (const float*)op.src1; a.C0 = op.C0; a.C1 = op.C1;
  a.wpk = (const __half*)op.w; a.bias = (const float*)op.bias; a.res = (const float*)op.aux0;
  a.tab = (const float4*)op.aux1; a.dst = (float*)op.dst;
  a.Pimg = op.H * op.W; a.B = op.B; a.Cout = op.Cout;
  a.Qtot = (long long)op.B * a.Pimg;
  a.NT = NT; a.tiles_n = op.Cout / NT; a.nKB = nKB; a.NA = NA;
  int NB = (int)((limit - fixed) / b_stage);
  a.NB = NB > 8 ? 8 : NB;
  a.nsets = (2 * NT <= 512) ? 2 : 1;
  int p2 = 32;
  while (p2 < a.nsets * NT) p2 <<= 1;
  a.tmem_cols = p2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long m_tiles = (a.Qtot + K1_MT - 1) / K1_MT;
  if (m_tiles > (1 << 23) || ((uintptr_t)op.src0 & 15) || ((uintptr_t)op.src1 & 15)) return false;
  // few m-tiles: spread the n-tiles of one m-tile over several CTAs (each converts its own copy of the input tile)
  int groups = 1;
  if (m_tiles < sms) groups = (int)(sms / m_tiles);
  if (groups > a.tiles_n || NA < nKB) groups = a.tiles_n;      // ring: one n-tile per item
  a.npg = (a.tiles_n + groups - 1) / groups;
  a.n_groups = (a.tiles_n + a.npg - 1) / a.npg;
  a.n_items = (int)m_tiles * a.n_groups;
  a.act_in = (op.flags & MCVD_F_ACT_IN) ? 1 : 0;
  a.act_out = (op.flags & MCVD_F_ACT_OUT) ? 1 : 0;
  a.wscale = op.f1; a.oscale = op.f0;
  a.split = (op.i3 >= 1 && op.i3 <= 3) ? op.i3 : (op.i3 == 4 ? 0 : 3);
  const size_t smem = fixed + (size_t)a.NB * b_stage;
  a.dbg = (long long*)op.aux2;
  { const char* e2 = getenv("MCVD_K1_DBGF"); a.dbgf = e2 ? atoi(e2) : 0; }
  auto kern = a.dbg ? k_conv1x1_umma<true> : k_conv1x1_umma<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
  if (e != cudaSuccess) { set_error("CONV1X1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); rc = -1; return true; }
  const int grid = a.n_items < sms ? a.n_items : sms;
  CUtensorMap map0, map1;
  if (!k1_make_map(&map0, op.src0, a.Qtot, op.C0) || !k1_make_map(&map1, op.C1 ? op.src1 : op.src0, a.Qtot, op.C1 ? op.C1 : op.C0)) {
    set_error("CONV1X1: cuTensorMapEncodeTiled failed for %lld x %d (+%d) fp32", a.Qtot, op.C0, op.C1);
    rc = -3;
    return true;
  }
  kern<<<grid, K1_THREADS, smem, s>>>(a, map0, map1);
  if (cudaGetLastError() != cudaSuccess) { set_error("CONV1X1: launch failed"); rc = -2; }
  return true;
}

}  // namespace mcvd
