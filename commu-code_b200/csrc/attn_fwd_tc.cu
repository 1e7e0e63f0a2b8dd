// Fused relative-position attention forward on tcgen05 tensor cores (v2).
//
// One CTA = 128 query rows of one (batch, head); key tiles of 128.  Per key tile t three MMAs run on
// the 5th-gen tensor cores with accumulators in TMEM:
//     S(t)    = (q+u) K_t^T                 [128 x 128]   A,B from 128B-swizzled smem (TMA-staged)
//     BD(b)   = (q+v) R_b^T                 [128 x 128]   one new 128-distance block of R per tile
//     Opart   = P(t) V_t                    [128 x 64]    A = P in TMEM (bf16), B = V tile (MN-major)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator (+ P~ store), warp 3 = P~ store, warps 4-19 = four softmax warpgroups (thread = one query row x 32 key columns,
// tcgen05.ld 32x32b; the threads of a row exchange their partial max through shared memory).
//
// Relative shift: the distances a (query tile, key tile) pair needs are the 255-row band
// delta = dlo + c, c = li + 127 - lj.  It is covered by two 128-row blocks: "lo" (new for this key
// tile) and "hi" (the previous tile's "lo").  A softmax thread copies ITS OWN row of each BD block
// from TMEM to an fp16 row in shared memory and re-reads it at c = li + 127 - lj (a row is shared only
// by the two threads that own it).  Which block a 32-key chunk needs is warp-uniform except on the
// diagonal chunk, so the re-read is `base - 2*lj` with immediate offsets.  T x K never exists in HBM.
//
// STORE (training): the probabilities are kept for the backward instead of being recomputed there.  Per key tile
// the softmax threads also write P~ = bf16(exp2(s - m_tile)) - the very operand of the PV product, before the dropout
// mask, with the SIGN bit set on dropped entries - into a 32 KB staging tile that warp 3 stores by TMA to
// p_save [B*H, Tpad, Kp], and the running maximum m_tile the tile was taken at to mt_save [B*H, Kp/128, Tpad]; the
// backward rebuilds P = P~ * exp2(m_tile - LSE) (attn_bwd_mat.cu).  2 bytes per score element, 2.1 GB per layer at
// the benchmark shape, of the 180 GB.
//
// Replaces commu/model/model.py:312-345 (AC, BD, _rel_shift, mask, softmax, AV).
#include <cuda_fp16.h>
#include "api_common.h"
#include "attn_common.cuh"
#include "attn_tc_common.cuh"

namespace cb_host {
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);
int check_attn_common(const attn::Params& p, const char* who);
}

namespace {
using attn::Params;
using attn::key_lo;
using namespace attn_tc;

constexpr int TM = 128;   // query rows per CTA
constexpr int TN = 128;   // keys per tile
constexpr int DH = 64;
constexpr int NWG = 4;                    // softmax warpgroups (column quarters of a row)
constexpr int SOFT = 128 * NWG;           // softmax threads
constexpr int NTHREADS = 128 + SOFT;      // 4 control warps + NWG softmax warpgroups
constexpr int CPT = TN / NWG;             // key columns per softmax thread
constexpr int OPT = DH / NWG;             // output dims per softmax thread
constexpr int STAGE_ROW = 592;            // bytes per staged fp16 BD row: [half 0 | half 1 | first 64 B of half 0 again] + 16 pad
constexpr int TILE_BYTES = TN * DH * 2;   // 16 KB
// TMEM columns
constexpr int COL_S = 0;      // 2 x 128
constexpr int COL_BD = 256;   // 128
constexpr int COL_O = 384;    // 64
constexpr int COL_P = 448;    // 64 (bf16 pairs)

struct Smem {
  uint8_t qu[TILE_BYTES];
  uint8_t qv[TILE_BYTES];
  uint8_t k[2][TILE_BYTES];
  uint8_t v[TILE_BYTES];          // single buffer: V(t) is only needed at the end of tile t's softmax
  uint8_t r[2][TILE_BYTES];
  uint8_t bd[TM * STAGE_ROW];
  uint8_t pst[2 * TILE_BYTES];    // STORE: P~ tile [key half][128 q rows][128 B] (128B swizzle), source of the TMA store
  float xch[4][TM];   // per-row partial max exchanged between the softmax warpgroups
  float xsum[4][TM];  // per-row partial sums (end of kernel)
  uint64_t q_ready;
  uint64_t k_full[2], k_empty[2], v_full, v_empty, r_full[2], r_empty[2];
  uint64_t s_full[2], s_empty[2], bd_full, bd_empty, p_full, o_full, pst_full[4], pst_free[4];
  uint32_t tmem_base;
};

struct StoreArgs {
  float* mt;        // [B*H, nkt, Tpad]
  int Tpad, nkt;
};

template <bool DROP, bool STORE>
__global__ void __launch_bounds__(NTHREADS, 1)
relattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_ps,
                      const Params p, const StoreArgs sa) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * TM;
  const bool reset = p.reset && p.reset[b];
  const int i_last = min(p.T - 1, i0 + TM - 1);
  const int jt_first = key_lo(i0, p.M, p.same_length, p.shift, reset) / TN;
  const int jt_last = (i_last + p.M) / TN;
  const int nt = jt_last - jt_first + 1;
  // distance of band column c for tile t: delta = i0 + M - (j0 + 127) + c ; block beta covers
  // R rows [dbase - 128*beta, +128) with dbase = "hi" block of the first tile
  const int dbase = i0 + p.M - (jt_first * TN + TN - 1) + TN;

  if (threadIdx.x == 0) {
    cb::mbar_init(&sm.q_ready, SOFT);
    for (int s = 0; s < 2; ++s) {
      cb::mbar_init(&sm.k_full[s], 1); cb::mbar_init(&sm.k_empty[s], 1);
      cb::mbar_init(&sm.r_full[s], 1); cb::mbar_init(&sm.r_empty[s], 1);
      cb::mbar_init(&sm.s_full[s], 1); cb::mbar_init(&sm.s_empty[s], SOFT);
    }
    cb::mbar_init(&sm.v_full, 1); cb::mbar_init(&sm.v_empty, 1);
    for (int s = 0; s < 4; ++s) { cb::mbar_init(&sm.pst_full[s], NWG * 32); cb::mbar_init(&sm.pst_free[s], 1); }
    cb::mbar_init(&sm.bd_full, 1); cb::mbar_init(&sm.bd_empty, SOFT);
    cb::mbar_init(&sm.p_full, SOFT);
    cb::mbar_init(&sm.o_full, 1);
    cb::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    cb::tma_prefetch_desc(&tm_k);
    cb::tma_prefetch_desc(&tm_v);
    cb::tma_prefetch_desc(&tm_r);
  }
  if (warp == 2) {
    cb::tmem_alloc(&sm.tmem_base, 512);
    cb::tmem_relinquish();
  }
  cb::tc_fence_before();
  __syncthreads();
  cb::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (cb::elect_one()) {
      Ring rk, rr;
      auto load_r = [&](int beta) {
        cb::mbar_wait(&sm.r_empty[rr.idx], rr.phase ^ 1);
        cb::mbar_arrive_expect_tx(&sm.r_full[rr.idx], TILE_BYTES);
        cb::tma_load_2d(sm.r[rr.idx], &tm_r, &sm.r_full[rr.idx], h * DH, dbase - TN * beta);
        rr.advance();
      };
      auto load_k = [&](int t) {
        cb::mbar_wait(&sm.k_empty[rk.idx], rk.phase ^ 1);
        cb::mbar_arrive_expect_tx(&sm.k_full[rk.idx], TILE_BYTES);
        cb::tma_load_3d(sm.k[rk.idx], &tm_k, &sm.k_full[rk.idx], h * DH, b, (jt_first + t) * TN);
        rk.advance();
      };
      load_r(0);
      load_k(0);
      load_r(1);
      for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) {          // the next tile's K and R never wait behind the single V buffer
          load_k(t + 1);
          load_r(t + 2);
        }
        cb::mbar_wait(&sm.v_empty, (t & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.v_full, TILE_BYTES);
        cb::tma_load_3d(sm.v, &tm_v, &sm.v_full, h * DH, b, (jt_first + t) * TN);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one()) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);   // S, BD: both operands K-major
      const uint32_t idesc_o = cb::umma_idesc_bf16(TM, DH, 0, 1);   // PV: A in TMEM, B = V (MN-major)
      Ring rk, rr, rs;
      uint32_t bd_phase = 0, p_phase = 0;
      const uint32_t a_qu = cb::smem_u32(sm.qu), a_qv = cb::smem_u32(sm.qv);
      cb::mbar_wait(&sm.q_ready, 0);
      cb::tc_fence_after();
      auto issue_bd = [&]() {
        cb::mbar_wait(&sm.r_full[rr.idx], rr.phase);
        cb::mbar_wait(&sm.bd_empty, bd_phase ^ 1);
        cb::tc_fence_after();
        const uint64_t ad = cb::umma_smem_desc(a_qv, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(cb::smem_u32(sm.r[rr.idx]), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          cb::umma_bf16_ss(tmem + COL_BD, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
        cb::umma_commit(&sm.r_empty[rr.idx]);
        cb::umma_commit(&sm.bd_full);
        rr.advance();
        bd_phase ^= 1;
      };
      auto issue_s = [&]() {
        cb::mbar_wait(&sm.k_full[rk.idx], rk.phase);
        cb::mbar_wait(&sm.s_empty[rs.idx], rs.phase ^ 1);
        cb::tc_fence_after();
        const uint64_t ad = cb::umma_smem_desc(a_qu, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(cb::smem_u32(sm.k[rk.idx]), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          cb::umma_bf16_ss(tmem + COL_S + rs.idx * TN, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
        cb::umma_commit(&sm.k_empty[rk.idx]);
        cb::umma_commit(&sm.s_full[rs.idx]);
        rk.advance();
        rs.advance();
      };
      issue_bd();   // beta = 0 ("hi" of the first tile)
      issue_s();    // S(0)
      issue_bd();   // beta = 1
      for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) {
          issue_s();   // S(t+1)
          issue_bd();  // beta = t+2
        }
        cb::mbar_wait(&sm.p_full, p_phase);     // (every row's O has been rescaled by then, where its maximum moved)
        cb::mbar_wait(&sm.v_full, t & 1);
        cb::tc_fence_after();
        const uint64_t vd = cb::umma_smem_desc(cb::smem_u32(sm.v), 8192, 1024);
#pragma unroll
        for (int k = 0; k < TN / 16; ++k)   // O accumulates in TMEM over the key tiles
          umma_bf16_ts(tmem + COL_O, tmem + COL_P + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc_o, (t > 0 || k > 0));
        cb::umma_commit(&sm.v_empty);
        cb::umma_commit(&sm.o_full);
        p_phase ^= 1;
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ============================== P~ store (TMA, shared -> global) ==============================
    if (STORE && cb::elect_one()) {
      // one 32-row slab per row group (TMEM lane quadrant): the row groups stay independent of each other (a CTA-wide
      // hand-over would make all 16 softmax warps move in lock-step); 4 x 2 boxes of {64 keys, 32 rows} per tile.
      // Two store warps, two row groups each: a warp waits for its slab to be read before it signals it free, so one
      // warp for all four serialised the groups behind each other.
      const int row0 = (b * p.H + h) * sa.Tpad + i0;
      const int rg0 = (warp - 2) * 2;
      for (int t = 0; t < nt; ++t) {
        const int j0 = (jt_first + t) * TN;
#pragma unroll 1
        for (int rg = rg0; rg < rg0 + 2; ++rg) {
          cb::mbar_wait(&sm.pst_full[rg], t & 1);
          cb::tma_store_2d(&tm_ps, sm.pst + rg * 4096, j0, row0 + 32 * rg);
          cb::tma_store_2d(&tm_ps, sm.pst + TILE_BYTES + rg * 4096, j0 + 64, row0 + 32 * rg);
          cb::tma_store_commit();
          cb::tma_store_wait_read<0>();       // 8 KB: the slab is free again well before the group's next exp phase
          cb::mbar_arrive(&sm.pst_free[rg]);
        }
      }
      cb::tma_store_wait<0>();
    }
  } else if (warp >= 4) {
    // ============================== softmax warpgroups ==============================
    // thread = (query row li, column quarter g): owns key columns 32g..32g+31 and out dims 16g..16g+15
    const int g = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;              // TMEM lane quadrant
    const int li = wq * 32 + lane;
    const int i = i0 + li;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    // ---- stage (q + r_w_bias) [g = 0] / (q + r_r_bias) [g = 1] rows, UMMA K-major 128B-swizzled ----
    if (g < 2) {
      const bf16* qrow = p.q + ((long long)i * p.B + b) * p.ldq + h * DH;
      const float* bias = g == 0 ? p.u : p.vb;
      uint8_t* tile = g == 0 ? sm.qu : sm.qv;
      bf16* save = g == 0 ? p.qu_s : p.qv_s;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (i < p.T) raw = *reinterpret_cast<const uint4*>(qrow + ch * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t ob[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = h * DH + ch * 8 + e * 2;
          ob[e] = cb::pack_bf16(cb::bf16_lo(w[e]) + __ldg(bias + c), cb::bf16_hi(w[e]) + __ldg(bias + c + 1));
        }
        *reinterpret_cast<uint4*>(tile + attn::swz(li, ch)) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        if (save && i < p.T)
          *reinterpret_cast<uint4*>(save + ((long long)i * p.B + b) * p.ldq + h * DH + ch * 8) =
              make_uint4(ob[0], ob[1], ob[2], ob[3]);
      }
      cb::fence_proxy_async();
    }
    cb::mbar_arrive(&sm.q_ready);
    uint32_t bd_phase = 0;
    Ring rs;
    const uint32_t my_row = cb::smem_u32(sm.bd) + li * STAGE_ROW;
    // copy this thread's 32 columns of the BD block in TMEM into half `half` of its row's fp16 staging
    auto stage_bd = [&](int half) {
      cb::mbar_wait(&sm.bd_full, bd_phase);
      cb::tc_fence_after();
      uint32_t r0[32];
      cb::tmem_ld_32x32b_x32(lane_addr + COL_BD + g * 32, r0);
      cb::tmem_ld_wait();
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.bd_empty);
      bd_phase ^= 1;
      const uint32_t dst = my_row + half * 256 + g * 64;
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) pk[e / 2] = pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1]));
#pragma unroll
      for (int c = 0; c < 4; ++c) sts_v4(dst + 16 * c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      // the band a tile reads is [lo block | hi block]: contiguous when lo sits in half 0; when lo sits in half 1 the
      // first 32 entries of the hi block (all the diagonal chunk reaches) are found in the copy behind half 1
      if (half == 0 && g == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts_v4(my_row + 512 + 16 * c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
    };
    float m_run = -INFINITY, l_run = 0.f;
    const float sl2 = p.scale * 1.4426950408889634f;
    const drop::Keys dkeys = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)((b * p.H + h) * p.T + i));
    const int hi_i = i < p.T ? i + p.M : -1;
    const int lo_i = key_lo(i, p.M, p.same_length, p.shift, reset);

    stage_bd(0);  // beta = 0
    for (int t = 0; t < nt; ++t) {
      stage_bd((t + 1) & 1);  // beta = t+1 = "lo" of this tile; "hi" = beta t sits in half t&1
      named_bar(2 + wq, NWG * 32);   // every column quarter of this row group's staged rows is visible
      // band column of key lj is idx = li + 127 - lj: idx < 128 -> "lo" block [idx], else "hi" block [idx-128]
      const uint32_t lo_base = my_row + ((t + 1) & 1) * 256 + 2 * (li + TN - 1);
      const uint32_t hi_base = my_row + (t & 1) * 256 + 2 * (li - 1);
      cb::mbar_wait(&sm.s_full[rs.idx], rs.phase);
      cb::tc_fence_after();
      const int jc0 = (jt_first + t) * TN + g * CPT;
      float s[CPT];
      {
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_S + rs.idx * TN + g * 32, r0);
        cb::tmem_ld_wait();
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.s_empty[rs.idx]);
        rs.advance();
        // this thread's 32-key chunk is chunk g; the warp's rows are 32*wq .. 32*wq+31, so
        // g < wq: every lj < li -> "hi" block; g > wq: "lo" block; g == wq: both  (warp-uniform)
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] = __uint_as_float(r0[e]);
        // packed 32-bit reads of the sheared window + FHADD (attn_tc_common.cuh); the diagonal chunk's window runs from
        // the lo block into the first entries of the hi block: contiguous in the staged row (see stage_bd)
        shear_add32(s, (g < wq ? hi_base : lo_base) - 2 * (g * 32));
      }
      if (!(jc0 + CPT - 1 <= hi_i && jc0 >= lo_i)) {   // boundary tile: analytic mask
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          const int j = jc0 + e;
          if (j > hi_i || j < lo_i) s[e] = -INFINITY;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < CPT; ++e) mx = fmaxf(mx, s[e]);
      sm.xch[g][li] = mx;
      named_bar(2 + wq, NWG * 32);
      mx = fmaxf(fmaxf(sm.xch[0][li], sm.xch[1][li]), fmaxf(sm.xch[2][li], sm.xch[3][li]));
      // LAZY running maximum: the reference point of the exponentials only moves when the row's maximum grew by more
      // than 2^8 (or when the row sees its first visible key); otherwise P~ = exp2(s - m_run) may reach 256, which bf16 /
      // fp32 hold without loss, the normaliser and the stored (P~, m_tile) pairs stay consistent, and O - which
      // accumulates in TMEM across the key tiles - needs no rescale.  The four threads of a row take the same decision.
      mx = fmaxf(mx * sl2, m_run);
      const bool raise = mx > m_run + 8.f;            // (-inf + 8 = -inf: the first finite maximum raises)
      const float m_use = raise ? mx : m_run;
      const float msafe = m_use == -INFINITY ? 0.f : m_use;
      const float corr = raise ? ex2(m_run - msafe) : 1.f;
      m_run = m_use;
      float rsum = 0.f;
      uint32_t pk[CPT / 2];
      uint32_t a_pst = 0;
      if (STORE) {
        if (t > 0) cb::mbar_wait(&sm.pst_free[wq], (t - 1) & 1);  // the previous tile's TMA store has read this row group's slab
        a_pst = cb::smem_u32(sm.pst) + (g >> 1) * TILE_BYTES + li * 128;
        if (g == 0 && i < p.T) sa.mt[((long long)(b * p.H + h) * sa.nkt + (jt_first + t)) * sa.Tpad + i] = msafe;
      }
#pragma unroll
      for (int c = 0; c < CPT / 8; ++c) {     // 8 keys = one 16-byte chunk of packed probabilities
#pragma unroll
        for (int e = c * 8; e < c * 8 + 8; e += 2) {
          const float p0 = ex2(fmaf(s[e], sl2, -msafe)), p1 = ex2(fmaf(s[e + 1], sl2, -msafe));
          rsum += p0 + p1;
          pk[e / 2] = cb::pack_bf16(p0, p1);
        }
        uint32_t ps0 = pk[4 * c], ps1 = pk[4 * c + 1], ps2 = pk[4 * c + 2], ps3 = pk[4 * c + 3];
        if (DROP) {   // dropped probabilities feed P V only; the normaliser keeps every term (model.py:336-337)
          const uint2 ra = drop::rand64((uint32_t)(jc0 >> 2) + 2 * c, dkeys);
          const uint2 rb = drop::rand64((uint32_t)(jc0 >> 2) + 2 * c + 1, dkeys);
          const uint32_t f0 = drop::keep_flags(ra.x, p.drop_thr2), f1 = drop::keep_flags(ra.y, p.drop_thr2);
          const uint32_t f2 = drop::keep_flags(rb.x, p.drop_thr2), f3 = drop::keep_flags(rb.y, p.drop_thr2);
          if (STORE) {  // sign bit = dropped (P~ >= 0, so the bit is free)
            ps0 |= ~f0 & 0x80008000u; ps1 |= ~f1 & 0x80008000u;
            ps2 |= ~f2 & 0x80008000u; ps3 |= ~f3 & 0x80008000u;
          }
          pk[4 * c] &= drop::mask16x2(f0); pk[4 * c + 1] &= drop::mask16x2(f1);
          pk[4 * c + 2] &= drop::mask16x2(f2); pk[4 * c + 3] &= drop::mask16x2(f3);
        }
        if (STORE) sts_v4(a_pst + ((((g & 1) * 4 + c) ^ (li & 7)) << 4), ps0, ps1, ps2, ps3);
      }
      if (STORE) {
        cb::fence_proxy_async();
        cb::mbar_arrive(&sm.pst_full[wq]);
      }
      l_run = l_run * corr + rsum;
      // ---- rescale O in TMEM where a row of this warp moved its reference point (rare after the first tiles) ----
      // (PV(t-1) must have completed before P(t) overwrites its operand columns - normally long done)
      if (t > 0) cb::mbar_wait(&sm.o_full, (t - 1) & 1);
      if (t > 0 && __any_sync(0xffffffffu, raise)) {
        cb::tc_fence_after();
        uint32_t r[OPT];
        tmem_ld_32x32b_x16(lane_addr + COL_O + g * OPT, r);
        cb::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < OPT; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * corr);
        tmem_st_32x32b_x16(lane_addr + COL_O + g * OPT, r);
      }
      // ---- P(t) -> TMEM (bf16 pairs; this thread's 32 keys = 16 columns) ----
      tmem_st_32x32b_x16(lane_addr + COL_P + g * (CPT / 2), pk);
      tmem_st_wait();
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.p_full);
    }
    // the accumulated O
    cb::mbar_wait(&sm.o_full, (nt - 1) & 1);
    cb::tc_fence_after();
    float o[OPT];
    {
      uint32_t r[OPT];
      tmem_ld_32x32b_x16(lane_addr + COL_O + g * OPT, r);
      cb::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < OPT; ++e) o[e] = __uint_as_float(r[e]);
    }
    cb::tc_fence_before();
    // ---- finalize: the column quarters add their partial sums ----
    sm.xsum[g][li] = l_run;
    named_bar(2 + wq, NWG * 32);
    const float l_tot = (sm.xsum[0][li] + sm.xsum[1][li]) + (sm.xsum[2][li] + sm.xsum[3][li]);
    if (i < p.T) {
      const float inv = l_tot > 0.f ? 1.f / (DROP ? l_tot * p.drop_keep : l_tot) : 0.f;
      bf16* orow = p.out + ((long long)i * p.B + b) * p.ldo + h * DH + g * OPT;
#pragma unroll
      for (int ch = 0; ch < OPT / 8; ++ch) {
        uint4 q;
        q.x = cb::pack_bf16(o[ch * 8 + 0] * inv, o[ch * 8 + 1] * inv);
        q.y = cb::pack_bf16(o[ch * 8 + 2] * inv, o[ch * 8 + 3] * inv);
        q.z = cb::pack_bf16(o[ch * 8 + 4] * inv, o[ch * 8 + 5] * inv);
        q.w = cb::pack_bf16(o[ch * 8 + 6] * inv, o[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = q;
      }
      if (p.lse && g == 0)
        p.lse[((long long)b * p.H + h) * p.T + i] = (m_run + log2f(l_tot)) * 0.6931471805599453f;
    }
  }
  cb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    cb::tc_fence_after();
    cb::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

namespace {
attn_tc::DropState g_drop = {0.f, 0ull};
}
namespace attn_tc {
DropState drop_state() { return g_drop; }
}
// Dropout on the attention probabilities (reference: self.dropatt, commu/model/model.py:211, 337) for the
// subsequent tcgen05 attention calls of this process: p in [0, 1), p == 0 switches it off.  The forward and
// the backward of one layer must run under the same (p, seed); masks are recomputed, never stored.
extern "C" int commu_relattn_set_dropout(float p, unsigned long long seed) {
  CB_REQUIRE(p >= 0.f && p < 1.f, "relattn_set_dropout: p must be in [0, 1)");
  g_drop.p = p;
  g_drop.seed = seed;
  return 0;
}

namespace attn_tc {
void psave_geometry(int T, int M, int* Tpad, int* Kp, int* nkt);   // attn_bwd_mat.cu
}

// Same contract as commu_relattn_fwd (the v1 warp-MMA kernel); this is the tcgen05 implementation.
// p_save / mt_save (both or neither; sizes from commu_relattn_bwd_sizes): when given, the forward keeps the
// un-normalised probabilities and the per-tile maxima for the materialised backward (commu_relattn_bwd).
extern "C" int commu_relattn_fwd_tc(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                    const void* r, int64_t ldr, int kr, const float* r_w_bias,
                                    const float* r_r_bias, const unsigned char* reset, int T, int M, int B,
                                    int H, int same_length, int shift, float scale, void* out, int64_t ldo,
                                    float* lse, void* qu_save, void* qv_save, void* p_save, float* mt_save,
                                    void* stream) {
  attn::Params p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.r = (const bf16*)r;
  p.qu_s = (bf16*)qu_save; p.qv_s = (bf16*)qv_save;
  p.u = r_w_bias; p.vb = r_r_bias; p.reset = reset;
  p.ldq = ldq; p.ldkv = ldkv; p.ldr = ldr;
  p.T = T; p.M = M; p.B = B; p.H = H; p.Kr = kr;
  p.same_length = same_length; p.shift = shift; p.scale = scale;
  p.out = (bf16*)out; p.ldo = ldo; p.lse = lse;
  apply_drop_state(p);
  int rc = cb_host::check_attn_common(p, "relattn_fwd_tc");
  if (rc) return rc;
  CB_REQUIRE(out && (ldo % 8 == 0), "relattn_fwd_tc: bad output");
  CB_REQUIRE((qu_save == nullptr) == (qv_save == nullptr), "relattn_fwd_tc: qu_save/qv_save must both be set or null");
  CB_REQUIRE((p_save == nullptr) == (mt_save == nullptr), "relattn_fwd_tc: p_save/mt_save must both be set or null");
  CB_REQUIRE(!p_save || lse, "relattn_fwd_tc: storing the probabilities needs the LSE output");
  const int Ktot = T + M;
  CUtensorMap tk, tv, tr, tps;
  // K and V live in the same [K*B, ldkv] matrix; each map starts at its own base pointer.  The column
  // extent is one head-row group of H*64 columns reachable from that base.
  rc = make_tmap_rows3d(&tk, k, (uint64_t)H * 64, B, Ktot, ldkv);
  if (rc) return rc;
  rc = make_tmap_rows3d(&tv, v, (uint64_t)H * 64, B, Ktot, ldkv);
  if (rc) return rc;
  rc = cb_host::make_tmap_bf16_2d(&tr, r, (uint64_t)H * 64, kr, ldr, 64, 128);
  if (rc) return rc;
  StoreArgs sa = {};
  tps = tr;   // placeholder when nothing is stored (never dereferenced)
  if (p_save) {
    int Tpad, Kp, nkt;
    psave_geometry(T, M, &Tpad, &Kp, &nkt);
    sa.mt = mt_save; sa.Tpad = Tpad; sa.nkt = nkt;
    rc = cb_host::make_tmap_bf16_2d(&tps, p_save, (uint64_t)Kp, (uint64_t)B * H * Tpad, Kp, 64, 32);   // one row group per box
    if (rc) return rc;
  }
  static bool attr = false;
  const int smem_bytes = (int)sizeof(Smem) + 1024;
  if (!attr) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(T, TM), H, B);
  cudaStream_t st = (cudaStream_t)stream;
  cb_host::ProfScope prof(cb_host::PROF_ATTN_FWD, st);
  if (p.drop_thr2) {
    if (p_save) relattn_fwd_tc_kernel<true, true><<<grid, NTHREADS, smem_bytes, st>>>(tk, tv, tr, tps, p, sa);
    else relattn_fwd_tc_kernel<true, false><<<grid, NTHREADS, smem_bytes, st>>>(tk, tv, tr, tps, p, sa);
  } else {
    if (p_save) relattn_fwd_tc_kernel<false, true><<<grid, NTHREADS, smem_bytes, st>>>(tk, tv, tr, tps, p, sa);
    else relattn_fwd_tc_kernel<false, false><<<grid, NTHREADS, smem_bytes, st>>>(tk, tv, tr, tps, p, sa);
  }
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
