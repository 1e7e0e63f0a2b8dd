// Relative-position attention backward on tcgen05 tensor cores: the dq pass (+ the column sums for d r_w_bias).
//
// One CTA = 128 query rows of one (batch, head); it walks the key tiles like the forward kernel.  Per key
// tile t six products run on the tensor cores (accumulators in TMEM):
//     S      = (q+u) K_t^T                           [128 x 128]
//     BD     = (q+v) R_blk^T   "lo" block (beta = t+1) and "hi" block (beta = t) of the 255-distance band
//     dP     = dO V_t^T                              [128 x 128]   (issued into the "hi" columns once they are staged)
//     dq    += dS K_t                                [128 x 64]    A = dS (bf16) in TMEM, B = K tile (MN-major)
//     dq    += C_beta R_beta                         [128 x 64]    A = band block beta in shared memory
// Band blocks: dBD is the inverse relative shift of dS - row li holds dS[li, lj] at band column
// c = li + 127 - lj of the 255-column band.  Columns c < 128 belong to distance block beta = t+1, the others to
// beta = t, and every element of block beta is produced exactly once: by tile beta-1 if idx >= li, by tile
// beta if idx < li.  The softmax threads therefore scatter dS straight into two 128 x 128 block buffers
// (ring of 2) and ONE K=128 product per tile consumes the block that just became complete.  A block is kept
// MN-major without swizzle ([16 groups of 8 rows][128 idx][8 rows x 2 B]), which makes the scatter address
// linear in idx: `base - 16*e` with immediate offsets, one 16-bit store per element.
// The software pipeline: S(t+1) / lo(t+1) / hi(t+1) are issued while the softmax threads still work on tile
// t (as soon as they have copied S / staged lo / loaded dP of tile t); the staged fp16 rows of the relative
// shift (attn_tc_common.cuh) alias the row group's piece of the block buffer that tile t starts to fill.
// dq = scale * acc.  d r_w_bias + d r_r_bias = colsum(dq) is added to du here; the dR pass computes
// d r_r_bias from the column sums of dBD and moves it from du to dvb (attn_bwd_dr_tc.cu).
//
// Autograd counterpart of commu/model/model.py:312-345 for d(queries) and the two global biases.
#include "api_common.h"
#include "attn_common.cuh"
#include "attn_tc_common.cuh"

namespace cb_host {
int check_attn_common(const attn::Params& p, const char* who);
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);
}

namespace {
using attn::Params;
using attn::key_lo;
using namespace attn_tc;

constexpr int TM = 128, TN = 128, DH = 64;
constexpr int NWG = 4;
constexpr int SOFT = 128 * NWG;
constexpr int NTHREADS = 64 + SOFT;        // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer, warps 2-17: softmax
constexpr int TILE_BYTES = 128 * DH * 2;      // 16 KB
constexpr int QG = 2576;                      // bytes between the 8-row groups of a band block (2048 used); 4 * QG >= kStageGroupBytes
constexpr int CB_BYTES = 16 * QG;             // one band block buffer
constexpr int BAND_THREADS = 32 * (NWG * (NWG + 1) / 2);   // threads (li, g) with g >= wq (or g <= wq): 320
constexpr int COL_S = 0, COL_X = 128, COL_LO = 256, COL_DS = 384, COL_DQ = 448;
static_assert(4 * QG >= kStageGroupBytes, "a row group's piece of a band block must hold its staged rows");

struct Smem {
  uint8_t qu[TILE_BYTES];
  uint8_t qv[TILE_BYTES];
  uint8_t dout[TILE_BYTES];
  uint8_t k[2][TILE_BYTES];
  uint8_t v[TILE_BYTES];
  uint8_t r[3][TILE_BYTES];
  uint8_t cb[2][CB_BYTES];      // band blocks beta (buffer beta & 1); ALSO the staged fp16 rows of tile t in buffer (t+1) & 1
  uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty, r_full[3], r_empty[3];
  uint64_t s_full, s_free, lo_full, lo_free, hi_full, hi_done, dp_full, x_free, ds_full, ds_free, cb_free[2], acc_full;
  uint32_t tmem_base;
};

template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1)
relattn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                         const __grid_constant__ CUtensorMap tm_qu, const __grid_constant__ CUtensorMap tm_qv,
                         const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_r,
                         const Params p) {
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
  // block beta covers R rows [dbase - 128*beta, +128); tile t: "hi" = beta t, "lo" = beta t+1
  const int dbase = i0 + p.M - (jt_first * TN + TN - 1) + TN;

  if (threadIdx.x == 0) {
    cb::mbar_init(&sm.q_full, 1);
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.k_full[s], 1); cb::mbar_init(&sm.k_empty[s], 1); }
    cb::mbar_init(&sm.v_full, 1); cb::mbar_init(&sm.v_empty, 1);
    for (int s = 0; s < 3; ++s) { cb::mbar_init(&sm.r_full[s], 1); cb::mbar_init(&sm.r_empty[s], 1); }
    cb::mbar_init(&sm.s_full, 1); cb::mbar_init(&sm.s_free, SOFT);
    cb::mbar_init(&sm.lo_full, 1); cb::mbar_init(&sm.lo_free, BAND_THREADS);
    cb::mbar_init(&sm.hi_full, 1); cb::mbar_init(&sm.hi_done, BAND_THREADS);
    cb::mbar_init(&sm.dp_full, 1); cb::mbar_init(&sm.x_free, SOFT);
    cb::mbar_init(&sm.ds_full, SOFT); cb::mbar_init(&sm.ds_free, 1);
    cb::mbar_init(&sm.cb_free[0], 1); cb::mbar_init(&sm.cb_free[1], 1);
    cb::mbar_init(&sm.acc_full, 1);
    cb::fence_barrier_init();
  }
  if (warp == 0) {
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
      cb::mbar_arrive_expect_tx(&sm.q_full, 3 * TILE_BYTES);
      cb::tma_load_3d(sm.qu, &tm_qu, &sm.q_full, h * DH, b, i0);
      cb::tma_load_3d(sm.qv, &tm_qv, &sm.q_full, h * DH, b, i0);
      cb::tma_load_3d(sm.dout, &tm_do, &sm.q_full, h * DH, b, i0);
      auto load_r = [&](int beta) {   // slot beta % 3, its (beta / 3)-th use
        const int sl = beta % 3;
        cb::mbar_wait(&sm.r_empty[sl], ((beta / 3) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.r_full[sl], TILE_BYTES);
        cb::tma_load_2d(sm.r[sl], &tm_r, &sm.r_full[sl], h * DH, dbase - TN * beta);
      };
      auto load_k = [&](int t) {      // buffer t & 1, its (t >> 1)-th use
        const int bi = t & 1;
        cb::mbar_wait(&sm.k_empty[bi], ((t >> 1) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.k_full[bi], TILE_BYTES);
        cb::tma_load_3d(sm.k[bi], &tm_k, &sm.k_full[bi], h * DH, b, (jt_first + t) * TN);
      };
      auto load_v = [&](int t) {
        cb::mbar_wait(&sm.v_empty, (t & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.v_full, TILE_BYTES);
        cb::tma_load_3d(sm.v, &tm_v, &sm.v_full, h * DH, b, (jt_first + t) * TN);
      };
      load_k(0);
      load_r(1);
      load_r(0);
      load_v(0);
      for (int t = 0; t + 1 < nt; ++t) {
        load_k(t + 1);
        load_r(t + 2);
        load_v(t + 1);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one()) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);   // S, BD, dP: K-major x K-major
      const uint32_t idesc_a = cb::umma_idesc_bf16(TM, DH, 0, 1);   // dq += dS K : A K-major (TMEM), B MN-major
      const uint32_t idesc_b = cb::umma_idesc_bf16(TM, DH, 1, 1);   // dq += C R  : A MN-major (no swizzle), B MN-major
      const uint32_t a_qu = cb::smem_u32(sm.qu), a_qv = cb::smem_u32(sm.qv), a_do = cb::smem_u32(sm.dout);
      const uint32_t a_v = cb::smem_u32(sm.v);
      auto kmajor_128 = [&](uint32_t col, uint32_t a_addr, uint32_t b_addr) {
        const uint64_t ad = cb::umma_smem_desc(a_addr, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(b_addr, 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + col, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
      };
      // dq += C_beta R_beta: block buffer beta & 1, R slot beta % 3
      auto issue_band = [&](int beta, bool first) {
        const uint32_t a_cb = cb::smem_u32(sm.cb[beta & 1]);
        const uint64_t br = cb::umma_smem_desc(cb::smem_u32(sm.r[beta % 3]), 8192, 1024);
#pragma unroll
        for (int k = 0; k < TN / 16; ++k)
          cb::umma_bf16_ss(tmem + COL_DQ, umma_smem_desc_nosw(a_cb + k * 256, 128, QG), br + (uint64_t)(k * 128), idesc_b,
                           !(first && k == 0));
      };
      cb::mbar_wait(&sm.q_full, 0);
      // front of tile 0
      cb::mbar_wait(&sm.k_full[0], 0);
      cb::tc_fence_after();
      kmajor_128(COL_S, a_qu, cb::smem_u32(sm.k[0]));
      cb::umma_commit(&sm.s_full);
      cb::mbar_wait(&sm.r_full[1], 0);
      kmajor_128(COL_LO, a_qv, cb::smem_u32(sm.r[1]));
      cb::umma_commit(&sm.lo_full);
      cb::mbar_wait(&sm.r_full[0], 0);
      kmajor_128(COL_X, a_qv, cb::smem_u32(sm.r[0]));
      cb::umma_commit(&sm.hi_full);
      for (int t = 0; t < nt; ++t) {
        const uint32_t ph = t & 1;
        // ---- dP(t) into the "hi" columns once every thread that needs them has staged them ----
        cb::mbar_wait(&sm.hi_done, ph);
        cb::mbar_wait(&sm.v_full, ph);
        cb::tc_fence_after();
        kmajor_128(COL_X, a_do, a_v);
        cb::umma_commit(&sm.dp_full);
        cb::umma_commit(&sm.v_empty);
        // ---- front of tile t+1 ----
        if (t + 1 < nt) {
          const int kb = (t + 1) & 1;
          cb::mbar_wait(&sm.k_full[kb], ((t + 1) >> 1) & 1);
          cb::mbar_wait(&sm.s_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_S, a_qu, cb::smem_u32(sm.k[kb]));
          cb::umma_commit(&sm.s_full);
          cb::mbar_wait(&sm.r_full[(t + 2) % 3], ((t + 2) / 3) & 1);
          cb::mbar_wait(&sm.lo_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_LO, a_qv, cb::smem_u32(sm.r[(t + 2) % 3]));
          cb::umma_commit(&sm.lo_full);
          cb::mbar_wait(&sm.x_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_X, a_qv, cb::smem_u32(sm.r[(t + 1) % 3]));
          cb::umma_commit(&sm.hi_full);
        }
        // ---- back of tile t: block beta = t is complete, dS(t) sits in TMEM ----
        cb::mbar_wait(&sm.ds_full, ph);
        cb::tc_fence_after();
        issue_band(t, t == 0);
        cb::umma_commit(&sm.cb_free[t & 1]);
        cb::umma_commit(&sm.r_empty[t % 3]);
        {
          const uint64_t bk = cb::umma_smem_desc(cb::smem_u32(sm.k[t & 1]), 8192, 1024);
#pragma unroll
          for (int k = 0; k < TN / 16; ++k)
            umma_bf16_ts(tmem + COL_DQ, tmem + COL_DS + 8 * k, bk + (uint64_t)(k * 128), idesc_a, 1);
        }
        cb::umma_commit(&sm.ds_free);
        cb::umma_commit(&sm.k_empty[t & 1]);
      }
      // tail: block beta = nt holds the "lo" part of the last tile (its "hi" part was zeroed)
      issue_band(nt, false);
      cb::umma_commit(&sm.acc_full);
    }
  } else if (warp >= 2) {
    // ============================== softmax warpgroups ==============================
    // thread = (query row li, 32-key chunk g); the four warps of a row group (same wq) share one scheduler
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;                     // TMEM lane quadrant of this warp (hardware: warp id % 4)
    const int li = wq * 32 + lane;
    const int i = i0 + li;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t cb0 = cb::smem_u32(sm.cb[0]);
    const uint32_t piece = wq * 4 * QG;                               // this row group's piece of a block buffer
    const uint32_t stg = piece + stage_row_off(lane) - 64 * wq;       // staged row: byte offset of position 0
    const uint32_t rowoff = (li >> 3) * QG + (li & 7) * 2;            // (row li, idx 0) inside a block buffer
    const int c0 = li + (TN - 1) - g * 32;                             // band column of this thread's first key
    const float sl2 = p.scale * 1.4426950408889634f;
    // under dropout P is scaled by 1 / keep (folded into the exponent) and Delta by keep: dS = (P / keep) * (m * dP - keep * Delta)
    const float lse2 = (i < p.T ? p.lse[((long long)b * p.H + h) * p.T + i] * 1.4426950408889634f : 0.f) +
                       (DROP ? log2f(p.drop_keep) : 0.f);
    const float delta = (i < p.T ? p.delta[((long long)b * p.H + h) * p.T + i] : 0.f) * (DROP ? p.drop_keep : 1.f);
    const drop::Keys dkeys = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)((b * p.H + h) * p.T + i));
    const int hi_i = i < p.T ? i + p.M : -1;
    const int lo_i = key_lo(i, p.M, p.same_length, p.shift, reset);
    // block 0 only ever receives its "hi" part (from tile 0): its "lo" part starts as zeros
    for (int x = (g * 32 + lane) * 16; x < 4 * QG; x += 128 * 16) sts_v4(cb0 + piece + x, 0, 0, 0, 0);

    for (int t = 0; t < nt; ++t) {
      const uint32_t ph = t & 1;
      const uint32_t buf_lo = cb0 + ((t + 1) & 1) * CB_BYTES;   // block t+1: staged rows now, "lo" scatter later
      const uint32_t buf_hi = cb0 + (t & 1) * CB_BYTES;         // block t: "hi" scatter
      const uint32_t row_v = buf_lo + stg;
      // ---- relative shift: copy the band columns this row needs (fp16 in registers), stage them once the buffer
      // of block t+1 is free (its previous occupant, block t-1, must have been consumed); read back sheared below.
      uint32_t sg[16], sg2[16];
      if (g >= wq) {
        cb::mbar_wait(&sm.lo_full, ph);
        cb::tc_fence_after();
        load_pack32(lane_addr + COL_LO + g * 32, sg);
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.lo_free);
      }
      if (g <= wq) {                            // the diagonal chunk (g == wq) needs both blocks
        cb::mbar_wait(&sm.hi_full, ph);
        cb::tc_fence_after();
        if (g < wq) load_pack32(lane_addr + COL_X + g * 32, sg);
        else load_pack32(lane_addr + COL_X + g * 32, sg2);
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.hi_done);
      }
      if (t > 0) cb::mbar_wait(&sm.cb_free[(t + 1) & 1], ((t - 1) >> 1) & 1);
      store_packed32(g >= wq ? row_v + 64 * g : row_v + 256 + 64 * g, sg);
      if (g == wq) store_packed32(row_v + 256 + 64 * g, sg2);
      named_bar(2 + wq, NWG * 32);              // the positions of this row group are staged
      // ---- P = exp2(score*log2e - LSE), dS = P * (dP - Delta), in two halves of 16 key columns (register peak) ----
      const int jc0 = (jt_first + t) * TN + g * 32;
      const bool full = __all_sync(0xffffffffu, (jc0 + 31 <= hi_i) && (jc0 >= lo_i));
      uint32_t dsk[16];
      cb::mbar_wait(&sm.s_full, ph);
      cb::tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float s[16];
        {
          uint32_t r0[16];
          tmem_ld_32x32b_x16(lane_addr + COL_S + g * 32 + hf * 16, r0);
          cb::tmem_ld_wait();
          if (hf == 1) {
            cb::tc_fence_before();
            cb::mbar_arrive(&sm.s_free);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) s[e] = __uint_as_float(r0[e]);
        }
        shear_add16(s, row_v + 2 * c0 - 32 * hf);
        if (hf == 0) {
          cb::mbar_wait(&sm.dp_full, ph);
          cb::tc_fence_after();
        }
        uint32_t r0[16];
        tmem_ld_32x32b_x16(lane_addr + COL_X + g * 32 + hf * 16, r0);
        cb::tmem_ld_wait();
        if (hf == 1) {
          cb::tc_fence_before();
          cb::mbar_arrive(&sm.x_free);
        }
        if (full)
          pds16<DROP, false, false>(s, r0, sl2, lse2, delta, jc0 + hf * 16, hi_i, lo_i, dkeys, p.drop_thr2, nullptr, dsk + hf * 8);
        else
          pds16<DROP, true, false>(s, r0, sl2, lse2, delta, jc0 + hf * 16, hi_i, lo_i, dkeys, p.drop_thr2, nullptr, dsk + hf * 8);
      }
      // ---- dS -> TMEM (A operand of dq += dS K); the previous tile's product must be done with those columns ----
      if (t > 0) {
        cb::mbar_wait(&sm.ds_free, ph ^ 1);
        cb::tc_fence_after();
      }
      tmem_st_32x32b_x16(lane_addr + COL_DS + g * 16, dsk);
      named_bar(2 + wq, NWG * 32);              // the row group is done with its staged rows (block t+1 aliases them)
      if (t == nt - 1) {                        // nobody will write the "hi" part of block nt: clear the piece first
        for (int x = (g * 32 + lane) * 16; x < 4 * QG; x += 128 * 16) sts_v4(buf_lo + piece + x, 0, 0, 0, 0);
        named_bar(2 + wq, NWG * 32);
      }
      // ---- inverse shift: dS[li, lj] -> band column c = c0 - e; c < 128: block t+1 at idx c, else block t at c - 128 ----
      if (g != wq) {
        const uint32_t base = (g > wq ? buf_lo + 16 * c0 : buf_hi + 16 * (c0 - 128)) + rowoff;
#pragma unroll
        for (int e = 0; e < 32; e += 2) sts_halves(base - 16 * e, base - 16 * (e + 1), dsk[e / 2]);
      } else {                                  // diagonal chunk: e < lane -> "hi", e >= lane -> "lo"
        const uint32_t base_lo = buf_lo + 16 * c0 + rowoff;
        const uint32_t base_hi = buf_hi + 16 * (c0 - 128) + rowoff;
#pragma unroll
        for (int e = 0; e < 32; e += 2)
          sts_halves((e < lane ? base_hi : base_lo) - 16 * e, (e + 1 < lane ? base_hi : base_lo) - 16 * (e + 1), dsk[e / 2]);
      }
      tmem_st_wait();
      cb::fence_proxy_async();
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.ds_full);
    }
    // ---- epilogue ----
    cb::mbar_wait(&sm.acc_full, 0);
    cb::tc_fence_after();
    uint32_t ra[16];
    tmem_ld_32x32b_x16(lane_addr + COL_DQ + g * 16, ra);
    cb::tmem_ld_wait();
    float fa[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) fa[e] = __uint_as_float(ra[e]) * p.scale;
    if (i < p.T) {
      bf16* dqr = p.dq + ((long long)i * p.B + b) * p.lddq + h * DH + g * 16;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint4 q;
        q.x = cb::pack_bf16(fa[ch * 8 + 0], fa[ch * 8 + 1]);
        q.y = cb::pack_bf16(fa[ch * 8 + 2], fa[ch * 8 + 3]);
        q.z = cb::pack_bf16(fa[ch * 8 + 4], fa[ch * 8 + 5]);
        q.w = cb::pack_bf16(fa[ch * 8 + 6], fa[ch * 8 + 7]);
        *reinterpret_cast<uint4*>(dqr + ch * 8) = q;
      }
    }
    // column sums over the 32 rows of this warp -> d r_w_bias + d r_r_bias (rows >= T hold exact zeros); the dR
    // pass later moves the d r_r_bias share from du to dvb
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float sa = cb::warp_sum(fa[e]);
      if (lane == 0) atomicAdd(p.du + h * DH + g * 16 + e, sa);
    }
  }
  cb::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    cb::tc_fence_after();
    cb::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// dq / (du + dvb) of commu_relattn_bwd on tcgen05 (same operand contract).  delta = rowsum(dO * O) [B,H,T].
// du receives colsum(dq) = d r_w_bias + d r_r_bias; commu_relattn_bwd_dr_tc subtracts the d r_r_bias share and
// adds it to dvb, so the two passes must run as a pair (commu_relattn_bwd does).
extern "C" int commu_relattn_bwd_dq_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                       int64_t ldkv, const void* r, int64_t ldr, int kr,
                                       const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                       int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                       const float* delta, void* dq, int64_t lddq, float* du, float* dvb,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  attn::Params p = {};
  p.q = (const bf16*)qu; p.k = (const bf16*)k; p.v = (const bf16*)v; p.r = (const bf16*)r;
  p.qu_s = (bf16*)const_cast<void*>(qu); p.qv_s = (bf16*)const_cast<void*>(qv);
  static const float dummy = 0.f;
  p.u = &dummy; p.vb = &dummy;
  p.reset = reset;
  p.ldq = ldq; p.ldkv = ldkv; p.ldr = ldr;
  p.T = T; p.M = M; p.B = B; p.H = H; p.Kr = kr;
  p.same_length = same_length; p.shift = shift; p.scale = scale;
  p.lse = const_cast<float*>(lse); p.delta = delta;
  p.dout = (const bf16*)dout; p.lddo = lddo;
  p.dq = (bf16*)dq; p.lddq = lddq; p.du = du; p.dvb = dvb;
  apply_drop_state(p);
  int rc = cb_host::check_attn_common(p, "relattn_bwd_dq_tc");
  if (rc) return rc;
  CB_REQUIRE(qv && lse && dout && delta && dq && du && dvb && lddq % 8 == 0 && lddo % 8 == 0, "relattn_bwd_dq_tc: bad args");
  const int Ktot = T + M;
  CUtensorMap tk, tv, tqu, tqv, tdo, tr;
  if ((rc = make_tmap_rows3d(&tk, k, (uint64_t)H * 64, B, Ktot, ldkv))) return rc;
  if ((rc = make_tmap_rows3d(&tv, v, (uint64_t)H * 64, B, Ktot, ldkv))) return rc;
  if ((rc = make_tmap_rows3d(&tqu, qu, (uint64_t)H * 64, B, T, ldq))) return rc;
  if ((rc = make_tmap_rows3d(&tqv, qv, (uint64_t)H * 64, B, T, ldq))) return rc;
  if ((rc = make_tmap_rows3d(&tdo, dout, (uint64_t)H * 64, B, T, lddo))) return rc;
  if ((rc = cb_host::make_tmap_bf16_2d(&tr, r, (uint64_t)H * 64, kr, ldr, 64, 128))) return rc;
  static bool attr = false;
  const int smem_bytes = (int)sizeof(Smem) + 1024;
  if (!attr) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(T, TM), H, B);
  if (p.drop_thr2) relattn_bwd_dq_tc_kernel<true><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p);
  else relattn_bwd_dq_tc_kernel<false><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
