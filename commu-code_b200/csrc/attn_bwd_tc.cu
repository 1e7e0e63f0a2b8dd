// Relative-position attention backward on tcgen05 tensor cores: the dk / dv pass.
//
// One CTA = 128 keys of one (batch, head); it walks the query tiles that can see those keys.  Per
// query tile five products run on the tensor cores with accumulators in TMEM:
//     S    = (q+u) K^T          [128 q x 128 keys]      A, B: TMA-staged, K-major
//     BD   = (q+v) R_blk^T      twice (the "lo" and "hi" 128-distance blocks of the band)
//     dP   = dO V^T             [128 q x 128 keys]
//     dV  += P^T dO             [128 keys x 64]          A = P  tile written by the softmax threads (MN-major)
//     dK  += dS^T (q+u)         [128 keys x 64]          A = dS tile (MN-major), B = (q+u) tile (MN-major)
// The softmax threads (thread = query row x 64 key columns) recompute P = exp2(score*log2e - LSE),
// form dS = P * (dP - Delta), and store both as bf16 rows of two shared-memory tiles whose layout is at
// once the K-major [q][key] tile and the MN-major operand the dV / dK products need - no transpose.
// The relative shift: every row group (32 query rows = one TMEM lane quadrant = the four warps of one
// scheduler) stages the band-block columns its rows need as fp16 "positions" (attn_tc_common.cuh) and reads
// them back sheared with packed 32-bit loads + FHADD.  The P / dS tiles are laid out so that a row group owns
// one contiguous 16 KB piece ([8-row group][P atom0 | P atom1 | dS atom0 | dS atom1]); the staged rows
// alias that piece, so the softmax loop synchronises row groups only (128-thread named barriers).
//
// Autograd counterpart of commu/model/model.py:312-345 for d(keys), d(values).
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
constexpr int BAND_THREADS = 32 * (NWG * (NWG + 1) / 2) * 4 / NWG;   // threads (li, g) with g >= wq (or g <= wq): 320
constexpr int COL_S = 0, COL_X = 128, COL_LO = 256, COL_DV = 384, COL_DK = 448;   // X: BD "hi", then dP

struct Smem {
  uint8_t k[TILE_BYTES];
  uint8_t v[TILE_BYTES];
  uint8_t qu[2][TILE_BYTES];    // query-side tiles are double buffered: the next query tile streams in while
  uint8_t qv[2][TILE_BYTES];    // the current one is being processed
  uint8_t dout[2][TILE_BYTES];
  uint8_t r[2][TILE_BYTES];
  uint8_t pds[4 * TILE_BYTES];  // [16 groups of 8 q rows][P atom0 | P atom1 | dS atom0 | dS atom1][8 rows x 128 B];
                                //   ALSO the fp16 staging rows of the BD blocks (row group wq: pds + 16 KB * wq)
  uint64_t kv_full, q_full[2], q_empty[2], r_full[2], r_empty[2];
  uint64_t s_full, s_free, lo_full, lo_free, hi_full, hi_done, dp_full, x_free, pds_full[4], pds_empty[4], acc_full;
  uint32_t tmem_base;
};

template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1)
relattn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                          const __grid_constant__ CUtensorMap tm_qu, const __grid_constant__ CUtensorMap tm_qv,
                          const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_r,
                          const Params p, bf16* __restrict__ dk_out, bf16* __restrict__ dv_out, long long lddkv) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * TN;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;
  // query tiles that can see a key of this tile (same bounds as the v1 dkv pass)
  int i_min = max(0, j0 - p.M);
  int i_max = p.T - 1;
  if (p.same_length) i_max = min(i_max, j0 + TN - 2 + p.shift);
  if (reset && j0 + TN - 1 < p.M) i_max = -1;
  const int it_first = i_min / TM;
  const int nq = i_max >= it_first * TM ? (i_max / TM - it_first + 1) : 0;
  // R block gamma covers rows [dlo0 + 128*gamma, +128), dlo0 = band start of the first query tile
  const int dlo0 = it_first * TM + p.M - j0 - (TN - 1);

  if (threadIdx.x == 0) {
    cb::mbar_init(&sm.kv_full, 1);
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.q_full[s], 1); cb::mbar_init(&sm.q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.r_full[s], 1); cb::mbar_init(&sm.r_empty[s], 1); }
    cb::mbar_init(&sm.s_full, 1); cb::mbar_init(&sm.s_free, SOFT);
    cb::mbar_init(&sm.lo_full, 1); cb::mbar_init(&sm.lo_free, BAND_THREADS);
    cb::mbar_init(&sm.hi_full, 1); cb::mbar_init(&sm.hi_done, BAND_THREADS);
    cb::mbar_init(&sm.dp_full, 1); cb::mbar_init(&sm.x_free, SOFT);
    for (int s = 0; s < 4; ++s) { cb::mbar_init(&sm.pds_full[s], NWG * 32); cb::mbar_init(&sm.pds_empty[s], 1); }
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
    if (cb::elect_one() && nq > 0) {
      cb::mbar_arrive_expect_tx(&sm.kv_full, 2 * TILE_BYTES);
      cb::tma_load_3d(sm.k, &tm_k, &sm.kv_full, h * DH, b, j0);
      cb::tma_load_3d(sm.v, &tm_v, &sm.kv_full, h * DH, b, j0);
      auto load_r = [&](int gamma) {   // slot gamma & 1, its (gamma >> 1)-th use
        const int sl = gamma & 1;
        cb::mbar_wait(&sm.r_empty[sl], ((gamma >> 1) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.r_full[sl], TILE_BYTES);
        cb::tma_load_2d(sm.r[sl], &tm_r, &sm.r_full[sl], h * DH, dlo0 + TN * gamma);
      };
      auto load_q = [&](int n) {   // buffer n&1, its (n>>1)-th use
        const int bi = n & 1;
        const int i0 = (it_first + n) * TM;
        cb::mbar_wait(&sm.q_empty[bi], ((n >> 1) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.q_full[bi], 3 * TILE_BYTES);
        cb::tma_load_3d(sm.qu[bi], &tm_qu, &sm.q_full[bi], h * DH, b, i0);
        cb::tma_load_3d(sm.qv[bi], &tm_qv, &sm.q_full[bi], h * DH, b, i0);
        cb::tma_load_3d(sm.dout[bi], &tm_do, &sm.q_full[bi], h * DH, b, i0);
      };
      load_q(0);
      load_r(0);
      load_r(1);
      for (int n = 0; n + 1 < nq; ++n) {
        load_q(n + 1);
        load_r(n + 2);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one() && nq > 0) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);   // S, BD, dP: K-major x K-major
      const uint32_t idesc_g = cb::umma_idesc_bf16(TN, DH, 1, 1);   // dV, dK: MN-major A (tile^T), MN-major B
      cb::mbar_wait(&sm.kv_full, 0);
      const uint32_t a_k = cb::smem_u32(sm.k), a_v = cb::smem_u32(sm.v);
      auto kmajor_128 = [&](uint32_t col, uint32_t a_addr, uint32_t b_addr) {
        const uint64_t ad = cb::umma_smem_desc(a_addr, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(b_addr, 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + col, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
      };
      // R blocks of query tile n: "lo" = gamma n, "hi" = gamma n+1 (slot gamma & 1).
      // "front" of a tile = S, BD "lo", BD "hi": all three are issued one tile ahead (software pipelining), each
      // as soon as the softmax threads have copied the previous contents of its TMEM columns; dP follows into
      // the "hi" columns once those are staged.
      cb::mbar_wait(&sm.q_full[0], 0);
      cb::tc_fence_after();
      kmajor_128(COL_S, cb::smem_u32(sm.qu[0]), a_k);
      cb::umma_commit(&sm.s_full);
      cb::mbar_wait(&sm.r_full[0], 0);
      kmajor_128(COL_LO, cb::smem_u32(sm.qv[0]), cb::smem_u32(sm.r[0]));
      cb::umma_commit(&sm.lo_full);
      cb::umma_commit(&sm.r_empty[0]);
      cb::mbar_wait(&sm.r_full[1], 0);
      kmajor_128(COL_X, cb::smem_u32(sm.qv[0]), cb::smem_u32(sm.r[1]));
      cb::umma_commit(&sm.hi_full);
      for (int n = 0; n < nq; ++n) {
        const int qb = n & 1;
        const uint32_t ph = n & 1;
        const uint32_t a_qu = cb::smem_u32(sm.qu[qb]), a_do = cb::smem_u32(sm.dout[qb]);
        // ---- dP(n) into the "hi" columns once every thread that needs them has copied them ----
        cb::mbar_wait(&sm.hi_done, ph);
        cb::tc_fence_after();
        kmajor_128(COL_X, a_do, a_v);
        cb::umma_commit(&sm.dp_full);
        // ---- front of tile n+1 ----
        if (n + 1 < nq) {
          const int qn = (n + 1) & 1;
          cb::mbar_wait(&sm.q_full[qn], ((n + 1) >> 1) & 1);
          cb::mbar_wait(&sm.s_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_S, cb::smem_u32(sm.qu[qn]), a_k);
          cb::umma_commit(&sm.s_full);
          cb::mbar_wait(&sm.lo_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_LO, cb::smem_u32(sm.qv[qn]), cb::smem_u32(sm.r[(n + 1) & 1]));   // gamma = n+1 (resident)
          cb::umma_commit(&sm.lo_full);
          cb::umma_commit(&sm.r_empty[(n + 1) & 1]);                                       // last use of gamma n+1
          cb::mbar_wait(&sm.r_full[n & 1], ((n + 2) >> 1) & 1);                             // gamma = n+2
          cb::mbar_wait(&sm.x_free, ph);
          cb::tc_fence_after();
          kmajor_128(COL_X, cb::smem_u32(sm.qv[qn]), cb::smem_u32(sm.r[n & 1]));
          cb::umma_commit(&sm.hi_full);
        }
        // ---- back of tile n: dV += P^T dO ; dK += dS^T (q+u), row group by row group (16 query rows per MMA) ----
        {
          // MN-major A: 64-key atoms 1024 B apart (LBO), 8-query-row groups 4096 B apart (SBO)
          const uint64_t ap = cb::umma_smem_desc(cb::smem_u32(sm.pds), 1024, 4096);
          const uint64_t as = cb::umma_smem_desc(cb::smem_u32(sm.pds) + 2048, 1024, 4096);
          const uint64_t bo = cb::umma_smem_desc(a_do, 8192, 1024), bq = cb::umma_smem_desc(a_qu, 8192, 1024);
#pragma unroll
          for (int rg = 0; rg < 4; ++rg) {
            cb::mbar_wait(&sm.pds_full[rg], ph);
            cb::tc_fence_after();
#pragma unroll
            for (int k = 2 * rg; k < 2 * rg + 2; ++k)
              cb::umma_bf16_ss(tmem + COL_DV, ap + (uint64_t)(k * 512), bo + (uint64_t)(k * 128), idesc_g, (n > 0 || k > 0));
#pragma unroll
            for (int k = 2 * rg; k < 2 * rg + 2; ++k)
              cb::umma_bf16_ss(tmem + COL_DK, as + (uint64_t)(k * 512), bq + (uint64_t)(k * 128), idesc_g, (n > 0 || k > 0));
            cb::umma_commit(&sm.pds_empty[rg]);
          }
          cb::umma_commit(&sm.q_empty[qb]);
        }
      }
      cb::umma_commit(&sm.acc_full);
    }
  } else if (warp >= 2) {
    // ============================== softmax warpgroups ==============================
    // thread = (query row li of the tile, 32-key chunk g)
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;                     // TMEM lane quadrant of this warp (hardware: warp id % 4)
    const int li = wq * 32 + lane;               // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    // staged row of this query row inside its row group's 16 KB piece; row_v = base of position 0
    const uint32_t row_v = cb::smem_u32(sm.pds) + wq * 16384 + stage_row_off(lane) - 64 * wq;
    const uint32_t shear0 = row_v + 2 * (li + 127 - 32 * g);      // position of this thread's first key column
    // rows of the P / dS tiles: 8-row group li>>3, key atom g>>1, chunks 4*(g&1)..+3 (128B swizzle)
    const uint32_t prow = cb::smem_u32(sm.pds) + (li >> 3) * 4096 + (g >> 1) * 1024 + (li & 7) * 128;
    const int cx = ((g & 1) * 4) ^ (li & 7);
    const float sl2 = p.scale * 1.4426950408889634f;
    const float* lse_p = p.lse + ((long long)b * p.H + h) * p.T;
    const float* del_p = p.delta + ((long long)b * p.H + h) * p.T;
    // per-row constants of the next tile are fetched one tile ahead
    float lse_n = 0.f, del_n = 0.f;
    {
      const int i = it_first * TM + li;
      if (nq > 0 && i < p.T) { lse_n = lse_p[i]; del_n = del_p[i]; }
    }

    for (int n = 0; n < nq; ++n) {
      const int i = (it_first + n) * TM + li;
      // under dropout P is scaled by 1 / keep (folded into the exponent) and Delta by keep
      const float lse2 = lse_n * 1.4426950408889634f + (DROP ? log2f(p.drop_keep) : 0.f);
      const float delta = DROP ? del_n * p.drop_keep : del_n;
      const drop::Keys dkeys = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)((b * p.H + h) * p.T + i));
      {
        const int inext = i + TM;
        lse_n = 0.f; del_n = 0.f;
        if (n + 1 < nq && inext < p.T) { lse_n = lse_p[inext]; del_n = del_p[inext]; }
      }
      const int hi_i = i < p.T ? i + p.M : -1;
      const int lo_i = key_lo(i, p.M, p.same_length, p.shift, reset);
      const uint32_t ph = n & 1;
      // ---- scores of this thread's 32 key columns ----
      cb::mbar_wait(&sm.s_full, ph);
      cb::tc_fence_after();
      float s[32];
      {
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_S + g * 32, r0);
        cb::tmem_ld_wait();
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.s_free);
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] = __uint_as_float(r0[e]);
      }
      // ---- relative shift: copy the band-block columns this row needs (fp16, in registers), then - once the
      // previous iteration's dV / dK products are done with the P / dS rows that the staged rows alias - put
      // them into the row group's staging piece and read them back sheared.
      uint32_t sg[16];
      if (g >= wq) {
        cb::mbar_wait(&sm.lo_full, ph);
        cb::tc_fence_after();
        load_pack32(lane_addr + COL_LO + g * 32, sg);
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.lo_free);
      } else {
        cb::mbar_wait(&sm.hi_full, ph);
        cb::tc_fence_after();
        load_pack32(lane_addr + COL_X + g * 32, sg);
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.hi_done);
      }
      if (n > 0) cb::mbar_wait(&sm.pds_empty[wq], ph ^ 1);
      store_packed32(g >= wq ? row_v + 64 * g : row_v + 256 + 64 * g, sg);
      if (g == wq) {                            // the diagonal chunk needs both blocks
        cb::mbar_wait(&sm.hi_full, ph);
        cb::tc_fence_after();
        load_pack32(lane_addr + COL_X + g * 32, sg);
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.hi_done);
        store_packed32(row_v + 256 + 64 * g, sg);
      }
      named_bar(2 + wq, NWG * 32);              // the positions of this row group are staged
      shear_add32(s, shear0);
      // ---- P = exp2(score*log2e - LSE), dS = P * (dP - Delta)  (the 1/sqrt(Dh) factor is applied to dK at the end)
      const int jc0 = j0 + g * 32;
      const bool full = (jc0 + 31 <= hi_i) && (jc0 >= lo_i);
      uint32_t pk[16], dsk[16];
      {
        cb::mbar_wait(&sm.dp_full, ph);
        cb::tc_fence_after();
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_X + g * 32, r0);
        cb::tmem_ld_wait();
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.x_free);
        if (__all_sync(0xffffffffu, full)) {
          pds16<DROP, false, true>(s, r0, sl2, lse2, delta, jc0, hi_i, lo_i, dkeys, p.drop_thr2, pk, dsk);
          pds16<DROP, false, true>(s + 16, r0 + 16, sl2, lse2, delta, jc0 + 16, hi_i, lo_i, dkeys, p.drop_thr2, pk + 8, dsk + 8);
        } else {
          pds16<DROP, true, true>(s, r0, sl2, lse2, delta, jc0, hi_i, lo_i, dkeys, p.drop_thr2, pk, dsk);
          pds16<DROP, true, true>(s + 16, r0 + 16, sl2, lse2, delta, jc0 + 16, hi_i, lo_i, dkeys, p.drop_thr2, pk + 8, dsk + 8);
        }
      }
      named_bar(2 + wq, NWG * 32);              // the row group is done with its staged rows (P / dS alias them)
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t a = prow + ((cx ^ c4) << 4);
        sts_v4(a, pk[c4 * 4], pk[c4 * 4 + 1], pk[c4 * 4 + 2], pk[c4 * 4 + 3]);
        sts_v4(a + 2048, dsk[c4 * 4], dsk[c4 * 4 + 1], dsk[c4 * 4 + 2], dsk[c4 * 4 + 3]);
      }
      cb::fence_proxy_async();
      cb::mbar_arrive(&sm.pds_full[wq]);
    }
    // ---- epilogue: dV, dK rows (thread = key row li, 16 of the 64 head dims) ----
    const int j = j0 + li;
    bf16* dvr = dv_out + ((long long)j * p.B + b) * lddkv + h * DH + g * 16;
    bf16* dkr = dk_out + ((long long)j * p.B + b) * lddkv + h * DH + g * 16;
    if (nq > 0) {
      cb::mbar_wait(&sm.acc_full, 0);
      cb::tc_fence_after();
      uint32_t rv[16], rk[16];
      tmem_ld_32x32b_x16(lane_addr + COL_DV + g * 16, rv);
      tmem_ld_32x32b_x16(lane_addr + COL_DK + g * 16, rk);
      cb::tmem_ld_wait();
      if (j < Ktot) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint4 a, c2;
          a.x = cb::pack_bf16(__uint_as_float(rv[ch * 8 + 0]), __uint_as_float(rv[ch * 8 + 1]));
          a.y = cb::pack_bf16(__uint_as_float(rv[ch * 8 + 2]), __uint_as_float(rv[ch * 8 + 3]));
          a.z = cb::pack_bf16(__uint_as_float(rv[ch * 8 + 4]), __uint_as_float(rv[ch * 8 + 5]));
          a.w = cb::pack_bf16(__uint_as_float(rv[ch * 8 + 6]), __uint_as_float(rv[ch * 8 + 7]));
          c2.x = cb::pack_bf16(__uint_as_float(rk[ch * 8 + 0]) * p.scale, __uint_as_float(rk[ch * 8 + 1]) * p.scale);
          c2.y = cb::pack_bf16(__uint_as_float(rk[ch * 8 + 2]) * p.scale, __uint_as_float(rk[ch * 8 + 3]) * p.scale);
          c2.z = cb::pack_bf16(__uint_as_float(rk[ch * 8 + 4]) * p.scale, __uint_as_float(rk[ch * 8 + 5]) * p.scale);
          c2.w = cb::pack_bf16(__uint_as_float(rk[ch * 8 + 6]) * p.scale, __uint_as_float(rk[ch * 8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dvr + ch * 8) = a;
          *reinterpret_cast<uint4*>(dkr + ch * 8) = c2;
        }
      }
    } else if (j < Ktot) {   // nothing attends to these keys: zero gradients
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        *reinterpret_cast<uint4*>(dvr + ch * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dkr + ch * 8) = make_uint4(0, 0, 0, 0);
      }
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

// dk / dv of commu_relattn_bwd on tcgen05 (same operand contract; dq, dr, du, dvb come from the other passes).
// delta must already hold rowsum(dO * O) [B,H,T].
extern "C" int commu_relattn_bwd_dkv_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                        int64_t ldkv, const void* r, int64_t ldr, int kr,
                                        const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                        int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                        const float* delta, void* dk, void* dv, int64_t lddkv, void* stream_) {
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
  apply_drop_state(p);
  int rc = cb_host::check_attn_common(p, "relattn_bwd_dkv_tc");
  if (rc) return rc;
  CB_REQUIRE(qv && lse && dout && delta && dk && dv && lddkv % 8 == 0 && lddo % 8 == 0, "relattn_bwd_dkv_tc: bad args");
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
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dkv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dkv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(Ktot, TN), H, B);
  if (p.drop_thr2)
    relattn_bwd_dkv_tc_kernel<true><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p, (bf16*)dk, (bf16*)dv, lddkv);
  else
    relattn_bwd_dkv_tc_kernel<false><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p, (bf16*)dk, (bf16*)dv, lddkv);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
