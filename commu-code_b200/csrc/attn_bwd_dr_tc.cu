// Relative-position attention backward on tcgen05 tensor cores: the dR pass (diagonal walk).
//
// One CTA = 128 distances d0 .. d0+127 of one (batch, head); it walks the query tiles, i.e. it follows the
// diagonal band j = i + M - delta of the score matrix, so the gradient of R (shared by every query and
// batch element) accumulates in TMEM and is added to global memory once per CTA.
// In (query, distance) coordinates the roles of K/V and R swap with respect to the other kernels:
//     S'   = (q+v) R_tile^T               natural                         [128 q x 128 distances]
//     AC   = (q+u) Kwin^T  (two 128-key blocks of the key window, read through the relative shift)
//     dP   = dO    Vwin^T  (same two blocks of the value window, read through the relative shift)
//     dR  += dS'^T (q+v)                  A = dS' tile written by the softmax threads (MN-major)
// Window of query tile i0: keys jw0 .. jw0+255 with jw0 = i0 + M - d0 - 127; consecutive query tiles
// share one 128-key block, which stays in shared memory.
//
//     g   += dS'^T 1                      column sums of dS' (N = 16 product against a tile of ones)
// The column sums give d r_r_bias = scale * sum_delta g[delta] R[delta]; the dq pass has put
// colsum(dq) = d r_w_bias + d r_r_bias into du, so this pass adds its share to dvb and subtracts it from du.
//
// Autograd counterpart of commu/model/model.py:312-345 for d(r_head_k) (then d r_net.weight by a GEMM) and r_r_bias.
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
constexpr int TILE_BYTES = 128 * DH * 2;
constexpr int BAND_THREADS = 32 * (NWG * (NWG + 1) / 2);   // threads (li, g) with g >= wq (or g <= wq): 320
constexpr int COL_S = 0, COL_LO = 128, COL_HI = 256, COL_DR = 384, COL_G = 448;

struct Smem {
  uint8_t r[TILE_BYTES];
  uint8_t qu[2][TILE_BYTES];    // query-side tiles are double buffered
  uint8_t qv[2][TILE_BYTES];
  uint8_t dout[2][TILE_BYTES];
  uint8_t k[2][TILE_BYTES];
  uint8_t v[2][TILE_BYTES];
  uint8_t ds[16 * 3072];        // dS' tile: [16 groups of 8 q rows][distance atom0 | atom1 | spare][8 rows x 128 B];
                                //   a row group (32 rows) owns 12 KB of it, which ALSO holds its staged fp16 rows
                                //   of the banded blocks earlier in the same iteration (attn_tc_common.cuh)
  uint8_t ones[512];            // bf16 1.0: B operand (K-major, no swizzle, N = 16) of the column-sum product
  float gsum[TM];               // scale * column sums, for the d r_r_bias epilogue
  uint64_t r_full, q_full[2], q_empty[2], kv_full[2], kv_empty[2];
  uint64_t s_full, s_empty, lo_full, lo_empty, hi_full, hi_empty, pds_full, pds_empty, acc_full;
  uint32_t tmem_base;
};

template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1)
relattn_bwd_dr_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                         const __grid_constant__ CUtensorMap tm_qu, const __grid_constant__ CUtensorMap tm_qv,
                         const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_r,
                         const Params p) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int d0 = blockIdx.x * TN;
  const bool reset = p.reset && p.reset[b];
  // queries that own a visible key at some distance of this tile: j = i + M - delta >= lo(i) >= 0
  int i_min = max(0, d0 - p.M);
  if (reset) i_min = max(i_min, d0);
  bool any = i_min < p.T;
  if (p.same_length && d0 >= p.M + p.shift) any = false;
  const int it_first = i_min / TM;
  const int nq = any ? ((p.T - 1) / TM - it_first + 1) : 0;
  // key block kappa covers keys [jw00 + 128*kappa, +128), jw00 = window start of the first query tile
  const int jw00 = it_first * TM + p.M - d0 - (TN - 1);

  if (threadIdx.x == 0) {
    cb::mbar_init(&sm.r_full, 1);
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.q_full[s], 1); cb::mbar_init(&sm.q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.kv_full[s], 1); cb::mbar_init(&sm.kv_empty[s], 1); }
    cb::mbar_init(&sm.s_full, 1); cb::mbar_init(&sm.s_empty, SOFT);
    cb::mbar_init(&sm.lo_full, 1); cb::mbar_init(&sm.lo_empty, BAND_THREADS);
    cb::mbar_init(&sm.hi_full, 1); cb::mbar_init(&sm.hi_empty, BAND_THREADS);
    cb::mbar_init(&sm.pds_full, SOFT); cb::mbar_init(&sm.pds_empty, 1);
    cb::mbar_init(&sm.acc_full, 1);
    cb::fence_barrier_init();
  }
  if (warp == 1) {   // 512 bytes of bf16 ones
    sts_v4(cb::smem_u32(sm.ones) + lane * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    cb::fence_proxy_async();
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
      cb::mbar_arrive_expect_tx(&sm.r_full, TILE_BYTES);
      cb::tma_load_2d(sm.r, &tm_r, &sm.r_full, h * DH, d0);
      auto load_kv = [&](int kappa) {   // buffer kappa&1, its (kappa>>1)-th use
        const int bi = kappa & 1;
        const uint32_t use = (kappa >> 1) & 1;
        cb::mbar_wait(&sm.kv_empty[bi], use ^ 1);
        cb::mbar_arrive_expect_tx(&sm.kv_full[bi], 2 * TILE_BYTES);
        cb::tma_load_3d(sm.k[bi], &tm_k, &sm.kv_full[bi], h * DH, b, jw00 + TN * kappa);
        cb::tma_load_3d(sm.v[bi], &tm_v, &sm.kv_full[bi], h * DH, b, jw00 + TN * kappa);
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
      load_kv(0);
      load_q(0);
      for (int n = 0; n < nq; ++n) {
        load_kv(n + 1);
        if (n + 1 < nq) load_q(n + 1);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one() && nq > 0) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_g = cb::umma_idesc_bf16(TN, DH, 1, 1);   // dR: MN-major A (dS'^T), MN-major B (q+v)
      const uint32_t idesc_1 = cb::umma_idesc_bf16(TN, 16, 1, 0);   // g : MN-major A (dS'^T), K-major B (ones)
      const uint64_t b_ones = umma_smem_desc_nosw(cb::smem_u32(sm.ones), 256, 128);
      uint32_t s_phase = 0, lo_phase = 0, hi_phase = 0, pds_phase = 0;
      const uint32_t a_r = cb::smem_u32(sm.r);
      cb::mbar_wait(&sm.r_full, 0);
      // banded products alternate between two accumulators: "lo" blocks (key block kappa = n) and "hi" blocks
      // (kappa = n + 1); each may be overwritten once the threads that stage it have read the previous one
      auto issue_band = [&](uint32_t a_addr, uint32_t b_addr, bool hi) {
        if (hi) { cb::mbar_wait(&sm.hi_empty, hi_phase ^ 1); hi_phase ^= 1; }
        else    { cb::mbar_wait(&sm.lo_empty, lo_phase ^ 1); lo_phase ^= 1; }
        cb::tc_fence_after();
        const uint64_t ad = cb::umma_smem_desc(a_addr, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(b_addr, 16, 1024);
        const uint32_t col = tmem + (hi ? COL_HI : COL_LO);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(col, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
        cb::umma_commit(hi ? &sm.hi_full : &sm.lo_full);
      };
      // "front" of query tile n: S' and the "lo" AC block; issued one tile ahead of the softmax threads
      auto issue_front = [&](int n) {
        const int qb = n & 1, lo_b = n & 1;
        cb::mbar_wait(&sm.q_full[qb], (n >> 1) & 1);
        cb::mbar_wait(&sm.s_empty, s_phase ^ 1);
        cb::tc_fence_after();
        const uint64_t aq = cb::umma_smem_desc(cb::smem_u32(sm.qv[qb]), 16, 1024), br = cb::umma_smem_desc(a_r, 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + COL_S, aq + 2 * k, br + 2 * k, idesc_s, k > 0);
        cb::umma_commit(&sm.s_full);
        s_phase ^= 1;
        cb::mbar_wait(&sm.kv_full[lo_b], (n >> 1) & 1);
        issue_band(cb::smem_u32(sm.qu[qb]), cb::smem_u32(sm.k[lo_b]), false);   // AC "lo"
      };
      issue_front(0);
      for (int n = 0; n < nq; ++n) {
        const int qb = n & 1;
        const int lo_b = n & 1, hi_b = (n + 1) & 1;   // key blocks kappa = n ("lo") and n+1 ("hi")
        const uint32_t a_qu = cb::smem_u32(sm.qu[qb]), a_qv = cb::smem_u32(sm.qv[qb]), a_do = cb::smem_u32(sm.dout[qb]);
        cb::mbar_wait(&sm.kv_full[hi_b], ((n + 1) >> 1) & 1);
        issue_band(a_qu, cb::smem_u32(sm.k[hi_b]), true);    // AC "hi"
        issue_band(a_do, cb::smem_u32(sm.v[lo_b]), false);   // dP "lo"
        issue_band(a_do, cb::smem_u32(sm.v[hi_b]), true);    // dP "hi"
        cb::umma_commit(&sm.kv_empty[lo_b]);                 // block kappa = n is dead after this tile
        if (n + 1 < nq) issue_front(n + 1);
        cb::mbar_wait(&sm.pds_full, pds_phase);
        cb::tc_fence_after();
        {
          // MN-major A: 64-distance atoms 1024 B apart (LBO), 8-query-row groups 3072 B apart (SBO)
          const uint64_t as = cb::umma_smem_desc(cb::smem_u32(sm.ds), 1024, 3072);
          const uint64_t bq = cb::umma_smem_desc(a_qv, 8192, 1024);
#pragma unroll
          for (int k = 0; k < TM / 16; ++k)
            cb::umma_bf16_ss(tmem + COL_DR, as + (uint64_t)(k * 384), bq + (uint64_t)(k * 128), idesc_g, (n > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < TM / 16; ++k)
            cb::umma_bf16_ss(tmem + COL_G, as + (uint64_t)(k * 384), b_ones, idesc_1, (n > 0 || k > 0));
          cb::umma_commit(&sm.pds_empty);
          cb::umma_commit(&sm.q_empty[qb]);
        }
        pds_phase ^= 1;
      }
      cb::umma_commit(&sm.acc_full);
    }
  } else if (warp >= 2) {
    // ============================== softmax warpgroups ==============================
    // thread = (query row li of the tile, 32-distance chunk g)
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;                     // TMEM lane quadrant of this warp (hardware: warp id % 4)
    const int li = wq * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    // staged row of this query row inside its row group's 12 KB piece; row_v = base of position 0
    const uint32_t row_v = cb::smem_u32(sm.ds) + wq * 12288 + stage_row_off(lane) - 64 * wq;
    const uint32_t shear0 = row_v + 2 * (li + 127 - 32 * g);
    const uint32_t drow = cb::smem_u32(sm.ds) + (li >> 3) * 3072 + (g >> 1) * 1024 + (li & 7) * 128;
    const int cx = ((g & 1) * 4) ^ (li & 7);
    const float sl2 = p.scale * 1.4426950408889634f;
    uint32_t s_phase = 0, band_phase = 0, pds_phase = 0;
    const float* lse_p = p.lse + ((long long)b * p.H + h) * p.T;
    const float* del_p = p.delta + ((long long)b * p.H + h) * p.T;
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
      const int lo_i = key_lo(i, p.M, p.same_length, p.shift, reset);
      cb::mbar_wait(&sm.s_full, s_phase);
      cb::tc_fence_after();
      float s[32], dp[32];
      {
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_S + g * 32, r0);
        cb::tmem_ld_wait();
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.s_empty);
        s_phase ^= 1;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          s[e] = __uint_as_float(r0[e]);
          dp[e] = 0.f;
        }
      }
      // two banded products (AC into s, dP into dp), each a "lo" and a "hi" block: the row group stages the block
      // columns its rows need as fp16 positions and reads them back sheared.  The staged rows alias the dS' tile:
      // the previous iteration's dR product must have consumed it first.
      cb::mbar_wait(&sm.pds_empty, pds_phase ^ 1);
#pragma unroll
      for (int prod = 0; prod < 2; ++prod) {
        if (prod) named_bar(2 + wq, NWG * 32);  // this row group finished reading the AC positions
        if (g >= wq) {
          cb::mbar_wait(&sm.lo_full, band_phase);
          cb::tc_fence_after();
          stage32_x16(lane_addr + COL_LO + g * 32, row_v + 64 * g);
          cb::tc_fence_before();
          cb::mbar_arrive(&sm.lo_empty);
        }
        if (g <= wq) {
          cb::mbar_wait(&sm.hi_full, band_phase);
          cb::tc_fence_after();
          stage32_x16(lane_addr + COL_HI + g * 32, row_v + 256 + 64 * g);
          cb::tc_fence_before();
          cb::mbar_arrive(&sm.hi_empty);
        }
        band_phase ^= 1;
        named_bar(2 + wq, NWG * 32);            // the positions of this row group are staged
        if (prod == 0) shear_add32(s, shear0);
        else shear_add32(dp, shear0);
      }
      // ---- P, dS' (columns are distances: key j = i + M - delta) ----
      uint32_t dsk[16];
      const int jk0 = i + p.M - (d0 + g * 32);   // key of column 0; column e is key jk0 - e
      const bool full = (i < p.T) && (jk0 - 31 >= lo_i);
      if (DROP) {   // dP of a dropped probability does not reach dS': mask it with the forward's keep flags
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t fl[4];
          drop_flags_desc8(fl, jk0, c, dkeys, p.drop_thr2);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            dp[8 * c + 2 * q] = __uint_as_float(__float_as_uint(dp[8 * c + 2 * q]) & drop::mask32_hi(fl[q]));
            dp[8 * c + 2 * q + 1] = __uint_as_float(__float_as_uint(dp[8 * c + 2 * q + 1]) & drop::mask32_lo(fl[q]));
          }
        }
      }
      if (__all_sync(0xffffffffu, full)) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2(fmaf(s[e], sl2, -lse2)), p1 = ex2(fmaf(s[e + 1], sl2, -lse2));
          dsk[e / 2] = cb::pack_bf16(p0 * (dp[e] - delta), p1 * (dp[e + 1] - delta));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = ex2(fmaf(s[e], sl2, -lse2)), p1 = ex2(fmaf(s[e + 1], sl2, -lse2));
          p0 = (i >= p.T || jk0 - e < lo_i) ? 0.f : p0;
          p1 = (i >= p.T || jk0 - e - 1 < lo_i) ? 0.f : p1;
          dsk[e / 2] = cb::pack_bf16(p0 * (dp[e] - delta), p1 * (dp[e + 1] - delta));
        }
      }
      named_bar(2 + wq, NWG * 32);              // the row group is done with its staged rows (dS' aliases them)
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        sts_v4(drow + ((cx ^ c4) << 4), dsk[c4 * 4], dsk[c4 * 4 + 1], dsk[c4 * 4 + 2], dsk[c4 * 4 + 3]);
      cb::fence_proxy_async();
      cb::mbar_arrive(&sm.pds_full);
      pds_phase ^= 1;
    }
    // ---- epilogue: dR rows (thread = distance row li, 16 of the 64 head dims), summed over batch with atomics ----
    if (nq > 0) {
      cb::mbar_wait(&sm.acc_full, 0);
      cb::tc_fence_after();
      uint32_t rr[16];
      tmem_ld_32x32b_x16(lane_addr + COL_DR + g * 16, rr);
      cb::tmem_ld_wait();
      const int dl = d0 + li;
      if (dl < p.Kr) {
        float* dst = p.dr + (long long)dl * p.H * DH + h * DH + g * 16;
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          cb::red_add_v4(dst + e, __uint_as_float(rr[e]) * p.scale, __uint_as_float(rr[e + 1]) * p.scale,
                         __uint_as_float(rr[e + 2]) * p.scale, __uint_as_float(rr[e + 3]) * p.scale);
      }
      // d r_r_bias share of this distance tile: scale * sum_li g[li] * R[d0 + li, :]  (rows >= Kr are zero-filled)
      if (g == 0) {
        uint32_t rg[16];
        tmem_ld_32x32b_x16(lane_addr + COL_G, rg);
        cb::tmem_ld_wait();
        sm.gsum[li] = __uint_as_float(rg[0]) * p.scale;
      }
      named_bar(1, SOFT);
      const int c = li;
      if (g == 0 && c < DH) {
        float acc = 0.f;
        const uint8_t* rt = sm.r;
#pragma unroll 8
        for (int rrow = 0; rrow < TN; ++rrow) {
          const bf16 rv = *reinterpret_cast<const bf16*>(rt + attn::swz(rrow, c >> 3) + (c & 7) * 2);
          acc = fmaf(sm.gsum[rrow], __bfloat162float(rv), acc);
        }
        atomicAdd(p.dvb + h * DH + c, acc);
        atomicAdd(p.du + h * DH + c, -acc);
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

// dR of commu_relattn_bwd on tcgen05: dr fp32 [kr, H*64] accumulated (+=) over batch.  delta = rowsum(dO*O).
// Also moves the d r_r_bias share of colsum(dq) (left in du by commu_relattn_bwd_dq_tc) from du to dvb.
extern "C" int commu_relattn_bwd_dr_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                       int64_t ldkv, const void* r, int64_t ldr, int kr,
                                       const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                       int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                       const float* delta, float* dr, float* du, float* dvb, void* stream_) {
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
  p.dr = dr; p.du = du; p.dvb = dvb;
  apply_drop_state(p);
  int rc = cb_host::check_attn_common(p, "relattn_bwd_dr_tc");
  if (rc) return rc;
  CB_REQUIRE(qv && lse && dout && delta && dr && du && dvb && lddo % 8 == 0, "relattn_bwd_dr_tc: bad args");
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
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dr_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dr_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(Ktot, TN), H, B);
  if (p.drop_thr2) relattn_bwd_dr_tc_kernel<true><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p);
  else relattn_bwd_dr_tc_kernel<false><<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
