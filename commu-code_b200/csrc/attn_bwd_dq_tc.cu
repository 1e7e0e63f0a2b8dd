// Relative-position attention backward on tcgen05 tensor cores: the dq pass (+ d r_w_bias, d r_r_bias).
//
// One CTA = 128 query rows of one (batch, head); it walks the key tiles like the forward kernel.  Per key
// tile t six products run on the tensor cores (accumulators in TMEM):
//     S      = (q+u) K_t^T                           [128 x 128]
//     BD     = (q+v) R_blk^T   ("lo" then "hi" 128-distance block, staged one at a time)
//     dP     = dO V_t^T                              [128 x 128]
//     dq_ac += dS K_t                                [128 x 64]   A = dS (bf16) in TMEM, B = K tile (MN-major)
//     dq_bd += dBD [R_lo ; R_hi]                     [128 x 64]   A = dBD band tile in shared memory
// dBD is the inverse relative shift of dS: row li of the 128 x 256 band tile holds dS[li, lj] at band
// column li + 127 - lj (every row owns a fixed run of 128 columns, the rest of the tile stays zero), so
// the product with the two R blocks the band spans is the position-term gradient.
// dq = scale * (dq_ac + dq_bd); d r_w_bias += colsum(scale * dq_ac); d r_r_bias += colsum(scale * dq_bd).
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
constexpr int NTHREADS = 128 + SOFT;
constexpr int TILE_BYTES = 128 * DH * 2;      // 16 KB
constexpr int STAGE_ROW = 272;
constexpr int COL_S = 0, COL_DP = 128, COL_BD = 256, COL_DQA = 384, COL_DQB = 448;

struct Smem {
  uint8_t qu[TILE_BYTES];
  uint8_t qv[TILE_BYTES];
  uint8_t dout[TILE_BYTES];
  uint8_t k[TILE_BYTES];
  uint8_t v[TILE_BYTES];
  uint8_t r[2][TILE_BYTES];
  uint8_t dbd[4 * TILE_BYTES];  // band tile: 4 K-atoms of 64 band columns, [128 rows][128 B] each
  uint8_t bd[TM * STAGE_ROW];
  uint64_t q_full, k_full, k_empty, r_full[2], r_empty[2];
  uint64_t s_full, s_empty, bd_full, bd_empty, ds_full, ds_empty, acc_full;
  uint32_t tmem_base;
};

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
    cb::mbar_init(&sm.k_full, 1); cb::mbar_init(&sm.k_empty, 1);
    for (int s = 0; s < 2; ++s) { cb::mbar_init(&sm.r_full[s], 1); cb::mbar_init(&sm.r_empty[s], 1); }
    cb::mbar_init(&sm.s_full, 1); cb::mbar_init(&sm.s_empty, SOFT);
    cb::mbar_init(&sm.bd_full, 1); cb::mbar_init(&sm.bd_empty, SOFT);
    cb::mbar_init(&sm.ds_full, SOFT); cb::mbar_init(&sm.ds_empty, 1);
    cb::mbar_init(&sm.acc_full, 1);
    cb::fence_barrier_init();
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
      cb::mbar_arrive_expect_tx(&sm.q_full, 3 * TILE_BYTES);
      cb::tma_load_3d(sm.qu, &tm_qu, &sm.q_full, h * DH, b, i0);
      cb::tma_load_3d(sm.qv, &tm_qv, &sm.q_full, h * DH, b, i0);
      cb::tma_load_3d(sm.dout, &tm_do, &sm.q_full, h * DH, b, i0);
      auto load_r = [&](int beta) {   // buffer beta&1, its (beta>>1)-th use
        const int bi = beta & 1;
        const uint32_t use = (beta >> 1) & 1;
        cb::mbar_wait(&sm.r_empty[bi], use ^ 1);
        cb::mbar_arrive_expect_tx(&sm.r_full[bi], TILE_BYTES);
        cb::tma_load_2d(sm.r[bi], &tm_r, &sm.r_full[bi], h * DH, dbase - TN * beta);
      };
      load_r(0);
      load_r(1);
      uint32_t k_phase = 0;
      for (int t = 0; t < nt; ++t) {
        const int j0 = (jt_first + t) * TN;
        if (t > 0) load_r(t + 1);
        cb::mbar_wait(&sm.k_empty, k_phase ^ 1);
        cb::mbar_arrive_expect_tx(&sm.k_full, 2 * TILE_BYTES);
        cb::tma_load_3d(sm.k, &tm_k, &sm.k_full, h * DH, b, j0);
        cb::tma_load_3d(sm.v, &tm_v, &sm.k_full, h * DH, b, j0);
        k_phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one()) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);   // S, BD, dP
      const uint32_t idesc_q = cb::umma_idesc_bf16(TM, DH, 0, 1);   // dq: A K-major (TMEM / band tile), B MN-major
      uint32_t k_phase = 0, s_phase = 0, bd_phase = 0, ds_phase = 0;
      const uint32_t a_qu = cb::smem_u32(sm.qu), a_qv = cb::smem_u32(sm.qv), a_do = cb::smem_u32(sm.dout);
      const uint32_t a_k = cb::smem_u32(sm.k), a_v = cb::smem_u32(sm.v), a_dbd = cb::smem_u32(sm.dbd);
      cb::mbar_wait(&sm.q_full, 0);
      auto issue_bd = [&](int bi) {
        cb::mbar_wait(&sm.bd_empty, bd_phase ^ 1);
        cb::tc_fence_after();
        const uint64_t ad = cb::umma_smem_desc(a_qv, 16, 1024);
        const uint64_t bd = cb::umma_smem_desc(cb::smem_u32(sm.r[bi]), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + COL_BD, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
        cb::umma_commit(&sm.bd_full);
        bd_phase ^= 1;
      };
      for (int t = 0; t < nt; ++t) {
        const int lo_b = (t + 1) & 1, hi_b = t & 1;
        cb::mbar_wait(&sm.k_full, k_phase);
        cb::mbar_wait(&sm.s_empty, s_phase ^ 1);
        cb::tc_fence_after();
        {
          const uint64_t aq = cb::umma_smem_desc(a_qu, 16, 1024), bk = cb::umma_smem_desc(a_k, 16, 1024);
          const uint64_t ad = cb::umma_smem_desc(a_do, 16, 1024), bv = cb::umma_smem_desc(a_v, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + COL_S, aq + 2 * k, bk + 2 * k, idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) cb::umma_bf16_ss(tmem + COL_DP, ad + 2 * k, bv + 2 * k, idesc_s, k > 0);
          cb::umma_commit(&sm.s_full);
        }
        // beta = t+1 is the ((t+1)>>1)-th use of buffer lo_b; beta = t the (t>>1)-th use of hi_b
        cb::mbar_wait(&sm.r_full[lo_b], ((t + 1) >> 1) & 1);
        issue_bd(lo_b);
        cb::mbar_wait(&sm.r_full[hi_b], (t >> 1) & 1);
        issue_bd(hi_b);
        // dq products
        cb::mbar_wait(&sm.ds_full, ds_phase);
        cb::tc_fence_after();
        {
          const uint64_t bk = cb::umma_smem_desc(a_k, 8192, 1024);
#pragma unroll
          for (int k = 0; k < TN / 16; ++k)
            umma_bf16_ts(tmem + COL_DQA, tmem + COL_BD + 8 * k, bk + (uint64_t)(k * 128), idesc_q, (t > 0 || k > 0));
          const uint64_t blo = cb::umma_smem_desc(cb::smem_u32(sm.r[lo_b]), 8192, 1024);
          const uint64_t bhi = cb::umma_smem_desc(cb::smem_u32(sm.r[hi_b]), 8192, 1024);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t ad = cb::umma_smem_desc(a_dbd + (k >> 2) * TILE_BYTES, 16, 1024) + 2 * (k & 3);
            const uint64_t bd = (k < 8 ? blo : bhi) + (uint64_t)((k & 7) * 128);
            cb::umma_bf16_ss(tmem + COL_DQB, ad, bd, idesc_q, (t > 0 || k > 0));
          }
          cb::umma_commit(&sm.k_empty);
          cb::umma_commit(&sm.r_empty[hi_b]);
          cb::umma_commit(&sm.ds_empty);
        }
        k_phase ^= 1;
        s_phase ^= 1;
        ds_phase ^= 1;
      }
      cb::umma_commit(&sm.acc_full);
    }
  } else if (warp >= 4) {
    // ============================== softmax warpgroups ==============================
    const int g = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;
    const int li = wq * 32 + lane;
    const int i = i0 + li;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t my_row = cb::smem_u32(sm.bd) + li * STAGE_ROW;
    const uint32_t dbd_base = cb::smem_u32(sm.dbd);
    const float sl2 = p.scale * 1.4426950408889634f;
    uint32_t s_phase = 0, bd_phase = 0, ds_phase = 0;
    const float lse2 = i < p.T ? p.lse[((long long)b * p.H + h) * p.T + i] * 1.4426950408889634f : 0.f;
    const float delta = i < p.T ? p.delta[((long long)b * p.H + h) * p.T + i] : 0.f;
    const int hi_i = i < p.T ? i + p.M : -1;
    const int lo_i = key_lo(i, p.M, p.same_length, p.shift, reset);
    // the band tile starts as zeros; each row only ever rewrites its own run of 128 band columns
    for (int idx = threadIdx.x - 128; idx < 4 * TILE_BYTES / 16; idx += SOFT) sts_v4(dbd_base + idx * 16, 0, 0, 0, 0);

    for (int t = 0; t < nt; ++t) {
      cb::mbar_wait(&sm.s_full, s_phase);
      cb::tc_fence_after();
      float s[32];
      {
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_S + g * 32, r0);
        cb::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) s[e] = __uint_as_float(r0[e]);
      }
      cb::mbar_wait(&sm.bd_full, bd_phase);
      cb::tc_fence_after();
      named_bar(1, SOFT);
      stage32(lane_addr + COL_BD + g * 32, my_row + g * 64);
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.bd_empty);
      bd_phase ^= 1;
      named_bar(2, SOFT);
      band_add<0, true>(s, my_row, li, g, wq);
      cb::mbar_wait(&sm.bd_full, bd_phase);
      cb::tc_fence_after();
      named_bar(1, SOFT);
      stage32(lane_addr + COL_BD + g * 32, my_row + g * 64);
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.bd_empty);
      bd_phase ^= 1;
      named_bar(2, SOFT);     // also: every thread of the row has finished reading BD from TMEM (dS aliases it)
      band_add<1, true>(s, my_row, li, g, wq);
      const int jc0 = (jt_first + t) * TN + g * 32;
      const bool full = (jc0 + 31 <= hi_i) && (jc0 >= lo_i);
      float ds[32];
      {
        uint32_t r0[32];
        cb::tmem_ld_32x32b_x32(lane_addr + COL_DP + g * 32, r0);
        cb::tmem_ld_wait();
        cb::tc_fence_before();
        cb::mbar_arrive(&sm.s_empty);
        s_phase ^= 1;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float pv = ex2(fmaf(s[e], sl2, -lse2));
          if (!full) {
            const int j = jc0 + e;
            if (j > hi_i || j < lo_i) pv = 0.f;
          }
          ds[e] = pv * (__uint_as_float(r0[e]) - delta);
        }
      }
      // ---- dS -> TMEM (A operand of dq_ac) and, inverse-shifted, -> the band tile (A operand of dq_bd) ----
      cb::mbar_wait(&sm.ds_empty, ds_phase ^ 1);
      cb::tc_fence_after();
      {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) pk[e / 2] = cb::pack_bf16(ds[e], ds[e + 1]);
        tmem_st_32x32b_x16(lane_addr + COL_BD + g * 16, pk);
        const int c0 = li + (TN - 1) - g * 32;       // band column of this thread's first key
        const uint32_t rowb = dbd_base + li * 128;
        const int sw = li & 7;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int c = c0 - e;                      // 0 .. 254
          const uint32_t addr = rowb + (c >> 6) * TILE_BYTES + ((((c & 63) >> 3) ^ sw) << 4) + (c & 7) * 2;
          const unsigned short hv = __bfloat16_as_ushort(__float2bfloat16_rn(ds[e]));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(hv) : "memory");
        }
        tmem_st_wait();
      }
      cb::fence_proxy_async();
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.ds_full);
      ds_phase ^= 1;
    }
    // ---- epilogue ----
    cb::mbar_wait(&sm.acc_full, 0);
    cb::tc_fence_after();
    uint32_t ra[16], rb[16];
    tmem_ld_32x32b_x16(lane_addr + COL_DQA + g * 16, ra);
    tmem_ld_32x32b_x16(lane_addr + COL_DQB + g * 16, rb);
    cb::tmem_ld_wait();
    float fa[16], fb[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      fa[e] = __uint_as_float(ra[e]) * p.scale;
      fb[e] = __uint_as_float(rb[e]) * p.scale;
    }
    if (i < p.T) {
      bf16* dqr = p.dq + ((long long)i * p.B + b) * p.lddq + h * DH + g * 16;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint4 q;
        q.x = cb::pack_bf16(fa[ch * 8 + 0] + fb[ch * 8 + 0], fa[ch * 8 + 1] + fb[ch * 8 + 1]);
        q.y = cb::pack_bf16(fa[ch * 8 + 2] + fb[ch * 8 + 2], fa[ch * 8 + 3] + fb[ch * 8 + 3]);
        q.z = cb::pack_bf16(fa[ch * 8 + 4] + fb[ch * 8 + 4], fa[ch * 8 + 5] + fb[ch * 8 + 5]);
        q.w = cb::pack_bf16(fa[ch * 8 + 6] + fb[ch * 8 + 6], fa[ch * 8 + 7] + fb[ch * 8 + 7]);
        *reinterpret_cast<uint4*>(dqr + ch * 8) = q;
      }
    }
    // column sums over the 32 rows of this warp -> d r_w_bias / d r_r_bias (rows >= T hold exact zeros)
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float sa = cb::warp_sum(fa[e]);
      const float sb = cb::warp_sum(fb[e]);
      if (lane == 0) {
        atomicAdd(p.du + h * DH + g * 16 + e, sa);
        atomicAdd(p.dvb + h * DH + g * 16 + e, sb);
      }
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

// dq / du / dvb of commu_relattn_bwd on tcgen05 (same operand contract).  delta = rowsum(dO * O) [B,H,T].
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
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(T, TM), H, B);
  relattn_bwd_dq_tc_kernel<<<grid, NTHREADS, smem_bytes, stream>>>(tk, tv, tqu, tqv, tdo, tr, p);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
