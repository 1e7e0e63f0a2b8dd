// Relative-position attention backward with the score gradient MATERIALISED once (product path of commu_relattn_bwd).
//
// The three recompute passes (attn_bwd_{dq_,,dr_}tc.cu) each rebuilt S, both position blocks, exp2 and the dropout
// mask (76 thread instructions per score element).  Here every score element is touched once:
//
//   forward (attn_fwd_tc.cu)  stores P~ = bf16(exp2(s - m_tile)) (sign bit = "dropped") and m_tile per (row, key tile)
//   pass 1  (key-tile CTA)    P = P~ * exp2(m_tile - LSE) ; dP = dO V^T ; dS = P (dP - Delta)
//                             dV += P^T dO ; dK += dS^T (q+u)   (TMEM accumulators, as in the dk/dv pass)
//                             dS (bf16) is written to HBM once, by TMA, in a COARSE-SHEARED layout
//   band GEMMs                dq_A = dS K            (rows i0..i0+127 of one (b,h), key tiles walked)
//                             dq_C = dB Rrev         (rows of one residue i mod 8, distance blocks walked)
//                             dR  += dB^T (q+v)      (one distance block, all (b, rows) walked)
//                             pure TMA + tcgen05.mma streams: no exp, no shear, no RNG, no score recompute.
//                             dq_A and dq_C share ONE ticket-ordered launch (the second view of the dS rows comes from
//                             L2); two ring stages and two CTAs per SM, so one CTA's epilogue overlaps the other's stream.
//
// The relative shift costs nothing in this layout.  Element (i, j) of one (b,h) lives at
//     row i, column  c = j + X - 8*(i >> 3)            (pitch P elements, X = Tpad)
// so an 8-row group is one contiguous TMA box {64 cols, 8 rows} at column j0 + X - 8*(i>>3) (what pass 1 stores and
// the dq_A GEMM loads back as the plain key-indexed tile), while the rows of ONE residue r = i & 7, read with a row
// stride of 8 rows, see column c' <-> key j = c' - X + 8a (a = i >> 3), i.e. distance
//     delta = i + M - j = r + M + X - c'        -- the same for every row of the residue class.
// A [128 rows of residue r] x [128 columns c'] tile of that view is therefore a plain GEMM operand against 128
// consecutive rows of the REVERSED distance table (Rrev[y] = R[Y0 - y], Y0 = M + X + 8, y = c' + 8 - r) - the
// reference's _rel_shift (commu/model/model.py:251-265) becomes a TMA stride.
// Untouched parts of the workspace must read as zero: the caller zeroes it once per (T, M, B, H); pass 1 rewrites
// every causal tile on every call (zeros where the mask hides a tile), so the invariant holds from call to call.
//
// Autograd counterpart of commu/model/model.py:312-345.
#include "api_common.h"
#include "attn_common.cuh"
#include "attn_tc_common.cuh"

namespace cb_host {
int check_attn_common(const attn::Params& p, const char* who);
}

namespace {
using attn::key_lo;
using namespace attn_tc;

constexpr int TM = 128, TN = 128, DH = 64;
constexpr int TILE16 = 16384, TILE32 = 32768;
constexpr float LOG2E = 1.4426950408889634f;

struct MatParams {
  int T, M, B, H, Tpad, Kp, P, X, nkt, kr;
  int same_length, shift;
  int ds4d;             // 1: the dS tile moves as one 4-D TMA box per key half (group stride 8P - 8), 0: 16 eight-row boxes
  const unsigned char* reset;
  float scale, drop_keep;
  const float* lse;     // [B,H,T]
  const float* delta;   // [B,H,T]
  const float* mt;      // [B*H, nkt, Tpad]
  bf16* dk;
  bf16* dv;
  long long lddkv;
  float* da;            // fp32 scratch [T*B, H*64]
  bf16* dq;
  long long lddq;
  float* du;
  float* dvb;
  float* dr;            // fp32 [kr, H*64]
  unsigned* sync;       // merged dq_A / dq_C launch: [0] ticket counter, [1 + bh * a_blocks + ab] finished dq_A tiles
};

// ================================================================================================================
// pass 1
// ================================================================================================================
constexpr int P1_SOFT = 512;
constexpr int P1_THREADS = 64 + P1_SOFT;   // warp 0: TMA loads + dS stores (+ TMEM alloc), warp 1: MMA issuer, warps 2-17: compute
constexpr int COL_DP = 0, COL_DV = 256, COL_DK = 320;

struct P1Smem {
  uint8_t v[TILE16];
  uint8_t pt[3][TILE32];     // P~ tile ring (HBM stream, 3 deep): [key half][128 q rows][128 B], 128B swizzle; rewritten in place as P
  uint8_t ds[TILE32];        // dS tile, same layout: MN-major operand of dK and the source of the TMA stores
  uint8_t dout[2][TILE16];   // dO / (q+u) tiles: small, L2-resident, 2 deep
  uint8_t qu[2][TILE16];
  uint64_t v_full, p_full[3], p_empty[3], q_full[2], q_empty[2], dp_full[2], dp_free[2], pds_full[2], ds_mma, ds_st, acc_full;
  uint32_t tmem_base;
};

template <bool DROP>
__global__ void __launch_bounds__(P1_THREADS, 1)
relattn_bwd_p1_kernel(const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_qu,
                      const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_p,
                      const __grid_constant__ CUtensorMap tm_ds, const __grid_constant__ CUtensorMap tm_ds4,
                      const MatParams p) {
  extern __shared__ uint8_t smem_raw[];
  P1Smem& sm = *reinterpret_cast<P1Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int jt = blockIdx.x;
  const int j0 = jt * TN;
  const int bh = b * p.H + h;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;
  // every query tile that can see a key of this tile under the CAUSAL bound (the workspace invariant); tiles the
  // actual mask hides (reset / same_length) were skipped by the forward: they are processed with P = 0
  const int it_first = max(0, j0 - p.M) / TM;
  const int nq = (p.T - 1) / TM - it_first + 1;

  if (threadIdx.x == 0) {
    cb::mbar_init(&sm.v_full, 1);
    for (int s = 0; s < 3; ++s) { cb::mbar_init(&sm.p_full[s], 1); cb::mbar_init(&sm.p_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      cb::mbar_init(&sm.q_full[s], 1); cb::mbar_init(&sm.q_empty[s], 1);
      cb::mbar_init(&sm.dp_full[s], 1); cb::mbar_init(&sm.dp_free[s], P1_SOFT);
      cb::mbar_init(&sm.pds_full[s], P1_SOFT);
    }
    cb::mbar_init(&sm.ds_mma, 1); cb::mbar_init(&sm.ds_st, 1);
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
    // ============================== TMA loads and dS stores ==============================
    if (cb::elect_one()) {
      cb::mbar_arrive_expect_tx(&sm.v_full, TILE16);
      cb::tma_load_3d(sm.v, &tm_v, &sm.v_full, h * DH, b, j0);
      auto load_p = [&](int n) {   // P~ ring: stage n % 3, its (n / 3)-th use
        const int st = n % 3;
        const int i0 = (it_first + n) * TM;
        cb::mbar_wait(&sm.p_empty[st], ((n / 3) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.p_full[st], TILE32);
        cb::tma_load_2d(sm.pt[st], &tm_p, &sm.p_full[st], j0, bh * p.Tpad + i0);
        cb::tma_load_2d(sm.pt[st] + TILE16, &tm_p, &sm.p_full[st], j0 + 64, bh * p.Tpad + i0);
      };
      auto load_q = [&](int n) {   // dO / (q+u) ring: stage n & 1, its (n >> 1)-th use
        const int bi = n & 1;
        const int i0 = (it_first + n) * TM;
        cb::mbar_wait(&sm.q_empty[bi], ((n >> 1) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.q_full[bi], 2 * TILE16);
        cb::tma_load_3d(sm.dout[bi], &tm_do, &sm.q_full[bi], h * DH, b, i0);
        cb::tma_load_3d(sm.qu[bi], &tm_qu, &sm.q_full[bi], h * DH, b, i0);
      };
      // the P~ stream comes from HBM: besides the 3-deep ring, the tiles a few steps further ahead are pulled into L2
      auto prefetch_p = [&](int n) {
        const int i0 = (it_first + n) * TM;
        cb::tma_prefetch_2d(&tm_p, j0, bh * p.Tpad + i0);
        cb::tma_prefetch_2d(&tm_p, j0 + 64, bh * p.Tpad + i0);
      };
      load_p(0);
      load_q(0);
      if (nq > 1) load_p(1);
      for (int n = 3; n < 6 && n < nq; ++n) prefetch_p(n);
      for (int n = 0; n < nq; ++n) {
        if (n + 2 < nq) load_p(n + 2);
        if (n + 1 < nq) load_q(n + 1);
        if (n + 6 < nq) prefetch_p(n + 6);
        const int i0 = (it_first + n) * TM;
        cb::mbar_wait(&sm.pds_full[n & 1], (n >> 1) & 1);
        // coarse-sheared store: the 8-row group g of the tile goes to column j0 + X - (i0 + 8g) - as ONE 4-D box per
        // key half when the driver accepted the group stride 8P - 8 (36 -> 6 TMA operations per tile), else 16 boxes
        if (p.ds4d) {
          cb::tma_store_4d(&tm_ds4, sm.ds, j0 + p.X, 0, i0 >> 3, bh);
          cb::tma_store_4d(&tm_ds4, sm.ds + TILE16, j0 + 64 + p.X, 0, i0 >> 3, bh);
        } else {
#pragma unroll 1
          for (int half = 0; half < 2; ++half)
#pragma unroll 4
            for (int g = 0; g < 16; ++g)
              cb::tma_store_2d(&tm_ds, sm.ds + half * TILE16 + g * 1024, j0 + 64 * half + p.X - (i0 + 8 * g),
                               bh * p.Tpad + i0 + 8 * g);
        }
        cb::tma_store_commit();
        cb::tma_store_wait_read<0>();       // the single dS tile may be rewritten (the next tile needs it ~1000 cycles later)
        cb::mbar_arrive(&sm.ds_st);
      }
      cb::tma_store_wait<0>();
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one()) {
      const uint32_t idesc_s = cb::umma_idesc_bf16(TM, TN, 0, 0);   // dP = dO V^T: K-major x K-major
      const uint32_t idesc_g = cb::umma_idesc_bf16(TN, DH, 1, 1);   // dV, dK: MN-major A (tile^T), MN-major B
      cb::mbar_wait(&sm.v_full, 0);
      const uint32_t a_v = cb::smem_u32(sm.v);
      auto back = [&](int m) {   // dV += P^T dO ; dK += dS^T (q+u) of tile m
        const int bj = m & 1, sp = m % 3;
        cb::mbar_wait(&sm.pds_full[bj], (m >> 1) & 1);
        cb::tc_fence_after();
        // MN-major A: 64-key atoms 16 KB apart (LBO), 8-query-row groups 1 KB apart (SBO); 16 query rows per MMA
        const uint64_t ap = cb::umma_smem_desc(cb::smem_u32(sm.pt[sp]), TILE16, 1024);
        const uint64_t as = cb::umma_smem_desc(cb::smem_u32(sm.ds), TILE16, 1024);
        const uint64_t bo = cb::umma_smem_desc(cb::smem_u32(sm.dout[bj]), 8192, 1024);
        const uint64_t bq = cb::umma_smem_desc(cb::smem_u32(sm.qu[bj]), 8192, 1024);
#pragma unroll
        for (int k = 0; k < TM / 16; ++k)
          cb::umma_bf16_ss(tmem + COL_DV, ap + (uint64_t)(k * 128), bo + (uint64_t)(k * 128), idesc_g, (m > 0 || k > 0));
        cb::umma_commit(&sm.p_empty[sp]);
#pragma unroll
        for (int k = 0; k < TM / 16; ++k)
          cb::umma_bf16_ss(tmem + COL_DK, as + (uint64_t)(k * 128), bq + (uint64_t)(k * 128), idesc_g, (m > 0 || k > 0));
        cb::umma_commit(&sm.q_empty[bj]);
        cb::umma_commit(&sm.ds_mma);
      };
      for (int n = 0; n < nq; ++n) {
        const int bi = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        cb::mbar_wait(&sm.q_full[bi], ph);
        cb::mbar_wait(&sm.dp_free[bi], ph ^ 1);
        cb::tc_fence_after();
        {
          const uint64_t ad = cb::umma_smem_desc(cb::smem_u32(sm.dout[bi]), 16, 1024);
          const uint64_t bd = cb::umma_smem_desc(a_v, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            cb::umma_bf16_ss(tmem + COL_DP + bi * TN, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
        }
        cb::umma_commit(&sm.dp_full[bi]);
        if (n >= 1) back(n - 1);
      }
      back(nq - 1);
      cb::umma_commit(&sm.acc_full);
    }
  } else {
    // ============================== compute threads ==============================
    // thread = (query row li of the tile, 32-key chunk g)
    const int g = (warp - 2) >> 2;
    const int wq = warp & 3;                     // TMEM lane quadrant of this warp (hardware: warp id % 4)
    const int li = wq * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    // this thread's 64 bytes of a tile row: key half g >> 1, 16-byte chunks 4*(g&1) .. +3 (128B swizzle)
    const uint32_t rowoff = (uint32_t)((g >> 1) * TILE16 + li * 128);
    const int cx = (g & 1) * 4;
    const int sw = li & 7;
    const float* lse_p = p.lse + (long long)bh * p.T;
    const float* del_p = p.delta + (long long)bh * p.T;
    const float* mt_p = p.mt + ((long long)bh * p.nkt + jt) * p.Tpad;
    const float lkeep = DROP ? log2f(p.drop_keep) : 0.f;
    const uint32_t a_ds = cb::smem_u32(sm.ds) + rowoff;
    // per-row constants of the NEXT TWO tiles are always in flight: under the bulk traffic of this kernel a plain
    // global load takes longer than one tile (ncu: 19 % of the stall samples sat on the first use with a one-tile lead)
    float lse_q[2] = {0.f, 0.f}, del_q[2] = {0.f, 0.f}, mt_q[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = (it_first + u) * TM + li;
      if (u < nq && i < p.T) { lse_q[u] = lse_p[i]; del_q[u] = del_p[i]; mt_q[u] = mt_p[i]; }
    }
    for (int n = 0; n < nq; ++n) {
      const int i0 = (it_first + n) * TM;
      const int i = i0 + li;
      // P = P~ * exp2(m_tile - LSE) ; under dropout the kept P is scaled by 1 / keep (folded into the exponent) and
      // Delta by keep:  dS = (P / keep) (mask dP - keep Delta)
      // (a tile in which this row saw no visible key yet was taken at m_tile = 0 with P~ = 0: clamp, so that a very
      // negative LSE cannot turn 0 * exp2(-LSE) into NaN; tiles the forward skipped carry no m_tile at all)
      const bool vis = jt >= key_lo(i0, p.M, p.same_length, p.shift, reset) / TN;   // did the forward visit this tile?
      const float lse_n = (n & 1) ? lse_q[1] : lse_q[0], del_n = (n & 1) ? del_q[1] : del_q[0];
      const float mt_n = (n & 1) ? mt_q[1] : mt_q[0];
      const float f = vis ? ex2(fminf(mt_n - lse_n * LOG2E, 0.f) - lkeep) : 0.f;
      const float ndelta = -(DROP ? del_n * p.drop_keep : del_n);
      if (g == 0 && lane == 0) {   // ... and the lines of the tile five ahead are pulled into L2 (one 128-byte line per warp)
        const int i5 = i + 5 * TM;
        if (n + 5 < nq && i5 < p.T) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(lse_p + i5));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(del_p + i5));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(mt_p + i5));
        }
      }
      {
        const int i2 = i + 2 * TM;
        float a = 0.f, bq = 0.f, c = 0.f;
        if (n + 2 < nq && i2 < p.T) { a = lse_p[i2]; bq = del_p[i2]; c = mt_p[i2]; }
        if (n & 1) { lse_q[1] = a; del_q[1] = bq; mt_q[1] = c; } else { lse_q[0] = a; del_q[0] = bq; mt_q[0] = c; }
      }
      const int bi = n & 1;
      const uint32_t ph = (n >> 1) & 1;
      const uint32_t a_pt = cb::smem_u32(sm.pt[n % 3]) + rowoff;
      cb::mbar_wait(&sm.p_full[n % 3], (n / 3) & 1);
      // this thread's 32 stored probabilities (64 bytes) and its 32 dP columns, all in flight before the first use
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t a = a_pt + (((cx + c) ^ sw) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[c * 4]), "=r"(w[c * 4 + 1]), "=r"(w[c * 4 + 2]), "=r"(w[c * 4 + 3]) : "r"(a));
      }
      cb::mbar_wait(&sm.dp_full[bi], ph);
      cb::tc_fence_after();
      uint32_t dp[32];
      cb::tmem_ld_32x32b_x32(lane_addr + COL_DP + bi * TN + g * 32, dp);
      cb::tmem_ld_wait();
      cb::tc_fence_before();
      cb::mbar_arrive(&sm.dp_free[bi]);
      uint32_t pk[16], dsk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const uint32_t ww = vis ? w[e] : 0u;       // a tile the forward skipped (reset / same_length): P = dS = 0
        const float x0 = cb::bf16_lo(ww) * f, x1 = cb::bf16_hi(ww) * f;     // signed: negative = dropped
        const float d0 = __uint_as_float(dp[2 * e]), d1 = __uint_as_float(dp[2 * e + 1]);
        float k0, k1, s0, s1;
        if (DROP) {
          k0 = fmaxf(x0, 0.f); k1 = fmaxf(x1, 0.f);
          s0 = fmaf(k0, d0, fabsf(x0) * ndelta);
          s1 = fmaf(k1, d1, fabsf(x1) * ndelta);
        } else {
          k0 = x0; k1 = x1;
          s0 = x0 * (d0 + ndelta);
          s1 = x1 * (d1 + ndelta);
        }
        pk[e] = cb::pack_bf16(k0, k1);
        dsk[e] = cb::pack_bf16(s0, s1);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts_v4(a_pt + (((cx + c) ^ sw) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      if (n >= 1) {   // the single dS tile: the previous tile's dK product and TMA stores are done with it
        cb::mbar_wait(&sm.ds_mma, (n - 1) & 1);
        cb::mbar_wait(&sm.ds_st, (n - 1) & 1);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts_v4(a_ds + (((cx + c) ^ sw) << 4), dsk[c * 4], dsk[c * 4 + 1], dsk[c * 4 + 2], dsk[c * 4 + 3]);
      cb::fence_proxy_async();
      cb::mbar_arrive(&sm.pds_full[bi]);
    }
    // ---- epilogue: dV, dK rows (thread = key row li, 16 of the 64 head dims) ----
    const int j = j0 + li;
    cb::mbar_wait(&sm.acc_full, 0);
    cb::tc_fence_after();
    uint32_t rv[16], rk[16];
    tmem_ld_32x32b_x16(lane_addr + COL_DV + g * 16, rv);
    tmem_ld_32x32b_x16(lane_addr + COL_DK + g * 16, rk);
    cb::tmem_ld_wait();
    if (j < Ktot) {
      bf16* dvr = p.dv + ((long long)j * p.B + b) * p.lddkv + h * DH + g * 16;
      bf16* dkr = p.dk + ((long long)j * p.B + b) * p.lddkv + h * DH + g * 16;
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
  }
  cb::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    cb::tc_fence_after();
    cb::tmem_dealloc(tmem, 512);
  }
}

// ================================================================================================================
// band GEMMs over the materialised dS
// ================================================================================================================
constexpr int MODE_A = 0, MODE_C = 1, MODE_R = 2;
constexpr int G_THREADS = 64 + 128;   // warp 0: TMA producer (+ TMEM alloc), warp 1: MMA issuer, warps 2-5: epilogue
constexpr int G_STAGES = 2;   // x 2 CTAs per SM: one CTA's epilogue overlaps the other's stream (4 stages x 1 CTA: same bytes in flight)

struct GSmem {
  uint8_t a[G_STAGES][TILE32];   // dS tile: [column half][128 rows][128 B], 128B swizzle
  uint8_t b[G_STAGES][TILE16];   // K / Rrev / (q+v) tile: [128 rows][64 dims]
  float csum[4][64];
  uint64_t full[G_STAGES], empty[G_STAGES], acc_full;
  uint32_t tmem_base;
  uint32_t ticket;
};

// One CTA's share of a band GEMM.  x / nx / h / b are what blockIdx.x / gridDim.x / blockIdx.y / blockIdx.z are in the
// one-kernel-per-GEMM launches; the flags carry the dq_A -> dq_C hand-over of the merged launch (below).
struct BandWork {
  int x, nx, h, b;
  const unsigned* wait_flag;   // MODE_C: the fp32 dq_A rows of this CTA are complete once *wait_flag == wait_need
  unsigned wait_need;
  unsigned* signal_flag;       // MODE_A: incremented once this CTA's dq_A rows are visible
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

struct Maps8 {
  CUtensorMap m[8];
};

// number of 128-row blocks of the residue-class view (rows a = i >> 3)
__host__ __device__ inline int a_blocks(int Tpad) { return (Tpad / 8 + 127) / 128; }

template <int MODE>
__device__ __forceinline__ void band_body(GSmem& sm,
                                          const CUtensorMap& tm_a,    // MODE_A: 2-D dS view, box {64, 8}; else the residue view, box {64,1,128,1}
                                          const CUtensorMap& tm_a4,   // MODE_A with p.ds4d: 4-D dS view, box {64, 8, 16, 1}
                                          const CUtensorMap& tm_b,    // MODE_A: K rows3d; MODE_C: Rrev 2-D; MODE_R: unused
                                          const CUtensorMap* tm_qv,   // MODE_R: (q+v) rows of this CTA's residue
                                          const MatParams& p, const BandWork w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nab = a_blocks(p.Tpad);
  const int amax = p.Tpad / 8 - 1;
  // ---- work decomposition ----
  int h = w.h, b = 0, bh = 0, i0 = 0, r = 0, a0 = 0, c0 = 0;
  int s_first = 0, nsteps = 0, a_first = 0, nper = 0;
  if (MODE == MODE_A) {
    b = w.b;
    bh = b * p.H + h;
    i0 = (w.nx - 1 - w.x) * TM;
    nsteps = (min(p.T - 1, i0 + TM - 1) + p.M) / TN + 1;          // every causal key tile (masked ones hold zeros)
  } else if (MODE == MODE_C) {
    b = w.b;
    bh = b * p.H + h;
    r = w.x & 7;
    a0 = (nab - 1 - (w.x >> 3)) * 128;
    const int lo = p.X - 8 * min(a0 + 127, amax);                 // first column any row of the block uses
    s_first = max(lo, 0) / 128;
    nsteps = (p.X + p.M + 7) / 128 - s_first + 1;                 // ... up to distance 0 of the largest residue
  } else {
    r = w.x & 7;
    c0 = (w.x >> 3) * 128;
    const int t = p.X - c0 - 127;                                  // rows a with X - 8a <= c0 + 127 hold data here
    a_first = t > 0 ? (t + 7) / 8 : 0;                            // first row a (NOT a block index): the sweep starts
    const int rows = p.Tpad / 8;                                   // there, rows past the end are TMA zero fill (no traffic)
    nper = a_first < rows ? (rows - a_first + 127) / 128 : 0;
    nsteps = p.B * nper;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) { cb::mbar_init(&sm.full[s], 1); cb::mbar_init(&sm.empty[s], 1); }
    cb::mbar_init(&sm.acc_full, 1);
    cb::fence_barrier_init();
  }
  if (warp == 0) {
    cb::tmem_alloc(&sm.tmem_base, 64);
    cb::tmem_relinquish();
  }
  cb::tc_fence_before();
  __syncthreads();
  cb::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (cb::elect_one()) {
      for (int s = 0; s < nsteps; ++s) {
        const int st = s % G_STAGES;
        cb::mbar_wait(&sm.empty[st], ((s / G_STAGES) & 1) ^ 1);
        cb::mbar_arrive_expect_tx(&sm.full[st], TILE32 + TILE16);
        if (MODE == MODE_A) {
          const int j0 = s * TN;
          if (p.ds4d) {
            cb::tma_load_4d(sm.a[st], &tm_a4, &sm.full[st], j0 + p.X, 0, i0 >> 3, bh);
            cb::tma_load_4d(sm.a[st] + TILE16, &tm_a4, &sm.full[st], j0 + 64 + p.X, 0, i0 >> 3, bh);
          } else {
#pragma unroll 1
            for (int half = 0; half < 2; ++half)
#pragma unroll 4
              for (int g = 0; g < 16; ++g)
                cb::tma_load_2d(sm.a[st] + half * TILE16 + g * 1024, &tm_a, &sm.full[st],
                                j0 + 64 * half + p.X - (i0 + 8 * g), bh * p.Tpad + i0 + 8 * g);
          }
          cb::tma_load_3d(sm.b[st], &tm_b, &sm.full[st], h * DH, b, j0);
        } else if (MODE == MODE_C) {
          const int cc = (s_first + s) * 128;
          cb::tma_load_4d(sm.a[st], &tm_a, &sm.full[st], cc, r, a0, bh);
          cb::tma_load_4d(sm.a[st] + TILE16, &tm_a, &sm.full[st], cc + 64, r, a0, bh);
          cb::tma_load_2d(sm.b[st], &tm_b, &sm.full[st], h * DH, cc + 8 - r);
        } else {
          const int bb = s / nper;
          const int aa = a_first + (s % nper) * 128;
          const int bhh = bb * p.H + h;
          cb::tma_load_4d(sm.a[st], &tm_a, &sm.full[st], c0, r, aa, bhh);
          cb::tma_load_4d(sm.a[st] + TILE16, &tm_a, &sm.full[st], c0 + 64, r, aa, bhh);
          cb::tma_load_3d(sm.b[st], tm_qv, &sm.full[st], h * DH, bb, aa);
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (cb::elect_one()) {
      // A / C: acc[row, d] += dS[row, col] B[col, d]   (A K-major, B MN-major)
      // R    : acc[col, d] += dS[row, col] (q+v)[row, d] (A MN-major = the same tile transposed, B MN-major)
      const uint32_t idesc = cb::umma_idesc_bf16(128, DH, MODE == MODE_R ? 1 : 0, 1);
      for (int s = 0; s < nsteps; ++s) {
        const int st = s % G_STAGES;
        cb::mbar_wait(&sm.full[st], (s / G_STAGES) & 1);
        cb::tc_fence_after();
        const uint32_t a_a = cb::smem_u32(sm.a[st]);
        const uint64_t bd = cb::umma_smem_desc(cb::smem_u32(sm.b[st]), 8192, 1024);
        if (MODE == MODE_R) {
          const uint64_t ad = cb::umma_smem_desc(a_a, TILE16, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            cb::umma_bf16_ss(tmem, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc, (s > 0 || k > 0));
        } else {
          const uint64_t ad0 = cb::umma_smem_desc(a_a, 16, 1024);
          const uint64_t ad1 = cb::umma_smem_desc(a_a + TILE16, 16, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            cb::umma_bf16_ss(tmem, (k < 4 ? ad0 : ad1) + (uint64_t)(2 * (k & 3)), bd + (uint64_t)(k * 128), idesc,
                             (s > 0 || k > 0));
        }
        cb::umma_commit(&sm.empty[st]);
      }
      cb::umma_commit(&sm.acc_full);
    }
  } else {
    // ============================== epilogue (thread = accumulator lane) ==============================
    const int wq = warp & 3;
    const int li = wq * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    float acc[64];
    if (nsteps > 0) {
      cb::mbar_wait(&sm.acc_full, 0);
      cb::tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t rr[16];
        tmem_ld_32x32b_x16(lane_addr + q * 16, rr);
        cb::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[q * 16 + e] = __uint_as_float(rr[e]) * p.scale;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 64; ++e) acc[e] = 0.f;
    }
    if (MODE == MODE_R) {
      const int delta = r + p.M + p.X - (c0 + li);
      if (nsteps > 0 && delta >= 0 && delta < p.kr) {
        float* dst = p.dr + (long long)delta * (p.H * DH) + h * DH;
#pragma unroll
        for (int e = 0; e < 64; e += 4) cb::red_add_v4(dst + e, acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
      }
    } else {
      const int i = MODE == MODE_A ? i0 + li : 8 * (a0 + li) + r;
      if (MODE == MODE_C && w.wait_flag) {   // merged launch: the dq_A tiles of these rows come from CTAs that started earlier
        if (threadIdx.x == 64)
          while (ld_acquire_gpu(w.wait_flag) < w.wait_need) __nanosleep(64);
        named_bar(1, 128);
      }
      if (i < p.T) {
        float* da = p.da + ((long long)i * p.B + b) * (p.H * DH) + h * DH;
        if (MODE == MODE_A) {
#pragma unroll
          for (int e = 0; e < 64; e += 4) *reinterpret_cast<float4*>(da + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
        } else {
          bf16* dq = p.dq + ((long long)i * p.B + b) * p.lddq + h * DH;
#pragma unroll
          for (int e = 0; e < 64; e += 8) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(da + e));       // (written by another CTA of this launch
            const float4 y = __ldcg(reinterpret_cast<const float4*>(da + e + 4));   //  in the merged form: never through L1)
            uint4 o;
            o.x = cb::pack_bf16(acc[e] + x.x, acc[e + 1] + x.y);
            o.y = cb::pack_bf16(acc[e + 2] + x.z, acc[e + 3] + x.w);
            o.z = cb::pack_bf16(acc[e + 4] + y.x, acc[e + 5] + y.y);
            o.w = cb::pack_bf16(acc[e + 6] + y.z, acc[e + 7] + y.w);
            *reinterpret_cast<uint4*>(dq + e) = o;
          }
        }
      }
      if (MODE == MODE_A && w.signal_flag) __threadfence();
      // column sums: d r_w_bias (MODE_A) / d r_r_bias (MODE_C); rows >= T hold exact zeros
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        const float sa = cb::warp_sum(acc[e]);
        if (lane == 0) sm.csum[warp - 2][e] = sa;
      }
      named_bar(1, 128);
      const int t = threadIdx.x - 64;
      if (t < 64) {
        const float sa = (sm.csum[0][t] + sm.csum[1][t]) + (sm.csum[2][t] + sm.csum[3][t]);
        atomicAdd((MODE == MODE_A ? p.du : p.dvb) + h * DH + t, sa);
      }
      if (MODE == MODE_A && w.signal_flag && t == 64) atomicAdd(w.signal_flag, 1u);   // (after the bar: every row is fenced)
    }
  }
  cb::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    cb::tc_fence_after();
    cb::tmem_dealloc(tmem, 64);
  }
}

template <int MODE>
__global__ void __launch_bounds__(G_THREADS, 2)
relattn_bwd_band_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a4,
                        const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ Maps8 tm_qv,
                        const MatParams p) {
  extern __shared__ uint8_t smem_raw[];
  GSmem& sm = *reinterpret_cast<GSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const BandWork w = {(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, (int)blockIdx.z, nullptr, 0u, nullptr};
  band_body<MODE>(sm, tm_a, tm_a4, tm_b, MODE == MODE_R ? &tm_qv.m[blockIdx.x & 7] : nullptr, p, w);
}

// dq_A and dq_C in ONE launch, ordered so that the two views of the same dS rows are read back to back: both GEMMs
// stream the whole workspace (1.6 GB per layer at the benchmark shape), and as separate launches each streams it from
// HBM.  Here a CTA takes a ticket and derives its tile from it: per (b,h), per block of 1024 query rows (from the last
// block down, heaviest first), the <= 8 key-indexed row tiles (dq_A) and then the 8 residue classes of the same rows
// (dq_C) - the second view finds the rows in L2.  The ticket (not blockIdx) fixes the order, so the one dependency -
// dq_C adds the fp32 dq_A rows and writes the bf16 result - only ever waits on CTAs that already run.
__global__ void __launch_bounds__(G_THREADS, 2)
relattn_bwd_band_ac_kernel(const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_a4,
                           const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_k,
                           const __grid_constant__ CUtensorMap tm_rr, const MatParams p) {
  extern __shared__ uint8_t smem_raw[];
  GSmem& sm = *reinterpret_cast<GSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) sm.ticket = atomicAdd(p.sync, 1u);
  __syncthreads();
  const int nab = a_blocks(p.Tpad);
  const int nta = p.Tpad / TM;
  const int per_bh = nta + 8 * nab;
  const int bh = (int)(sm.ticket / (unsigned)per_bh);
  int u = (int)(sm.ticket % (unsigned)per_bh);
  int ab = nab - 1, na = 0;
  bool is_a = false;
  for (; ab >= 0; --ab) {
    na = min(8, nta - 8 * ab);
    if (u < na) { is_a = true; break; }
    u -= na;
    if (u < 8) break;
    u -= 8;
  }
  unsigned* flag = p.sync + 1 + bh * nab + ab;
  BandWork w;
  w.h = bh % p.H;
  w.b = bh / p.H;
  if (is_a) {
    w.x = nta - 1 - (8 * ab + na - 1 - u);      // row tile 8ab + na-1-u, in band_body's reversed numbering
    w.nx = nta;
    w.wait_flag = nullptr; w.wait_need = 0u; w.signal_flag = flag;
    band_body<MODE_A>(sm, tm_a2, tm_a4, tm_k, nullptr, p, w);
  } else {
    w.x = (nab - 1 - ab) * 8 + u;               // residue u of row block ab
    w.nx = nab * 8;
    w.wait_flag = flag; w.wait_need = (unsigned)na; w.signal_flag = nullptr;
    band_body<MODE_C>(sm, tm_res, tm_a4, tm_rr, nullptr, p, w);
  }
}

// Rrev[y, :] = R[Y0 - y, :] (zero outside [0, kr)), 16 bytes per thread
__global__ void rrev_kernel(const bf16* __restrict__ r, long long ldr, int kr, int Y0, int rows, int cols8,
                            bf16* __restrict__ out, unsigned* __restrict__ sync, int nsync) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long j = idx; j < nsync; j += (long long)gridDim.x * blockDim.x) sync[j] = 0u;   // ticket + hand-over flags
  if (idx >= (long long)rows * cols8) return;
  const int y = (int)(idx / cols8), c = (int)(idx % cols8);
  const int d = Y0 - y;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (d >= 0 && d < kr) v = *reinterpret_cast<const uint4*>(r + (long long)d * ldr + c * 8);
  *reinterpret_cast<uint4*>(out + (long long)y * (cols8 * 8) + c * 8) = v;
}

// generic bf16 tiled tensor map (128B swizzle)
int make_tmap(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr_fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr_fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return cb_host::fail(COMMU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(ptr_fn);
  }
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, s, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    return cb_host::fail(COMMU_ERR_CUDA, "cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)rc);
  return 0;
}

struct Geo {
  int Tpad, Kp, P, X, nkt, Y0, rrev_rows;
  int64_t ds_bytes, rrev_bytes, da_bytes, sync_bytes, p_bytes, mt_bytes;
  int nsync;
};
Geo geometry(int T, int M, int B, int H) {
  Geo g;
  g.Tpad = (T + 127) / 128 * 128;
  g.Kp = (T + M + 127) / 128 * 128;
  g.X = g.Tpad;
  g.P = g.Kp + g.Tpad + 64;                       // columns [X - 8a, X - 8a + Kp) of every row, 8 spare on either side
  g.nkt = g.Kp / 128;
  g.Y0 = M + g.X + 8;
  g.rrev_rows = g.Y0 + 1;
  g.ds_bytes = (int64_t)B * H * g.Tpad * g.P * 2;
  g.rrev_bytes = ((int64_t)g.rrev_rows * H * 64 * 2 + 1023) / 1024 * 1024;
  g.da_bytes = ((int64_t)T * B * H * 64 * 4 + 1023) / 1024 * 1024;
  g.nsync = 1 + B * H * a_blocks(g.Tpad);
  g.sync_bytes = ((int64_t)g.nsync * 4 + 1023) / 1024 * 1024;
  g.p_bytes = (int64_t)B * H * g.Tpad * g.Kp * 2;
  g.mt_bytes = (int64_t)B * H * g.nkt * g.Tpad * 4;
  return g;
}

}  // namespace

namespace attn_tc {
// layout of the stored probabilities, shared with the forward kernel's entry point (attn_fwd_tc.cu)
void psave_geometry(int T, int M, int* Tpad, int* Kp, int* nkt) {
  *Tpad = (T + 127) / 128 * 128;
  *Kp = (T + M + 127) / 128 * 128;
  *nkt = *Kp / 128;
}
}  // namespace attn_tc

// Sizes of the buffers of the materialised backward for one attention call of shape (T, M, B, H):
//   p_bytes / mt_bytes : the forward's p_save (bf16 [B*H, Tpad, Kp]) and mt_save (fp32 [B*H, Kp/128, Tpad])
//   ws_bytes           : workspace of commu_relattn_bwd; its first ws_zero_bytes must be zero when the workspace is
//                        first used with this shape (the kernels keep that part valid from call to call)
extern "C" int commu_relattn_bwd_sizes(int T, int M, int B, int H, int64_t* p_bytes, int64_t* mt_bytes,
                                       int64_t* ws_bytes, int64_t* ws_zero_bytes) {
  CB_REQUIRE(T > 0 && M >= 0 && B > 0 && H > 0, "relattn_bwd_sizes: bad shape");
  const Geo g = geometry(T, M, B, H);
  if (p_bytes) *p_bytes = g.p_bytes;
  if (mt_bytes) *mt_bytes = g.mt_bytes;
  if (ws_bytes) *ws_bytes = g.ds_bytes + g.rrev_bytes + g.da_bytes + g.sync_bytes;
  if (ws_zero_bytes) *ws_zero_bytes = g.ds_bytes;
  return 0;
}

// Materialised backward (see the header of this file).  delta = rowsum(dO * O) [B,H,T] must already be computed.
extern "C" int commu_relattn_bwd_mat(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                     int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset,
                                     int T, int M, int B, int H, int same_length, int shift, float scale,
                                     const float* lse, const void* dout, int64_t lddo, const float* delta,
                                     const void* p_save, const float* mt_save, void* ws, int64_t ws_bytes, void* dq,
                                     int64_t lddq, void* dk, void* dv, int64_t lddkv, float* dr, float* du,
                                     float* dvb, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CB_REQUIRE(qu && qv && k && v && r && lse && dout && delta && p_save && mt_save && ws && dq && dk && dv && dr && du && dvb,
             "relattn_bwd_mat: null arg");
  CB_REQUIRE(T > 0 && M >= 0 && B > 0 && H > 0 && kr >= T + M, "relattn_bwd_mat: bad shape");
  CB_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldr % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddkv % 8 == 0,
             "relattn_bwd_mat: leading dims must be multiples of 8");
  const Geo g = geometry(T, M, B, H);
  const int64_t ws_need = g.ds_bytes + g.rrev_bytes + g.da_bytes + g.sync_bytes;
  CB_REQUIRE(ws_bytes >= ws_need, "relattn_bwd_mat: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)ws_need);
  CB_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 127) == 0, "relattn_bwd_mat: workspace must be 128-byte aligned");
  bf16* ds = (bf16*)ws;
  bf16* rrev = (bf16*)((uint8_t*)ws + g.ds_bytes);
  float* da = (float*)((uint8_t*)ws + g.ds_bytes + g.rrev_bytes);
  unsigned* sync = (unsigned*)((uint8_t*)ws + g.ds_bytes + g.rrev_bytes + g.da_bytes);
  const int Ktot = T + M;
  const int BH = B * H;

  MatParams p = {};
  p.T = T; p.M = M; p.B = B; p.H = H; p.Tpad = g.Tpad; p.Kp = g.Kp; p.P = g.P; p.X = g.X; p.nkt = g.nkt; p.kr = kr;
  p.same_length = same_length; p.shift = shift; p.reset = reset; p.scale = scale;
  p.lse = lse; p.delta = delta; p.mt = mt_save;
  p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddkv = lddkv;
  p.da = da; p.sync = sync; p.dq = (bf16*)dq; p.lddq = lddq; p.du = du; p.dvb = dvb; p.dr = dr;
  const DropState dst = drop_state();
  const uint32_t thr = dst.p > 0.f ? drop::thr15_of(dst.p) : 0u;
  p.drop_keep = 1.f - (float)thr / 32768.f;

  int rc;
  CUtensorMap tk, tv, tqu, tdo, tp, tds2, tdsg, trr;
  static Maps8 tqv;   // 1 KB of kernel parameters, rebuilt per call (the launch copies it)
  if ((rc = make_tmap_rows3d(&tk, k, (uint64_t)H * 64, B, Ktot, ldkv))) return rc;
  if ((rc = make_tmap_rows3d(&tv, v, (uint64_t)H * 64, B, Ktot, ldkv))) return rc;
  if ((rc = make_tmap_rows3d(&tqu, qu, (uint64_t)H * 64, B, T, ldq))) return rc;
  if ((rc = make_tmap_rows3d(&tdo, dout, (uint64_t)H * 64, B, T, lddo))) return rc;
  {
    const uint64_t dims[2] = {(uint64_t)g.Kp, (uint64_t)BH * g.Tpad};
    const uint64_t str[1] = {(uint64_t)g.Kp * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_tmap(&tp, p_save, 2, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)g.P, (uint64_t)BH * g.Tpad};
    const uint64_t str[1] = {(uint64_t)g.P * 2};
    const uint32_t box[2] = {64, 8};
    if ((rc = make_tmap(&tds2, ds, 2, dims, str, box))) return rc;
  }
  CUtensorMap tds4 = tds2;
  {   // key-indexed view as one box per key half: {column, row in group, 8-row group (stride 8P - 8: the coarse shear), (b,h)}
    const uint64_t dims[4] = {(uint64_t)g.P, 8, (uint64_t)g.Tpad / 8, (uint64_t)BH};
    const uint64_t str[3] = {(uint64_t)g.P * 2, (uint64_t)(8 * g.P - 8) * 2, (uint64_t)g.Tpad * g.P * 2};
    const uint32_t box[4] = {64, 8, 16, 1};
    static int use4d = -1;        // COMMU_ATTN_DS4D=0 forces the 16-box path
    if (use4d < 0) {
      const char* e = getenv("COMMU_ATTN_DS4D");
      use4d = (e && e[0] == '0') ? 0 : 1;
    }
    p.ds4d = use4d && make_tmap(&tds4, ds, 4, dims, str, box) == 0;
    if (!p.ds4d) tds4 = tds2;     // (a driver that rejects the overlapping group stride: keep the 2-D boxes)
  }
  {   // residue view: {column, residue r, row block a, (b,h)}
    const uint64_t dims[4] = {(uint64_t)g.P, 8, (uint64_t)g.Tpad / 8, (uint64_t)BH};
    const uint64_t str[3] = {(uint64_t)g.P * 2, (uint64_t)g.P * 16, (uint64_t)g.Tpad * g.P * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    if ((rc = make_tmap(&tdsg, ds, 4, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)H * 64, (uint64_t)g.rrev_rows};
    const uint64_t str[1] = {(uint64_t)H * 64 * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_tmap(&trr, rrev, 2, dims, str, box))) return rc;
  }
  for (int rr = 0; rr < 8; ++rr) {   // (q+v) rows i = 8a + r: {column, batch, a}
    const bool any = rr < T;
    const bf16* base = (const bf16*)qv + (any ? (long long)rr * B * ldq : 0);
    const uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)B, (uint64_t)(any ? (T - rr + 7) / 8 : 1)};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)ldq * 2 * B * 8};
    const uint32_t box[3] = {64, 1, 128};
    if ((rc = make_tmap(&tqv.m[rr], base, 3, dims, str, box))) return rc;
  }

  static bool attr = false;
  const int p1_smem = (int)sizeof(P1Smem) + 1024, g_smem = (int)sizeof(GSmem) + 1024;
  if (!attr) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_p1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p1_smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_p1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p1_smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_band_kernel<MODE_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_band_kernel<MODE_C>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_band_kernel<MODE_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_band_ac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem));
    attr = true;
  }
  {
    const int cols8 = H * 8;
    const long long n = (long long)g.rrev_rows * cols8;
    rrev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const bf16*)r, ldr, kr, g.Y0, g.rrev_rows, cols8, rrev, sync, g.nsync);
  }
  {
    dim3 grid(cb_host::ceil_div(Ktot, TN), H, B);
    if (thr) relattn_bwd_p1_kernel<true><<<grid, P1_THREADS, p1_smem, stream>>>(tv, tqu, tdo, tp, tds2, tds4, p);
    else relattn_bwd_p1_kernel<false><<<grid, P1_THREADS, p1_smem, stream>>>(tv, tqu, tdo, tp, tds2, tds4, p);
  }
  const int nab = a_blocks(g.Tpad);
  static int merge = -1;          // COMMU_ATTN_BAND_MERGE=0: dq_A and dq_C as two launches (each streams dS from HBM)
  if (merge < 0) {
    const char* e = getenv("COMMU_ATTN_BAND_MERGE");
    merge = (e && e[0] == '0') ? 0 : 1;
  }
  if (merge) {
    relattn_bwd_band_ac_kernel<<<dim3((unsigned)(BH * (g.Tpad / TM + 8 * nab))), G_THREADS, g_smem, stream>>>(tds2, tds4, tdsg, tk, trr, p);
  } else {
    relattn_bwd_band_kernel<MODE_A><<<dim3(cb_host::ceil_div(T, TM), H, B), G_THREADS, g_smem, stream>>>(tds2, tds4, tk, tqv, p);
    relattn_bwd_band_kernel<MODE_C><<<dim3(nab * 8, H, B), G_THREADS, g_smem, stream>>>(tdsg, tds4, trr, tqv, p);
  }
  const int ncb = (g.X + M + 7) / 128 + 1;
  relattn_bwd_band_kernel<MODE_R><<<dim3(ncb * 8, H, 1), G_THREADS, g_smem, stream>>>(tdsg, tds4, trr, tqv, p);
  cb_host::count_launch(merge ? 4 : 5);
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
