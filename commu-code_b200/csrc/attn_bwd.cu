// Relative-position attention backward, v1 (warp-level bf16 tensor-core MMA, recompute-based,
// atomic-free on the large tensors).  Three passes share one tile routine that recomputes
// P = softmax probabilities and dS = P * (dP - Delta) * scale for a 16-row warp slice:
//
//   dq  pass : CTA = 64 query rows; walks key tiles.       dq, d r_w_bias, d r_r_bias
//   dkv pass : CTA = 64 keys;       walks query tiles.     dk, dv   (written once, bf16)
//   dr  pass : CTA = 64 distances;  walks query tiles along the diagonal j = i + M - delta, so the
//              gradient of R (shared by all queries and batch elements) accumulates in registers.
//              In this pass the roles of K/V and R swap: R is the natural operand and the K / V
//              windows are read through the anti-diagonal (relative-shift) re-indexing.
//
// Autograd counterpart of commu/model/model.py:312-345 (the reference relies on torch autograd).
#include <stdlib.h>
#include "api_common.h"
#include "attn_common.cuh"

namespace {
using namespace attn;
constexpr int NTHREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

// ---- warp-level building blocks; all tiles are [rows][64] bf16, 128 B rows, XOR-swizzled ----

// A fragment (16 rows starting at row0, k-step ks) of a row-major tile
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const uint8_t* tile, int row0, int ks, int lane) {
  const int row = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int ch = 2 * ks + (lane >> 4);
  cb::ldmatrix_x4(a, cb::smem_u32(tile + swz(row, ch)));
}
// acc[NB][4] += A(16 x 64, rows a_row0.. of a_tile) * B^T where B = rows b_row0 .. b_row0+8*NB of b_tile
template <int NB>
__device__ __forceinline__ void mma_rows(float (&acc)[NB][4], const uint8_t* a_tile, int a_row0,
                                         const uint8_t* b_tile, int b_row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    load_a(a, a_tile, a_row0, ks, lane);
#pragma unroll
    for (int np = 0; np < NB / 2; ++np) {
      uint32_t bb[4];
      const int row = b_row0 + np * 16 + (lane & 7) + (lane >> 4) * 8;
      const int ch = 2 * ks + ((lane >> 3) & 1);
      cb::ldmatrix_x4(bb, cb::smem_u32(b_tile + swz(row, ch)));
      const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
      cb::mma_bf16_16816(acc[2 * np], a, b0);
      cb::mma_bf16_16816(acc[2 * np + 1], a, b1);
    }
  }
}
// out[li, lc] (+)= band[li, li + BN-1 - lc] through the warp's scratch
template <bool ACCUM>
__device__ __forceinline__ void shear(float (&out)[8][4], const float (&band)[10][4], float* scr, int g, int q4) {
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 10; ++n) {
    *reinterpret_cast<float2*>(scr + g * SW + n * 8 + 2 * q4) = make_float2(band[n][0], band[n][1]);
    *reinterpret_cast<float2*>(scr + (g + 8) * SW + n * 8 + 2 * q4) = make_float2(band[n][2], band[n][3]);
  }
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int lc = n * 8 + 2 * q4;
    const float* r0 = scr + g * SW + (g + BN - 1 - lc);
    const float* r1 = scr + (g + 8) * SW + (g + 8 + BN - 1 - lc);
    if (ACCUM) {
      out[n][0] += r0[0]; out[n][1] += r0[-1]; out[n][2] += r1[0]; out[n][3] += r1[-1];
    } else {
      out[n][0] = r0[0]; out[n][1] = r0[-1]; out[n][2] = r1[0]; out[n][3] = r1[-1];
    }
  }
}

struct RowInfo {
  float lse2[2];   // lse * log2(e) of rows g, g+8
  float delta[2];  // rowsum(dO * O)
  int i[2];        // global query rows
};

// Tile routine.  DIAG = false: columns are keys j0 + lc (K/V natural, R banded).
//                DIAG = true : columns are distances d0 + lc (R natural, K/V windows banded).
template <bool DIAG>
__device__ __forceinline__ void tile_p_ds(const Params& p, const uint8_t* qu, const uint8_t* qv,
                                          const uint8_t* dout, const uint8_t* kt, const uint8_t* vt,
                                          const uint8_t* rt, int warp, int lane, float* scr,
                                          const RowInfo& ri, int col0, bool reset,
                                          float (&pr)[8][4], float (&ds)[8][4]) {
  const int g = lane >> 2, q4 = lane & 3;
  float (&sc)[8][4] = pr;  // scores are overwritten by probabilities in place
  float (&dp)[8][4] = ds;  // dP is overwritten by dS in place
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
    dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
  }
  if (!DIAG) {
    mma_rows<8>(sc, qu, warp * 16, kt, 0, lane);
    {
      float bd[10][4];
#pragma unroll
      for (int n = 0; n < 10; ++n) bd[n][0] = bd[n][1] = bd[n][2] = bd[n][3] = 0.f;
      mma_rows<10>(bd, qv, warp * 16, rt, warp * 16, lane);
      shear<true>(sc, bd, scr, g, q4);
    }
    mma_rows<8>(dp, dout, warp * 16, vt, 0, lane);
  } else {
    mma_rows<8>(sc, qv, warp * 16, rt, 0, lane);
    {
      float bd[10][4];
#pragma unroll
      for (int n = 0; n < 10; ++n) bd[n][0] = bd[n][1] = bd[n][2] = bd[n][3] = 0.f;
      mma_rows<10>(bd, qu, warp * 16, kt, warp * 16, lane);
      shear<true>(sc, bd, scr, g, q4);
    }
    {
      float bd[10][4];
#pragma unroll
      for (int n = 0; n < 10; ++n) bd[n][0] = bd[n][1] = bd[n][2] = bd[n][3] = 0.f;
      mma_rows<10>(bd, dout, warp * 16, vt, warp * 16, lane);
      shear<false>(dp, bd, scr, g, q4);
    }
  }
  const float sl2 = p.scale * LOG2E;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = ri.i[r];
    const int hi = i < p.T ? i + p.M : -1;
    const int lo = key_lo(i, p.M, p.same_length, p.shift, reset);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = col0 + n * 8 + 2 * q4 + e;
        const int j = DIAG ? (i + p.M - c) : c;
        const bool ok = (j <= hi) && (j >= lo);
        const float pv = ok ? exp2f(sc[n][2 * r + e] * sl2 - ri.lse2[r]) : 0.f;
        pr[n][2 * r + e] = pv;
        ds[n][2 * r + e] = pv * (dp[n][2 * r + e] - ri.delta[r]) * p.scale;
      }
    }
  }
}

__device__ __forceinline__ RowInfo load_rowinfo(const Params& p, int b, int h, int iw, int g) {
  RowInfo ri;
  const float* lp = p.lse + ((long long)b * p.H + h) * p.T;
  const float* dl = p.delta + ((long long)b * p.H + h) * p.T;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = iw + g + 8 * r;
    ri.i[r] = i;
    ri.lse2[r] = i < p.T ? lp[i] * LOG2E : 0.f;
    ri.delta[r] = i < p.T ? dl[i] : 0.f;
  }
  return ri;
}

// accumulator (m16n8 layout, 8 n-blocks = 64 cols) -> bf16 rows in a swizzled tile
__device__ __forceinline__ void store_acc_bf16(uint8_t* tile, int row0, const float (&a)[8][4], int g, int q4) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(tile + swz(row0 + g, n) + q4 * 4) = cb::pack_bf16(a[n][0], a[n][1]);
    *reinterpret_cast<uint32_t*>(tile + swz(row0 + g + 8, n) + q4 * 4) = cb::pack_bf16(a[n][2], a[n][3]);
  }
}

// acc[8][4] (16 x 64) += A^T-slice * B where A is stored [k rows][m cols] (take m cols m0..m0+15
// as the 16 output rows) and B is stored [k rows][64 n cols]; k runs over KROWS rows.
template <int KROWS>
__device__ __forceinline__ void mma_tn(float (&acc)[8][4], const uint8_t* a_tile, int m0,
                                       const uint8_t* b_tile, int lane) {
#pragma unroll
  for (int kq = 0; kq < KROWS / 16; ++kq) {
    uint32_t a[4];
    {
      const int row = kq * 16 + (lane & 7) + (lane >> 4) * 8;
      const int ch = (m0 >> 3) + ((lane >> 3) & 1);
      cb::ldmatrix_x4_trans(a, cb::smem_u32(a_tile + swz(row, ch)));
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      const int row = kq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int ch = 2 * np + (lane >> 4);
      cb::ldmatrix_x4_trans(bb, cb::smem_u32(b_tile + swz(row, ch)));
      const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
      cb::mma_bf16_16816(acc[2 * np], a, b0);
      cb::mma_bf16_16816(acc[2 * np + 1], a, b1);
    }
  }
}

// acc[8][4] += A(regs, 16 x 16*KS from accumulators) * B where B stored [k rows][64 n cols]
__device__ __forceinline__ void mma_pv(float (&acc)[8][4], const float (&pa)[8][4], const uint8_t* b_tile,
                                       int b_row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = cb::pack_bf16(pa[2 * ks][0], pa[2 * ks][1]);
    a[1] = cb::pack_bf16(pa[2 * ks][2], pa[2 * ks][3]);
    a[2] = cb::pack_bf16(pa[2 * ks + 1][0], pa[2 * ks + 1][1]);
    a[3] = cb::pack_bf16(pa[2 * ks + 1][2], pa[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      const int row = b_row0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int ch = 2 * np + (lane >> 4);
      cb::ldmatrix_x4_trans(bb, cb::smem_u32(b_tile + swz(row, ch)));
      const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
      cb::mma_bf16_16816(acc[2 * np], a, b0);
      cb::mma_bf16_16816(acc[2 * np + 1], a, b1);
    }
  }
}

struct BwdSmem {
  uint8_t qu[BM * 128];
  uint8_t qv[BM * 128];
  uint8_t dout[BM * 128];
  uint8_t kt[BAND * 128];  // key tile (64 rows) or key window (128 rows, dr pass)
  uint8_t vt[BAND * 128];
  uint8_t rt[BAND * 128];  // R band (128 rows) or R tile (64 rows, dr pass)
  float scratch[4][16 * SW];
};
// transposition area carved from the scratch: P rows then dS rows, 128 B per row, per warp
__device__ __forceinline__ uint8_t* xpose_row_base(BwdSmem& sm, int which) {
  return reinterpret_cast<uint8_t*>(&sm.scratch[0][0]) + which * 2048;
}

// ------------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dO[i,b,h,d] * O[i,b,h,d]
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ dout,
                                  long long lddo, int T, int B, int H, float* __restrict__ delta) {
  // 8 lanes per (i, b, h) row of 64 values (16 bytes each), four rows per warp
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int part = threadIdx.x & 7;
  const long long total = (long long)T * B * H;
  const bool ok = w < total;
  const long long wc = ok ? w : 0;
  const int h = wc % H;
  const long long ib = wc / H;  // i*B + b
  const int b = ib % B;
  const int i = ib / B;
  const uint4 a = *reinterpret_cast<const uint4*>(o + ib * ldo + h * DH + part * 8);
  const uint4 d = *reinterpret_cast<const uint4*>(dout + ib * lddo + h * DH + part * 8);
  float s = cb::bf16_lo(a.x) * cb::bf16_lo(d.x) + cb::bf16_hi(a.x) * cb::bf16_hi(d.x);
  s += cb::bf16_lo(a.y) * cb::bf16_lo(d.y) + cb::bf16_hi(a.y) * cb::bf16_hi(d.y);
  s += cb::bf16_lo(a.z) * cb::bf16_lo(d.z) + cb::bf16_hi(a.z) * cb::bf16_hi(d.z);
  s += cb::bf16_lo(a.w) * cb::bf16_lo(d.w) + cb::bf16_hi(a.w) * cb::bf16_hi(d.w);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && part == 0) delta[((long long)b * H + h) * T + i] = s;
}

// ------------------------------------------------------------------------------------------------
// dq pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2) relattn_bwd_dq_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * BM;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;
  const bf16* qub = p.qu_s + (long long)b * p.ldq + h * DH;
  const bf16* qvb = p.qv_s + (long long)b * p.ldq + h * DH;
  const bf16* kb = p.k + (long long)b * p.ldkv + h * DH;
  const bf16* vb_ = p.v + (long long)b * p.ldkv + h * DH;
  const bf16* rb = p.r + h * DH;
  const bf16* dob = p.dout + (long long)b * p.lddo + h * DH;

  load_tile_async<BM, NTHREADS>(sm.qu, qub, p.ldq, p.B, i0, p.T, tid);
  load_tile_async<BM, NTHREADS>(sm.qv, qvb, p.ldq, p.B, i0, p.T, tid);
  load_tile_async<BM, NTHREADS>(sm.dout, dob, p.lddo, p.B, i0, p.T, tid);
  cb::cp_async_commit();

  const int iw = i0 + warp * 16;
  const RowInfo ri = load_rowinfo(p, b, h, iw, g);
  const int i_last = min(p.T - 1, i0 + BM - 1);
  const int jt_last = (i_last + p.M) / BN;
  const int jt_first = key_lo(i0, p.M, p.same_length, p.shift, reset) / BN;
  float dq_ac[8][4], dq_bd[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    dq_ac[n][0] = dq_ac[n][1] = dq_ac[n][2] = dq_ac[n][3] = 0.f;
    dq_bd[n][0] = dq_bd[n][1] = dq_bd[n][2] = dq_bd[n][3] = 0.f;
  }
  float* scr = sm.scratch[warp];

  for (int jt = jt_first; jt <= jt_last; ++jt) {
    __syncthreads();  // previous tile fully consumed
    load_tile_async<BN, NTHREADS>(sm.kt, kb, p.ldkv, p.B, jt * BN, Ktot, tid);
    load_tile_async<BN, NTHREADS>(sm.vt, vb_, p.ldkv, p.B, jt * BN, Ktot, tid);
    const int dlo = i0 + p.M - (jt * BN + BN - 1);
    load_tile_async<BAND, NTHREADS>(sm.rt, rb, p.ldr, 1, dlo, p.Kr, tid);
    cb::cp_async_commit();
    cb::cp_async_wait<0>();
    __syncthreads();

    float pr[8][4], ds[8][4];
    tile_p_ds<false>(p, sm.qu, sm.qv, sm.dout, sm.kt, sm.vt, sm.rt, warp, lane, scr, ri, jt * BN, reset, pr, ds);
    // dq_ac += dS K
    mma_pv(dq_ac, ds, sm.kt, 0, lane);
    // dBDraw[li, c] = dS[li, li + BN-1 - c]  (inverse shift), then dq_bd += dBDraw Rband
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<float2*>(scr + g * SW + n * 8 + 2 * q4) = make_float2(ds[n][0], ds[n][1]);
      *reinterpret_cast<float2*>(scr + (g + 8) * SW + n * 8 + 2 * q4) = make_float2(ds[n][2], ds[n][3]);
    }
    __syncwarp();
#pragma unroll
    for (int kc = 0; kc < 5; ++kc) {
      uint32_t a[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {      // k columns c .. c+1 and c+8 .. c+9
#pragma unroll
        for (int r = 0; r < 2; ++r) {             // rows g, g+8
          const int li = g + 8 * r;
          const int c = kc * 16 + half * 8 + 2 * q4;
          const int lj0 = li + BN - 1 - c, lj1 = lj0 - 1;
          const float v0 = (lj0 >= 0 && lj0 < BN) ? scr[li * SW + lj0] : 0.f;
          const float v1 = (lj1 >= 0 && lj1 < BN) ? scr[li * SW + lj1] : 0.f;
          a[half * 2 + r] = cb::pack_bf16(v0, v1);
        }
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        const int row = warp * 16 + kc * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = 2 * np + (lane >> 4);
        cb::ldmatrix_x4_trans(bb, cb::smem_u32(sm.rt + swz(row, ch)));
        const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
        cb::mma_bf16_16816(dq_bd[2 * np], a, b0);
        cb::mma_bf16_16816(dq_bd[2 * np + 1], a, b1);
      }
    }
  }

  // ---- outputs: dq (bf16), d r_w_bias += colsum(dq_ac), d r_r_bias += colsum(dq_bd) ----
  bf16* dqb = p.dq + (long long)b * p.lddq + h * DH;
  const int ia = iw + g, ib = iw + g + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + 2 * q4;
    if (ia < p.T)
      *reinterpret_cast<uint32_t*>(dqb + ((long long)ia * p.B) * p.lddq + c) =
          cb::pack_bf16(dq_ac[n][0] + dq_bd[n][0], dq_ac[n][1] + dq_bd[n][1]);
    if (ib < p.T)
      *reinterpret_cast<uint32_t*>(dqb + ((long long)ib * p.B) * p.lddq + c) =
          cb::pack_bf16(dq_ac[n][2] + dq_bd[n][2], dq_ac[n][3] + dq_bd[n][3]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float sa = (ia < p.T ? dq_ac[n][e] : 0.f) + (ib < p.T ? dq_ac[n][2 + e] : 0.f);
      float sb = (ia < p.T ? dq_bd[n][e] : 0.f) + (ib < p.T ? dq_bd[n][2 + e] : 0.f);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
      }
      if (g == 0) {
        atomicAdd(p.du + h * DH + n * 8 + 2 * q4 + e, sa);
        atomicAdd(p.dvb + h * DH + n * 8 + 2 * q4 + e, sb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dk / dv pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2) relattn_bwd_dkv_kernel(const Params p, bf16* dk_out,
                                                                      bf16* dv_out, long long lddkv) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * BN;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;
  const bf16* qub = p.qu_s + (long long)b * p.ldq + h * DH;
  const bf16* qvb = p.qv_s + (long long)b * p.ldq + h * DH;
  const bf16* kb = p.k + (long long)b * p.ldkv + h * DH;
  const bf16* vb_ = p.v + (long long)b * p.ldkv + h * DH;
  const bf16* rb = p.r + h * DH;
  const bf16* dob = p.dout + (long long)b * p.lddo + h * DH;

  load_tile_async<BN, NTHREADS>(sm.kt, kb, p.ldkv, p.B, j0, Ktot, tid);
  load_tile_async<BN, NTHREADS>(sm.vt, vb_, p.ldkv, p.B, j0, Ktot, tid);
  cb::cp_async_commit();

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  // query rows that can see any key of this tile
  int i_min = max(0, j0 - p.M);
  int i_max = p.T - 1;
  if (p.same_length) i_max = min(i_max, j0 + BN - 2 + p.shift);
  if (reset && j0 + BN - 1 < p.M) i_max = -1;
  float* scr = sm.scratch[warp];
  uint8_t* xp = xpose_row_base(sm, 0);
  uint8_t* xds = reinterpret_cast<uint8_t*>(&sm.scratch[0][0]) + 8192;

  for (int it = i_min / BM; it * BM <= i_max; ++it) {
    const int i0 = it * BM;
    __syncthreads();
    load_tile_async<BM, NTHREADS>(sm.qu, qub, p.ldq, p.B, i0, p.T, tid);
    load_tile_async<BM, NTHREADS>(sm.qv, qvb, p.ldq, p.B, i0, p.T, tid);
    load_tile_async<BM, NTHREADS>(sm.dout, dob, p.lddo, p.B, i0, p.T, tid);
    const int dlo = i0 + p.M - (j0 + BN - 1);
    load_tile_async<BAND, NTHREADS>(sm.rt, rb, p.ldr, 1, dlo, p.Kr, tid);
    cb::cp_async_commit();
    cb::cp_async_wait<0>();
    __syncthreads();
    const int iw = i0 + warp * 16;
    const RowInfo ri = load_rowinfo(p, b, h, iw, g);
    float pr[8][4], ds[8][4];
    tile_p_ds<false>(p, sm.qu, sm.qv, sm.dout, sm.kt, sm.vt, sm.rt, warp, lane, scr, ri, j0, reset, pr, ds);
    __syncthreads();  // every warp is done with its scratch before it becomes the transposition area
    store_acc_bf16(xp, warp * 16, pr, g, q4);
    store_acc_bf16(xds, warp * 16, ds, g, q4);
    __syncthreads();
    mma_tn<BM>(dv, xp, warp * 16, sm.dout, lane);  // dV[keys_w] += P^T dO
    mma_tn<BM>(dk, xds, warp * 16, sm.qu, lane);   // dK[keys_w] += dS^T (q + u)
  }
  cb::cp_async_wait<0>();
  const int ja = j0 + warp * 16 + g, jb = ja + 8;
  bf16* dkb = dk_out + (long long)b * lddkv + h * DH;
  bf16* dvb2 = dv_out + (long long)b * lddkv + h * DH;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + 2 * q4;
    if (ja < Ktot) {
      *reinterpret_cast<uint32_t*>(dkb + ((long long)ja * p.B) * lddkv + c) = cb::pack_bf16(dk[n][0], dk[n][1]);
      *reinterpret_cast<uint32_t*>(dvb2 + ((long long)ja * p.B) * lddkv + c) = cb::pack_bf16(dv[n][0], dv[n][1]);
    }
    if (jb < Ktot) {
      *reinterpret_cast<uint32_t*>(dkb + ((long long)jb * p.B) * lddkv + c) = cb::pack_bf16(dk[n][2], dk[n][3]);
      *reinterpret_cast<uint32_t*>(dvb2 + ((long long)jb * p.B) * lddkv + c) = cb::pack_bf16(dv[n][2], dv[n][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dR pass (diagonal walk)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2) relattn_bwd_dr_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int d0 = blockIdx.x * BN;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;
  const bf16* qub = p.qu_s + (long long)b * p.ldq + h * DH;
  const bf16* qvb = p.qv_s + (long long)b * p.ldq + h * DH;
  const bf16* kb = p.k + (long long)b * p.ldkv + h * DH;
  const bf16* vb_ = p.v + (long long)b * p.ldkv + h * DH;
  const bf16* rb = p.r + h * DH;
  const bf16* dob = p.dout + (long long)b * p.lddo + h * DH;

  load_tile_async<BN, NTHREADS>(sm.rt, rb, p.ldr, 1, d0, p.Kr, tid);
  cb::cp_async_commit();
  float dr[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) dr[n][0] = dr[n][1] = dr[n][2] = dr[n][3] = 0.f;
  // queries that own a visible key at some distance of this tile: j = i + M - delta >= lo(i)
  int i_min = max(0, d0 - p.M);
  if (reset) i_min = max(i_min, d0);
  bool any = true;
  if (p.same_length && d0 >= p.M + p.shift) any = false;
  float* scr = sm.scratch[warp];
  uint8_t* xds = xpose_row_base(sm, 0);

  for (int it = i_min / BM; any && it * BM < p.T; ++it) {
    const int i0 = it * BM;
    __syncthreads();
    load_tile_async<BM, NTHREADS>(sm.qu, qub, p.ldq, p.B, i0, p.T, tid);
    load_tile_async<BM, NTHREADS>(sm.qv, qvb, p.ldq, p.B, i0, p.T, tid);
    load_tile_async<BM, NTHREADS>(sm.dout, dob, p.lddo, p.B, i0, p.T, tid);
    const int jw0 = i0 + p.M - d0 - (BN - 1);   // key window start; band row c <-> key jw0 + c
    load_tile_async<BAND, NTHREADS>(sm.kt, kb, p.ldkv, p.B, jw0, Ktot, tid);
    load_tile_async<BAND, NTHREADS>(sm.vt, vb_, p.ldkv, p.B, jw0, Ktot, tid);
    cb::cp_async_commit();
    cb::cp_async_wait<0>();
    __syncthreads();
    const int iw = i0 + warp * 16;
    const RowInfo ri = load_rowinfo(p, b, h, iw, g);
    float pr[8][4], ds[8][4];
    tile_p_ds<true>(p, sm.qu, sm.qv, sm.dout, sm.kt, sm.vt, sm.rt, warp, lane, scr, ri, d0, reset, pr, ds);
    __syncthreads();
    store_acc_bf16(xds, warp * 16, ds, g, q4);
    __syncthreads();
    mma_tn<BM>(dr, xds, warp * 16, sm.qv, lane);  // dR[delta_w] += dS^T (q + v)
  }
  cb::cp_async_wait<0>();
  const int da = d0 + warp * 16 + g, db = da + 8;
  float* drb = p.dr + h * DH;
  const long long ldr32 = (long long)p.H * DH;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + 2 * q4;
    if (da < p.Kr) {
      atomicAdd(drb + da * ldr32 + c, dr[n][0]);
      atomicAdd(drb + da * ldr32 + c + 1, dr[n][1]);
    }
    if (db < p.Kr) {
      atomicAdd(drb + db * ldr32 + c, dr[n][2]);
      atomicAdd(drb + db * ldr32 + c + 1, dr[n][3]);
    }
  }
}

}  // namespace

namespace cb_host {
int check_attn_common(const attn::Params& p, const char* who);
}
extern "C" int commu_relattn_bwd_dr_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                       int64_t ldkv, const void* r, int64_t ldr, int kr,
                                       const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                       int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                       const float* delta, float* dr, float* du, float* dvb, void* stream_);
extern "C" int commu_relattn_bwd_dq_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                       int64_t ldkv, const void* r, int64_t ldr, int kr,
                                       const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                       int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                       const float* delta, void* dq, int64_t lddq, float* du, float* dvb,
                                       void* stream_);
extern "C" int commu_relattn_bwd_dkv_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                        int64_t ldkv, const void* r, int64_t ldr, int kr,
                                        const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                        int shift, float scale, const float* lse, const void* dout, int64_t lddo,
                                        const float* delta, void* dk, void* dv, int64_t lddkv, void* stream_);

namespace {
// pass implementations: 1 = tcgen05 kernel, 0 = v1 warp-MMA kernel; -1 = not yet read from the environment
int g_impl[3] = {-1, -1, -1};   // dq, dkv, dr
int pass_impl(int which, const char* env) {
  if (g_impl[which] < 0) {
    const char* e = getenv(env);
    g_impl[which] = (e && e[0] == 'v') ? 0 : 1;
  }
  return g_impl[which];
}
}  // namespace

// Selects the implementation of each backward pass at run time (1 = tcgen05, 0 = v1 warp-MMA, <0 = keep).
// Defaults come from COMMU_ATTN_BWD_{DQ,DKV,DR} (unset or "tc" = tcgen05, "v1" = warp-MMA).
extern "C" int commu_relattn_bwd_set_impl(int dq_tc, int dkv_tc, int dr_tc) {
  if (dq_tc >= 0) g_impl[0] = dq_tc ? 1 : 0;
  if (dkv_tc >= 0) g_impl[1] = dkv_tc ? 1 : 0;
  if (dr_tc >= 0) g_impl[2] = dr_tc ? 1 : 0;
  return 0;
}

// Backward of commu_relattn_fwd.  Inputs are the forward's operands plus the saved (q+r_w_bias),
// (q+r_r_bias) bf16 tensors (layout of q), the forward output `out`, its LSE and dout.
//   dq   : bf16 [T*B, lddq]        (same column layout as q)
//   dk,dv: bf16 [K*B, lddkv]       (written for every key row, zero where nothing attends)
//   dr   : fp32 [kr, H*64]         accumulated (+=) over batch; caller zeroes
//   du,dvb: fp32 [H,64]            accumulated (+=); caller zeroes
//   delta_ws: fp32 [B,H,T] workspace
extern "C" int commu_relattn_bwd(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                                 int64_t ldkv, const void* r, int64_t ldr, int kr,
                                 const unsigned char* reset, int T, int M, int B, int H, int same_length,
                                 int shift, float scale, const void* out, int64_t ldo, const float* lse,
                                 const void* dout, int64_t lddo, float* delta_ws, void* dq, int64_t lddq,
                                 void* dk, void* dv, int64_t lddkv, float* dr, float* du, float* dvb,
                                 const void* p_save, const float* mt_save, void* ws, int64_t ws_bytes,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  attn::Params p = {};
  p.q = (const bf16*)qu; p.k = (const bf16*)k; p.v = (const bf16*)v; p.r = (const bf16*)r;
  p.qu_s = (bf16*)const_cast<void*>(qu); p.qv_s = (bf16*)const_cast<void*>(qv);
  static const float dummy = 0.f;
  p.u = &dummy; p.vb = &dummy;  // biases are already folded into qu / qv
  p.reset = reset;
  p.ldq = ldq; p.ldkv = ldkv; p.ldr = ldr;
  p.T = T; p.M = M; p.B = B; p.H = H; p.Kr = kr;
  p.same_length = same_length; p.shift = shift; p.scale = scale;
  p.lse = const_cast<float*>(lse);
  p.dout = (const bf16*)dout; p.lddo = lddo; p.delta = delta_ws;
  p.dq = (bf16*)dq; p.lddq = lddq;
  p.dr = dr; p.du = du; p.dvb = dvb;
  int rc = cb_host::check_attn_common(p, "relattn_bwd");
  if (rc) return rc;
  CB_REQUIRE(qv && out && lse && dout && delta_ws && dq && dk && dv && dr && du && dvb, "relattn_bwd: null arg");
  CB_REQUIRE(ldo % 2 == 0 && lddo % 8 == 0 && lddq % 2 == 0 && lddkv % 2 == 0, "relattn_bwd: bad leading dims");
  static bool attr = false;
  if (!attr) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_bwd_dr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    attr = true;
  }
  cb_host::ProfScope prof(cb_host::PROF_ATTN_BWD, stream);
  {
    const long long rows8 = (long long)T * B * H * 8;      // 8 lanes per row
    CB_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0,
               "relattn_bwd: out / dout must be 16-byte aligned with leading dims that are multiples of 8");
    attn_delta_kernel<<<(unsigned)((rows8 + 255) / 256), 256, 0, stream>>>(
        (const bf16*)out, ldo, (const bf16*)dout, lddo, T, B, H, delta_ws);
  }
  if (p_save) {   // product path: probabilities stored by the forward, dS materialised once (attn_bwd_mat.cu)
    CB_REQUIRE(mt_save && ws, "relattn_bwd: p_save needs mt_save and a workspace");
    int rcm = commu_relattn_bwd_mat(qu, qv, ldq, k, v, ldkv, r, ldr, kr, reset, T, M, B, H, same_length, shift, scale,
                                    lse, dout, lddo, delta_ws, p_save, mt_save, ws, ws_bytes, dq, lddq, dk, dv,
                                    lddkv, dr, du, dvb, stream_);
    if (rcm) return rcm;
    cb_host::count_launch(1);
    return 0;
  }
  const int Ktot = T + M;
  // the tcgen05 dq and dR passes split d r_w_bias / d r_r_bias between them, so they are selected as a pair
  const bool dq_tc = pass_impl(0, "COMMU_ATTN_BWD_DQ") && pass_impl(2, "COMMU_ATTN_BWD_DR");
  if (dq_tc) {
    int rc3 = commu_relattn_bwd_dq_tc(qu, qv, ldq, k, v, ldkv, r, ldr, kr, reset, T, M, B, H, same_length, shift, scale,
                                      lse, dout, lddo, delta_ws, dq, lddq, du, dvb, stream_);
    if (rc3) return rc3;
  } else {
    relattn_bwd_dq_kernel<<<dim3(cb_host::ceil_div(T, attn::BM), H, B), NTHREADS, sizeof(BwdSmem), stream>>>(p);
  }
  // dk / dv pass: tcgen05 kernel (default), or the warp-MMA pass with COMMU_ATTN_BWD_DKV=v1
  const bool dkv_tc = pass_impl(1, "COMMU_ATTN_BWD_DKV");
  if (dkv_tc) {
    int rc2 = commu_relattn_bwd_dkv_tc(qu, qv, ldq, k, v, ldkv, r, ldr, kr, reset, T, M, B, H, same_length, shift,
                                       scale, lse, dout, lddo, delta_ws, dk, dv, lddkv, stream_);
    if (rc2) return rc2;
  } else {
    relattn_bwd_dkv_kernel<<<dim3(cb_host::ceil_div(Ktot, attn::BN), H, B), NTHREADS, sizeof(BwdSmem), stream>>>(
        p, (bf16*)dk, (bf16*)dv, lddkv);
  }
  const bool dr_tc = dq_tc;
  if (dr_tc) {
    int rc4 = commu_relattn_bwd_dr_tc(qu, qv, ldq, k, v, ldkv, r, ldr, kr, reset, T, M, B, H, same_length, shift, scale,
                                      lse, dout, lddo, delta_ws, dr, du, dvb, stream_);
    if (rc4) return rc4;
  } else {
    relattn_bwd_dr_kernel<<<dim3(cb_host::ceil_div(Ktot, attn::BN), H, B), NTHREADS, sizeof(BwdSmem), stream>>>(p);
  }
  cb_host::count_launch(4);
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
