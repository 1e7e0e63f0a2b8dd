// Fused relative-position attention forward (training / prefill path), v1: warp-level bf16 tensor
// core MMA (mma.sync m16n8k16), flash-style online softmax, analytic mask, relative shift as an
// anti-diagonal re-read of a banded product through shared memory.  Never materialises T x K.
// Replaces commu/model/model.py:312-345 (AC, BD, _rel_shift, mask, softmax, AV).
#include "api_common.h"
#include "attn_common.cuh"

namespace {
using namespace attn;

constexpr int NTHREADS = 128;

struct FwdSmem {
  uint8_t k[2][BN * 128];
  uint8_t v[2][BN * 128];
  uint8_t r[2][BAND * 128];
  float scratch[4][16 * SW];
};

__global__ void __launch_bounds__(NTHREADS, 2) relattn_fwd_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * BM;
  const bool reset = p.reset && p.reset[b];
  const int Ktot = p.T + p.M;

  // ---- stage (q + r_w_bias) and (q + r_r_bias) as bf16 tiles, then hold them as A fragments ----
  {
    const bf16* qb = p.q + (long long)b * p.ldq + h * DH;
#pragma unroll
    for (int it = 0; it < (BM * 8) / NTHREADS; ++it) {
      const int idx = it * NTHREADS + tid;
      const int row = idx >> 3, ch = idx & 7;
      const int i = i0 + row;
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (i < p.T) raw = *reinterpret_cast<const uint4*>(qb + ((long long)i * p.B) * p.ldq + ch * 8);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t ou[4], ov[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = cb::bf16_lo(w[e]), hi = cb::bf16_hi(w[e]);
        const int c = h * DH + ch * 8 + e * 2;
        ou[e] = cb::pack_bf16(lo + __ldg(p.u + c), hi + __ldg(p.u + c + 1));
        ov[e] = cb::pack_bf16(lo + __ldg(p.vb + c), hi + __ldg(p.vb + c + 1));
      }
      *reinterpret_cast<uint4*>(sm.k[0] + swz(row, ch)) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<uint4*>(sm.v[0] + swz(row, ch)) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      if (p.qu_s && i < p.T) {  // saved for the backward passes
        const long long off = (long long)b * p.ldq + h * DH + ((long long)i * p.B) * p.ldq + ch * 8;
        *reinterpret_cast<uint4*>(p.qu_s + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<uint4*>(p.qv_s + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      }
    }
  }
  __syncthreads();
  uint32_t qu[4][4], qv[4][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int ch = 2 * ks + (lane >> 4);
      cb::ldmatrix_x4(qu[ks], cb::smem_u32(sm.k[0] + swz(row, ch)));
      cb::ldmatrix_x4(qv[ks], cb::smem_u32(sm.v[0] + swz(row, ch)));
    }
  }
  __syncthreads();

  // ---- key tile range of this query tile ----
  const int i_last = min(p.T - 1, i0 + BM - 1);
  const int jt_last = (i_last + p.M) / BN;
  const int jt_first = key_lo(i0, p.M, p.same_length, p.shift, reset) / BN;

  const bf16* kb = p.k + (long long)b * p.ldkv + h * DH;
  const bf16* vbse = p.v + (long long)b * p.ldkv + h * DH;
  const bf16* rb = p.r + h * DH;

  auto issue = [&](int jt, int s) {
    load_tile_async<BN, NTHREADS>(sm.k[s], kb, p.ldkv, p.B, jt * BN, Ktot, tid);
    load_tile_async<BN, NTHREADS>(sm.v[s], vbse, p.ldkv, p.B, jt * BN, Ktot, tid);
    const int dlo = i0 + p.M - (jt * BN + BN - 1);
    load_tile_async<BAND, NTHREADS>(sm.r[s], rb, p.ldr, 1, dlo, p.Kr, tid);
    cb::cp_async_commit();
  };

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  float* scr = sm.scratch[warp];
  const int iw = i0 + warp * 16;  // first query row of this warp

  issue(jt_first, 0);
  for (int jt = jt_first; jt <= jt_last; ++jt) {
    const int s = (jt - jt_first) & 1;
    if (jt + 1 <= jt_last) {
      issue(jt + 1, s ^ 1);
      cb::cp_async_wait<1>();
    } else {
      cb::cp_async_wait<0>();
    }
    __syncthreads();

    // ---- content scores: S = (q+u) K^T ----
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bk[4];
        const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int ch = 2 * ks + ((lane >> 3) & 1);
        cb::ldmatrix_x4(bk, cb::smem_u32(sm.k[s] + swz(row, ch)));
        const uint32_t b0[2] = {bk[0], bk[1]}, b1[2] = {bk[2], bk[3]};
        cb::mma_bf16_16816(sc[2 * np], qu[ks], b0);
        cb::mma_bf16_16816(sc[2 * np + 1], qu[ks], b1);
      }
    }
    // ---- position scores on the warp's band: BDraw[li, c] = (q+v)[li] . Rband[16w + c] ----
    {
      float bd[10][4];
#pragma unroll
      for (int n = 0; n < 10; ++n) bd[n][0] = bd[n][1] = bd[n][2] = bd[n][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 5; ++np) {
          uint32_t br[4];
          const int row = warp * 16 + np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int ch = 2 * ks + ((lane >> 3) & 1);
          cb::ldmatrix_x4(br, cb::smem_u32(sm.r[s] + swz(row, ch)));
          const uint32_t b0[2] = {br[0], br[1]}, b1[2] = {br[2], br[3]};
          cb::mma_bf16_16816(bd[2 * np], qv[ks], b0);
          cb::mma_bf16_16816(bd[2 * np + 1], qv[ks], b1);
        }
      }
#pragma unroll
      for (int n = 0; n < 10; ++n) {
        *reinterpret_cast<float2*>(scr + g * SW + n * 8 + 2 * q4) = make_float2(bd[n][0], bd[n][1]);
        *reinterpret_cast<float2*>(scr + (g + 8) * SW + n * 8 + 2 * q4) = make_float2(bd[n][2], bd[n][3]);
      }
    }
    __syncwarp();
    // relative shift: S[li, lj] += BDraw[li, li + BN-1 - lj]
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int lj = n * 8 + 2 * q4;
      const float* r0 = scr + g * SW + (g + BN - 1 - lj);
      const float* r1 = scr + (g + 8) * SW + (g + 8 + BN - 1 - lj);
      sc[n][0] += r0[0];
      sc[n][1] += r0[-1];
      sc[n][2] += r1[0];
      sc[n][3] += r1[-1];
    }
    // ---- mask (analytic) ----
    const int j0 = jt * BN;
    {
      const int ia = iw + g, ib = iw + g + 8;
      const bool full = (j0 + BN - 1 <= iw + p.M) && (iw + 15 < p.T) &&
                        (j0 >= key_lo(iw + 15, p.M, p.same_length, p.shift, reset));
      if (!full) {
        const int hia = ia < p.T ? ia + p.M : -1, hib = ib < p.T ? ib + p.M : -1;
        const int loa = key_lo(ia, p.M, p.same_length, p.shift, reset);
        const int lob = key_lo(ib, p.M, p.same_length, p.shift, reset);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int j = j0 + n * 8 + 2 * q4;
          if (j > hia || j < loa) sc[n][0] = -INFINITY;
          if (j + 1 > hia || j + 1 < loa) sc[n][1] = -INFINITY;
          if (j > hib || j < lob) sc[n][2] = -INFINITY;
          if (j + 1 > hib || j + 1 < lob) sc[n][3] = -INFINITY;
        }
      }
    }
    // ---- online softmax (base-2 domain) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      mx[0] = fmaxf(mx[0], fmaxf(sc[n][0], sc[n][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sc[n][2], sc[n][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(m_run[r], mx[r] * sl2);
      const float msafe = mnew[r] == -INFINITY ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);  // m_run = -inf -> 0
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      sc[n][0] = exp2f(sc[n][0] * sl2 - mnew[0]);
      sc[n][1] = exp2f(sc[n][1] * sl2 - mnew[0]);
      sc[n][2] = exp2f(sc[n][2] * sl2 - mnew[1]);
      sc[n][3] = exp2f(sc[n][3] * sl2 - mnew[1]);
      rs[0] += sc[n][0] + sc[n][1];
      rs[1] += sc[n][2] + sc[n][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= corr[0];
      o[n][1] *= corr[0];
      o[n][2] *= corr[1];
      o[n][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      a[0] = cb::pack_bf16(sc[2 * ks][0], sc[2 * ks][1]);
      a[1] = cb::pack_bf16(sc[2 * ks][2], sc[2 * ks][3]);
      a[2] = cb::pack_bf16(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
      a[3] = cb::pack_bf16(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bv[4];
        const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = 2 * np + (lane >> 4);
        cb::ldmatrix_x4_trans(bv, cb::smem_u32(sm.v[s] + swz(row, ch)));
        const uint32_t b0[2] = {bv[0], bv[1]}, b1[2] = {bv[2], bv[3]};
        cb::mma_bf16_16816(o[2 * np], a, b0);
        cb::mma_bf16_16816(o[2 * np + 1], a, b1);
      }
    }
    __syncthreads();
  }

  // ---- finalize: O /= l, LSE (natural log of sum exp(scaled score)) ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int ia = iw + g, ib = iw + g + 8;
  bf16* ob = p.out + (long long)b * p.ldo + h * DH;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + 2 * q4;
    if (ia < p.T)
      *reinterpret_cast<uint32_t*>(ob + ((long long)ia * p.B) * p.ldo + c) =
          cb::pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
    if (ib < p.T)
      *reinterpret_cast<uint32_t*>(ob + ((long long)ib * p.B) * p.ldo + c) =
          cb::pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
  }
  if (p.lse && q4 == 0) {
    float* lp = p.lse + ((long long)b * p.H + h) * p.T;
    if (ia < p.T) lp[ia] = (m_run[0] + log2f(l_run[0])) * 0.6931471805599453f;
    if (ib < p.T) lp[ib] = (m_run[1] + log2f(l_run[1])) * 0.6931471805599453f;
  }
}

}  // namespace

namespace cb_host {
int check_attn_common(const attn::Params& p, const char* who) {
  CB_REQUIRE(p.q && p.k && p.v && p.r && p.u && p.vb, "%s: null input", who);
  CB_REQUIRE(p.T > 0 && p.M >= 0 && p.B > 0 && p.H > 0, "%s: bad shape", who);
  CB_REQUIRE(p.Kr >= p.T + p.M, "%s: R has %d rows, needs >= %d", who, p.Kr, p.T + p.M);
  CB_REQUIRE(p.ldq % 8 == 0 && p.ldkv % 8 == 0 && p.ldr % 8 == 0, "%s: leading dims must be multiples of 8", who);
  return 0;
}
}  // namespace cb_host

extern "C" int commu_relattn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                 const void* r, int64_t ldr, int kr, const float* r_w_bias,
                                 const float* r_r_bias, const unsigned char* reset, int T, int M, int B,
                                 int H, int same_length, int shift, float scale, void* out, int64_t ldo,
                                 float* lse, void* qu_save, void* qv_save, void* stream) {
  attn::Params p = {};
  p.qu_s = (bf16*)qu_save; p.qv_s = (bf16*)qv_save;
  CB_REQUIRE((qu_save == nullptr) == (qv_save == nullptr), "relattn_fwd: qu_save/qv_save must both be set or null");
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.r = (const bf16*)r;
  p.u = r_w_bias; p.vb = r_r_bias; p.reset = reset;
  p.ldq = ldq; p.ldkv = ldkv; p.ldr = ldr;
  p.T = T; p.M = M; p.B = B; p.H = H; p.Kr = kr;
  p.same_length = same_length; p.shift = shift; p.scale = scale;
  p.out = (bf16*)out; p.ldo = ldo; p.lse = lse;
  int rc = cb_host::check_attn_common(p, "relattn_fwd");
  if (rc) return rc;
  CB_REQUIRE(out && (ldo % 2 == 0), "relattn_fwd: bad output");
  static bool attr = false;
  if (!attr) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(relattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FwdSmem)));
    attr = true;
  }
  dim3 grid(cb_host::ceil_div(T, attn::BM), H, B);
  cb_host::ProfScope prof(cb_host::PROF_ATTN_FWD, (cudaStream_t)stream);
  relattn_fwd_kernel<<<grid, NTHREADS, sizeof(FwdSmem), (cudaStream_t)stream>>>(p);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
