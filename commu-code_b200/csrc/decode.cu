// Token-by-token decode path (generate.py / InferenceTask hot loop): projected K/V ring cache,
// HBM-streaming single-query relative attention, skinny linear layers, on-device sampler.
//
// The reference re-runs the whole model over [memory ; token] for every token and re-copies the
// entire [L+1, M, B, d] memory (commu/midi_generator/midi_inferrer.py:199-207 ->
// commu/model/model.py:606-628, :507-538).  Here each layer keeps the PROJECTED keys / values of
// its last mem_len inputs in a ring buffer and the projected sinusoid table R by distance, so a
// step streams K, V and R once (the roofline of SURVEY.md section 8d) and touches every weight once.
//
// Precision: cache / weight element type is a template parameter (float for bit-faithful greedy
// parity with the fp32 reference, bf16 for throughput); all accumulation is fp32.
#include <math.h>
#include "api_common.h"
#include "common.cuh"

namespace {

template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct Elem<bf16> {
  static __device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = cb::bf16_lo(a.x); v[1] = cb::bf16_hi(a.x); v[2] = cb::bf16_lo(a.y); v[3] = cb::bf16_hi(a.y);
    v[4] = cb::bf16_lo(a.z); v[5] = cb::bf16_hi(a.z); v[6] = cb::bf16_lo(a.w); v[7] = cb::bf16_hi(a.w);
  }
  static __device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = cb::bf16_lo(a.x); v[1] = cb::bf16_hi(a.x); v[2] = cb::bf16_lo(a.y); v[3] = cb::bf16_hi(a.y);
  }
  static __device__ __forceinline__ void store(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------------------------------------
// out[b, n] = act( sum_k x[b,k] W[n,k] + bias[n] ) + res[b,n]      b < B <= 64
// One CTA (8 warps) owns NT output columns; warps split K; lanes own batch rows (lane, lane+32).
// ---------------------------------------------------------------------------------------------
constexpr int LIN_NT = 4;
constexpr int LIN_WARPS = 8;

template <typename WT>
__global__ void __launch_bounds__(LIN_WARPS * 32) decode_linear_kernel(
    const float* __restrict__ x, long long ldx, const WT* __restrict__ W, long long ldw,
    const float* __restrict__ bias, int relu, const float* __restrict__ res, long long ldr,
    float* __restrict__ out, long long ldo, int B, int N, int K) {
  __shared__ float part[LIN_WARPS][LIN_NT][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * LIN_NT;
  const int kchunks = (K + 31) / 32;
  const int per = (kchunks + LIN_WARPS - 1) / LIN_WARPS;
  const int c0 = warp * per, c1 = min(kchunks, c0 + per);
  float acc[LIN_NT][2];
#pragma unroll
  for (int j = 0; j < LIN_NT; ++j) acc[j][0] = acc[j][1] = 0.f;
  const bool r0 = lane < B, r1 = lane + 32 < B;
  for (int c = c0; c < c1; ++c) {
    const int k0 = c * 32;
    float xa[32], xb[32];
#pragma unroll
    for (int t = 0; t < 32; t += 4) {
      float4 va = make_float4(0, 0, 0, 0), vb2 = make_float4(0, 0, 0, 0);
      if (k0 + t < K) {
        if (r0) va = *reinterpret_cast<const float4*>(x + (long long)lane * ldx + k0 + t);
        if (r1) vb2 = *reinterpret_cast<const float4*>(x + (long long)(lane + 32) * ldx + k0 + t);
      }
      xa[t] = va.x; xa[t + 1] = va.y; xa[t + 2] = va.z; xa[t + 3] = va.w;
      xb[t] = vb2.x; xb[t + 1] = vb2.y; xb[t + 2] = vb2.z; xb[t + 3] = vb2.w;
    }
#pragma unroll
    for (int j = 0; j < LIN_NT; ++j) {
      const int n = n0 + j;
      if (n >= N) break;
      const WT* wr = W + (long long)n * ldw + k0;
#pragma unroll
      for (int t = 0; t < 32; t += 4) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (k0 + t < K) Elem<WT>::load4(wr + t, w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[j][0] = fmaf(xa[t + e], w[e], acc[j][0]);
          acc[j][1] = fmaf(xb[t + e], w[e], acc[j][1]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LIN_NT; ++j) {
    part[warp][j][lane] = acc[j][0];
    part[warp][j][lane + 32] = acc[j][1];
  }
  __syncthreads();
  const int j = threadIdx.x >> 6, b = threadIdx.x & 63;
  const int n = n0 + j;
  if (b < B && n < N) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < LIN_WARPS; ++w) v += part[w][j][b];
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    if (res) v += res[(long long)b * ldr + n];
    out[(long long)b * ldo + n] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled variant for the fp32 engine: out[64, N] = x[64, K] W^T as a register-tiled SIMT GEMM (fp32 FMA, the
// arithmetic the bit-faithful greedy parity needs).  One CTA = all batch rows x 32 output columns x one K split;
// thread = 4 rows x 2 columns; x and W chunks of 32 reduction steps are staged in shared memory (x transposed so the
// 4 rows are one 16-byte read, W rows padded to 36 floats: conflict-free 16-byte reads), the next chunk is fetched
// into registers while the current one is multiplied.  With K splits the partial sums go to a scratch buffer and
// the LAST CTA of a column tile (device counter) adds them in split order - a fixed summation order, so repeated
// runs are bit-identical - and applies bias / ReLU / residual.  The per-column-tile kernel above re-read the whole
// x for every 4 columns (200 MB of L2 traffic per layer at the benchmark shape).
// ---------------------------------------------------------------------------------------------
constexpr int TL_BN = 32, TL_KC = 32, TL_THREADS = 256, TL_WP = TL_KC + 4, TL_MAXS = 8;
struct QkvDest {      // optional epilogue of the qkv projection (fp32 engine): scatter into the padded head layouts
  float* q;
  float* k;
  float* v;
  int H, Dh, C, slot;
  const int* dstate;
};

template <typename WT>
__global__ void __launch_bounds__(TL_THREADS) decode_linear_tiled_kernel(
    const float* __restrict__ x, long long ldx, const WT* __restrict__ W, long long ldw,
    const float* __restrict__ bias, int relu, const float* __restrict__ res, long long ldr,
    float* __restrict__ out, long long ldo, int B, int N, int K, int klen, float* __restrict__ partial,
    int* __restrict__ counters, QkvDest qd) {
  __shared__ __align__(16) float xs[2][TL_KC][64];
  __shared__ __align__(16) float ws[2][TL_BN][TL_WP];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;           // rows 4*ty .. +3, columns tx and tx + 16 of the tile
  const int n0 = blockIdx.x * TL_BN;
  const int S = gridDim.y, ks = blockIdx.y;
  const int k_begin = ks * klen, k_end = min(K, k_begin + klen);
  const int nchunks = (k_end - k_begin + TL_KC - 1) / TL_KC;
  // loader roles: x: rows (tid & 63), 16-byte column groups (tid >> 6) and +4; W: column tid >> 3, group tid & 7
  const int xr = tid & 63, xk = tid >> 6;
  const int wn = tid >> 3, wk = tid & 7;
  float4 px[2], pw;
  auto fetch = [&](int c) {
    const int k0 = k_begin + c * TL_KC;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + 4 * (xk + 4 * i);
      px[i] = (xr < B && k < k_end) ? *reinterpret_cast<const float4*>(x + (long long)xr * ldx + k) : make_float4(0, 0, 0, 0);
    }
    const int k = k0 + 4 * wk;
    float w4[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 + wn < N && k < k_end) Elem<WT>::load4(W + (long long)(n0 + wn) * ldw + k, w4);
    pw = make_float4(w4[0], w4[1], w4[2], w4[3]);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kk = 4 * (xk + 4 * i);
      xs[buf][kk][xr] = px[i].x; xs[buf][kk + 1][xr] = px[i].y; xs[buf][kk + 2][xr] = px[i].z; xs[buf][kk + 3][xr] = px[i].w;
    }
    *reinterpret_cast<float4*>(&ws[buf][wn][4 * wk]) = pw;
  };
  float acc[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
  if (nchunks > 0) {
    fetch(0);
    stash(0);
  }
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) fetch(c + 1);
#pragma unroll
    for (int k4 = 0; k4 < TL_KC / 4; ++k4) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[buf][tx][4 * k4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[buf][tx + 16][4 * k4]);
      const float wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[buf][4 * k4 + e][4 * ty]);
        acc[0][0] = fmaf(xv.x, wa[e], acc[0][0]); acc[0][1] = fmaf(xv.x, wb[e], acc[0][1]);
        acc[1][0] = fmaf(xv.y, wa[e], acc[1][0]); acc[1][1] = fmaf(xv.y, wb[e], acc[1][1]);
        acc[2][0] = fmaf(xv.z, wa[e], acc[2][0]); acc[2][1] = fmaf(xv.z, wb[e], acc[2][1]);
        acc[3][0] = fmaf(xv.w, wa[e], acc[3][0]); acc[3][1] = fmaf(xv.w, wb[e], acc[3][1]);
      }
    }
    if (c + 1 < nchunks) stash(buf ^ 1);
    __syncthreads();
  }
  const int slot = qd.q ? (qd.dstate ? qd.dstate[0] : qd.slot) : 0;
  auto finish = [&](float v, int b, int n) {
    if (b < B && n < N) {
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (res) v += res[(long long)b * ldr + n];
      if (qd.q) {   // qkv_net: q -> staged query [B,H,64], k / v -> ring slot of the caches [B,H,C,64] (head dim padded to 64)
        const int hd = qd.H * qd.Dh;
        const int which = n / hd, hh = (n % hd) / qd.Dh, e = n % qd.Dh;
        if (which == 0) qd.q[((long long)b * qd.H + hh) * 64 + e] = v;
        else (which == 1 ? qd.k : qd.v)[(((long long)b * qd.H + hh) * qd.C + slot) * 64 + e] = v;
      } else {
        out[(long long)b * ldo + n] = v;
      }
    }
  };
  if (S == 1) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      finish(acc[r][0], 4 * ty + r, n0 + tx);
      finish(acc[r][1], 4 * ty + r, n0 + tx + 16);
    }
    return;
  }
  // partial[ks][row][tile column]: [S][gridDim.x][64][32]
  float* mine = partial + (((long long)ks * gridDim.x + blockIdx.x) * 64) * TL_BN;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    mine[(4 * ty + r) * TL_BN + tx] = acc[r][0];
    mine[(4 * ty + r) * TL_BN + tx + 16] = acc[r][1];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(counters + blockIdx.x, 1);
    is_last = prev == S - 1;
    if (is_last) counters[blockIdx.x] = 0;          // ready for the next launch (launches are stream-ordered)
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int row = 4 * ty + r, col = tx + 16 * cc;
      // all partials of this output are requested before the first add (one L2 round trip instead of S), then added
      // in split order
      float pv[TL_MAXS];
#pragma unroll
      for (int s2 = 0; s2 < TL_MAXS; ++s2)
        pv[s2] = s2 < S ? __ldcg(partial + (((long long)s2 * gridDim.x + blockIdx.x) * 64 + row) * TL_BN + col) : 0.f;
      float v = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < TL_MAXS; ++s2) v += pv[s2];
      finish(v, row, n0 + col);
    }
}

// src f32 [rows, ld_src] (+ col_off) with H heads of Dh columns -> dst[row*rs + h*hs + off + e],
// e < 64, zero-padded beyond Dh.  Serves q staging, K/V ring append and the R table.
template <typename CT>
__global__ void pad_heads_kernel(const float* __restrict__ src, long long ld_src, int col_off, int rows,
                                 int H, int Dh, CT* __restrict__ dst, long long rs, long long hs, long long off,
                                 const int* __restrict__ dstate) {
  if (dstate) off = (long long)dstate[0] * 64;   // ring slot kept on the device (CUDA-graph replay)
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * H * 64) return;
  const int e = idx & 63;
  const int h = (idx >> 6) % H;
  const int r = idx / (64LL * H);
  const float v = e < Dh ? src[(long long)r * ld_src + col_off + h * Dh + e] : 0.f;
  Elem<CT>::store(dst + r * rs + h * hs + off + e, v);
}

// ---------------------------------------------------------------------------------------------
// single-query relative attention over the ring cache
//   score_a = scale * ( (q+u).k_a + (q+vb).R[a] ),  a = age (0 = current token) < n_vis
// One CTA per (h, b); 8 lanes share a key (8 dims each), 4 keys per warp iteration.
// ---------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 8;
constexpr int DA_UNROLL = 4;
template <typename CT>
__global__ void __launch_bounds__(DA_WARPS * 32, 2) decode_attn_kernel(
    const float* __restrict__ q, const CT* __restrict__ kc, const CT* __restrict__ vc,
    const CT* __restrict__ rt, const float* __restrict__ u, const float* __restrict__ vb, int H, int C,
    int n_vis, int cur_slot, float scale, float* __restrict__ out, long long ldo, const int* __restrict__ dstate,
    bf16* __restrict__ out_bf16, float* __restrict__ partial, int* __restrict__ counters) {
  __shared__ float sh_m[DA_WARPS], sh_l[DA_WARPS], sh_o[DA_WARPS][64];
  __shared__ int is_last;
  if (dstate) {
    cur_slot = dstate[0];
    n_vis = dstate[1];
  }
  const int h = blockIdx.x, b = blockIdx.y;
  // key splits (gridDim.z): B*H CTAs do not fill the 2 x 148 resident slots evenly (512 = 1.73 waves); with 4 splits
  // the grid is 6.9 waves.  The partial (max, sum, out) triples are merged in split order by the last CTA of a (b, h).
  const int nsplit = gridDim.z, zi = blockIdx.z;
  const int chunk = ((n_vis + nsplit - 1) / nsplit + 31) & ~31;
  const int a_begin = zi * chunk, a_end = min(n_vis, a_begin + chunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, part = lane & 7;
  float qu[8], qv[8];
  {
    const float* qp = q + ((long long)b * H + h) * 64 + part * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      qu[e] = qp[e] + u[h * 64 + part * 8 + e];
      qv[e] = qp[e] + vb[h * 64 + part * 8 + e];
    }
  }
  const CT* kbase = kc + ((long long)b * H + h) * C * 64 + part * 8;
  const CT* vbase = vc + ((long long)b * H + h) * C * 64 + part * 8;
  const CT* rbase = rt + (long long)h * 64 + part * 8;
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;
  // DA_UNROLL keys per lane group and iteration: every load of the batch is issued before the first use, so one
  // thread keeps 3 * DA_UNROLL 32-byte reads in flight (the stream is HBM-latency bound otherwise: 52 % -> of peak)
  for (int a0 = a_begin + warp * 4; a0 < a_end; a0 += DA_WARPS * 4 * DA_UNROLL) {
    float kk[DA_UNROLL][8], rr[DA_UNROLL][8], vv[DA_UNROLL][8];
    bool ok[DA_UNROLL];
#pragma unroll
    for (int uu = 0; uu < DA_UNROLL; ++uu) {
      const int a = a0 + uu * DA_WARPS * 4 + sub;
      ok[uu] = a < a_end;
#pragma unroll
      for (int e = 0; e < 8; ++e) kk[uu][e] = rr[uu][e] = vv[uu][e] = 0.f;
      if (ok[uu]) {
        int slot = cur_slot - a;
        if (slot < 0) slot += C;
        Elem<CT>::load8(kbase + (long long)slot * 64, kk[uu]);
        Elem<CT>::load8(rbase + (long long)a * H * 64, rr[uu]);
        Elem<CT>::load8(vbase + (long long)slot * 64, vv[uu]);
      }
    }
#pragma unroll
    for (int uu = 0; uu < DA_UNROLL; ++uu) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qu[e], kk[uu][e], fmaf(qv[e], rr[uu][e], d));
      // reduce the partial dot over the 8 lanes of the key (inactive keys carry -inf)
      float t = ok[uu] ? d : 0.f;
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      const float s = ok[uu] ? t * sl2 : -INFINITY;
      const float mn = fmaxf(m, s);
      const float msafe = mn == -INFINITY ? 0.f : mn;
      const float corr = exp2f(m - msafe);
      const float p = exp2f(s - msafe);
      l = l * corr + p;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(p, vv[uu][e], o[e] * corr);
      m = mn;
    }
  }
  // combine the 4 key sub-groups of the warp
#pragma unroll
  for (int off = 8; off < 32; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float msafe = mn == -INFINITY ? 0.f : mn;
    const float c1 = exp2f(m - msafe), c2 = exp2f(m2 - msafe);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[e], off);
      o[e] = o[e] * c1 + o2 * c2;
    }
    l = l * c1 + l2 * c2;
    m = mn;
  }
  if (sub == 0) {
    if (part == 0) {
      sh_m[warp] = m;
      sh_l[warp] = l;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sh_o[warp][part * 8 + e] = o[e];
  }
  __syncthreads();
  float mm = -INFINITY, ll = 0.f, oo = 0.f;
  if (threadIdx.x < 64) {
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) mm = fmaxf(mm, sh_m[w]);
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float c = sh_m[w] == -INFINITY ? 0.f : exp2f(sh_m[w] - mm);
      ll += sh_l[w] * c;
      oo += sh_o[w][threadIdx.x] * c;
    }
  }
  const long long orow = (long long)b * ldo + h * 64 + threadIdx.x;
  if (nsplit == 1) {
    if (threadIdx.x < 64) {
      out[orow] = oo / ll;
      if (out_bf16) out_bf16[orow] = __float2bfloat16_rn(oo / ll);
    }
    return;
  }
  const int bh = b * H + h;
  float* mine = partial + ((long long)bh * nsplit + zi) * 66;
  if (threadIdx.x < 64) {
    mine[threadIdx.x] = oo;
    if (threadIdx.x == 0) { mine[64] = mm; mine[65] = ll; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(counters + bh, 1);
    is_last = prev == nsplit - 1;
    if (is_last) counters[bh] = 0;
  }
  __syncthreads();
  if (!is_last || threadIdx.x >= 64) return;
  __threadfence();
  float m_all = -INFINITY;
  for (int z = 0; z < nsplit; ++z) m_all = fmaxf(m_all, __ldcg(partial + ((long long)bh * nsplit + z) * 66 + 64));
  float l_all = 0.f, o_all = 0.f;
  for (int z = 0; z < nsplit; ++z) {
    const float* pz = partial + ((long long)bh * nsplit + z) * 66;
    const float mz = __ldcg(pz + 64);
    const float c = mz == -INFINITY ? 0.f : exp2f(mz - m_all);
    l_all += __ldcg(pz + 65) * c;
    o_all += __ldcg(pz + threadIdx.x) * c;
  }
  out[orow] = o_all / l_all;
  if (out_bf16) out_bf16[orow] = __float2bfloat16_rn(o_all / l_all);
}

// ---------------------------------------------------------------------------------------------
// sampler: one CTA (1024 threads) per sequence.  Reference semantics of calc_probs / apply_sampling
// / infer_token (midi_inferrer.py:209-237) + top-p (new capability, BASELINE config 4).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
constexpr int SAMP_N = 1024;

__global__ void __launch_bounds__(SAMP_N) sampler_kernel(
    const float* __restrict__ logits, long long ld, int V, float temperature, int top_k, float top_p,
    const unsigned char* __restrict__ wrong, unsigned long long seed, unsigned long long offset,
    long long* __restrict__ tokens, float* __restrict__ probs_out, long long ldp, const int* __restrict__ dstate) {
  __shared__ float key[SAMP_N];
  if (dstate) offset += (unsigned long long)dstate[3];
  __shared__ int idx[SAMP_N];
  __shared__ float red[32];
  __shared__ float prob[SAMP_N];
  __shared__ float keep[SAMP_N];
  const int row = blockIdx.x, t = threadIdx.x;
  const float* lg = logits + (long long)row * ld;
  const bool live = t >= 1 && t < V;  // token 0 is never sampled (midi_inferrer.py:206, :220)

  auto block_reduce = [&](float v, bool is_max) -> float {
    for (int o = 16; o > 0; o >>= 1) {
      const float w = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, w) : v + w;
    }
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < SAMP_N / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
  };

  float p = 0.f;
  if (temperature == 0.f) {
    // one-hot at the first arg-max of logits[1:]
    const float v = live ? lg[t] : -INFINITY;
    const float mx = block_reduce(v, true);
    key[t] = (live && v == mx) ? (float)t : 1e9f;
    __syncthreads();
    for (int s = SAMP_N / 2; s > 0; s >>= 1) {
      if (t < s) key[t] = fminf(key[t], key[t + s]);
      __syncthreads();
    }
    p = (live && (float)t == key[0]) ? 1.f : 0.f;
    __syncthreads();
  } else {
    const float v = live ? lg[t] / temperature : -INFINITY;
    const float mx = block_reduce(v, true);
    const float e = live ? expf(v - mx) : 0.f;
    const float sum = block_reduce(e, false);
    p = e / sum;
  }
  prob[t] = p;
  // sort (prob desc, index asc) with a bitonic network over 1024 slots
  key[t] = live ? p : -1.f;
  idx[t] = t;
  __syncthreads();
  for (int k2 = 2; k2 <= SAMP_N; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const int ixj = t ^ j;
      if (ixj > t) {
        const bool desc = (t & k2) == 0;
        const float a = key[t], bb = key[ixj];
        const int ia = idx[t], ib = idx[ixj];
        const bool a_first = (a > bb) || (a == bb && ia < ib);  // a should precede b in desc order
        if (desc ? !a_first : a_first) {
          key[t] = bb; key[ixj] = a; idx[t] = ib; idx[ixj] = ia;
        }
      }
      __syncthreads();
    }
  }
  // keep mask in sorted order: top-k keeps ranks < k; top-p keeps ranks whose preceding mass < top_p
  {
    // inclusive scan of sorted probabilities (Hillis-Steele in shared memory via `keep` as scratch)
    keep[t] = key[t] > 0.f ? key[t] : 0.f;
    __syncthreads();
    for (int o = 1; o < SAMP_N; o <<= 1) {
      const float add = t >= o ? keep[t - o] : 0.f;
      __syncthreads();
      keep[t] += add;
      __syncthreads();
    }
    const float pk = key[t] > 0.f ? key[t] : 0.f;
    const float before = keep[t] - pk;
    bool kp = key[t] >= 0.f;
    if (top_k > 0) kp = kp && (t < top_k);
    if (top_p > 0.f) kp = kp && (before < top_p);
    __syncthreads();
    const int tokid = idx[t];
    if (wrong && tokid < V && wrong[(long long)row * V + tokid]) kp = false;
    keep[tokid] = kp ? 1.f : 0.f;   // scatter back to token order (idx is a permutation)
    __syncthreads();
  }
  const float pm = prob[t] * keep[t];
  const float tot = block_reduce(pm, false);
  const float pf = pm / tot;
  if (probs_out && t < V) probs_out[(long long)row * ldp + t] = pf;
  if (tokens) {
    // inverse-CDF sampling in token order with a counter-based uniform (seed, offset, row)
    prob[t] = pf;
    __syncthreads();
    for (int o = 1; o < SAMP_N; o <<= 1) {
      const float add = t >= o ? prob[t - o] : 0.f;
      __syncthreads();
      prob[t] += add;
      __syncthreads();
    }
    uint32_t hsh = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9e3779b9U));
    hsh = mix32(hsh ^ mix32((uint32_t)offset + 0x85ebca6bU) ^ mix32((uint32_t)(offset >> 32) ^ (uint32_t)row * 0xc2b2ae35U));
    const float uni = ((hsh >> 8) + 0.5f) * (1.0f / 16777216.0f);
    // first token whose inclusive CDF exceeds u and that has non-zero probability
    const float cdf = prob[t], prev = t > 0 ? prob[t - 1] : 0.f;
    key[t] = (pf > 0.f && uni < cdf && uni >= prev) ? (float)t : 1e9f;
    __syncthreads();
    for (int s = SAMP_N / 2; s > 0; s >>= 1) {
      if (t < s) key[t] = fminf(key[t], key[t + s]);
      __syncthreads();
    }
    if (t == 0) {
      long long tok = (long long)key[0];
      if (key[0] > 1e8f) {  // numerical corner: u beyond the last cdf step -> last non-zero token
        tok = -1;
      }
      tokens[row] = tok;
    }
    if (key[0] > 1e8f) {
      __syncthreads();
      key[t] = pf > 0.f ? (float)(SAMP_N - t) : 1e9f;  // pick the largest index with pf > 0
      __syncthreads();
      for (int s = SAMP_N / 2; s > 0; s >>= 1) {
        if (t < s) key[t] = fminf(key[t], key[t + s]);
        __syncthreads();
      }
      if (t == 0) tokens[row] = key[0] > 1e8f ? -1 : (long long)(SAMP_N - (int)key[0]);
    }
  }
}

// state = {slot, n_vis, cached, step}: advanced once at the start of every decode step
__global__ void decode_advance_kernel(int* state, int C, int mem_len, int extra_visible) {
  const int slot = (state[0] + 1) % C;
  const int cached = min(state[2], mem_len);
  state[0] = slot;
  state[1] = min(cached + 1, mem_len + extra_visible);
  state[2] = min(cached + 1, mem_len);
  state[3] = state[3] + 1;
}

}  // namespace

extern "C" {

// Device-resident decode bookkeeping for CUDA-graph replay: state = int32[4] {slot, n_vis, cached, step}
// (initialise to {-1, 0, 0, 0}).  extra_visible = 0 for same_length models, 1 otherwise.
int commu_decode_advance(int* state, int C, int mem_len, int extra_visible, void* stream) {
  CB_REQUIRE(state && C > 0 && mem_len > 0, "decode_advance: bad args");
  decode_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, C, mem_len, extra_visible);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out[b,n] = act(x W^T + bias) + res.  W is fp32 (w_bf16 = 0) or bf16 (w_bf16 = 1), [N, K] row-major.
// Replaces the nn.Linear calls of the reference on the T = 1 decode step.
int commu_decode_linear(const float* x, int64_t ldx, const void* w, int64_t ldw, int w_bf16, const float* bias,
                        int relu, const float* res, int64_t ldr, float* out, int64_t ldo, int B, int N, int K,
                        void* stream) {
  CB_REQUIRE(x && w && out && B >= 1 && B <= 64 && N > 0 && K > 0, "decode_linear: bad args (B=%d must be <= 64)", B);
  CB_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, "decode_linear: K, ldx, ldw must be multiples of 4");
  const int grid = cb_host::ceil_div(N, LIN_NT);
  cudaStream_t s = (cudaStream_t)stream;
  if (w_bf16)
    decode_linear_kernel<bf16><<<grid, LIN_WARPS * 32, 0, s>>>(x, ldx, (const bf16*)w, ldw, bias, relu, res, ldr, out, ldo, B, N, K);
  else
    decode_linear_kernel<float><<<grid, LIN_WARPS * 32, 0, s>>>(x, ldx, (const float*)w, ldw, bias, relu, res, ldr, out, ldo, B, N, K);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Same contract on the register-tiled kernel (fp32 engine).  splits >= 1 K splits (0 = chosen here so that the grid
// fills the SMs); scratch: fp32 [splits * ceil(N / 32) * 64 * 32], counters: int32 [ceil(N / 32)] zero-initialised
// once (both may be NULL when splits == 1).
int commu_decode_linear_tiled(const float* x, int64_t ldx, const void* w, int64_t ldw, int w_bf16, const float* bias,
                              int relu, const float* res, int64_t ldr, float* out, int64_t ldo, int B, int N, int K,
                              int splits, float* scratch, int* counters, void* stream) {
  CB_REQUIRE(x && w && out && B >= 1 && B <= 64 && N > 0 && K > 0, "decode_linear_tiled: bad args (B=%d must be <= 64)", B);
  CB_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, "decode_linear_tiled: K, ldx, ldw must be multiples of 4");
  const int tiles = cb_host::ceil_div(N, TL_BN);
  int S = splits;
  if (S <= 0) {
    S = cb_host::ceil_div(2 * cb_host::num_sms(), tiles);
    if (!scratch || !counters) S = 1;
  }
  const int kchunks = cb_host::ceil_div(K, TL_KC);
  if (S > kchunks) S = kchunks;
  if (S > TL_MAXS) S = TL_MAXS;
  if (S < 1) S = 1;
  const int klen = cb_host::ceil_div(kchunks, S) * TL_KC;
  S = cb_host::ceil_div(K, klen);
  CB_REQUIRE(S == 1 || (scratch && counters), "decode_linear_tiled: K splits need scratch and counters");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(tiles, S);
  QkvDest qd = {};
  if (w_bf16)
    decode_linear_tiled_kernel<bf16><<<grid, TL_THREADS, 0, s>>>(x, ldx, (const bf16*)w, ldw, bias, relu, res, ldr, out, ldo,
                                                                 B, N, K, klen, scratch, counters, qd);
  else
    decode_linear_tiled_kernel<float><<<grid, TL_THREADS, 0, s>>>(x, ldx, (const float*)w, ldw, bias, relu, res, ldr, out,
                                                                  ldo, B, N, K, klen, scratch, counters, qd);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// qkv_net of the fp32 engine on the same kernel: [q | k | v] = x W^T (W fp32 [3*H*Dh, K], model.py:283-310 at T = 1) with
// the results scattered by the epilogue - q to the staged query [B,H,64], k / v into ring slot `slot` (or dev_state[0])
// of the fp32 caches [B,H,C,64]; the padding columns Dh..63 of all three are never written (they stay zero).
int commu_decode_qkv_tiled(const float* x, int64_t ldx, const float* w, int64_t ldw, int B, int H, int Dh, int K, float* q_out,
                           float* k_cache, float* v_cache, int C, int slot, const int* dev_state, float* scratch,
                           int* counters, void* stream) {
  CB_REQUIRE(x && w && q_out && k_cache && v_cache && B >= 1 && B <= 64 && H > 0 && Dh > 0 && Dh <= 64 && K > 0,
             "decode_qkv_tiled: bad args");
  CB_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, "decode_qkv_tiled: K, ldx, ldw must be multiples of 4");
  CB_REQUIRE(dev_state || (slot >= 0 && slot < C), "decode_qkv_tiled: bad slot");
  const int N = 3 * H * Dh;
  const int tiles = cb_host::ceil_div(N, TL_BN);
  int S = (scratch && counters) ? cb_host::ceil_div(2 * cb_host::num_sms(), tiles) : 1;
  const int kchunks = cb_host::ceil_div(K, TL_KC);
  if (S > kchunks) S = kchunks;
  if (S > TL_MAXS) S = TL_MAXS;
  if (S < 1) S = 1;
  const int klen = cb_host::ceil_div(kchunks, S) * TL_KC;
  S = cb_host::ceil_div(K, klen);
  QkvDest qd = {q_out, k_cache, v_cache, H, Dh, C, slot, dev_state};
  decode_linear_tiled_kernel<float><<<dim3(tiles, S), TL_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, w, ldw, nullptr, 0, nullptr, 0, nullptr, 0, B, N, K, klen, scratch, counters, qd);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dst[row*row_stride + h*head_stride + offset + e] = src[row, col_off + h*Dh + e] (zero for e >= Dh),
// dst element type fp32 (dst_bf16 = 0) or bf16.
int commu_pad_heads(const float* src, int64_t ld_src, int col_off, int rows, int H, int Dh, void* dst,
                    int dst_bf16, int64_t row_stride, int64_t head_stride, int64_t offset, const int* dev_state,
                    void* stream) {
  CB_REQUIRE(src && dst && rows > 0 && H > 0 && Dh > 0 && Dh <= 64, "pad_heads: bad args");
  const long long n = (long long)rows * H * 64;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dst_bf16)
    pad_heads_kernel<bf16><<<grid, 256, 0, s>>>(src, ld_src, col_off, rows, H, Dh, (bf16*)dst, row_stride, head_stride, offset, dev_state);
  else
    pad_heads_kernel<float><<<grid, 256, 0, s>>>(src, ld_src, col_off, rows, H, Dh, (float*)dst, row_stride, head_stride, offset, dev_state);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// q: fp32 [B,H,64]; kcache/vcache: [B,H,C,64]; rtab: [>= n_vis, H, 64] by distance; out fp32 [B, ldo].
// The n_vis most recent ring entries (ages 0..n_vis-1, age 0 at slot cur_slot) are attended.
int commu_decode_attn(const float* q, const void* kcache, const void* vcache, const void* rtab, int cache_bf16,
                      const float* r_w_bias, const float* r_r_bias, int B, int H, int C, int n_vis, int cur_slot,
                      float scale, float* out, int64_t ldo, const int* dev_state, void* out_bf16, int splits,
                      float* partial, int* counters, void* stream) {
  CB_REQUIRE(q && kcache && vcache && rtab && out, "decode_attn: null arg");
  CB_REQUIRE(dev_state || (n_vis >= 1 && n_vis <= C && cur_slot >= 0 && cur_slot < C),
             "decode_attn: bad args (n_vis=%d C=%d slot=%d)", n_vis, C, cur_slot);
  cudaStream_t s = (cudaStream_t)stream;
  cb_host::ProfScope prof(cb_host::PROF_DECODE_ATTN, s);
  if (splits < 1 || !partial || !counters) splits = 1;
  CB_REQUIRE(splits <= 16, "decode_attn: at most 16 key splits");
  dim3 grid(H, B, splits);
  if (cache_bf16)
    decode_attn_kernel<bf16><<<grid, DA_WARPS * 32, 0, s>>>(q, (const bf16*)kcache, (const bf16*)vcache, (const bf16*)rtab,
                                                            r_w_bias, r_r_bias, H, C, n_vis, cur_slot, scale, out, ldo, dev_state,
                                                            (bf16*)out_bf16, partial, counters);
  else
    decode_attn_kernel<float><<<grid, DA_WARPS * 32, 0, s>>>(q, (const float*)kcache, (const float*)vcache, (const float*)rtab,
                                                             r_w_bias, r_r_bias, H, C, n_vis, cur_slot, scale, out, ldo, dev_state,
                                                             (bf16*)out_bf16, partial, counters);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Sampler over B rows of raw logits [B, ld] (token 0 included; it is never sampled).
//   temperature == 0 -> greedy one-hot;  top_k > 0 keeps the k most probable;  top_p > 0 keeps the
//   smallest prefix whose mass reaches top_p;  wrong: uint8 [B,V] tokens to zero (or NULL).
//   tokens: int64 [B] sampled ids (or NULL);  probs_out: fp32 [B, ldp] final distribution (or NULL).
int commu_sample(const float* logits, int64_t ld, int B, int V, float temperature, int top_k, float top_p,
                 const unsigned char* wrong, uint64_t seed, uint64_t offset, int64_t* tokens, float* probs_out,
                 int64_t ldp, const int* dev_state, void* stream) {
  CB_REQUIRE(logits && B > 0 && V > 1 && V <= SAMP_N, "sample: vocabulary must be <= %d", SAMP_N);
  CB_REQUIRE(tokens || probs_out, "sample: no output requested");
  sampler_kernel<<<B, SAMP_N, 0, (cudaStream_t)stream>>>(logits, ld, V, temperature, top_k, top_p, wrong, seed,
                                                         offset, (long long*)tokens, probs_out, ldp, dev_state);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
