// Fused token-step kernels of the bf16 decode engine (generate.py hot loop, BASELINE configs[3]).
//
// One decoder layer of the reference's T = 1 forward (commu/model/model.py:283-352 attention block,
// :174-181 position-wise FF, driven per token by commu/midi_generator/midi_inferrer.py:199-207) is five
// launches here instead of ten:
//   1. dec_linear<EMBED|LN, QKV>   x = LN(prev layer's FF sum) (or the scaled embedding), q/k/v = x Wqkv^T,
//                                  q staged per head, k / v appended to the layer's ring cache in place
//   2. dec_attn_split              single-query relative attention over the ring cache, keys split over
//                                  CTAs (grid ~ 7 CTAs per SM), last-arriving CTA merges the partials
//   3. dec_linear<BF16, RES>       z1 = x + att Wo^T
//   4. dec_linear<LN, RELU>        y = LN(z1), h = relu(y W1^T + b1)
//   5. dec_linear<BF16, RES>       z2 = y + h W2^T + b2, K split over a thread-block cluster (DSMEM reduce)
// plus dec_linear<LN, LOGITS> for the tied output layer.  The batch (<= 64 sequences) is the M = 64 side
// of warp-level bf16 MMAs (m16n8k16, fp32 accumulate); every CTA owns 16 output columns, so a layer's
// weights are streamed from HBM exactly once per token step and the launch grid is 32..128 CTAs.  LayerNorm is
// recomputed by every CTA in its prologue (64 x d fp32 from L2) instead of being a launch of its own.
// All kernels call griddepcontrol.wait before touching activations and prefetch their weight slice
// before it, so with programmatic dependent launch the weight fetch overlaps the previous kernel's tail.
#include <cooperative_groups.h>
#include <math.h>
#include "api_common.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace cb_host {
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer);   // gemm.cu
}

namespace {

constexpr int DF_THREADS = 256;
constexpr int DF_M = 64;     // batch rows of the MMA tile (rows >= B are zero)
constexpr int DF_NT = 16;    // output columns per CTA
constexpr int DF_PAD = 8;    // bf16 elements of row padding (conflict-free ldmatrix)
constexpr int DF_PART_LD = 20;

enum { PRO_EMBED = 0, PRO_LN = 1, PRO_BF16 = 2 };
enum { EPI_QKV = 0, EPI_RES = 1, EPI_RELU = 2, EPI_LOGITS = 3 };

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

struct LinP {
  // prologue
  const long long* tokens; const float* emb; float emb_scale;
  const float* z; long long ldz; const float* gamma; const float* beta; float eps;
  const bf16* a; long long lda;
  float* x_out; long long ldx;
  int B, K, KC, N, n_pad, d_true;
  // weights
  const bf16* w; long long ldw; const float* bias;
  // epilogue
  const float* res; long long ldr;
  float* out_f32; long long ldo;
  bf16* out_bf16; long long ldob;
  float* q_out; bf16* k_cache; bf16* v_cache; int H, C, slot; const int* dstate;
};

// Embedding / LayerNorm prologue shared by a cluster of PCL = 8 CTAs (8 neighbouring column slices): CTA `rank`
// normalises rows 8*rank .. 8*rank+7 (one row per warp) and stores the bf16 row into the A tile of every CTA of the
// cluster through distributed shared memory, so the 64 x d LayerNorm is computed once per cluster instead of once
// per CTA (graph-chain probe: the per-CTA version made these kernels 9.3 us vs 3.3 us for the plain-bf16 prologue).
constexpr int PCL = 8;
template <int PRO, int MAXV>
__device__ __forceinline__ void prologue_row_cluster(const LinP& p, bf16* sA, int lds, int warp, int lane) {
  cg::cluster_group cluster = cg::this_cluster();
  const int d = p.d_true, KC = p.KC;
  const int r = (int)cluster.block_rank() * 8 + warp;
  float4 v[MAXV], gm[MAXV], bt[MAXV];
  const float* src = nullptr;
  if (r < p.B) src = PRO == PRO_EMBED ? p.emb + p.tokens[r] * (long long)d : p.z + (long long)r * p.ldz;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = 4 * (lane + 32 * i);
    v[i] = gm[i] = bt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src && c < d) v[i] = *reinterpret_cast<const float4*>(src + c);
    if (PRO == PRO_LN && c < d) {
      gm[i] = *reinterpret_cast<const float4*>(p.gamma + c);
      bt[i] = *reinterpret_cast<const float4*>(p.beta + c);
    }
  }
  if (PRO == PRO_EMBED) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      v[i].x *= p.emb_scale; v[i].y *= p.emb_scale; v[i].z *= p.emb_scale; v[i].w *= p.emb_scale;
    }
  } else {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = cb::warp_sum(s) / (float)d;
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (4 * (lane + 32 * i) < d) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        qv += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(cb::warp_sum(qv) / (float)d + p.eps);
    if (r < p.B) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (4 * (lane + 32 * i) < d) {
          v[i].x = (v[i].x - mean) * rstd * gm[i].x + bt[i].x;
          v[i].y = (v[i].y - mean) * rstd * gm[i].y + bt[i].y;
          v[i].z = (v[i].z - mean) * rstd * gm[i].z + bt[i].z;
          v[i].w = (v[i].w - mean) * rstd * gm[i].w + bt[i].w;
        }
      }
    }
  }
  const bool writer = p.x_out && r < p.B && blockIdx.x < PCL;     // the first cluster also emits the fp32 rows
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < KC) {
      uint2 pk;
      pk.x = cb::pack_bf16(v[i].x, v[i].y);
      pk.y = cb::pack_bf16(v[i].z, v[i].w);
#pragma unroll
      for (int dr = 0; dr < PCL; ++dr) *reinterpret_cast<uint2*>(cluster.map_shared_rank(sA, dr) + r * lds + c) = pk;
    }
    if (writer && c < d) *reinterpret_cast<float4*>(p.x_out + (long long)r * p.ldx + c) = v[i];
  }
}

// A tile [64][KC + 8] bf16, W slice [16][KC + 8] bf16, partial sums [2][64][20] fp32
template <int PRO, int EPI, bool CLUSTER>
__global__ void __launch_bounds__(DF_THREADS) dec_linear_kernel(const LinP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KC = p.KC, lds = KC + DF_PAD;
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sW = sA + DF_M * lds;
  float* sP = reinterpret_cast<float*>(sW + DF_NT * lds);   // [2][64][DF_PART_LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * DF_NT;
  const int k0 = CLUSTER ? blockIdx.y * KC : 0;
  constexpr bool PROC = PRO != PRO_BF16;          // prologue shared by a cluster of PCL CTAs along x
  const bool active = n0 < p.n_pad;               // the grid is padded to a multiple of PCL

  // ---- weight slice (independent of the previous kernel): cp.async, 16 B per request ----
  if (active) {
    const bf16* wsrc = p.w + (long long)n0 * p.ldw + k0;
    for (int r = warp; r < DF_NT; r += DF_THREADS / 32)
      for (int kk = lane * 8; kk < KC; kk += 256) cb::cp_async16(sW + r * lds + kk, wsrc + (long long)r * p.ldw + kk, true);
    cb::cp_async_commit();
  }
  // epilogue operands of this thread (row, 4 columns): parameters now, activations right after the wait, so
  // that their latency hides behind the tile loads and the MMAs
  const int row = tid >> 2, cq = (tid & 3) * 4;
  float ebias[4] = {0.f, 0.f, 0.f, 0.f}, eres[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI != EPI_QKV && p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + cq + j < p.N) ebias[j] = p.bias[n0 + cq + j];
  }
  pdl_wait();
  pdl_launch();
  if (EPI == EPI_RES && row < p.B && (!CLUSTER || blockIdx.y == 0)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + cq + j < p.N) eres[j] = p.res[(long long)row * p.ldr + n0 + cq + j];
  }

  // ---- A tile ----
  if (PRO == PRO_BF16) {
    const bf16* asrc = p.a + k0;
    for (int r = warp; r < DF_M; r += DF_THREADS / 32) {
      const bool ok = r < p.B;
      for (int kk = lane * 8; kk < KC; kk += 256) cb::cp_async16(sA + r * lds + kk, asrc + (long long)(ok ? r : 0) * p.lda + kk, ok);
    }
    cb::cp_async_commit();
  } else {
    if (p.d_true <= 512) prologue_row_cluster<PRO, 4>(p, sA, lds, warp, lane);
    else prologue_row_cluster<PRO, 8>(p, sA, lds, warp, lane);
  }
  cb::cp_async_wait<0>();
  if (PROC) {
    cg::this_cluster().sync();                    // every CTA's A tile is complete (no remote access after this point)
    if (!active) return;
  } else {
    __syncthreads();
  }

  // ---- MMA: warp = (m tile, K half); both 8-column n tiles per warp ----
  {
    const int mt = warp & 3, kh = warp >> 2;
    const int ksteps = KC >> 4;
    const int ks0 = kh * (ksteps >> 1), ks1 = kh ? ksteps : (ksteps >> 1);
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t aBase = cb::smem_u32(sA + (mt * 16 + (lane & 15)) * lds + (lane >> 4) * 8);
    const uint32_t bBase = cb::smem_u32(sW + ((lane & 7) + (lane >> 4) * 8) * lds + ((lane >> 3) & 1) * 8);
#pragma unroll 4
    for (int ks = ks0; ks < ks1; ++ks) {
      uint32_t a[4], b[4];
      cb::ldmatrix_x4(a, aBase + ks * 32);
      cb::ldmatrix_x4(b, bBase + ks * 32);
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      cb::mma_bf16_16816(acc0, a, b0);
      cb::mma_bf16_16816(acc1, a, b1);
    }
    float* dst = sP + kh * DF_M * DF_PART_LD;
    const int r = mt * 16 + (lane >> 2), c = 2 * (lane & 3);
    dst[r * DF_PART_LD + c] = acc0[0];
    dst[r * DF_PART_LD + c + 1] = acc0[1];
    dst[(r + 8) * DF_PART_LD + c] = acc0[2];
    dst[(r + 8) * DF_PART_LD + c + 1] = acc0[3];
    dst[r * DF_PART_LD + 8 + c] = acc1[0];
    dst[r * DF_PART_LD + 8 + c + 1] = acc1[1];
    dst[(r + 8) * DF_PART_LD + 8 + c] = acc1[2];
    dst[(r + 8) * DF_PART_LD + 8 + c + 1] = acc1[3];
  }
  __syncthreads();

  // ---- epilogue: thread = (row, 4 consecutive columns) ----
  float v[4];
  {
    const float4 a = *reinterpret_cast<const float4*>(sP + row * DF_PART_LD + cq);
    const float4 b = *reinterpret_cast<const float4*>(sP + DF_M * DF_PART_LD + row * DF_PART_LD + cq);
    v[0] = a.x + b.x; v[1] = a.y + b.y; v[2] = a.z + b.z; v[3] = a.w + b.w;
  }
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    // every CTA of the cluster publishes its [64][16] sum in the first partial buffer; rank 0 adds them in rank order
    __syncthreads();
    *reinterpret_cast<float4*>(sP + row * DF_PART_LD + cq) = make_float4(v[0], v[1], v[2], v[3]);
    cluster.sync();
    if (cluster.block_rank() == 0) {
      const unsigned nr = cluster.num_blocks();
      for (unsigned r = 1; r < nr; ++r) {
        const float* remote = cluster.map_shared_rank(sP, r);
        const float4 o = *reinterpret_cast<const float4*>(remote + row * DF_PART_LD + cq);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
    }
    cluster.sync();
    if (cluster.block_rank() != 0) return;
  }
  if (row >= p.B) return;
  const int n = n0 + cq;
  if (EPI == EPI_QKV) {
    // padded column layout [3][H][64]
    const int hw = p.H * 64;
    const int which = n / hw, rem = n - which * hw;
    const int h = rem >> 6, e = rem & 63;
    if (which == 0) {
      *reinterpret_cast<float4*>(p.q_out + ((long long)row * p.H + h) * 64 + e) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      const int slot = p.dstate ? p.dstate[0] : p.slot;
      bf16* dst = (which == 1 ? p.k_cache : p.v_cache) + (((long long)row * p.H + h) * p.C + slot) * 64 + e;
      uint2 pk;
      pk.x = cb::pack_bf16(v[0], v[1]);
      pk.y = cb::pack_bf16(v[2], v[3]);
      *reinterpret_cast<uint2*>(dst) = pk;
    }
  } else if (EPI == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += ebias[j];
    uint2 pk;
    pk.x = cb::pack_bf16(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f));
    pk.y = cb::pack_bf16(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
    *reinterpret_cast<uint2*>(p.out_bf16 + (long long)row * p.ldob + n) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < p.N) p.out_f32[(long long)row * p.ldo + n + j] = v[j] + ebias[j] + eres[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// single-query relative attention over the bf16 ring cache, keys split over gridDim.z CTAs
//   score_a = scale * ( (q+u).k_a + (q+vb).R[a] ),  a = age (0 = current token) < n_vis
// 8 lanes share a key (8 dims each), 16 keys per warp iteration (4 independent load groups in flight).
// ---------------------------------------------------------------------------------------------
constexpr int DA2_WARPS = 8;
constexpr int DA2_U = 4;

__device__ __forceinline__ uint4 ld_stream(const bf16* p) {   // K / V rows: read once, keep them out of L1
  uint4 a;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
  return a;
}
__device__ __forceinline__ float dot8(const uint4 a, const float (&x)[8], float d) {
  d = fmaf(x[0], cb::bf16_lo(a.x), d); d = fmaf(x[1], cb::bf16_hi(a.x), d);
  d = fmaf(x[2], cb::bf16_lo(a.y), d); d = fmaf(x[3], cb::bf16_hi(a.y), d);
  d = fmaf(x[4], cb::bf16_lo(a.z), d); d = fmaf(x[5], cb::bf16_hi(a.z), d);
  d = fmaf(x[6], cb::bf16_lo(a.w), d); d = fmaf(x[7], cb::bf16_hi(a.w), d);
  return d;
}
__device__ __forceinline__ void axpy8(float pr, const uint4 a, float (&o)[8]) {
  o[0] = fmaf(pr, cb::bf16_lo(a.x), o[0]); o[1] = fmaf(pr, cb::bf16_hi(a.x), o[1]);
  o[2] = fmaf(pr, cb::bf16_lo(a.y), o[2]); o[3] = fmaf(pr, cb::bf16_hi(a.y), o[3]);
  o[4] = fmaf(pr, cb::bf16_lo(a.z), o[4]); o[5] = fmaf(pr, cb::bf16_hi(a.z), o[5]);
  o[6] = fmaf(pr, cb::bf16_lo(a.w), o[6]); o[7] = fmaf(pr, cb::bf16_hi(a.w), o[7]);
}

__global__ void __launch_bounds__(DA2_WARPS * 32, 3) dec_attn_split_kernel(
    const float* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
    const bf16* __restrict__ rt, const float* __restrict__ u, const float* __restrict__ vb, int H, int C,
    int n_vis, int cur_slot, float scale, float* __restrict__ partial, int* __restrict__ counters,
    bf16* __restrict__ out_bf16, float* __restrict__ out_f32, long long ldo, const int* __restrict__ dstate) {
  __shared__ float sh_m[DA2_WARPS], sh_l[DA2_WARPS], sh_o[DA2_WARPS][64];
  __shared__ int sh_last;
  pdl_wait();
  pdl_launch();
  if (dstate) {
    cur_slot = dstate[0];
    n_vis = dstate[1];
  }
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, S = gridDim.z;
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
  const bf16* kbase = kc + ((long long)b * H + h) * C * 64 + part * 8;
  const bf16* vbase = vc + ((long long)b * H + h) * C * 64 + part * 8;
  const bf16* rbase = rt + (long long)h * 64 + part * 8;
  // this CTA's ages: [a_lo, a_hi)
  const int chunk = (n_vis + S - 1) / S;
  const int a_lo = sp * chunk, a_hi = min(n_vis, a_lo + chunk);
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;
  for (int a0 = a_lo + warp * (4 * DA2_U); a0 < a_hi; a0 += DA2_WARPS * 4 * DA2_U) {
    uint4 kk[DA2_U], rr[DA2_U], vv[DA2_U];
    bool ok[DA2_U];
#pragma unroll
    for (int t = 0; t < DA2_U; ++t) {
      const int a = a0 + t * 4 + sub;
      ok[t] = a < a_hi;
      const int ac = ok[t] ? a : a_lo;
      int slot = cur_slot - ac;
      if (slot < 0) slot += C;
      kk[t] = ld_stream(kbase + (long long)slot * 64);
      rr[t] = __ldg(reinterpret_cast<const uint4*>(rbase + (long long)ac * H * 64));
      vv[t] = ld_stream(vbase + (long long)slot * 64);
    }
    float s[DA2_U];
#pragma unroll
    for (int t = 0; t < DA2_U; ++t) {
      float d = dot8(kk[t], qu, dot8(rr[t], qv, 0.f));
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      s[t] = ok[t] ? d * sl2 : -INFINITY;
    }
    float mn = m;
#pragma unroll
    for (int t = 0; t < DA2_U; ++t) mn = fmaxf(mn, s[t]);
    const float msafe = mn == -INFINITY ? 0.f : mn;
    const float corr = exp2f(m - msafe);
    l *= corr;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] *= corr;
#pragma unroll
    for (int t = 0; t < DA2_U; ++t) {
      const float pr = exp2f(s[t] - msafe);
      l += pr;
      axpy8(pr, vv[t], o);
    }
    m = mn;
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
  const int bh = b * H + h;
  float mm = -INFINITY, ll = 0.f, oo = 0.f;
  if (threadIdx.x < 64) {
#pragma unroll
    for (int w = 0; w < DA2_WARPS; ++w) mm = fmaxf(mm, sh_m[w]);
#pragma unroll
    for (int w = 0; w < DA2_WARPS; ++w) {
      const float c = sh_m[w] == -INFINITY ? 0.f : exp2f(sh_m[w] - mm);
      ll += sh_l[w] * c;
      oo += sh_o[w][threadIdx.x] * c;
    }
  }
  if (S > 1) {
    // publish (m, l, o[64]) of this split; the CTA that arrives last merges all of them in split order
    float* mine = partial + ((long long)bh * S + sp) * 66;
    if (threadIdx.x < 64) {
      mine[2 + threadIdx.x] = oo;
      if (threadIdx.x == 0) {
        mine[0] = mm;
        mine[1] = ll;
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sh_last = (atomicAdd(counters + bh, 1) == S - 1);
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    if (threadIdx.x < 64) {
      const float* all = partial + (long long)bh * S * 66;
      mm = -INFINITY;
      for (int s2 = 0; s2 < S; ++s2) mm = fmaxf(mm, __ldcg(all + s2 * 66));
      ll = 0.f;
      oo = 0.f;
      for (int s2 = 0; s2 < S; ++s2) {
        const float ms = __ldcg(all + s2 * 66);
        const float c = ms == -INFINITY ? 0.f : exp2f(ms - mm);
        ll += __ldcg(all + s2 * 66 + 1) * c;
        oo += __ldcg(all + s2 * 66 + 2 + threadIdx.x) * c;
      }
    }
    if (threadIdx.x == 0) counters[bh] = 0;   // ready for the next launch
  }
  if (threadIdx.x < 64) {
    const float r = oo / ll;
    if (out_f32) out_f32[(long long)b * ldo + h * 64 + threadIdx.x] = r;
    if (out_bf16) out_bf16[(long long)b * ldo + h * 64 + threadIdx.x] = __float2bfloat16_rn(r);
  }
}

// ---------------------------------------------------------------------------------------------
// Stream-K, TMA-fed form of the tensor-core kernel (the product path).
//   * persistent grid (2 CTAs per SM): the B*H*nt slot tiles of the launch are one flat list cut into equal contiguous
//     runs, one per CTA, so every CTA streams the same bytes and the fixed costs are paid once per run; a
//     (sequence, head) covered by several runs is merged by its last-arriving CTA (fixed order -> deterministic);
//     (measured: one CTA per (sequence, head, split) reached 61 us per launch, this form 51 us in a replayed graph =
//     5.2 TB/s = 81 % of the measured copy bandwidth; sharing the R tile between 4 sequences or splitting the MMA
//     accumulate chains changed nothing, i.e. what is left is ramp-up / tail, not instruction issue or L2 traffic);
//   * a producer warp issues, per 64-slot tile, three TMA tile loads (K, V from HBM; R from L2; 128-byte swizzle done
//     by the copy engine) and one 256-byte bulk copy of the query row into a ring of STAGES stages guarded by
//     full / empty mbarriers; the four consumer warps never compute an address and never meet at a CTA barrier
//     inside a (sequence, head) (ncu of the cp.async version: 104 of 258 loop instructions were address arithmetic);
//   * tiles are aligned to ring SLOTS (C is a multiple of 64), so a tile is one in-bounds box; the relative-position
//     operand comes from the reversed, doubled table rt2[h][j] = R[h][C-1 - (j mod C)], 2C rows, where the tile
//     of slots s0.. is rows j0.. with j0 = (C-1-cur+s0) mod C - contiguous even across the age wrap.
// ---------------------------------------------------------------------------------------------
constexpr int AM_WARPS = 4;                    // consumer warps: 16 of a tile's 64 slots each
constexpr int AM_THREADS = AM_WARPS * 32;
constexpr int AM_TILE = 64;
constexpr int AM_MAT = AM_TILE * 128;          // one 64 x 64 bf16 matrix
constexpr int AM_STAGE_BYTES = 3 * AM_MAT;     // K, R, V
constexpr int TK_THREADS = AM_THREADS + 32;     // 4 consumer warps + 1 producer warp

// byte offset of 16-byte chunk `chunk` of row `row` in a 128-byte-swizzled tile (what the TMA writes)
__device__ __forceinline__ uint32_t am_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(cb::smem_u32(bar))
               : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(TK_THREADS) dec_attn_tma_kernel(
    const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
    const __grid_constant__ CUtensorMap tm_r, const float* __restrict__ q, const float* __restrict__ u,
    const float* __restrict__ vb, int BH, int H, int C, int n_vis, int cur_slot, float scale,
    float* __restrict__ partial, int max_parts, int* __restrict__ counters, bf16* __restrict__ out_bf16,
    float* __restrict__ out_f32, long long ldo, const int* __restrict__ dstate) {
  extern __shared__ unsigned char tk_raw[];
  __shared__ float sh_m[AM_WARPS], sh_l[AM_WARPS], sh_o[AM_WARPS][64];
  __shared__ int sh_last;
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (cb::smem_u32(tk_raw) + 1023u) & ~1023u;       // tiles need 1024-byte alignment (swizzle)
  unsigned char* sgen = tk_raw + (sbase - cb::smem_u32(tk_raw));
  float* s_q = reinterpret_cast<float*>(sgen + STAGES * AM_STAGE_BYTES);  // [STAGES][64]
  float* s_uv = s_q + STAGES * 64;                                        // [2][H][64]
  for (int i = tid; i < H * 64; i += TK_THREADS) {   // parameters: independent of the previous kernel
    s_uv[i] = u[i];
    s_uv[H * 64 + i] = vb[i];
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      cb::mbar_init(&full_bar[i], 1);
      cb::mbar_init(&empty_bar[i], AM_WARPS);
    }
    cb::fence_barrier_init();
    cb::tma_prefetch_desc(&tm_k);
    cb::tma_prefetch_desc(&tm_v);
    cb::tma_prefetch_desc(&tm_r);
  }
  pdl_wait();
  pdl_launch();
  __syncthreads();
  if (dstate) {
    cur_slot = dstate[0];
    n_vis = dstate[1];
  }
  // slot tiles that hold visible keys: the n_vis newest slots (cur - n_vis, cur] of the ring
  const int NTS = C >> 6;
  int lo = cur_slot - n_vis + 1;
  if (lo < 0) lo += C;
  const int t_first = lo >> 6;
  const int nt = min(NTS, ((lo & 63) + n_vis + 63) >> 6);
  const long long T = (long long)BH * nt;
  const int quota = (int)((T + gridDim.x - 1) / gridDim.x);
  const long long T0 = (long long)blockIdx.x * quota;
  const long long T1 = T0 + quota < T ? T0 + quota : T;
  const int ntiles = T1 > T0 ? (int)(T1 - T0) : 0;
  if (ntiles == 0) return;
  int bh = (int)(T0 / nt), kt = (int)(T0 - (long long)bh * nt);

  if (warp == AM_WARPS) {
    // ===== producer =====
    if (lane == 0) {
      for (int it = 0; it < ntiles; ++it) {
        const int st = it % STAGES;
        if (it >= STAGES) cb::mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);
        int tt = t_first + kt;
        if (tt >= NTS) tt -= NTS;
        const int s0 = tt << 6;
        int j0 = C - 1 - cur_slot + s0;
        if (j0 >= C) j0 -= C;
        const uint32_t dst = sbase + st * AM_STAGE_BYTES;
        cb::mbar_arrive_expect_tx(&full_bar[st], AM_STAGE_BYTES + 256);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
                     "l"(reinterpret_cast<uint64_t>(&tm_k)), "r"(cb::smem_u32(&full_bar[st])), "r"(0), "r"(bh * C + s0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst + AM_MAT),
                     "l"(reinterpret_cast<uint64_t>(&tm_r)), "r"(cb::smem_u32(&full_bar[st])), "r"(0), "r"((bh % H) * 2 * C + j0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst + 2 * AM_MAT),
                     "l"(reinterpret_cast<uint64_t>(&tm_v)), "r"(cb::smem_u32(&full_bar[st])), "r"(0), "r"(bh * C + s0) : "memory");
        bulk_load_1d(cb::smem_u32(s_q + st * 64), q + (long long)bh * 64, 256, &full_bar[st]);
        if (++kt == nt) {
          kt = 0;
          ++bh;
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const int g = lane >> 2, t = lane & 3;
  bool fresh = true;
  uint32_t aqu[4][2], aqv[4][2];
  float m = -INFINITY, l = 0.f;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  const float sl2 = scale * 1.4426950408889634f;
  const int row_b = 16 * warp + (lane & 7) + (lane >> 4) * 8, ch_b = (lane >> 3) & 1;
  const int row_v = 16 * warp + (lane & 7) + ((lane >> 3) & 1) * 8, ch_v = lane >> 4;

  for (int it = 0; it < ntiles; ++it) {
    const int stg = it % STAGES;
    cb::mbar_wait(&full_bar[stg], (it / STAGES) & 1);
    const uint32_t st = sbase + stg * AM_STAGE_BYTES;
    if (fresh) {   // first tile of a (sequence, head) in this run: query fragments (row 0 of the A operand)
      const float* qp = s_q + stg * 64;
      const float* up = s_uv + (bh % H) * 64;
      const float* vp = up + H * 64;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = 16 * ks + 8 * hf + 2 * t;
          const float q0 = qp[c], q1 = qp[c + 1];
          aqu[ks][hf] = g == 0 ? cb::pack_bf16(q0 + up[c], q1 + up[c + 1]) : 0u;
          aqv[ks][hf] = g == 0 ? cb::pack_bf16(q0 + vp[c], q1 + vp[c + 1]) : 0u;
        }
      }
      fresh = false;
    }
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bk[4], br[4];
      cb::ldmatrix_x4(bk, st + am_swz(row_b, 2 * ks + ch_b));
      cb::ldmatrix_x4(br, st + AM_MAT + am_swz(row_b, 2 * ks + ch_b));
      const uint32_t au[4] = {aqu[ks][0], 0u, aqu[ks][1], 0u}, av[4] = {aqv[ks][0], 0u, aqv[ks][1], 0u};
      const uint32_t k0[2] = {bk[0], bk[1]}, k1[2] = {bk[2], bk[3]}, r0[2] = {br[0], br[1]}, r1[2] = {br[2], br[3]};
      cb::mma_bf16_16816(acc0, au, k0);
      cb::mma_bf16_16816(acc1, au, k1);
      cb::mma_bf16_16816(acc0, av, r0);
      cb::mma_bf16_16816(acc1, av, r1);
    }
    uint32_t bvv[4][4];     // V operand fragments: independent of the softmax, fetched while the score MMAs drain
#pragma unroll
    for (int np = 0; np < 4; ++np) cb::ldmatrix_x4_trans(bvv[np], st + 2 * AM_MAT + am_swz(row_v, 2 * np + ch_v));
    // row 0 of the score tile: this lane holds slots s0 + 16*warp + {2t, 2t+1, 8+2t, 9+2t}; age = (cur - slot) mod C
    int tt = t_first + kt;
    if (tt >= NTS) tt -= NTS;
    int age0 = cur_slot - ((tt << 6) + 16 * warp + 2 * t);    // ages of the four keys: age0, age0-1, age0-8, age0-9 (mod C)
    auto vis = [&](int a) { return (a < 0 ? a + C : a) < n_vis; };
    const float s0 = vis(age0) ? acc0[0] * sl2 : -INFINITY;
    const float s1 = vis(age0 - 1) ? acc0[1] * sl2 : -INFINITY;
    const float s2 = vis(age0 - 8) ? acc1[0] * sl2 : -INFINITY;
    const float s3 = vis(age0 - 9) ? acc1[1] * sl2 : -INFINITY;
    float tm = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
    tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
    tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
    const float mn = fmaxf(m, tm);
    const float msafe = mn == -INFINITY ? 0.f : mn;
    const float corr = exp2f(m - msafe);
    const float p0 = exp2f(s0 - msafe), p1 = exp2f(s1 - msafe), p2 = exp2f(s2 - msafe), p3 = exp2f(s3 - msafe);
    l = l * corr + ((p0 + p1) + (p2 + p3));
    m = mn;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= corr;
      o[j][1] *= corr;
    }
    const uint32_t ap[4] = {cb::pack_bf16(p0, p1), 0u, cb::pack_bf16(p2, p3), 0u};
    __syncwarp();
    if (lane == 0) cb::mbar_arrive(&empty_bar[stg]);      // every operand of this stage is in registers
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const uint32_t v0[2] = {bvv[np][0], bvv[np][1]}, v1[2] = {bvv[np][2], bvv[np][3]};
      cb::mma_bf16_16816(o[2 * np], ap, v0);
      cb::mma_bf16_16816(o[2 * np + 1], ap, v1);
    }
    // ---- end of this run's share of the (sequence, head): publish ----
    if (kt == nt - 1 || it == ntiles - 1) {
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      asm volatile("bar.sync 1, 128;\n" ::: "memory");    // the previous publication has been read
      if (g == 0) {
        if (t == 0) {
          sh_m[warp] = m;
          sh_l[warp] = l;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sh_o[warp][8 * j + 2 * t] = o[j][0];
          sh_o[warp][8 * j + 2 * t + 1] = o[j][1];
        }
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      float mm = -INFINITY, ll = 0.f, oo = 0.f;
      if (tid < 64) {
#pragma unroll
        for (int w = 0; w < AM_WARPS; ++w) mm = fmaxf(mm, sh_m[w]);
#pragma unroll
        for (int w = 0; w < AM_WARPS; ++w) {
          const float c = sh_m[w] == -INFINITY ? 0.f : exp2f(sh_m[w] - mm);
          ll += sh_l[w] * c;
          oo += sh_o[w][tid] * c;
        }
      }
      // runs covering this (sequence, head): CTAs first .. last
      const int first = (int)(((long long)bh * nt) / quota), last = (int)(((long long)(bh + 1) * nt - 1) / quota);
      const int parts = last - first + 1;
      bool write = true;
      if (parts > 1) {
        float* mine = partial + ((long long)bh * max_parts + ((int)blockIdx.x - first)) * 66;
        if (tid < 64) {
          mine[2 + tid] = oo;
          if (tid == 0) {
            mine[0] = mm;
            mine[1] = ll;
          }
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (tid == 0) sh_last = (atomicAdd(counters + bh, 1) == parts - 1);
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        write = sh_last != 0;
        if (write) {
          __threadfence();
          if (tid < 64) {
            const float* all = partial + (long long)bh * max_parts * 66;
            mm = -INFINITY;
            for (int s2 = 0; s2 < parts; ++s2) mm = fmaxf(mm, __ldcg(all + s2 * 66));
            ll = 0.f;
            oo = 0.f;
            for (int s2 = 0; s2 < parts; ++s2) {
              const float ms = __ldcg(all + s2 * 66);
              const float c = ms == -INFINITY ? 0.f : exp2f(ms - mm);
              ll += __ldcg(all + s2 * 66 + 1) * c;
              oo += __ldcg(all + s2 * 66 + 2 + tid) * c;
            }
          }
          if (tid == 0) counters[bh] = 0;
        }
      }
      if (write && tid < 64) {
        const float r = oo / ll;
        const long long off = (long long)(bh / H) * ldo + (bh % H) * 64 + tid;
        if (out_f32) out_f32[off] = r;
        if (out_bf16) out_bf16[off] = __float2bfloat16_rn(r);
      }
      m = -INFINITY;
      l = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = 0.f;
      fresh = true;
    }
    if (++kt == nt) {
      kt = 0;
      ++bh;
    }
  }
}

// tensor maps of the decode caches are encoded once per (pointer, rows) and reused
struct TmapSlot { const void* ptr; uint64_t rows; CUtensorMap map; };
static TmapSlot g_tmaps[256];
static int g_ntmaps = 0;
int cached_tmap(const void* ptr, uint64_t rows, const CUtensorMap** out) {
  for (int i = 0; i < g_ntmaps; ++i)
    if (g_tmaps[i].ptr == ptr && g_tmaps[i].rows == rows) {
      *out = &g_tmaps[i].map;
      return 0;
    }
  TmapSlot& s = g_tmaps[g_ntmaps < 256 ? g_ntmaps : 255];
  const int rc = cb_host::make_tmap_bf16_2d(&s.map, ptr, 64, rows, 64, 64, 64);
  if (rc) return rc;
  s.ptr = ptr;
  s.rows = rows;
  if (g_ntmaps < 256) ++g_ntmaps;
  *out = &s.map;
  return 0;
}

template <int PRO, int EPI, bool CLUSTER>
int launch_linear(const LinP& p, int grid_x, int split, size_t smem, int pdl, cudaStream_t s) {
  auto kern = dec_linear_kernel<PRO, EPI, CLUSTER>;
  static size_t configured = 0;
  if (smem > configured) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  constexpr bool PROC = PRO != PRO_BF16;
  if (PROC) grid_x = cb_host::ceil_div(grid_x, PCL) * PCL;
  cfg.gridDim = dim3(grid_x, CLUSTER ? split : 1, 1);
  cfg.blockDim = dim3(DF_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CLUSTER || PROC) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = PROC ? PCL : 1;
    attrs[na].val.clusterDim.y = CLUSTER ? split : 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  cb_host::count_launch();
  return 0;
}

}  // namespace

extern "C" {

int commu_decode_fused_linear(const CommuDecLinear* a, void* stream) {
  CB_REQUIRE(a && a->w, "decode_fused_linear: null args");
  CB_REQUIRE(a->B >= 1 && a->B <= DF_M, "decode_fused_linear: B=%d must be in [1, 64]", a->B);
  CB_REQUIRE(a->K > 0 && a->K % 64 == 0 && a->N > 0, "decode_fused_linear: K=%d must be a positive multiple of 64", a->K);
  const int split = a->split_k > 0 ? a->split_k : 1;
  CB_REQUIRE(split == 1 || split == 2 || split == 4 || split == 8, "decode_fused_linear: split_k must be 1, 2, 4 or 8");
  CB_REQUIRE(a->K % (split * 32) == 0, "decode_fused_linear: K=%d not divisible by 32 * split_k", a->K);
  const int KC = a->K / split;
  CB_REQUIRE(KC <= 1024, "decode_fused_linear: K / split_k = %d exceeds 1024", KC);
  CB_REQUIRE(a->ldw % 8 == 0, "decode_fused_linear: ldw must be a multiple of 8");
  LinP p = {};
  p.tokens = (const long long*)a->tokens; p.emb = a->emb; p.emb_scale = a->emb_scale;
  p.z = a->z; p.ldz = a->ldz; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps;
  p.a = (const bf16*)a->a_bf16; p.lda = a->lda;
  p.x_out = a->x_out; p.ldx = a->ldx;
  p.B = a->B; p.K = a->K; p.KC = KC; p.N = a->N; p.d_true = a->d_true;
  p.n_pad = cb_host::ceil_div(a->N, DF_NT) * DF_NT;
  p.w = (const bf16*)a->w; p.ldw = a->ldw; p.bias = a->bias;
  p.res = a->res; p.ldr = a->ldr; p.out_f32 = a->out_f32; p.ldo = a->ldo;
  p.out_bf16 = (bf16*)a->out_bf16; p.ldob = a->ldob;
  p.q_out = a->q_out; p.k_cache = (bf16*)a->k_cache; p.v_cache = (bf16*)a->v_cache;
  p.H = a->H; p.C = a->C; p.slot = a->slot; p.dstate = a->dev_state;
  if (a->prologue != PRO_BF16) {
    CB_REQUIRE(split == 1, "decode_fused_linear: split_k needs the bf16-rows prologue");
    CB_REQUIRE(a->d_true > 0 && a->d_true <= a->K && a->d_true % 4 == 0, "decode_fused_linear: d_true=%d must be a multiple of 4 <= K", a->d_true);
    if (a->prologue == PRO_LN) CB_REQUIRE(a->z && a->gamma && a->beta && a->ldz % 4 == 0, "decode_fused_linear: layernorm prologue needs z, gamma, beta");
    if (a->prologue == PRO_EMBED) CB_REQUIRE(a->tokens && a->emb, "decode_fused_linear: embedding prologue needs tokens, emb");
    if (a->x_out) CB_REQUIRE(a->ldx % 4 == 0, "decode_fused_linear: ldx must be a multiple of 4");
  } else {
    CB_REQUIRE(a->a_bf16 && a->lda % 8 == 0, "decode_fused_linear: bf16-rows prologue needs a_bf16 with lda %% 8 == 0");
  }
  const int n_pad = cb_host::ceil_div(a->N, DF_NT) * DF_NT;   // the caller pads the weight rows to this
  const int grid_x = n_pad / DF_NT;
  const size_t smem = (size_t)(DF_M + DF_NT) * (KC + DF_PAD) * 2 + 2 * DF_M * DF_PART_LD * 4;
  cudaStream_t s = (cudaStream_t)stream;
  const int pdl = a->pdl;
  switch (a->epilogue) {
    case EPI_QKV:
      CB_REQUIRE(a->q_out && a->k_cache && a->v_cache && a->H > 0 && a->C > 0 && a->N == 3 * a->H * 64,
                 "decode_fused_linear: qkv epilogue needs q_out, caches and N = 3*H*64");
      CB_REQUIRE(a->dev_state || (a->slot >= 0 && a->slot < a->C), "decode_fused_linear: bad ring slot %d", a->slot);
      if (a->prologue == PRO_EMBED) return launch_linear<PRO_EMBED, EPI_QKV, false>(p, grid_x, 1, smem, pdl, s);
      if (a->prologue == PRO_LN) return launch_linear<PRO_LN, EPI_QKV, false>(p, grid_x, 1, smem, pdl, s);
      break;
    case EPI_RES:
      CB_REQUIRE(a->res && a->out_f32, "decode_fused_linear: residual epilogue needs res and out_f32");
      if (a->prologue == PRO_BF16)
        return split > 1 ? launch_linear<PRO_BF16, EPI_RES, true>(p, grid_x, split, smem, pdl, s)
                         : launch_linear<PRO_BF16, EPI_RES, false>(p, grid_x, 1, smem, pdl, s);
      break;
    case EPI_RELU:
      CB_REQUIRE(a->out_bf16 && a->ldob % 4 == 0 && a->N % DF_NT == 0, "decode_fused_linear: relu epilogue needs out_bf16, N %% 16 == 0");
      if (a->prologue == PRO_LN) return launch_linear<PRO_LN, EPI_RELU, false>(p, grid_x, 1, smem, pdl, s);
      break;
    case EPI_LOGITS:
      CB_REQUIRE(a->out_f32, "decode_fused_linear: logits epilogue needs out_f32");
      if (a->prologue == PRO_LN) return launch_linear<PRO_LN, EPI_LOGITS, false>(p, grid_x, 1, smem, pdl, s);
      break;
  }
  return cb_host::fail(COMMU_ERR_UNSUPPORTED, "decode_fused_linear: prologue %d / epilogue %d combination is not built",
                       a->prologue, a->epilogue);
}

int commu_decode_attn_split(const float* q, const void* kcache, const void* vcache, const void* rtab,
                            const float* r_w_bias, const float* r_r_bias, int B, int H, int C, int n_vis, int cur_slot,
                            float scale, int splits, float* partial, int* counters, void* out_bf16, float* out_f32,
                            int64_t ldo, const int* dev_state, int pdl, int impl, void* stream) {
  CB_REQUIRE(q && kcache && vcache && rtab && (out_bf16 || out_f32), "decode_attn_split: null arg");
  CB_REQUIRE(splits >= 1 && splits <= 16 && ((splits == 1 && !(impl & 72)) || (partial && counters)),
             "decode_attn_split: splits=%d needs partial / counters scratch", splits);
  CB_REQUIRE(dev_state || (n_vis >= 1 && n_vis <= C && cur_slot >= 0 && cur_slot < C),
             "decode_attn_split: bad args (n_vis=%d C=%d slot=%d)", n_vis, C, cur_slot);
  cudaStream_t s = (cudaStream_t)stream;
  cb_host::ProfScope prof(cb_host::PROF_DECODE_ATTN, s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H, B, splits);
  cfg.blockDim = dim3(DA2_WARPS * 32, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl ? 1 : 0;
  if (impl & 8) {    // stream-K TMA kernel (product path): `splits` = CTAs per SM, partial = [B*H][C/64][66]
    CB_REQUIRE(C % 64 == 0, "decode_attn_split: the TMA kernel needs a ring capacity that is a multiple of 64 (C=%d)", C);
    static int configured_h = 0;
    const int stages = (impl & 2) ? 3 : 4;
    const int smem = 1024 + stages * (AM_STAGE_BYTES + 256) + 2 * H * 64 * 4;
    if (H > configured_h) {
      const int big = 1024 + 4 * (AM_STAGE_BYTES + 256) + 2 * H * 64 * 4;
      CB_REQUIRE(big <= 227 * 1024, "decode_attn_split: too many heads (%d) for the shared-memory parameter block", H);
      CB_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      CB_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
      configured_h = H;
    }
    const CUtensorMap *tk, *tv, *tr;
    int rc = cached_tmap(kcache, (uint64_t)B * H * C, &tk);
    if (rc) return rc;
    if ((rc = cached_tmap(vcache, (uint64_t)B * H * C, &tv))) return rc;
    if ((rc = cached_tmap(rtab, (uint64_t)H * 2 * C, &tr))) return rc;
    const int max_parts = C / AM_TILE;
    cfg.gridDim = dim3(splits * cb_host::num_sms(), 1, 1);
    cfg.blockDim = dim3(TK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    if (stages == 3)
      CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_attn_tma_kernel<3>, *tk, *tv, *tr, q, r_w_bias, r_r_bias, B * H, H, C, n_vis,
                                       cur_slot, scale, partial, max_parts, counters, (bf16*)out_bf16, out_f32,
                                       (long long)ldo, dev_state));
    else
      CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_attn_tma_kernel<4>, *tk, *tv, *tr, q, r_w_bias, r_r_bias, B * H, H, C, n_vis,
                                       cur_slot, scale, partial, max_parts, counters, (bf16*)out_bf16, out_f32,
                                       (long long)ldo, dev_state));
    cb_host::count_launch();
    return 0;
  }
  CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_attn_split_kernel, q, (const bf16*)kcache, (const bf16*)vcache,
                                   (const bf16*)rtab, r_w_bias, r_r_bias, H, C, n_vis, cur_slot, scale, partial, counters,
                                   (bf16*)out_bf16, out_f32, (long long)ldo, dev_state));
  cb_host::count_launch();
  return 0;
}

}  // extern "C"
