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
  int B, K, KC, N, d_true;
  // weights
  const bf16* w; long long ldw; const float* bias;
  // epilogue
  const float* res; long long ldr;
  float* out_f32; long long ldo;
  bf16* out_bf16; long long ldob;
  float* q_out; bf16* k_cache; bf16* v_cache; int H, C, slot; const int* dstate;
};

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

  // ---- weight slice (independent of the previous kernel): cp.async, 16 B per request ----
  {
    const int chunks = KC >> 3;
    for (int c = tid; c < DF_NT * chunks; c += DF_THREADS) {
      const int r = c / chunks, kk = (c - r * chunks) << 3;
      cb::cp_async16(sW + r * lds + kk, p.w + (long long)(n0 + r) * p.ldw + k0 + kk, true);
    }
    cb::cp_async_commit();
  }
  pdl_wait();
  pdl_launch();

  // ---- A tile ----
  if (PRO == PRO_BF16) {
    const int chunks = KC >> 3;
    for (int c = tid; c < DF_M * chunks; c += DF_THREADS) {
      const int r = c / chunks, kk = (c - r * chunks) << 3;
      const bool ok = r < p.B;
      cb::cp_async16(sA + r * lds + kk, p.a + (long long)(ok ? r : 0) * p.lda + k0 + kk, ok);
    }
    cb::cp_async_commit();
  } else {
    // each warp owns 8 rows; a lane holds columns 4*(lane + 32*i) .. +3
    constexpr int MAXV = 8;   // d <= 1024
    const int d = p.d_true;
    for (int rr = 0; rr < 8; ++rr) {
      const int r = warp * 8 + rr;
      float4 v[MAXV];
      const float* src = nullptr;
      float scale = 1.f;
      if (r < p.B) {
        if (PRO == PRO_EMBED) {
          src = p.emb + p.tokens[r] * (long long)d;
          scale = p.emb_scale;
        } else {
          src = p.z + (long long)r * p.ldz;
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = 4 * (lane + 32 * i);
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src && c < d) v[i] = *reinterpret_cast<const float4*>(src + c);
        if (PRO == PRO_EMBED) { v[i].x *= scale; v[i].y *= scale; v[i].z *= scale; v[i].w *= scale; }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      if (PRO == PRO_LN) {
        const float mean = cb::warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = 4 * (lane + 32 * i);
          if (c < d) {
            const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
            q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
          }
        }
        const float rstd = rsqrtf(cb::warp_sum(q) / (float)d + p.eps);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = 4 * (lane + 32 * i);
          if (src && c < d) {
            const float4 g = *reinterpret_cast<const float4*>(p.gamma + c);
            const float4 b = *reinterpret_cast<const float4*>(p.beta + c);
            v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
            v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
            v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
            v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
          }
        }
      }
      const bool writer = p.x_out && r < p.B && (r % (int)gridDim.x) == (int)blockIdx.x;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = 4 * (lane + 32 * i);
        if (c < KC) {
          uint2 pk;
          pk.x = cb::pack_bf16(v[i].x, v[i].y);
          pk.y = cb::pack_bf16(v[i].z, v[i].w);
          *reinterpret_cast<uint2*>(sA + r * lds + c) = pk;
        }
        if (writer && c < d) *reinterpret_cast<float4*>(p.x_out + (long long)r * p.ldx + c) = v[i];
      }
    }
  }
  cb::cp_async_wait<0>();
  __syncthreads();

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
  const int row = tid >> 2, cq = (tid & 3) * 4;
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
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += p.bias[n + j];
    }
    uint2 pk;
    pk.x = cb::pack_bf16(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f));
    pk.y = cb::pack_bf16(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
    *reinterpret_cast<uint2*>(p.out_bf16 + (long long)row * p.ldob + n) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < p.N) {
        float o = v[j];
        if (p.bias) o += p.bias[n + j];
        if (EPI == EPI_RES) o += p.res[(long long)row * p.ldr + n + j];
        p.out_f32[(long long)row * p.ldo + n + j] = o;
      }
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

template <int PRO, int EPI, bool CLUSTER>
int launch_linear(const LinP& p, int grid_x, int split, size_t smem, int pdl, cudaStream_t s) {
  auto kern = dec_linear_kernel<PRO, EPI, CLUSTER>;
  static size_t configured = 0;
  if (smem > configured) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
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
  if (CLUSTER) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 1;
    attrs[na].val.clusterDim.y = split;
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
                            int64_t ldo, const int* dev_state, int pdl, void* stream) {
  CB_REQUIRE(q && kcache && vcache && rtab && (out_bf16 || out_f32), "decode_attn_split: null arg");
  CB_REQUIRE(splits >= 1 && splits <= 16 && (splits == 1 || (partial && counters)),
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
  CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_attn_split_kernel, q, (const bf16*)kcache, (const bf16*)vcache,
                                   (const bf16*)rtab, r_w_bias, r_r_bias, H, C, n_vis, cur_slot, scale, partial, counters,
                                   (bf16*)out_bf16, out_f32, (long long)ldo, dev_state));
  cb_host::count_launch();
  return 0;
}

}  // extern "C"
