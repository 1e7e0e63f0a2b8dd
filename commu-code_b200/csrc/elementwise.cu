// HBM-bound kernels of the training path: embedding gather/scatter, sinusoid table, LayerNorm
// forward/backward, log-softmax NLL forward/backward, bias-gradient column sums, fp32->bf16 weight
// shadow casts (with head/row padding and optional transpose), gradient un-padding, grad-norm and
// the fused clip+Adam update.  All are vectorised, coalesced, grid sized from the SM count.
#include "api_common.h"
#include "common.cuh"
#include "dropout.cuh"
#include <math.h>

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

// ---------------------------------------------------------------------------------------------
// embedding: x = E[tok] * scale  (model.py:409-420)
// ---------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ tok, const float* __restrict__ table,
                                 int d, int dp, float scale, long long n, float* __restrict__ out_f32,
                                 long long ldf, bf16* __restrict__ out_bf16, long long ldb) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    const long long t = tok[r];
    const float* src = table + t * d;
    for (int c = lane; c < dp; c += 32) {
      float v = c < d ? __ldg(src + c) * scale : 0.f;
      if (out_f32) out_f32[r * ldf + c] = v;
      if (out_bf16) out_bf16[r * ldb + c] = __float2bfloat16_rn(v);
    }
  }
}

// dE[tok] += dx * scale.  729-row table -> heavy address reuse; one warp per row with fp32 atomics
// (red.global.add) spread over d columns keeps contention per address low.
__global__ void embed_bwd_kernel(const long long* __restrict__ tok, const float* __restrict__ dx,
                                 long long ld, int d, float scale, long long n,
                                 float* __restrict__ dtable) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    const long long t = tok[r];
    for (int c = lane; c < d; c += 32) atomicAdd(dtable + t * d + c, dx[r * ld + c] * scale);
  }
}

// P[delta, :] = [sin(delta' f) | cos(delta' f)], delta' = min(delta, clamp) (model.py:142-147,578-583)
__global__ void pos_table_kernel(const float* __restrict__ inv_freq, int K, int clamp_len, int d,
                                 int dp, bf16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)K * dp) return;
  const int delta = idx / dp, c = idx % dp;
  float v = 0.f;
  if (c < d) {
    float pos = (float)delta;
    if (clamp_len > 0) pos = fminf(pos, (float)clamp_len);
    const int half = d / 2;
    const float ang = pos * inv_freq[c < half ? c : c - half];
    v = c < half ? sinf(ang) : cosf(ang);
  }
  if (out_bf16) out_bf16[idx] = __float2bfloat16_rn(v);
  if (out_f32) out_f32[idx] = v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the d real columns of z (model.py:352, :179; eps 1e-5), one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int d, int dp, float eps, long long rows,
                              float* __restrict__ y_f32, long long ldy, bf16* __restrict__ y_bf16,
                              long long ldyb, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* zr = z + r * ldz;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += zr[c];
    const float mean = cb::warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float t = zr[c] - mean;
      q += t * t;
    }
    const float rstd = rsqrtf(cb::warp_sum(q) / d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    for (int c = lane; c < dp; c += 32) {
      float v = 0.f;
      if (c < d) v = (zr[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (y_f32) y_f32[r * ldy + c] = v;
      if (y_bf16) y_bf16[r * ldyb + c] = __float2bfloat16_rn(v);
    }
  }
}

// dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = gamma*dy; dgamma += dy*xhat; dbeta += dy.
template <int MAXC>  // MAXC = ceil(d/32) upper bound
__global__ void ln_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ z,
                              long long ldz, const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ gamma, int d, int dp, long long rows,
                              float* __restrict__ dz_f32, long long lddz, bf16* __restrict__ dz_bf16,
                              long long lddzb, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sh_g[32 * MAXC];
  __shared__ float sh_b[32 * MAXC];
  for (int c = threadIdx.x; c < 32 * MAXC; c += blockDim.x) sh_g[c] = sh_b[c] = 0.f;
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float acc_g[MAXC], acc_b[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) acc_g[i] = acc_b[i] = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[MAXC], g[MAXC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      xh[i] = 0.f;
      g[i] = 0.f;
      if (c < d) {
        const float dyv = dy[r * lddy + c];
        xh[i] = (z[r * ldz + c] - mu) * rs;
        g[i] = dyv * __ldg(gamma + c);
        acc_g[i] += dyv * xh[i];
        acc_b[i] += dyv;
        s1 += g[i];
        s2 += g[i] * xh[i];
      }
    }
    s1 = cb::warp_sum(s1) / d;
    s2 = cb::warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < dp) {
        float v = c < d ? rs * (g[i] - s1 - xh[i] * s2) : 0.f;
        if (dz_f32) dz_f32[r * lddz + c] = v;
        if (dz_bf16) dz_bf16[r * lddzb + c] = __float2bfloat16_rn(v);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    atomicAdd(&sh_g[lane + 32 * i], acc_g[i]);
    atomicAdd(&sh_b[lane + 32 * i], acc_b[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(dgamma + c, sh_g[c]);
    atomicAdd(dbeta + c, sh_b[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// Vectorised LayerNorm (d, dp, leading dims multiples of 4, 16-byte aligned rows): one warp per row, the row lives in
// registers as float4 (a lane owns columns 4*(lane + 32 i) .. +3), so z / dy are read once with 16-byte accesses
// (the scalar kernels above read z three times with 4-byte loads: 41 / 51 us per launch = 63-70 % of HBM peak).
// The backward can apply the inverted dropout of the sub-block output to its bf16 copy (the GEMM operand of the
// sub-block's weight / input gradients); the fp32 copy (residual branch) stays undropped.
// ---------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void ln_fwd_v4_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int d, int dp, float eps, long long rows,
                                 float* __restrict__ y_f32, long long ldy, bf16* __restrict__ y_bf16, long long ldyb,
                                 float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 gm[MAXV], bt[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = 4 * (lane + 32 * i);
    gm[i] = bt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d) {
      gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
      bt[i] = __ldg(reinterpret_cast<const float4*>(beta + c));
    }
  }
  for (long long r = warp; r < rows; r += nwarps) {
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) v[i] = *reinterpret_cast<const float4*>(z + r * ldz + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = cb::warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (4 * (lane + 32 * i) < d) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(cb::warp_sum(q) / d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < dp) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) {
          o.x = (v[i].x - mean) * rstd * gm[i].x + bt[i].x;
          o.y = (v[i].y - mean) * rstd * gm[i].y + bt[i].y;
          o.z = (v[i].z - mean) * rstd * gm[i].z + bt[i].z;
          o.w = (v[i].w - mean) * rstd * gm[i].w + bt[i].w;
        }
        if (y_f32) *reinterpret_cast<float4*>(y_f32 + r * ldy + c) = o;
        if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + r * ldyb + c) = make_uint2(cb::pack_bf16(o.x, o.y), cb::pack_bf16(o.z, o.w));
      }
    }
  }
}

template <int MAXV>
__global__ void __launch_bounds__(THREADS, MAXV <= 4 ? 2 : 1)
ln_bwd_v4_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ z,
                                 long long ldz, const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const float* __restrict__ gamma, int d, int dp, long long rows,
                                 float* __restrict__ dz_f32, long long lddz, bf16* __restrict__ dz_bf16,
                                 long long lddzb, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                 uint32_t drop_thr2, uint32_t drop_ka, uint32_t drop_kb, float drop_inv) {
  __shared__ float sh_g[128 * MAXV];
  __shared__ float sh_b[128 * MAXV];
  __shared__ __align__(16) float sh_gamma[128 * MAXV];     // gamma is read from shared memory: 4 * MAXV registers fewer
  for (int c = threadIdx.x; c < 128 * MAXV; c += blockDim.x) {
    sh_g[c] = sh_b[c] = 0.f;
    sh_gamma[c] = c < d ? gamma[c] : 0.f;
  }
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 acc_g[MAXV], acc_b[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) acc_g[i] = acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // software pipeline: the loads of this warp's NEXT row are issued as soon as the current row's values have been
  // consumed, so they are in flight during the two warp reductions and the output phase (one row per warp at a time left
  // 16 warps x 2 KB in flight per SM: 63 % of the copy bandwidth)
  float4 dyv[MAXV], zv[MAXV];
  float mu = 0.f, rs = 0.f;
  auto load_row = [&](long long r) {
    mu = mean[r]; rs = rstd[r];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < d) {
        dyv[i] = *reinterpret_cast<const float4*>(dy + r * lddy + c);
        zv[i] = *reinterpret_cast<const float4*>(z + r * ldz + c);
      }
    }
  };
  if (warp < rows) load_row(warp);
  for (long long r = warp; r < rows; r += nwarps) {
    const float rs_r = rs;
    float4 xh[MAXV], g[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
      xh[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        const float4 dv = dyv[i];
        xh[i] = make_float4((zv[i].x - mu) * rs, (zv[i].y - mu) * rs, (zv[i].z - mu) * rs, (zv[i].w - mu) * rs);
        const float4 gmv = *reinterpret_cast<const float4*>(sh_gamma + c);
        g[i] = make_float4(dv.x * gmv.x, dv.y * gmv.y, dv.z * gmv.z, dv.w * gmv.w);
        acc_g[i].x += dv.x * xh[i].x; acc_g[i].y += dv.y * xh[i].y; acc_g[i].z += dv.z * xh[i].z; acc_g[i].w += dv.w * xh[i].w;
        acc_b[i].x += dv.x; acc_b[i].y += dv.y; acc_b[i].z += dv.z; acc_b[i].w += dv.w;
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      }
    }
    if (r + nwarps < rows) load_row(r + nwarps);
    s1 = cb::warp_sum(s1) / d;
    s2 = cb::warp_sum(s2) / d;
    const drop::Keys dk = drop::row_keys(drop_ka, drop_kb, (uint32_t)r);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < dp) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) {
          o.x = rs_r * (g[i].x - s1 - xh[i].x * s2);
          o.y = rs_r * (g[i].y - s1 - xh[i].y * s2);
          o.z = rs_r * (g[i].z - s1 - xh[i].z * s2);
          o.w = rs_r * (g[i].w - s1 - xh[i].w * s2);
        }
        if (dz_f32) *reinterpret_cast<float4*>(dz_f32 + r * lddz + c) = o;
        if (dz_bf16) {
          if (drop_thr2) {
            const uint2 rnd = drop::rand64((uint32_t)(c >> 2), dk);
            const uint32_t f0 = drop::keep_flags(rnd.x, drop_thr2), f1 = drop::keep_flags(rnd.y, drop_thr2);
            o.x = (f0 & 0x8000u) ? o.x * drop_inv : 0.f;
            o.y = (f0 & 0x80000000u) ? o.y * drop_inv : 0.f;
            o.z = (f1 & 0x8000u) ? o.z * drop_inv : 0.f;
            o.w = (f1 & 0x80000000u) ? o.w * drop_inv : 0.f;
          }
          *reinterpret_cast<uint2*>(dz_bf16 + r * lddzb + c) = make_uint2(cb::pack_bf16(o.x, o.y), cb::pack_bf16(o.z, o.w));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = 4 * (lane + 32 * i);
    atomicAdd(&sh_g[c + 0], acc_g[i].x); atomicAdd(&sh_g[c + 1], acc_g[i].y);
    atomicAdd(&sh_g[c + 2], acc_g[i].z); atomicAdd(&sh_g[c + 3], acc_g[i].w);
    atomicAdd(&sh_b[c + 0], acc_b[i].x); atomicAdd(&sh_b[c + 1], acc_b[i].y);
    atomicAdd(&sh_b[c + 2], acc_b[i].z); atomicAdd(&sh_b[c + 3], acc_b[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(dgamma + c, sh_g[c]);
    atomicAdd(dbeta + c, sh_b[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// NLL = -log_softmax(logits)[target]  (model.py:71-73), one warp per row; backward writes
// dlogits = (softmax - onehot) * dloss[row] as bf16 (operand of the logits dgrad / wgrad GEMMs).
// ---------------------------------------------------------------------------------------------
__global__ void nll_fwd_kernel(const float* __restrict__ logits, long long ld, int V,
                               const long long* __restrict__ target, long long rows,
                               float* __restrict__ nll, float* __restrict__ lse_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* lr = logits + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, lr[c]);
    mx = cb::warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(lr[c] - mx);
    s = cb::warp_sum(s);
    const float lse = mx + logf(s);
    if (lane == 0) {
      if (lse_out) lse_out[r] = lse;
      if (nll) nll[r] = lse - lr[target[r]];
    }
  }
}

__global__ void nll_bwd_kernel(const float* __restrict__ logits, long long ld, int V, int Vp,
                               const float* __restrict__ lse, const long long* __restrict__ target,
                               const float* __restrict__ dloss, long long rows,
                               bf16* __restrict__ dlogits, long long ldd) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* lr = logits + r * ld;
    const float l = lse[r], g = dloss[r];
    const int t = (int)target[r];
    for (int c = lane; c < Vp; c += 32) {
      float v = 0.f;
      if (c < V) v = (expf(lr[c] - l) - (c == t ? 1.f : 0.f)) * g;
      dlogits[r * ldd + c] = __float2bfloat16_rn(v);
    }
  }
}

// out[c] += sum_r x[r, c]   (bias gradients), x bf16.
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long ld, int ncols, long long rows,
                                   long long rows_per_block, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
  atomicAdd(out + c, s);
}

// The same with 8 columns (16 bytes) per thread: a CTA is 32 column groups x 8 row lanes, four rows in flight per
// thread, the row lanes are reduced through shared memory so that a CTA issues one atomic per column (the scalar
// version read 2 bytes per thread and row: 67 us for the 134 MB FF hidden gradient = 2 TB/s).
__global__ void __launch_bounds__(256) colsum_bf16_v8_kernel(const bf16* __restrict__ x, long long ld, int ncols, long long rows,
                                                             long long rows_per_block, float* __restrict__ out) {
  __shared__ float red[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 8;
  const bool live = c < ncols;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto acc = [&](const uint4 a) {
    s[0] += cb::bf16_lo(a.x); s[1] += cb::bf16_hi(a.x); s[2] += cb::bf16_lo(a.y); s[3] += cb::bf16_hi(a.y);
    s[4] += cb::bf16_lo(a.z); s[5] += cb::bf16_hi(a.z); s[6] += cb::bf16_lo(a.w); s[7] += cb::bf16_hi(a.w);
  };
  if (live) {
    long long r = r0 + ty;
    const bf16* px = x + r * ld + c;
    for (; r + 24 < r1; r += 32, px += 32 * ld) {
      const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(px));
      const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(px + 8 * ld));
      const uint4 a2 = __ldg(reinterpret_cast<const uint4*>(px + 16 * ld));
      const uint4 a3 = __ldg(reinterpret_cast<const uint4*>(px + 24 * ld));
      acc(a0); acc(a1); acc(a2); acc(a3);
    }
    for (; r < r1; r += 8, px += 8 * ld) acc(__ldg(reinterpret_cast<const uint4*>(px)));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx][j] = s[j];
  __syncthreads();
  // 256 threads -> 256 columns of the CTA
  const int col = threadIdx.x;             // column inside the CTA's 256
  const int gx = col >> 3, gj = col & 7;
  float t = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) t += red[y][gx][gj];
  const int cc = blockIdx.x * 256 + col;
  if (cc < ncols) atomicAdd(out + cc, t);
}

// ---------------------------------------------------------------------------------------------
// fp32 master weight [R, C] -> bf16 shadow with segment padding on rows / cols, optional transpose.
//   dst_r = (r / rseg) * rseg_pad + r % rseg,  dst_c likewise.  Padding is never written (the
//   shadow arena is zero-initialised once).
// ---------------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, long long ld_src, int R, int C, int rseg,
                                int rseg_pad, int cseg, int cseg_pad, bf16* __restrict__ dst,
                                long long ld_dst, int transpose) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r, c;
    float v;
    if (!transpose) {
      r = r0 + i;
      c = c0 + threadIdx.x;
      v = tile[i][threadIdx.x];
    } else {
      r = r0 + threadIdx.x;
      c = c0 + i;
      v = tile[threadIdx.x][i];
    }
    if (r < R && c < C) {
      const long long dr = (long long)(r / rseg) * rseg_pad + r % rseg;
      const long long dc = (long long)(c / cseg) * cseg_pad + c % cseg;
      if (!transpose)
        dst[dr * ld_dst + dc] = __float2bfloat16_rn(v);
      else
        dst[dc * ld_dst + dr] = __float2bfloat16_rn(v);
    }
  }
}

// dst[r, c] += src[pad(r), pad(c)]   (padded fp32 gradient -> reference-layout gradient)
__global__ void unpad_accum_kernel(const float* __restrict__ src, long long ld_src, int R, int C, int rseg,
                                   int rseg_pad, int cseg, int cseg_pad, float* __restrict__ dst,
                                   long long ld_dst, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * C) return;
  const int r = idx / C, c = idx % C;
  const long long sr = (long long)(r / rseg) * rseg_pad + r % rseg;
  const long long sc = (long long)(c / cseg) * cseg_pad + c % cseg;
  dst[(long long)r * ld_dst + c] += scale * src[sr * ld_src + sc];
}

// ---------------------------------------------------------------------------------------------
// optimizer: sum of squares, then clip + Adam on flat fp32 arenas (train.py:159-169)
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float sh[WARPS];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    s += g[i] * g[i];
  s = cb::warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < WARPS ? sh[threadIdx.x] : 0.f;
    t = cb::warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// grad is first scaled by grad_scale (1/world), then by the clip coefficient
// min(1, clip / (||grad_scale*g|| + 1e-6)) (torch.nn.utils.clip_grad_norm_), then Adam
// (torch.optim.Adam, no amsgrad, weight_decay 0).  gnorm_sq holds sum(g^2) of the UNSCALED grads.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2, const float* __restrict__ gnorm_sq, float clip,
                            float grad_scale, float weight_decay, float* __restrict__ gnorm_out,
                            bf16* __restrict__ p_bf16) {
  float coef = grad_scale;
  const float norm = sqrtf(*gnorm_sq) * grad_scale;
  if (clip > 0.f) coef *= fminf(1.f, clip / (norm + 1e-6f));
  if (gnorm_out && blockIdx.x == 0 && threadIdx.x == 0) *gnorm_out = norm;
  const float step = lr / bc1;
  const float rsq_bc2 = rsqrtf(bc2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = fmaf(weight_decay, p[i], g[i] * coef);   // torch.optim.Adam: grad += weight_decay * param (after the clip)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pn = p[i] - step * mi / (sqrtf(vi) * rsq_bc2 + eps);
    p[i] = pn;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pn);   // bf16 operand shadow of the updated master, same flat layout
  }
}

int grid_for_rows(long long rows) {
  long long want = (rows + WARPS - 1) / WARPS;
  long long cap = (long long)cb_host::num_sms() * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" {

int commu_embed_fwd(const int64_t* tokens, const float* table, int d, int dp, float scale, int64_t n,
                    float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, void* stream) {
  CB_REQUIRE(tokens && table && n > 0, "embed_fwd: bad args");
  embed_fwd_kernel<<<grid_for_rows(n), THREADS, 0, (cudaStream_t)stream>>>(
      (const long long*)tokens, table, d, dp, scale, n, out_f32, ld_f32, (bf16*)out_bf16, ld_bf16);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_embed_bwd(const int64_t* tokens, const float* dx, int64_t ld, int d, float scale, int64_t n,
                    float* dtable, void* stream) {
  CB_REQUIRE(tokens && dx && dtable && n > 0, "embed_bwd: bad args");
  embed_bwd_kernel<<<grid_for_rows(n), THREADS, 0, (cudaStream_t)stream>>>(
      (const long long*)tokens, dx, ld, d, scale, n, dtable);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_pos_table(const float* inv_freq, int klen, int clamp_len, int d, int dp, void* out_bf16,
                    float* out_f32, void* stream) {
  CB_REQUIRE(inv_freq && klen > 0 && (out_bf16 || out_f32), "pos_table: bad args");
  const long long n = (long long)klen * dp;
  pos_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      inv_freq, klen, clamp_len, d, dp, (bf16*)out_bf16, out_f32);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_layernorm_fwd(const float* z, int64_t ldz, const float* gamma, const float* beta, int d, int dp,
                        float eps, int64_t rows, float* y_f32, int64_t ldy, void* y_bf16, int64_t ldyb,
                        float* mean, float* rstd, void* stream) {
  CB_REQUIRE(z && gamma && beta && rows > 0, "layernorm_fwd: bad args");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (d % 4 == 0 && dp % 4 == 0 && dp <= 1024 && ldz % 4 == 0 && al16(z) && al16(gamma) && al16(beta) &&
      (!y_f32 || (ldy % 4 == 0 && al16(y_f32))) && (!y_bf16 || (ldyb % 4 == 0 && al16(y_bf16)))) {
    if (dp <= 512)
      ln_fwd_v4_kernel<4><<<grid_for_rows(rows), THREADS, 0, (cudaStream_t)stream>>>(
          z, ldz, gamma, beta, d, dp, eps, rows, y_f32, ldy, (bf16*)y_bf16, ldyb, mean, rstd);
    else
      ln_fwd_v4_kernel<8><<<grid_for_rows(rows), THREADS, 0, (cudaStream_t)stream>>>(
          z, ldz, gamma, beta, d, dp, eps, rows, y_f32, ldy, (bf16*)y_bf16, ldyb, mean, rstd);
    cb_host::count_launch();
    CB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  ln_fwd_kernel<<<grid_for_rows(rows), THREADS, 0, (cudaStream_t)stream>>>(
      z, ldz, gamma, beta, d, dp, eps, rows, y_f32, ldy, (bf16*)y_bf16, ldyb, mean, rstd);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_layernorm_bwd(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* mean,
                        const float* rstd, const float* gamma, int d, int dp, int64_t rows, float* dz_f32,
                        int64_t lddz, void* dz_bf16, int64_t lddzb, float* dgamma, float* dbeta,
                        float drop_p, uint64_t drop_seed, void* stream) {
  CB_REQUIRE(dy && z && mean && rstd && gamma && dgamma && dbeta && rows > 0, "layernorm_bwd: bad args");
  CB_REQUIRE(dp <= 1024, "layernorm_bwd: d_model (padded) %d > 1024 unsupported", dp);
  CB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "layernorm_bwd: drop_p must be in [0, 1)");
  const int grid = cb_host::num_sms() * 2;
  cudaStream_t s = (cudaStream_t)stream;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = d % 4 == 0 && dp % 4 == 0 && lddy % 4 == 0 && ldz % 4 == 0 && al16(dy) && al16(z) && al16(gamma) &&
                   (!dz_f32 || (lddz % 4 == 0 && al16(dz_f32))) && (!dz_bf16 || (lddzb % 4 == 0 && al16(dz_bf16)));
  CB_REQUIRE(vec || drop_p == 0.f, "layernorm_bwd: the fused dropout needs 16-byte aligned rows and d, dp multiples of 4");
  if (vec) {
    const uint32_t thr = drop_p > 0.f ? drop::thr15_of(drop_p) : 0u;
    const uint32_t thr2 = thr * 0x00010001u;
    const float inv = 1.f / (1.f - (float)thr / 32768.f);
    const uint32_t ka = (uint32_t)drop_seed, kb = (uint32_t)(drop_seed >> 32);
    if (dp <= 512)
      ln_bwd_v4_kernel<4><<<grid, THREADS, 0, s>>>(dy, lddy, z, ldz, mean, rstd, gamma, d, dp, rows, dz_f32, lddz,
                                                   (bf16*)dz_bf16, lddzb, dgamma, dbeta, thr2, ka, kb, inv);
    else
      ln_bwd_v4_kernel<8><<<grid, THREADS, 0, s>>>(dy, lddy, z, ldz, mean, rstd, gamma, d, dp, rows, dz_f32, lddz,
                                                   (bf16*)dz_bf16, lddzb, dgamma, dbeta, thr2, ka, kb, inv);
    cb_host::count_launch();
    CB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (dp <= 512)
    ln_bwd_kernel<16><<<grid, THREADS, 0, s>>>(dy, lddy, z, ldz, mean, rstd, gamma, d, dp, rows, dz_f32,
                                               lddz, (bf16*)dz_bf16, lddzb, dgamma, dbeta);
  else
    ln_bwd_kernel<32><<<grid, THREADS, 0, s>>>(dy, lddy, z, ldz, mean, rstd, gamma, d, dp, rows, dz_f32,
                                               lddz, (bf16*)dz_bf16, lddzb, dgamma, dbeta);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_nll_fwd(const float* logits, int64_t ld, int V, const int64_t* target, int64_t rows, float* nll,
                  float* lse, void* stream) {
  CB_REQUIRE(logits && target && rows > 0, "nll_fwd: bad args");
  nll_fwd_kernel<<<grid_for_rows(rows), THREADS, 0, (cudaStream_t)stream>>>(
      logits, ld, V, (const long long*)target, rows, nll, lse);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_nll_bwd(const float* logits, int64_t ld, int V, int Vp, const float* lse, const int64_t* target,
                  const float* dloss, int64_t rows, void* dlogits_bf16, int64_t ldd, void* stream) {
  CB_REQUIRE(logits && lse && target && dloss && dlogits_bf16 && rows > 0, "nll_bwd: bad args");
  nll_bwd_kernel<<<grid_for_rows(rows), THREADS, 0, (cudaStream_t)stream>>>(
      logits, ld, V, Vp, lse, (const long long*)target, dloss, rows, (bf16*)dlogits_bf16, ldd);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_colsum_bf16(const void* x, int64_t ld, int ncols, int64_t rows, float* out, void* stream) {
  CB_REQUIRE(x && out && rows > 0 && ncols > 0, "colsum: bad args");
  if (ncols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int bxv = cb_host::ceil_div(ncols, 256);
    int byv = (cb_host::num_sms() * 4) / bxv;
    if (byv < 1) byv = 1;
    if ((long long)byv * 32 > rows) byv = (int)((rows + 31) / 32);
    const long long rpbv = (rows + byv - 1) / byv;
    colsum_bf16_v8_kernel<<<dim3(bxv, byv), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ld, ncols, rows, rpbv, out);
    cb_host::count_launch();
    CB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int bx = cb_host::ceil_div(ncols, 128);
  int by = (cb_host::num_sms() * 4) / bx;
  if (by < 1) by = 1;
  if (by > rows) by = (int)rows;
  const long long rpb = (rows + by - 1) / by;
  colsum_bf16_kernel<<<dim3(bx, by), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, ld, ncols, rows, rpb, out);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_cast_pad(const float* src, int64_t ld_src, int R, int C, int rseg, int rseg_pad, int cseg,
                   int cseg_pad, void* dst_bf16, int64_t ld_dst, int transpose, void* stream) {
  CB_REQUIRE(src && dst_bf16 && R > 0 && C > 0 && rseg > 0 && cseg > 0, "cast_pad: bad args");
  dim3 grid(cb_host::ceil_div(C, 32), cb_host::ceil_div(R, 32));
  cast_pad_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, ld_src, R, C, rseg, rseg_pad, cseg,
                                                                 cseg_pad, (bf16*)dst_bf16, ld_dst, transpose);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_unpad_accum(const float* src, int64_t ld_src, int R, int C, int rseg, int rseg_pad, int cseg,
                      int cseg_pad, float* dst, int64_t ld_dst, float scale, void* stream) {
  CB_REQUIRE(src && dst && R > 0 && C > 0, "unpad_accum: bad args");
  const long long n = (long long)R * C;
  unpad_accum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, R, C, rseg, rseg_pad, cseg, cseg_pad, dst, ld_dst, scale);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_sumsq(const float* g, int64_t n, float* out_accum, void* stream) {
  CB_REQUIRE(g && out_accum && n > 0, "sumsq: bad args");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: buffer must be 16-byte aligned");
  sumsq_kernel<<<cb_host::num_sms() * 4, THREADS, 0, (cudaStream_t)stream>>>(g, n, out_accum);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int commu_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, int step, const float* gnorm_sq, float clip, float grad_scale,
                    float weight_decay, float* gnorm_out, void* p_bf16, void* stream) {
  CB_REQUIRE(p && g && m && v && gnorm_sq && n > 0 && step >= 1, "clip_adam: bad args");
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<cb_host::num_sms() * 8, THREADS, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, gnorm_sq, clip, grad_scale, weight_decay, gnorm_out, (bf16*)p_bf16);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
