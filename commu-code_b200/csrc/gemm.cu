// tcgen05 / TMA / TMEM dense contraction for sm_100a.
//
//   C[m,n] = alpha * sum_k A[m,k] * B[n,k]      bf16 operands, fp32 accumulation in TMEM
//
// Persistent, warp-specialised kernel (one CTA per SM, 256 threads):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected lane issues tcgen05.mma, tcgen05.commit frees slots)
//   warp 2   : TMEM allocator (2 accumulator buffers so the epilogue overlaps the next tile)
//   warps 4-11: epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> global; two warps per TMEM lane
//                              quadrant, the load of the next 32-column chunk in flight while one is stored: with
//                              four warps and a wait after every load the K = 512 shapes were epilogue-bound)
//
// Both operand majors are supported so forward (K-major x K-major), dgrad (K-major x K-major on a
// transposed weight shadow) and wgrad (MN-major x MN-major: reduction over tokens) all run here.
// Replaces the nn.Linear calls of the reference (commu/model/model.py:285-286, 348, 163-169, 46)
// and their autograd backward.
#include <stdlib.h>
#include "api_common.h"
#include "common.cuh"
#include "dropout.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;   // warps 0-3: producer / MMA / TMEM alloc / idle, warps 4-11: epilogue
constexpr int EPI_WARPS = 8;       // two warps per TMEM lane quadrant, each half of the tile's columns
constexpr int SMEM_BUDGET = 200 * 1024;

struct GemmKernelParams {
  int M, N, K;
  int a_mn, b_mn;
  int split_k;
  float alpha;
  const float* bias;
  int relu;
  const bf16* relu_mask;
  long long ld_mask;
  const float* add_f32;
  long long ld_add;
  bf16* out_bf16;
  long long ld_out_bf16;
  float* out_f32;
  long long ld_out_f32;
  int f32_atomic;
  int vec_ok;  // all leading dims / bases allow 16-byte vector stores
  uint32_t drop_thr2, drop_ka, drop_kb;   // drop_thr2 != 0: inverted dropout of the result (counter-based mask)
  float drop_inv;
  int dbg;   // COMMU_GEMM_EPI_DEBUG (measurement only): 1 = no global stores on the bf16 path, 2 = no epilogue work at all
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * 3072 /*epilogue staging + bias*/;
};

// One warp's 32 rows x 32 columns of the accumulator (lane = row).  The fused element-wise part runs with a row per
// lane; the result is then transposed through a 2.5 KB per-warp staging tile so that the global accesses of a warp
// cover whole 64-byte row segments (a row per lane made every 16-byte store instruction touch 32 different lines:
// the epilogue was bound by the LSU, not by its warps).
constexpr int EPI_PITCH = 80;                          // bytes per staged row: 64 data + 16 pad (conflict-free 16 B accesses)
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH;        // per epilogue warp
constexpr int EPI_BIAS_BYTES = 512;                    // the warp's 128 bias values of the current tile
constexpr int EPI_WARP_BYTES = EPI_STAGE_BYTES + EPI_BIAS_BYTES;

// Stages the bias of the 128 columns this epilogue warp owns in the current tile (one guarded load per lane and
// value, issued while the warp still waits for the accumulator; the chunks then read it back as shared-memory
// broadcasts: 32 scalar __ldg per chunk went to L2 every time, the 28 KB of L1 left beside the smem ring being
// thrashed by the output stream - 35 % of the executed instructions and 32 % of the stall samples of the FF1 GEMM).
__device__ __forceinline__ void epilogue_stage_bias(const GemmKernelParams& p, int nbase, int lane, unsigned char* ws) {
  if (!p.bias) return;
  float4 bv;
  const int c = nbase + lane * 4;
  bv.x = c + 0 < p.N ? __ldg(p.bias + c + 0) : 0.f;
  bv.y = c + 1 < p.N ? __ldg(p.bias + c + 1) : 0.f;
  bv.z = c + 2 < p.N ? __ldg(p.bias + c + 2) : 0.f;
  bv.w = c + 3 < p.N ? __ldg(p.bias + c + 3) : 0.f;
  __syncwarp();
  *reinterpret_cast<float4*>(ws + EPI_STAGE_BYTES + lane * 16) = bv;
  __syncwarp();
}

// chunk: index of the 32-column chunk inside the warp's 128 columns (selects the staged bias values)
template <bool DROP>
__device__ __forceinline__ void epilogue_store(const GemmKernelParams& p, const uint32_t (&r)[32],
                                               int row0, int lane, int n0, int chunk, unsigned char* stage) {
  const int row = row0 + lane;
  const bool rowok = row < p.M;
  const bool full = (n0 + 32 <= p.N) && p.vec_ok;      // warp-uniform
  const int tr = lane >> 2, tc = lane & 3;             // transposed role: rows tr + 8 i, 16-byte chunk tc of a staged row
  const bool path_a = p.out_bf16 && !p.add_f32 && !p.out_f32;
  // operands of the transposed (coalesced) phase do not depend on the accumulator: fetch them first
  uint4 mk_a[4];
  float4 ad[2][4];
  uint2 mk_b[2][4];
  if (full) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long rr = row0 + tr + 8 * i;
      const bool ok = rr < p.M;
      if (path_a) {
        mk_a[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        if (p.relu_mask && ok) mk_a[i] = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + rr * p.ld_mask + n0 + tc * 8));
      } else {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col = n0 + hf * 16 + tc * 4;
          ad[hf][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          mk_b[hf][i] = make_uint2(0x3f803f80u, 0x3f803f80u);
          if (p.add_f32 && ok) ad[hf][i] = __ldg(reinterpret_cast<const float4*>(p.add_f32 + rr * p.ld_add + col));
          if (p.relu_mask && ok) mk_b[hf][i] = __ldg(reinterpret_cast<const uint2*>(p.relu_mask + rr * p.ld_mask + col));
        }
      }
    }
  }
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
  if (p.bias) {
    const float4* bs = reinterpret_cast<const float4*>(stage + EPI_STAGE_BYTES + chunk * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 bq = bs[i];
      v[i * 4 + 0] += bq.x; v[i * 4 + 1] += bq.y; v[i * 4 + 2] += bq.z; v[i * 4 + 3] += bq.w;
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (DROP) {   // same (seed, row, column) -> keep function as dropout.cu
    const drop::Keys dk = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)row);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const uint2 rnd = drop::rand64((uint32_t)((n0 >> 2) + c4), dk);
      const uint32_t f0 = drop::keep_flags(rnd.x, p.drop_thr2), f1 = drop::keep_flags(rnd.y, p.drop_thr2);
      v[c4 * 4 + 0] = (f0 & 0x8000u) ? v[c4 * 4 + 0] * p.drop_inv : 0.f;
      v[c4 * 4 + 1] = (f0 & 0x80000000u) ? v[c4 * 4 + 1] * p.drop_inv : 0.f;
      v[c4 * 4 + 2] = (f1 & 0x8000u) ? v[c4 * 4 + 2] * p.drop_inv : 0.f;
      v[c4 * 4 + 3] = (f1 & 0x80000000u) ? v[c4 * 4 + 3] * p.drop_inv : 0.f;
    }
  }
  if (!full) {       // ragged tile edge / unaligned buffers: a row per lane, scalar accesses
    if (!rowok) return;
    if (p.relu_mask) {
      const bf16* mrow = p.relu_mask + (long long)row * p.ld_mask + n0;
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N && !(__bfloat162float(mrow[i]) > 0.f)) v[i] = 0.f;
    }
    if (p.add_f32) {
      const float* arow = p.add_f32 + (long long)row * p.ld_add + n0;
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) v[i] += arow[i];
    }
    if (p.out_bf16) {
      bf16* orow = p.out_bf16 + (long long)row * p.ld_out_bf16 + n0;
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) orow[i] = __float2bfloat16_rn(v[i]);
    }
    if (p.out_f32) {
      float* orow = p.out_f32 + (long long)row * p.ld_out_f32 + n0;
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.N) {
          if (p.f32_atomic) atomicAdd(orow + i, v[i]);
          else orow[i] = v[i];
        }
    }
    return;
  }
  unsigned char* mine = stage + lane * EPI_PITCH;
  if (path_a) {
    // bf16 only: stage 32 columns (64 B per row), every store instruction writes 8 rows x 64 contiguous bytes
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(mine + c * 16) =
          make_uint4(cb::pack_bf16(v[c * 8 + 0], v[c * 8 + 1]), cb::pack_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                     cb::pack_bf16(v[c * 8 + 4], v[c * 8 + 5]), cb::pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = tr + 8 * i;
      uint4 q = *reinterpret_cast<const uint4*>(stage + rr * EPI_PITCH + tc * 16);
      if (p.relu_mask) {   // keep where the saved activation is > 0 (positive bf16: sign clear and not zero)
        const uint32_t w[4] = {mk_a[i].x, mk_a[i].y, mk_a[i].z, mk_a[i].w};
        uint32_t* qq = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!(cb::bf16_lo(w[j]) > 0.f)) qq[j] &= 0xFFFF0000u;
          if (!(cb::bf16_hi(w[j]) > 0.f)) qq[j] &= 0x0000FFFFu;
        }
      }
      if (row0 + rr < p.M && !(p.dbg & 1))
        *reinterpret_cast<uint4*>(p.out_bf16 + (long long)(row0 + rr) * p.ld_out_bf16 + n0 + tc * 8) = q;
    }
    return;
  }
  // fp32 result (+ residual, + optional bf16 copy): two halves of 16 columns (64 B per row) through the staging tile
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(mine + c * 16) =
          make_float4(v[hf * 16 + c * 4], v[hf * 16 + c * 4 + 1], v[hf * 16 + c * 4 + 2], v[hf * 16 + c * 4 + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = tr + 8 * i;
      if (row0 + rr >= p.M) continue;
      float4 q = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + tc * 16);
      const long long col = n0 + hf * 16 + tc * 4;
      if (p.relu_mask) {
        if (!(cb::bf16_lo(mk_b[hf][i].x) > 0.f)) q.x = 0.f;
        if (!(cb::bf16_hi(mk_b[hf][i].x) > 0.f)) q.y = 0.f;
        if (!(cb::bf16_lo(mk_b[hf][i].y) > 0.f)) q.z = 0.f;
        if (!(cb::bf16_hi(mk_b[hf][i].y) > 0.f)) q.w = 0.f;
      }
      q.x += ad[hf][i].x; q.y += ad[hf][i].y; q.z += ad[hf][i].z; q.w += ad[hf][i].w;
      if (p.out_bf16)
        *reinterpret_cast<uint2*>(p.out_bf16 + (long long)(row0 + rr) * p.ld_out_bf16 + col) =
            make_uint2(cb::pack_bf16(q.x, q.y), cb::pack_bf16(q.z, q.w));
      if (p.out_f32) {
        float* o = p.out_f32 + (long long)(row0 + rr) * p.ld_out_f32 + col;
        if (p.f32_atomic) cb::red_add_v4(o, q.x, q.y, q.z, q.w);
        else *reinterpret_cast<float4*>(o) = q;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Whole-tile epilogues of one warp for FULL tiles (its 32 rows < M, its CH*32 columns < N, vector-aligned buffers).
// The general routine above re-derives every option, bound and 64-bit address per 32-column chunk: ~560 instructions
// per chunk and warp, 18 k per 128 x 256 tile - more issue slots than the tile's MMAs take (K = 512: the GEMMs ran at
// half the speed they reach with the epilogue switched off).  Here the options are hoisted per tile, the pointers
// advance by constants, shared memory is addressed in its own window and the ReLU mask is applied on packed pairs.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// keep mask (0xFFFF per half) of a packed pair of saved activations: positive <=> sign clear and magnitude non-zero
__device__ __forceinline__ uint32_t positive16x2(uint32_t w) {
  return drop::mask16x2(((w & 0x7FFF7FFFu) + 0x7FFF7FFFu) & ~w);
}

// v = alpha * acc (+ bias) (relu) (dropout): the part of the epilogue that runs with a row per lane, NV columns from n0
template <bool DROP, int NV>
__device__ __forceinline__ void epilogue_rowmath(const GemmKernelParams& p, const uint32_t (&r)[NV], float (&v)[NV],
                                                 bool scale, bool has_bias, bool relu, uint32_t st_bias, int n0,
                                                 const drop::Keys& dk) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __uint_as_float(r[i]);
  if (scale) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= p.alpha;
  }
  if (has_bias) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const uint4 bq = lds128(st_bias + i * 16);
      v[i * 4 + 0] += __uint_as_float(bq.x); v[i * 4 + 1] += __uint_as_float(bq.y);
      v[i * 4 + 2] += __uint_as_float(bq.z); v[i * 4 + 3] += __uint_as_float(bq.w);
    }
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (DROP) {   // same (seed, row, column) -> keep function as dropout.cu
#pragma unroll
    for (int c4 = 0; c4 < NV / 4; ++c4) {
      const uint2 rnd = drop::rand64((uint32_t)((n0 >> 2) + c4), dk);
      const uint32_t f0 = drop::keep_flags(rnd.x, p.drop_thr2), f1 = drop::keep_flags(rnd.y, p.drop_thr2);
      v[c4 * 4 + 0] = (f0 & 0x8000u) ? v[c4 * 4 + 0] * p.drop_inv : 0.f;
      v[c4 * 4 + 1] = (f0 & 0x80000000u) ? v[c4 * 4 + 1] * p.drop_inv : 0.f;
      v[c4 * 4 + 2] = (f1 & 0x8000u) ? v[c4 * 4 + 2] * p.drop_inv : 0.f;
      v[c4 * 4 + 3] = (f1 & 0x80000000u) ? v[c4 * 4 + 3] * p.drop_inv : 0.f;
    }
  }
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// bf16 result only (optional ReLU mask of a saved activation)
template <bool DROP, int CH>
__device__ __forceinline__ void epilogue_fast_bf16(const GemmKernelParams& p, uint32_t taddr, int row0, int nbase,
                                                   int lane, unsigned char* stage) {
  const int tr = lane >> 2, tc = lane & 3;   // transposed role: rows tr + 8 i, 16-byte chunk tc of a staged row
  const uint32_t st = cb::smem_u32(stage);
  const uint32_t st_mine = st + lane * EPI_PITCH, st_tr = st + tr * EPI_PITCH + tc * 16, st_bias = st + EPI_STAGE_BYTES;
  bf16* optr = p.out_bf16 + (long long)(row0 + tr) * p.ld_out_bf16 + nbase + tc * 8;
  const long long ostep = 8 * p.ld_out_bf16;
  const bf16* mptr = p.relu_mask ? p.relu_mask + (long long)(row0 + tr) * p.ld_mask + nbase + tc * 8 : nullptr;
  const long long mstep = 8 * p.ld_mask;
  const bool scale = p.alpha != 1.f, has_bias = p.bias != nullptr, relu = p.relu != 0;
  const drop::Keys dk = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)(row0 + lane));
  uint32_t r[2][32];
  cb::tmem_ld_32x32b_x32(taddr, r[0]);
  uint4 mk[2][4];                      // the saved activation of the NEXT chunk is in flight while one is processed
  auto load_mask = [&](int c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mk[c & 1][i] = __ldg(reinterpret_cast<const uint4*>(mptr + i * mstep + c * 32));
  };
  if (mptr) load_mask(0);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if (mptr && c + 1 < CH) load_mask(c + 1);
    cb::tmem_ld_wait();
    if (c + 1 < CH) cb::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
    float v[32];
    epilogue_rowmath<DROP, 32>(p, r[c & 1], v, scale, has_bias, relu, st_bias + c * 128, nbase + c * 32, dk);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      sts128(st_mine + q * 16, cb::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), cb::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
             cb::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), cb::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {      // every store instruction writes 8 rows x 64 contiguous bytes
      uint4 q = lds128(st_tr + i * 8 * EPI_PITCH);
      if (mptr) {
        q.x &= positive16x2(mk[c & 1][i].x); q.y &= positive16x2(mk[c & 1][i].y);
        q.z &= positive16x2(mk[c & 1][i].z); q.w &= positive16x2(mk[c & 1][i].w);
      }
      if (!(p.dbg & 1)) *reinterpret_cast<uint4*>(optr + i * ostep + c * 32) = q;
    }
  }
}

// fp32 result (+ fp32 addend, which may be the output itself), stored or - split-K - atomically accumulated
template <bool DROP, int CH>
__device__ __forceinline__ void epilogue_fast_f32(const GemmKernelParams& p, uint32_t taddr, int row0, int nbase,
                                                  int lane, unsigned char* stage) {
  const int tr = lane >> 2, tc = lane & 3;
  const uint32_t st = cb::smem_u32(stage);
  const uint32_t st_mine = st + lane * EPI_PITCH, st_tr = st + tr * EPI_PITCH + tc * 16, st_bias = st + EPI_STAGE_BYTES;
  float* optr = p.out_f32 + (long long)(row0 + tr) * p.ld_out_f32 + nbase + tc * 4;
  const long long ostep = 8 * p.ld_out_f32;
  const float* aptr = p.add_f32 ? p.add_f32 + (long long)(row0 + tr) * p.ld_add + nbase + tc * 4 : nullptr;
  const long long astep = 8 * p.ld_add;
  const bool scale = p.alpha != 1.f, has_bias = p.bias != nullptr, relu = p.relu != 0, atomic = p.f32_atomic != 0;
  const drop::Keys dk = drop::row_keys(p.drop_ka, p.drop_kb, (uint32_t)(row0 + lane));
  // the addend of the NEXT 32-column chunk is in flight while one chunk is processed (HBM latency; the TMEM load is
  // short and is simply waited for: double-buffering both does not fit the register file)
  float4 ad[2][2][4];
  auto load_addend = [&](int c) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ad[c & 1][hf][i] = __ldg(reinterpret_cast<const float4*>(aptr + i * astep + c * 32 + hf * 16));
  };
  if (aptr) load_addend(0);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    if (aptr && c + 1 < CH) load_addend(c + 1);
    uint32_t r[32];
    cb::tmem_ld_32x32b_x32(taddr + c * 32, r);
    cb::tmem_ld_wait();
    float v[32];
    epilogue_rowmath<DROP, 32>(p, r, v, scale, has_bias, relu, st_bias + c * 128, nbase + c * 32, dk);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {   // two halves of 16 columns (64 B per row) through the staging tile
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sts128(st_mine + q * 16, __float_as_uint(v[hf * 16 + q * 4]), __float_as_uint(v[hf * 16 + q * 4 + 1]),
               __float_as_uint(v[hf * 16 + q * 4 + 2]), __float_as_uint(v[hf * 16 + q * 4 + 3]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {    // every store instruction writes 8 rows x 64 contiguous bytes
        const uint4 w = lds128(st_tr + i * 8 * EPI_PITCH);
        float4 q = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
        if (aptr) {
          const float4 x = ad[c & 1][hf][i];
          q.x += x.x; q.y += x.y; q.z += x.z; q.w += x.w;
        }
        float* o = optr + i * ostep + c * 32 + hf * 16;
        if (atomic) cb::red_add_v4(o, q.x, q.y, q.z, q.w);     // split-K partial sums
        else *reinterpret_cast<float4*>(o) = q;
      }
    }
  }
}

// One warp's share of a finished accumulator tile: 32 rows (row0 ..) x CH*32 columns (nbase ..) at TMEM address taddr.
// FAST is chosen by the host (fast_epilogue_ok): every tile full, buffers vector-aligned, one of the two whole-tile
// epilogues applies.  The general routine lives in the other instantiation only - next to the fast paths (inlined or
// called) it made the register allocator spill in all of them.
template <bool DROP, int CH, bool FAST>
__device__ __forceinline__ void epilogue_tile(const GemmKernelParams& p, uint32_t taddr, int row0, int nbase, int lane,
                                              unsigned char* stage) {
  if (p.dbg & 2) return;
  if constexpr (FAST) {
    if (p.out_bf16) epilogue_fast_bf16<DROP, CH>(p, taddr, row0, nbase, lane, stage);
    else epilogue_fast_f32<DROP, CH>(p, taddr, row0, nbase, lane, stage);
  } else {
    uint32_t r[2][32];
    cb::tmem_ld_32x32b_x32(taddr, r[0]);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      cb::tmem_ld_wait();
      if (c + 1 < CH) cb::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
      const int n0 = nbase + c * 32;
      if (n0 < p.N) epilogue_store<DROP>(p, r[c & 1], row0, lane, n0, c, stage);
    }
  }
}

template <int BLOCK_N, bool DROP, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmKernelParams p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = (kb_total + p.split_k - 1) / p.split_k;
  const int num_items = m_tiles * n_tiles * p.split_k;

  if (warp == 0 && lane == 0) {
    cb::tma_prefetch_desc(&tmap_a);
    cb::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      cb::mbar_init(&full_bar[s], 1);
      cb::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      cb::mbar_init(&tmem_full[b], 1);
      cb::mbar_init(&tmem_empty[b], EPI_WARPS * 32);
    }
    cb::fence_barrier_init();
  }
  if (warp == 2) {
    cb::tmem_alloc(tmem_ptr, C::TMEM_COLS);
    cb::tmem_relinquish();
  }
  cb::tc_fence_before();
  __syncthreads();
  cb::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (cb::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int n_blk = item % n_tiles;
        const int m_blk = (item / n_tiles) % m_tiles;
        const int ks = item / (n_tiles * m_tiles);
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          cb::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          cb::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if (!p.a_mn) {
            cb::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a)
              cb::tma_load_2d(sa + a * (64 * BLOCK_K * 2), &tmap_a, &full_bar[stage],
                              m_blk * BLOCK_M + a * 64, kb * BLOCK_K);
          }
          if (!p.b_mn) {
            cb::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_N / 64; ++a)
              cb::tma_load_2d(sb + a * (64 * BLOCK_K * 2), &tmap_b, &full_bar[stage],
                              n_blk * BLOCK_N + a * 64, kb * BLOCK_K);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (cb::elect_one()) {
      const uint32_t idesc = cb::umma_idesc_bf16(BLOCK_M, BLOCK_N, p.a_mn, p.b_mn);
      // K-major: SBO = 8 rows * 128 B, LBO unused. MN-major: SBO = 8 k-rows * 128 B,
      // LBO = one 64-wide MN atom = BLOCK_K rows * 128 B.
      const uint32_t a_lbo = p.a_mn ? (BLOCK_K * 128) : 16;
      const uint32_t b_lbo = p.b_mn ? (BLOCK_K * 128) : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ks = item / (n_tiles * m_tiles);
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        cb::mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        cb::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          cb::mbar_wait(&full_bar[stage], phase);
          cb::tc_fence_after();
          const uint32_t sa = cb::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          uint64_t adesc = cb::umma_smem_desc(sa, a_lbo, 1024);
          uint64_t bdesc = cb::umma_smem_desc(sb, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            cb::umma_bf16_ss(tmem_d, adesc + (uint64_t)(k * a_kstep),
                             bdesc + (uint64_t)(k * b_kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          cb::umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        cb::umma_commit(&tmem_full[buf]);
        buf ^= 1;
        if (buf == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    unsigned char* epi_stage = smem + C::STAGES * C::STAGE_BYTES + 256 + (warp - 4) * EPI_WARP_BYTES;
    constexpr int CH = BLOCK_N / 64;       // 32-column chunks per epilogue warp
    int buf = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int n_blk = item % n_tiles;
      const int m_blk = (item / n_tiles) % m_tiles;
      const int ks = item / (n_tiles * m_tiles);
      const int kb0 = ks * kb_per;
      const int kb1 = min(kb_total, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      epilogue_stage_bias(p, n_blk * BLOCK_N + half * (CH * 32), lane, epi_stage);
      cb::mbar_wait(&tmem_full[buf], acc_phase);
      cb::tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BLOCK_N + half * (CH * 32);
      const int nbase = n_blk * BLOCK_N + half * (CH * 32);
      epilogue_tile<DROP, CH, FAST>(p, taddr, row - lane, nbase, lane, epi_stage);
      cb::tc_fence_before();
      cb::mbar_arrive(&tmem_empty[buf]);
      buf ^= 1;
      if (buf == 0) acc_phase ^= 1;
    }
  }

  cb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    cb::tc_fence_after();
    cb::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on the two SMs of a TPC owns a 256 x 256 output
// tile.  Each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 n-rows); the pair's MMA
// (M = 256) reads the other half from the peer's shared memory, so the shared-memory fill per CTA and k-block drops
// from 48 KB to 32 KB (at 128 x 256 x 64 per 0.37 us a single CTA needs ~130 GB/s of L2 -> SM bandwidth, which is what
// caps the one-CTA kernel on the model's K = 512 shapes).  Protocol:
//   * both producers issue their TMA loads with .cta_group::2 so that the bytes land on the LEADER's full barrier;
//   * only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2; its commits are multicast to the empty /
//     tmem_full barriers of both CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (512 arrivals).
// ---------------------------------------------------------------------------------------------
struct Cfg2 {
  static constexpr int BLOCK_N = 256;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // this CTA's 128 rows
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;    // this CTA's half of the n-rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_WARPS * 3072;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(cb::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on `bar` of both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          cb::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}

template <bool DROP, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmKernelParams p) {
  using C = Cfg2;
  constexpr int BLOCK_N = C::BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = (kb_total + p.split_k - 1) / p.split_k;
  const int num_items = m_pairs * n_tiles * p.split_k;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    cb::tma_prefetch_desc(&tmap_a);
    cb::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      cb::mbar_init(&full_bar[s], 1);
      cb::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      cb::mbar_init(&tmem_full[b], 1);
      cb::mbar_init(&tmem_empty[b], 2 * EPI_WARPS * 32);     // the epilogue threads of both CTAs (leader's copy is the one used)
    }
    cb::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(cb::smem_u32(tmem_ptr)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  cb::tc_fence_before();
  cluster_sync_all();
  cb::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (cb::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int n_blk = item % n_tiles;
        const int m_blk = (item / n_tiles) % m_pairs;
        const int ks = item / (n_tiles * m_pairs);
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        const int m0 = m_blk * 2 * BLOCK_M + (int)rank * BLOCK_M;
        const int n0 = n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          cb::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const uint32_t lbar = mapa_rank(cb::smem_u32(&full_bar[stage]), 0);
          if (rank == 0) cb::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tmap_a, lbar, kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a)
              tma_load_2d_2sm(sa + a * (64 * BLOCK_K * 2), &tmap_a, lbar, m0 + a * 64, kb * BLOCK_K);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &tmap_b, lbar, kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_N / 128; ++a)
              tma_load_2d_2sm(sb + a * (64 * BLOCK_K * 2), &tmap_b, lbar, n0 + a * 64, kb * BLOCK_K);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && cb::elect_one()) {
      const uint32_t idesc = cb::umma_idesc_bf16(2 * BLOCK_M, BLOCK_N, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? (BLOCK_K * 128) : 16;
      const uint32_t b_lbo = p.b_mn ? (BLOCK_K * 128) : 16;
      const uint32_t a_kstep = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t b_kstep = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t acc_phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int ks = item / (n_tiles * m_pairs);
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        cb::mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        cb::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          cb::mbar_wait(&full_bar[stage], phase);
          cb::tc_fence_after();
          const uint32_t sa = cb::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = cb::umma_smem_desc(sa, a_lbo, 1024);
          const uint64_t bdesc = cb::umma_smem_desc(sb, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_ss_2sm(tmem_d, adesc + (uint64_t)(k * a_kstep), bdesc + (uint64_t)(k * b_kstep), idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[buf]);
        buf ^= 1;
        if (buf == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: their own 128 rows) =====================
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    unsigned char* epi_stage = smem + C::STAGES * C::STAGE_BYTES + 256 + (warp - 4) * EPI_WARP_BYTES;
    constexpr int CH = BLOCK_N / 64;
    int buf = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int n_blk = item % n_tiles;
      const int m_blk = (item / n_tiles) % m_pairs;
      const int ks = item / (n_tiles * m_pairs);
      const int kb0 = ks * kb_per;
      const int kb1 = min(kb_total, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      epilogue_stage_bias(p, n_blk * BLOCK_N + half * (CH * 32), lane, epi_stage);
      cb::mbar_wait(&tmem_full[buf], acc_phase);
      cb::tc_fence_after();
      const int row = m_blk * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BLOCK_N + half * (CH * 32);
      const int nbase = n_blk * BLOCK_N + half * (CH * 32);
      epilogue_tile<DROP, CH, FAST>(p, taddr, row - lane, nbase, lane, epi_stage);
      cb::tc_fence_before();
      mbar_arrive_cluster(mapa_rank(cb::smem_u32(&tmem_empty[buf]), 0));
      buf ^= 1;
      if (buf == 0) acc_phase ^= 1;
    }
  }

  cb::tc_fence_before();
  cluster_sync_all();      // nobody leaves while the peer may still read this CTA's tiles or signal its barriers
  if (warp == 2) {
    cb::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// Naive SIMT cross-check kernel: same contract, one thread per output element.
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, long long lda, const bf16* __restrict__ B,
                                 long long ldb, GemmKernelParams p) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  const int ks = blockIdx.z;
  if (n >= p.N || m >= p.M) return;
  const int kper = (p.K + p.split_k - 1) / p.split_k;
  const int k0 = ks * kper, k1 = min(p.K, k0 + kper);
  float acc = 0.f;
  for (int k = k0; k < k1; ++k) {
    float a = __bfloat162float(p.a_mn ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
    float b = __bfloat162float(p.b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
    acc = fmaf(a, b, acc);
  }
  float v = acc * p.alpha;
  if (p.split_k > 1 && ks > 0) {
    atomicAdd(p.out_f32 + (long long)m * p.ld_out_f32 + n, v);
    return;
  }
  if (p.bias) v += p.bias[n];
  if (p.relu) v = fmaxf(v, 0.f);
  if (p.relu_mask && !(__bfloat162float(p.relu_mask[(long long)m * p.ld_mask + n]) > 0.f)) v = 0.f;
  if (p.add_f32) v += p.add_f32[(long long)m * p.ld_add + n];
  if (p.out_bf16) p.out_bf16[(long long)m * p.ld_out_bf16 + n] = __float2bfloat16_rn(v);
  if (p.out_f32) {
    if (p.f32_atomic)
      atomicAdd(p.out_f32 + (long long)m * p.ld_out_f32 + n, v);
    else
      p.out_f32[(long long)m * p.ld_out_f32 + n] = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

namespace cb_host {
// 2D bf16 tensor map with 128B swizzle: inner (contiguous) extent `inner`, outer extent `outer`,
// row stride `ld` elements, box {box_inner (=64), box_outer}.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer,
                      uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(COMMU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(COMMU_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu", (int)r,
                ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return 0;
}
}  // namespace cb_host

template <int BLOCK_N, bool DROP, bool FAST>
static int launch_tc(const commu_gemm_args* a, const GemmKernelParams& p, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap ta, tb;
  int rc;
  if (!a->a_mn_major)
    rc = cb_host::make_tmap_bf16_2d(&ta, a->a, a->k, a->m, a->lda, 64, BLOCK_M);
  else
    rc = cb_host::make_tmap_bf16_2d(&ta, a->a, a->m, a->k, a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!a->b_mn_major)
    rc = cb_host::make_tmap_bf16_2d(&tb, a->b, a->k, a->n, a->ldb, 64, BLOCK_N);
  else
    rc = cb_host::make_tmap_bf16_2d(&tb, a->b, a->n, a->k, a->ldb, 64, BLOCK_K);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, DROP, FAST>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = cb_host::ceil_div(a->m, BLOCK_M), n_tiles = cb_host::ceil_div(a->n, BLOCK_N);
  const int items = m_tiles * n_tiles * p.split_k;
  const int grid = items < cb_host::num_sms() ? items : cb_host::num_sms();
  cb_host::ProfScope prof(cb_host::PROF_GEMM, stream);
  gemm_tcgen05_kernel<BLOCK_N, DROP, FAST><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool DROP, bool FAST>
static int launch_tc_2cta(const commu_gemm_args* a, const GemmKernelParams& p, cudaStream_t stream) {
  using C = Cfg2;
  CUtensorMap ta, tb;
  int rc;
  if (!a->a_mn_major)
    rc = cb_host::make_tmap_bf16_2d(&ta, a->a, a->k, a->m, a->lda, 64, BLOCK_M);
  else
    rc = cb_host::make_tmap_bf16_2d(&ta, a->a, a->m, a->k, a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!a->b_mn_major)
    rc = cb_host::make_tmap_bf16_2d(&tb, a->b, a->k, a->n, a->ldb, 64, C::BLOCK_N / 2);
  else
    rc = cb_host::make_tmap_bf16_2d(&tb, a->b, a->n, a->k, a->ldb, 64, BLOCK_K);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<DROP, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_pairs = cb_host::ceil_div(a->m, 2 * BLOCK_M), n_tiles = cb_host::ceil_div(a->n, C::BLOCK_N);
  const int items = m_pairs * n_tiles * p.split_k;
  const int max_clusters = cb_host::num_sms() / 2;
  const int clusters = items < max_clusters ? items : max_clusters;
  cb_host::ProfScope prof(cb_host::PROF_GEMM, stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_2cta_kernel<DROP, FAST>, ta, tb, p));
  cb_host::count_launch();
  return 0;
}

extern "C" int commu_gemm_bf16(const commu_gemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CB_REQUIRE(a && a->a && a->b, "gemm: null operand");
  CB_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "gemm: bad shape m=%d n=%d k=%d", a->m, a->n, a->k);
  CB_REQUIRE(a->out_bf16 || a->out_f32, "gemm: no output");
  int split = a->split_k < 1 ? 1 : a->split_k;
  const int kb_total = cb_host::ceil_div(a->k, BLOCK_K);
  if (split > kb_total) split = kb_total;
  if (split > 1) {
    CB_REQUIRE(a->out_f32 && a->f32_atomic && !a->out_bf16 && !a->bias && !a->relu &&
                   !a->relu_mask && !a->add_f32,
               "gemm: split_k>1 needs a pure atomic fp32 epilogue");
    // make every split non-empty
    const int per = cb_host::ceil_div(kb_total, split);
    split = cb_host::ceil_div(kb_total, per);
  }
  GemmKernelParams p;
  p.M = a->m; p.N = a->n; p.K = a->k;
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.split_k = split;
  p.alpha = a->alpha;
  p.bias = a->bias;
  p.relu = a->relu;
  p.relu_mask = static_cast<const bf16*>(a->relu_mask);
  p.ld_mask = a->ld_mask;
  p.add_f32 = a->add_f32;
  p.ld_add = a->ld_add;
  p.out_bf16 = static_cast<bf16*>(a->out_bf16);
  p.ld_out_bf16 = a->ld_out_bf16;
  p.out_f32 = a->out_f32;
  p.ld_out_f32 = a->ld_out_f32;
  p.f32_atomic = a->f32_atomic;
  p.drop_thr2 = 0; p.drop_ka = p.drop_kb = 0; p.drop_inv = 1.f;
  if (a->drop_p > 0.f) {
    CB_REQUIRE(a->drop_p < 1.f && split == 1 && a->impl != 1, "gemm: fused dropout needs 0 < p < 1, no split-k, a tcgen05 kernel");
    const uint32_t thr = drop::thr15_of(a->drop_p);
    p.drop_thr2 = thr * 0x00010001u;
    p.drop_inv = 1.f / (1.f - (float)thr / 32768.f);
    p.drop_ka = (uint32_t)a->drop_seed;
    p.drop_kb = (uint32_t)(a->drop_seed >> 32);
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = 1;
  {
    static const int dbg = []() { const char* e = getenv("COMMU_GEMM_EPI_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
  }
  if (p.out_bf16 && (!al16(p.out_bf16) || (p.ld_out_bf16 % 8))) p.vec_ok = 0;
  if (p.out_f32 && (!al16(p.out_f32) || (p.ld_out_f32 % 4))) p.vec_ok = 0;
  if (p.relu_mask && (!al16(p.relu_mask) || (p.ld_mask % 8))) p.vec_ok = 0;
  if (p.add_f32 && (!al16(p.add_f32) || (p.ld_add % 4))) p.vec_ok = 0;

  if (a->impl == 1) {
    dim3 grid(cb_host::ceil_div(a->n, 128), a->m, split);
    gemm_simt_kernel<<<grid, 128, 0, stream>>>(static_cast<const bf16*>(a->a), a->lda,
                                               static_cast<const bf16*>(a->b), a->ldb, p);
    cb_host::count_launch();
    CB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  CB_REQUIRE(al16(a->a) && al16(a->b) && (a->lda % 8 == 0) && (a->ldb % 8 == 0),
             "gemm: TMA needs 16-byte aligned operands and leading dims that are multiples of 8 "
             "(lda=%lld ldb=%lld)", (long long)a->lda, (long long)a->ldb);
  // impl 2 forces the CTA-pair kernel, impl 3 the one-CTA kernel; impl 0 picks the pair kernel for wide outputs
  // with at least two row tiles and a long reduction (COMMU_GEMM_2CTA=0 disables it)
  static const int use_2cta = []() {
    const char* e = getenv("COMMU_GEMM_2CTA");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  // (measured with the whole-tile epilogues, profiles/r02_gemm_shapes.json: +10-14 % where a split's reduction length
  //  is >= 1024; on the K = 512 shapes the one-CTA kernel is 5-8 % faster)
  const int k_per_split = a->k / split;
  const bool pair = a->impl == 2 || (a->impl == 0 && use_2cta && a->n > 128 && a->m > BLOCK_M && k_per_split >= 1024);
  const int block_n = (pair || a->n > 128) ? 256 : 128;
  // whole-tile epilogues: every tile full, vector-aligned buffers, bf16-only result (optional ReLU mask) or fp32-only
  // result (optional addend / atomic accumulation)
  static const int fast_on = []() { const char* e = getenv("COMMU_GEMM_FAST_EPI"); return (e && e[0] == '0') ? 0 : 1; }();
  const bool path_a = p.out_bf16 && !p.add_f32 && !p.out_f32;
  const bool path_f = p.out_f32 && !p.out_bf16 && !p.relu_mask;
  const bool fast = fast_on && p.vec_ok && (a->m % (pair ? 2 * BLOCK_M : BLOCK_M) == 0) && (a->n % block_n == 0) && (path_a || path_f);
  const bool drop = p.drop_thr2 != 0;
#define COMMU_GEMM_DISPATCH(FN, ...)                                                        \
  (drop ? (fast ? FN<__VA_ARGS__ true, true>(a, p, stream) : FN<__VA_ARGS__ true, false>(a, p, stream)) \
        : (fast ? FN<__VA_ARGS__ false, true>(a, p, stream) : FN<__VA_ARGS__ false, false>(a, p, stream)))
  if (pair) return COMMU_GEMM_DISPATCH(launch_tc_2cta, );
  if (block_n == 256) return COMMU_GEMM_DISPATCH(launch_tc, 256, );
  return COMMU_GEMM_DISPATCH(launch_tc, 128, );
#undef COMMU_GEMM_DISPATCH
}
