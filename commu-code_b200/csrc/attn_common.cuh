// Shared pieces of the relative-position attention kernels (forward and backward).
//
// Semantics (reference commu/model/model.py:280-345 and :549-574):
//   score[b,h,i,j] = Dh^-0.5 * ( (q_i + r_w_bias_h) . k_j  +  (q_i + r_r_bias_h) . R[i + M - j] )
//   valid iff  j <= i + M,  and (same_length -> j > i - shift),  and (reset[b] -> j >= M)
//   P = softmax_j(score),  out_i = sum_j P_ij v_j
// R is indexed by DISTANCE (row delta <-> pos_emb row K-1-delta of the reference), so the
// reference's _rel_shift (model.py:251-265) becomes the index  delta = i + M - j.
//
// Tiling: a CTA owns BM=64 query rows of one (b,h); each of its 4 warps owns 16 rows and walks the
// key tiles (BN=64).  For a (query-tile, key-tile) pair the distances form a band of BM+BN-1
// consecutive rows of R; a warp multiplies its 16 rows of (q + r_r_bias) with its 80-row slice of
// that band and re-reads the product along anti-diagonals (c = li + BN-1 - lj) through a private
// shared-memory scratch: that is the whole "relative shift".
#pragma once
#include "common.cuh"

namespace attn {

constexpr int DH = 64;    // padded head dim
constexpr int BM = 64;    // query rows per CTA
constexpr int BN = 64;    // keys per tile
constexpr int BAND = 128; // R rows staged per key tile (>= BM + BN - 1)
constexpr int WBAND = 80; // band rows a warp multiplies (>= 16 + BN - 1, multiple of 8)
constexpr int SW = 88;    // scratch row stride in floats (bank-spread, even)

struct Params {
  const bf16* q;    // [T*B, ldq]   (+ h*64)   raw projected queries (forward only)
  bf16* qu_s;       // [T*B, ldq]   bf16(q + r_w_bias): written by forward, read by backward
  bf16* qv_s;       // [T*B, ldq]   bf16(q + r_r_bias)
  const bf16* k;    // [K*B, ldkv]
  const bf16* v;    // [K*B, ldkv]
  const bf16* r;    // [Kr, ldr]     by distance
  const float* u;   // r_w_bias [H,64] (padded, fp32)
  const float* vb;  // r_r_bias [H,64]
  const unsigned char* reset;  // [B] or null
  long long ldq, ldkv, ldr;
  int T, M, B, H, Kr;
  int same_length, shift;  // shift = mask_shift_len (valid iff j > i - shift)
  float scale;             // 1/sqrt(real Dh)
  // dropout on the attention probabilities (dropout.cuh); drop_thr2 == 0 -> none.  Set by commu_relattn_set_dropout.
  uint32_t drop_thr2;      // 15-bit threshold replicated in both halves
  uint32_t drop_ka, drop_kb;
  float drop_keep;         // effective keep probability 1 - thr15 / 32768
  // forward outputs
  bf16* out;  // [T*B, ldo]
  long long ldo;
  float* lse;  // [B, H, T]
  // backward inputs/outputs
  const bf16* dout;  // [T*B, lddo]
  long long lddo;
  const float* delta;  // [B,H,T] rowsum(dO * O)
  bf16* dq;            // [T*B, lddq]  (written once per row by the dq pass)
  long long lddq;
  float* dk;  // fp32 [K*B, lddkv] accumulated (zeroed by caller) -- or bf16 outputs, see bwd
  float* dv;
  long long lddkv;
  float* dr;   // fp32 [Kr, H*64] atomically accumulated over batch
  float* du;   // fp32 [H,64] atomically accumulated
  float* dvb;  // fp32 [H,64]
};

// swizzled byte offset inside a [rows][64] bf16 tile (128 B per row, 16 B chunks XOR row&7)
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// smallest visible key of query row i (keys below it are masked)
__device__ __forceinline__ int key_lo(int i, int M, int same_length, int shift, bool reset) {
  int lo = reset ? M : 0;
  if (same_length) lo = max(lo, i - shift + 1);
  return lo;
}

// cp.async a [rows x 64] bf16 tile from global (row r0+rr of a matrix with `n_rows` rows, row
// stride `ld` elements, already offset to the head's first column) into a swizzled smem tile.
// Rows outside [0, n_rows) are zero-filled.
template <int ROWS, int NTHREADS>
__device__ __forceinline__ void load_tile_async(uint8_t* smem_tile, const bf16* g, long long ld,
                                                long long row_mul, int r0, int n_rows, int tid) {
#pragma unroll
  for (int it = 0; it < (ROWS * 8) / NTHREADS; ++it) {
    const int idx = it * NTHREADS + tid;
    const int rr = idx >> 3, ch = idx & 7;
    const int r = r0 + rr;
    const bool ok = (r >= 0) && (r < n_rows);
    const bf16* src = g + (ok ? ((long long)r * row_mul) * ld + ch * 8 : 0);
    cb::cp_async16(smem_tile + swz(rr, ch), src, ok);
  }
}

}  // namespace attn
