// Helpers shared by the tcgen05 attention kernels (forward and backward): TMEM stores, TS-mode MMA,
// fast exp2, explicit shared-space accesses, named barriers, mbarrier ring bookkeeping and the 3-D
// tensor maps over [rows, B, cols] activations.
#pragma once
#include <cuda_fp16.h>
#include "api_common.h"
#include "common.cuh"
#include "dropout.cuh"

namespace attn_tc {

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// explicit shared-space accesses with 32-bit addresses: ptxas folds `base + constant` into the
// instruction immediate, so the unrolled shear reads cost one LDS each (no 64-bit pointer math)
__device__ __forceinline__ float lds_f16(uint32_t addr) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- shear helpers for a thread that owns (tile row li, 32-column chunk g); wq = li / 32 ----
// TMEM block columns [g*32, g*32+32) of this thread's lane -> fp16 at dst (64 bytes of its staging row)
__device__ __forceinline__ void stage32(uint32_t taddr, uint32_t dst) {
  uint32_t r0[32];
  cb::tmem_ld_32x32b_x32(taddr, r0);
  cb::tmem_ld_wait();
#pragma unroll
  for (int e = 0; e < 32; e += 8)
    sts_v4(dst + e * 2, pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1])),
           pack_f16(__uint_as_float(r0[e + 2]), __uint_as_float(r0[e + 3])),
           pack_f16(__uint_as_float(r0[e + 4]), __uint_as_float(r0[e + 5])),
           pack_f16(__uint_as_float(r0[e + 6]), __uint_as_float(r0[e + 7])));
}
// s[e] (+)= band value of tile column lc = 32g + e, read from ONE staged 128-wide block at `row`
// (row = this thread's staging row).  Band column of lc is li + 127 - lc: PASS 0 = "lo" block
// (holds lc >= li at index li+127-lc), PASS 1 = "hi" block (holds lc < li at index li-1-lc).
// Which block a chunk needs is warp-uniform except on the diagonal chunk g == wq.
template <int PASS, bool ACCUM>
__device__ __forceinline__ void band_add(float (&s)[32], uint32_t row, int li, int g, int wq) {
  const uint32_t base = PASS == 0 ? row + 2 * (li + 127) : row + 2 * (li - 1);
  const bool all = PASS == 0 ? g > wq : g < wq;
  if (all) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float v = lds_f16(base - 2 * (g * 32 + e));
      s[e] = ACCUM ? s[e] + v : v;
    }
  } else if (g == wq) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int lc = g * 32 + e;
      const bool use = PASS == 0 ? lc >= li : lc < li;
      const float v = lds_f16(use ? base - 2 * lc : row);
      if (use) s[e] = ACCUM ? s[e] + v : v;
    }
  }
}

// Variants for an UNPADDED 256-byte staging row (so the 32 KB buffer can alias a 128x128 bf16 tile): the row
// is rotated by rot = (li & 7) 16-byte chunks, which keeps the 16-byte stores of 8 consecutive rows on
// distinct bank groups; element idx lives at byte (2*idx + 16*rot) mod 256.
__device__ __forceinline__ void stage32_rot(uint32_t taddr, uint32_t row, int g, int rot) {
  uint32_t r0[32];
  cb::tmem_ld_32x32b_x32(taddr, r0);
  cb::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int e = c * 8;
    sts_v4(row + (((g * 4 + c + rot) & 15) << 4), pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1])),
           pack_f16(__uint_as_float(r0[e + 2]), __uint_as_float(r0[e + 3])),
           pack_f16(__uint_as_float(r0[e + 4]), __uint_as_float(r0[e + 5])),
           pack_f16(__uint_as_float(r0[e + 6]), __uint_as_float(r0[e + 7])));
  }
}
template <int PASS>
__device__ __forceinline__ void band_add_rot(float (&s)[32], uint32_t row, int li, int g, int wq, int rot) {
  // byte offset of tile column lc = 32g + e inside the rotated row: (A - 2e) mod 256
  const int A = (PASS == 0 ? 2 * (li + 127 - 32 * g) : 2 * (li - 1 - 32 * g)) + 16 * rot;
  const bool all = PASS == 0 ? g > wq : g < wq;
  if (all) {
#pragma unroll
    for (int e = 0; e < 32; ++e) s[e] += lds_f16(row + ((A - 2 * e) & 255));
  } else if (g == wq) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int lc = g * 32 + e;
      const bool use = PASS == 0 ? lc >= li : lc < li;
      const float v = lds_f16(row + ((A - 2 * e) & 255));
      if (use) s[e] += v;
    }
  }
}

// ---- combined staging ----
// Row li of a 128 x 128 tile needs band indices idx >= li from the "lo" block and idx < li from the "hi"
// block (idx = the block column): together exactly one 128-entry row.  Thread (li, g) therefore stages
// its 32 block columns c = 32g + e from ONE block (lo if g > wq, hi if g < wq) and only the diagonal
// chunk g == wq from both, into a private padded fp16 row (kStageRow bytes apart), and the shear read
// becomes a circular read of that row: tile column lc holds index (li + 127 - lc) mod 128.
constexpr int kStageRow = 272;
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// "hi" part of the diagonal chunk: overwrite the entries with block column c = 32g + e < li, i.e. e < lane
__device__ __forceinline__ void stage32_diag_hi(uint32_t taddr, uint32_t dst, int lane) {
  uint32_t r0[32];
  cb::tmem_ld_32x32b_x32(taddr, r0);
  cb::tmem_ld_wait();
#pragma unroll
  for (int e = 0; e < 31; ++e) {
    unsigned short hv;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(hv) : "f"(__uint_as_float(r0[e])));
    if (e < lane) sts_u16(dst + 2 * e, hv);
  }
}
// the same two helpers with 16-column TMEM reads (lower register peak, for the dR kernel)
__device__ __forceinline__ void stage32_h(uint32_t taddr, uint32_t dst) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[16];
    tmem_ld_32x32b_x16(taddr + half * 16, r0);
    cb::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; e += 8)
      sts_v4(dst + half * 32 + e * 2, pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1])),
             pack_f16(__uint_as_float(r0[e + 2]), __uint_as_float(r0[e + 3])),
             pack_f16(__uint_as_float(r0[e + 4]), __uint_as_float(r0[e + 5])),
             pack_f16(__uint_as_float(r0[e + 6]), __uint_as_float(r0[e + 7])));
  }
}
__device__ __forceinline__ void stage32_diag_hi_h(uint32_t taddr, uint32_t dst, int lane) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[16];
    tmem_ld_32x32b_x16(taddr + half * 16, r0);
    cb::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      unsigned short hv;
      asm("cvt.rn.f16.f32 %0, %1;" : "=h"(hv) : "f"(__uint_as_float(r0[e])));
      if (half * 16 + e < lane) sts_u16(dst + 2 * (half * 16 + e), hv);
    }
  }
}
// s[e] += band value of tile column lc = 32g + e out of the combined row
__device__ __forceinline__ void band_read(float (&s)[32], uint32_t row, int li, int g, int wq, int lane) {
  const uint32_t base1 = row + 2 * (li - 1) - 64 * g;     // idx = li - 1 - lc   (lc <  li)
  const uint32_t base0 = base1 + 256;                     // idx = li + 127 - lc (lc >= li)
  if (g > wq) {
#pragma unroll
    for (int e = 0; e < 32; ++e) s[e] += lds_f16(base0 - 2 * e);
  } else if (g < wq) {
#pragma unroll
    for (int e = 0; e < 32; ++e) s[e] += lds_f16(base1 - 2 * e);
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) s[e] += lds_f16((e >= lane ? base0 : base1) - 2 * e);
  }
}

// ---- v3 shear: positions, packed reads, mixed-precision adds ----
// A row group (the 32 query rows of TMEM lane quadrant wq) needs, for row li = 32*wq + l, the band indices
// idx in [li, li+127]: idx < 128 comes from the "lo" 128-distance block, idx >= 128 from the "hi" block at
// idx - 128.  Over the whole group that is the index window [32*wq, 32*wq + 160).  A staged row therefore
// holds 160 fp16 POSITIONS (320 bytes, placed by stage_row_off): thread (li, g) copies its 32 block columns
// 32g..32g+31 of "lo" to positions 32g (only if g >= wq) and of "hi" to positions 128 + 32g (only if
// g <= wq).  The relative shift is then the plain descending read  pos = li + 127 - lc  with no select:
// `row_v` below is the row's base minus 64*wq bytes, so that byte offset = 2 * pos.
// Placement of the 32 staged rows (320 bytes each) of a row group inside its private piece (kStageGroupBytes):
// row l = 8k + m sits in region m>>1 (a region = 8 rows + one 16-byte gap) at slot 2k + (m&1).  The 16-byte
// stores of 8 consecutive lanes then start in 8 different 16-byte columns of a 128-byte line, and the sheared
// 32-bit reads of 32 consecutive rows fall into 32 different banks: both directions are conflict-free.
constexpr int kStageGroupBytes = 10288;
__device__ __forceinline__ uint32_t stage_row_off(int l) {
  return 16u * (161u * ((l & 7) >> 1) + 20u * (2u * (l >> 3) + (l & 1)));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned short lds_u16(uint32_t addr) {
  unsigned short h;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
  return h;
}
__device__ __forceinline__ void fhadd1(float& x, unsigned short h) {
  asm("add.f32.f16 %0, %1, %0;" : "+f"(x) : "h"(h));
}
// x0 += f16 low half of pr, x1 += f16 high half (FHADD: one instruction each, no separate convert)
__device__ __forceinline__ void fhadd2(float& x0, float& x1, uint32_t pr) {
  asm("{\n\t.reg .b16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "add.f32.f16 %0, l, %0;\n\t"
      "add.f32.f16 %1, h, %1;\n\t}"
      : "+f"(x0), "+f"(x1)
      : "r"(pr));
}
// 32 TMEM values -> 32 fp16 at dst (64 bytes, 16-byte aligned)
__device__ __forceinline__ void pack_store32(uint32_t dst, const uint32_t (&r0)[32]) {
#pragma unroll
  for (int e = 0; e < 32; e += 8)
    sts_v4(dst + e * 2, pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1])),
           pack_f16(__uint_as_float(r0[e + 2]), __uint_as_float(r0[e + 3])),
           pack_f16(__uint_as_float(r0[e + 4]), __uint_as_float(r0[e + 5])),
           pack_f16(__uint_as_float(r0[e + 6]), __uint_as_float(r0[e + 7])));
}
// 32 TMEM columns -> 16 registers of fp16 pairs (the TMEM read completes here)
__device__ __forceinline__ void load_pack32(uint32_t taddr, uint32_t (&pk)[16]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {   // two 16-column reads: lower register peak
    uint32_t r0[16];
    tmem_ld_32x32b_x16(taddr + half * 16, r0);
    cb::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; e += 2)
      pk[half * 8 + e / 2] = pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1]));
  }
}
__device__ __forceinline__ void store_packed32(uint32_t dst, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) sts_v4(dst + 16 * c, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}
// s[e] += fp16 at (addr0 - 2e), e = 0..31.  addr0 is only 2-byte aligned: 17 aligned 32-bit words cover the
// 64-byte window and one PRMT per pair (selector by the alignment phase) puts (e, e+1) into (low, high).
__device__ __forceinline__ void shear_add32(float (&s)[32], uint32_t addr0) {
  const uint32_t aw = addr0 & ~3u;
  const uint32_t sel = (addr0 & 2u) ? 0x1032u : 0x7610u;
#pragma unroll
  for (int half = 0; half < 2; ++half) {   // two halves of 8 pairs: 9 words in flight keeps the register peak low
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = lds_u32(aw - 4 * (half * 8 + k));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t pr;
      asm("prmt.b32 %0, %1, %2, %3;" : "=r"(pr) : "r"(w[q]), "r"(w[q + 1]), "r"(sel));
      fhadd2(s[half * 16 + 2 * q], s[half * 16 + 2 * q + 1], pr);
    }
  }
}
// 16-column variant: s[e] += fp16 at (addr0 - 2e), e = 0..15
__device__ __forceinline__ void shear_add16(float (&s)[16], uint32_t addr0) {
  const uint32_t aw = addr0 & ~3u;
  const uint32_t sel = (addr0 & 2u) ? 0x1032u : 0x7610u;
  uint32_t w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = lds_u32(aw - 4 * k);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t pr;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(pr) : "r"(w[q]), "r"(w[q + 1]), "r"(sel));
    fhadd2(s[2 * q], s[2 * q + 1], pr);
  }
}
// 16-column variant of the TMEM -> fp16 staging copy (lower register peak)
__device__ __forceinline__ void stage32_x16(uint32_t taddr, uint32_t dst) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[16];
    tmem_ld_32x32b_x16(taddr + half * 16, r0);
    cb::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; e += 8)
      sts_v4(dst + half * 32 + e * 2, pack_f16(__uint_as_float(r0[e]), __uint_as_float(r0[e + 1])),
             pack_f16(__uint_as_float(r0[e + 2]), __uint_as_float(r0[e + 3])),
             pack_f16(__uint_as_float(r0[e + 4]), __uint_as_float(r0[e + 5])),
             pack_f16(__uint_as_float(r0[e + 6]), __uint_as_float(r0[e + 7])));
  }
}

// Shared-memory matrix descriptor without swizzle (layout type 0): core matrices of 8 rows x 16 bytes.
//   MN-major: a core matrix holds 8 MN-elements (16 B) x 8 K-rows (16 B apart, 128 B total); `k_group_bytes`
//             strides 8-K-row groups (LBO field), `mn_group_bytes` strides 8-element MN groups (SBO field).
//   K-major : a core matrix holds 8 MN-rows x 8 K-elements; LBO strides the K chunks, SBO the 8-row groups.
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version = 1
  return d;
}
// the two bf16 halves of `packed` to two shared-memory addresses
__device__ __forceinline__ void sts_halves(uint32_t addr_lo, uint32_t addr_hi, uint32_t packed) {
  asm volatile("{\n\t.reg .b16 l, h;\n\t"
               "mov.b32 {l, h}, %2;\n\t"
               "st.shared.b16 [%0], l;\n\t"
               "st.shared.b16 [%1], h;\n\t}" ::"r"(addr_lo), "r"(addr_hi), "r"(packed) : "memory");
}

// ---- backward: probabilities and score gradients of 16 consecutive key columns of one query row ----
//   P  = exp2(s * sl2 - lse2)                       (lse2 already holds -log2(1/keep) under dropout: P / keep)
//   dS = P * (keep-masked dP - delta)                (delta already scaled by keep under dropout)
// s: scores (AC + BD), dpr: raw dP bits from TMEM, j_first: key index of column 0 (multiple of 4).
// MASKED applies the analytic attention mask (boundary tiles), WANT_P also emits the (dropped) P pairs.
template <bool DROP, bool MASKED, bool WANT_P>
__device__ __forceinline__ void pds16(const float* s, const uint32_t* dpr, float sl2, float lse2, float delta, int j_first,
                                      int hi_i, int lo_i, drop::Keys dk, uint32_t thr2, uint32_t* pk8, uint32_t* dsk8) {
#pragma unroll
  for (int e = 0; e < 16; e += 4) {
    uint32_t f0 = 0, f1 = 0;
    if (DROP) {
      const uint2 rnd = drop::rand64((uint32_t)((j_first + e) >> 2), dk);
      f0 = drop::keep_flags(rnd.x, thr2);
      f1 = drop::keep_flags(rnd.y, thr2);
    }
    float pv[4], dsv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pv[u] = ex2(fmaf(s[e + u], sl2, -lse2));
      if (MASKED) {
        const int j = j_first + e + u;
        pv[u] = (j > hi_i || j < lo_i) ? 0.f : pv[u];
      }
      uint32_t dbits = dpr[e + u];
      if (DROP) dbits &= (u == 0 ? drop::mask32_lo(f0) : u == 1 ? drop::mask32_hi(f0) : u == 2 ? drop::mask32_lo(f1) : drop::mask32_hi(f1));
      dsv[u] = pv[u] * (__uint_as_float(dbits) - delta);
    }
    if (WANT_P) {
      uint32_t p01 = cb::pack_bf16(pv[0], pv[1]), p23 = cb::pack_bf16(pv[2], pv[3]);
      if (DROP) {
        p01 &= drop::mask16x2(f0);
        p23 &= drop::mask16x2(f1);
      }
      pk8[e / 2] = p01;
      pk8[e / 2 + 1] = p23;
    }
    dsk8[e / 2] = cb::pack_bf16(dsv[0], dsv[1]);
    dsk8[e / 2 + 1] = cb::pack_bf16(dsv[2], dsv[3]);
  }
}

// Keep flags for keys in DESCENDING order (the dR pass walks distances): element e of this thread is key
// jk0 - e.  Returns the flags of the 8 elements 8c .. 8c+7 as 4 words, element 2q in the HIGH half and 2q+1 in
// the low half (mask32_hi / mask32_lo).  The fields are the same bits the ascending kernels use: key j is field
// j & 3 of rand64(j >> 2); in descending order a group contributes [y.hi, y.lo, x.hi, x.lo].
__device__ __forceinline__ void drop_flags_desc8(uint32_t (&fl)[4], int jk0, int c, drop::Keys dk, uint32_t thr2) {
  const int sk = 3 - (jk0 & 3);          // stream position of element 0 (16-bit fields, high half first)
  const int g0 = jk0 >> 2;
  uint32_t st[6];                        // stream words 4c .. 4c+5 (three groups)
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const uint2 rnd = drop::rand64((uint32_t)(g0 - 2 * c - q), dk);
    st[2 * q] = rnd.y;
    st[2 * q + 1] = rnd.x;
  }
  const bool woff = (sk >> 1) != 0;
  const uint32_t sh = (sk & 1) ? 16u : 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t w0 = woff ? st[q + 1] : st[q];
    const uint32_t w1 = woff ? st[q + 2] : st[q + 1];
    fl[q] = drop::keep_flags(__funnelshift_l(w1, w0, sh), thr2);
  }
}

struct Ring {  // stage / parity bookkeeping of a 2-deep mbarrier ring
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    idx ^= 1;
    if (idx == 0) phase ^= 1;
  }
};


// current attention-dropout setting (commu_relattn_set_dropout, attn_fwd_tc.cu); copied into Params by the tcgen05 entry points
struct DropState { float p; unsigned long long seed; };
DropState drop_state();
template <class ParamsT>
inline void apply_drop_state(ParamsT& p) {
  const DropState d = drop_state();
  const uint32_t thr = d.p > 0.f ? drop::thr15_of(d.p) : 0u;
  p.drop_thr2 = thr * 0x00010001u;
  p.drop_ka = (uint32_t)d.seed;
  p.drop_kb = (uint32_t)(d.seed >> 32);
  p.drop_keep = 1.f - (float)thr / 32768.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3D map over a [rows, B, cols] bf16 tensor (row = i*B + b): dims {cols, B, rows}, box {64, 1, 128}
inline int make_tmap_rows3d(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t B, uint64_t rows, uint64_t ld) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr_fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr_fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return cb_host::fail(COMMU_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(ptr_fn);
  }
  cuuint64_t dims[3] = {cols, B, rows};
  cuuint64_t strides[2] = {ld * 2, ld * 2 * B};
  cuuint32_t box[3] = {64, 1, 128};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cb_host::fail(COMMU_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return 0;
}


}  // namespace attn_tc
