// Counter-based dropout masks shared by the elementwise dropout kernel and the attention kernels.
//
// The reference draws its masks from torch's generator (nn.Dropout, commu/model/model.py:166-168, 210-211,
// 337, 349, 454, 585-586, 600); bit-equal masks are out of reach, so the contract here is statistical:
// every element is kept independently with probability 1 - p and kept values are scaled by 1 / (1 - p).
// A mask is a pure function of (key, row, column): the backward kernels recompute it, nothing is stored.
//
// 64 bits per counter (three multiply-xor rounds, Philox-style) feed four consecutive columns with 15-bit
// fields: element (row, col) uses field col & 3 of rand64(col >> 2, row keys); keep <=> field >= thr15 with
// thr15 = round(p * 32768).  tests/helpers.py holds the same arithmetic in numpy (the parity tests apply the
// identical mask to the torch reference).
#pragma once
#include <stdint.h>

namespace drop {

struct Keys { uint32_t a, b; };

__host__ __device__ __forceinline__ Keys row_keys(uint32_t ka, uint32_t kb, uint32_t row) {
  Keys k;
  k.a = ka ^ (row * 0x9E3779B1u);
  k.b = kb + row * 0x85EBCA77u;
  return k;
}
// out.x: fields of columns 4c, 4c+1 (low, high half); out.y: columns 4c+2, 4c+3
__device__ __forceinline__ uint2 rand64(uint32_t ctr, Keys k) {
  const uint64_t x = (uint64_t)(ctr ^ k.a) * 0xD2511F53u;
  const uint32_t y = (uint32_t)(x >> 32) ^ (uint32_t)x ^ k.b;
  const uint64_t z = (uint64_t)y * 0xCD9E8D57u;
  const uint32_t w = (uint32_t)(z >> 32) ^ (uint32_t)z ^ k.a;
  const uint64_t r = (uint64_t)w * 0x9E3779B1u;
  return make_uint2((uint32_t)(r >> 32) ^ y, (uint32_t)r ^ (uint32_t)(z >> 32));
}
// Two 15-bit fields of `x` against the threshold replicated in both halves (thr2 = thr15 * 0x00010001):
// bit 15 / bit 31 of the result are the keep flags of the low / high column (no borrow between the halves).
__device__ __forceinline__ uint32_t keep_flags(uint32_t x, uint32_t thr2) {
  return ((x & 0x7FFF7FFFu) | 0x80008000u) - thr2;
}
// flags -> 0xFFFF per kept half (mask for a packed bf16 / fp16 pair)
__device__ __forceinline__ uint32_t mask16x2(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(flags));
  return m;
}
// flags -> 32-bit all-ones masks of the low / high column
__device__ __forceinline__ uint32_t mask32_lo(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(flags));
  return m;
}
__device__ __forceinline__ uint32_t mask32_hi(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBBBB;" : "=r"(m) : "r"(flags));
  return m;
}
__host__ __device__ __forceinline__ uint32_t thr15_of(float p) {
  const int t = (int)(p * 32768.f + 0.5f);
  return (uint32_t)(t < 0 ? 0 : (t > 32767 ? 32767 : t));
}

}  // namespace drop
