// Elementwise dropout (+ optional residual add): the nn.Dropout sites of the reference outside the attention
// probabilities - embedding, positional table, attention output, FF hidden, FF output, final hidden
// (commu/model/model.py:166-168, 349, 585-586, 600).  Masks come from dropout.cuh, so the backward pass calls
// the same kernel with the same seed on the incoming gradient.
#include "api_common.h"
#include "common.cuh"
#include "dropout.cuh"

namespace {

template <bool X_BF16>
__global__ void dropout_kernel(const void* __restrict__ x_, long long ldx, const float* __restrict__ res, long long ldres,
                               long long rows, int cols, uint32_t thr2, float inv, uint32_t ka, uint32_t kb,
                               float* __restrict__ out_f32, long long ldo, bf16* __restrict__ out_bf16, long long ldob) {
  const int c4n = cols >> 2;
  const long long total = rows * c4n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / c4n;
    const int c4 = (int)(idx - row * c4n);
    float v[4];
    if (X_BF16) {
      const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const bf16*>(x_) + row * ldx + c4 * 4);
      v[0] = cb::bf16_lo(raw.x); v[1] = cb::bf16_hi(raw.x); v[2] = cb::bf16_lo(raw.y); v[3] = cb::bf16_hi(raw.y);
    } else {
      const float4 raw = *reinterpret_cast<const float4*>(static_cast<const float*>(x_) + row * ldx + c4 * 4);
      v[0] = raw.x; v[1] = raw.y; v[2] = raw.z; v[3] = raw.w;
    }
    const uint2 rnd = drop::rand64((uint32_t)c4, drop::row_keys(ka, kb, (uint32_t)row));
    const uint32_t f0 = drop::keep_flags(rnd.x, thr2), f1 = drop::keep_flags(rnd.y, thr2);
    v[0] = (f0 & 0x8000u) ? v[0] * inv : 0.f;
    v[1] = (f0 & 0x80000000u) ? v[1] * inv : 0.f;
    v[2] = (f1 & 0x8000u) ? v[2] * inv : 0.f;
    v[3] = (f1 & 0x80000000u) ? v[3] * inv : 0.f;
    if (res) {
      const float4 rr = *reinterpret_cast<const float4*>(res + row * ldres + c4 * 4);
      v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * ldo + c4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
    if (out_bf16)
      *reinterpret_cast<uint2*>(out_bf16 + row * ldob + c4 * 4) = make_uint2(cb::pack_bf16(v[0], v[1]), cb::pack_bf16(v[2], v[3]));
  }
}

}  // namespace

// out = res + keep(x) / (1 - p) over [rows, cols]; x is fp32 or bf16 (x_is_bf16), res (fp32) optional, outputs
// fp32 and / or bf16, in place allowed.  cols and every leading dimension must be multiples of 4.
// seed: any 64-bit value; a (seed, row, column) triple always yields the same keep decision.
extern "C" int commu_dropout(const void* x, int x_is_bf16, int64_t ldx, const float* res, int64_t ldres, int64_t rows,
                             int cols, float p, uint64_t seed, float* out_f32, int64_t ldo, void* out_bf16,
                             int64_t ldob, void* stream) {
  CB_REQUIRE(x && rows > 0 && cols > 0 && (out_f32 || out_bf16), "dropout: bad args");
  CB_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  CB_REQUIRE(cols % 4 == 0 && ldx % 4 == 0 && (!res || ldres % 4 == 0) && (!out_f32 || ldo % 4 == 0) &&
                 (!out_bf16 || ldob % 4 == 0),
             "dropout: columns and leading dimensions must be multiples of 4");
  const uint32_t thr2 = drop::thr15_of(p) * 0x00010001u;
  const float inv = 1.f / (1.f - (float)drop::thr15_of(p) / 32768.f);
  const uint32_t ka = (uint32_t)seed, kb = (uint32_t)(seed >> 32);
  const long long total = rows * (cols / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)cb_host::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (x_is_bf16)
    dropout_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, res, ldres, rows, cols, thr2, inv, ka, kb,
                                                                             out_f32, ldo, (bf16*)out_bf16, ldob);
  else
    dropout_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, res, ldres, rows, cols, thr2, inv, ka, kb,
                                                                              out_f32, ldo, (bf16*)out_bf16, ldob);
  cb_host::count_launch();
  CB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
