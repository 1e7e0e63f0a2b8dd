/*
 * commu_b200.h -- C-ABI of libcommu_b200.so: hand-written sm_100a kernels for the ComMU
 * Transformer-XL hot path.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every entry point returns 0 on success and a negative code on failure; the message is
 *     available from commu_last_error() (thread-local).  Nothing throws.
 *   - every device pointer is owned by the caller (PyTorch's caching allocator in this repo);
 *     kernels never allocate.  `stream` is a cudaStream_t passed as void*.
 *   - "bf16" buffers are __nv_bfloat16, row-major with an explicit leading dimension in ELEMENTS.
 *   - the reference has no FFI of its own (it is pure Python on top of PyTorch); each entry point
 *     therefore cites the reference Python code it replaces (paths relative to the reference
 *     repo root, POZAlabs/ComMU-code @ 3949a5b).  INTEGRATION.md shows the ctypes binding.
 */
#ifndef COMMU_B200_H_
#define COMMU_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMMU_OK 0
#define COMMU_ERR_INVALID (-1)
#define COMMU_ERR_CUDA (-2)
#define COMMU_ERR_UNSUPPORTED (-3)
#define COMMU_ERR_NCCL (-4)

const char* commu_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int commu_abi_version(void);
/* Fills the compute capability of the current device; returns COMMU_ERR_CUDA without a GPU. */
int commu_device_info(int* sm_major, int* sm_minor, int* num_sms);

/* ------------------------------------------------------------------------------------------
 * Dense contraction  C[m,n] = alpha * sum_k A[m,k] * B[n,k]   (bf16 operands, fp32 accumulate)
 * on tcgen05 tensor cores with TMA-staged operands and a TMEM accumulator.
 * Replaces: nn.Linear / F.linear calls of commu/model/model.py:285-286 (qkv_net, r_net), :348
 * (o_net), :163-169 (CoreNet), :46 (logits) and their autograd backward (dgrad / wgrad).
 *
 *   a_mn_major = 0: A stored [m, k] row-major (k contiguous), lda = row stride.
 *   a_mn_major = 1: A stored [k, m] row-major (m contiguous), lda = row stride. (wgrad operands)
 *   same for B with n.
 * Epilogue (all optional, applied in this order):
 *   v = alpha*acc ; v += bias[n] ; v = relu(v) ; v *= (relu_mask[m,n] > 0) ; v += add_f32[m,n]
 *   out_bf16[m,n] = bf16(v) ; out_f32[m,n] = v (f32_atomic=0) or atomically += v (f32_atomic=1)
 * split_k > 1 partitions the k range over CTAs and requires f32_atomic = 1 and no bf16 output.
 * impl: 0 = tcgen05 kernel (product path), 1 = naive SIMT kernel (test cross-check only).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a;
  int64_t lda;
  int a_mn_major;
  const void* b;
  int64_t ldb;
  int b_mn_major;
  int m, n, k;
  int split_k;
  float alpha;
  const float* bias;
  int relu;
  const void* relu_mask;
  int64_t ld_mask;
  const float* add_f32;
  int64_t ld_add;
  void* out_bf16;
  int64_t ld_out_bf16;
  float* out_f32;
  int64_t ld_out_f32;
  int f32_atomic;
  int impl;
} commu_gemm_args;

int commu_gemm_bf16(const commu_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMMU_B200_H_ */
