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
 *   v = alpha*acc ; v += bias[n] ; v = relu(v) ; v *= (relu_mask[m,n] > 0) ; v = dropout(v) ; v += add_f32[m,n]
 *   out_bf16[m,n] = bf16(v) ; out_f32[m,n] = v (f32_atomic=0) or atomically += v (f32_atomic=1)
 * split_k > 1 partitions the k range over CTAs and requires f32_atomic = 1 and no bf16 output.
 * When every output tile is full (m a multiple of 128 - 256 for the CTA-pair kernel -, n of the tile width) and the
 * buffers allow 16-byte accesses, a whole-tile epilogue instantiation is launched for the bf16-only and the fp32-only
 * results (same arithmetic, same order; COMMU_GEMM_FAST_EPI=0 keeps the general per-chunk epilogue).
 * impl: 0 = tcgen05 kernels (product path: the CTA-pair kernel, tcgen05 cta_group::2 on 256 x 256 tiles with the B tile
 *       split across the two SMs of a TPC, for n > 128 and m > 128; else the one-CTA kernel), 1 = naive SIMT kernel
 *       (test cross-check only), 2 / 3 = force the CTA-pair / the one-CTA tcgen05 kernel.
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
  /* inverted dropout of the GEMM result (nn.Dropout after o_net / CoreNet.1 / CoreNet.3, model.py:349, 166-168):
   * with drop_p > 0, after bias / relu and before add_f32:  v = keep(drop_seed, m, n) ? v / keep_prob : 0, the same
   * counter-based mask as commu_dropout (element (row m, column n) of the [m, n] result). */
  float drop_p;
  uint64_t drop_seed;
} commu_gemm_args;

int commu_gemm_bf16(const commu_gemm_args* args, void* stream);

/* Per-kernel-class device timers (bench.py roofline line) and the launch counter.
 * classes: 0 = GEMM, 1 = attention fwd, 2 = attention bwd, 3 = decode attention. */
int commu_prof_arm(unsigned class_mask);
int commu_prof_read(int cls, float* total_ms, int* launches);
long long commu_launch_count(int reset);

/* ------------------------------------------------------------------------------------------
 * Embedding  x = E[tok] * scale   (AdaptiveEmbedding.forward, commu/model/model.py:409-420) and its
 * backward (scatter-add into the 729-row table, tied with the logits weight, model.py:480-481).
 * Outputs are [n, ld] with columns d..dp-1 zeroed.
 * ------------------------------------------------------------------------------------------ */
int commu_embed_fwd(const int64_t* tokens, const float* table, int d, int dp, float scale, int64_t n,
                    float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, void* stream);
int commu_embed_bwd(const int64_t* tokens, const float* dx, int64_t ld, int d, float scale, int64_t n,
                    float* dtable, void* stream);

/* Sinusoid table by DISTANCE: row delta = [sin(min(delta,clamp) f) | cos(...)]
 * (PositionalEmbedding.forward model.py:145-152 with pos_seq of :578-583; row delta of this table is
 * row klen-1-delta of the reference's pos_emb). */
int commu_pos_table(const float* inv_freq, int klen, int clamp_len, int d, int dp, void* out_bf16,
                    float* out_f32, void* stream);

/* LayerNorm over the d real columns (eps as given; reference uses nn.LayerNorm default 1e-5,
 * model.py:214, :171, applied at :352 and :179).  z already contains residual + sub-block output. */
int commu_layernorm_fwd(const float* z, int64_t ldz, const float* gamma, const float* beta, int d, int dp,
                        float eps, int64_t rows, float* y_f32, int64_t ldy, void* y_bf16, int64_t ldyb,
                        float* mean, float* rstd, void* stream);
int commu_layernorm_bwd(const float* dy, int64_t lddy, const float* z, int64_t ldz, const float* mean,
                        const float* rstd, const float* gamma, int d, int dp, int64_t rows, float* dz_f32,
                        int64_t lddz, void* dz_bf16, int64_t lddzb, float* dgamma, float* dbeta,
                        float drop_p, uint64_t drop_seed, void* stream);
/* (drop_p > 0: the bf16 copy dz_bf16 additionally carries the inverted dropout of the sub-block output, i.e. the
 * gradient that enters o_net / CoreNet.3 (model.py:349, 168); the fp32 copy is the undropped residual-branch gradient.) */

/* Per-token NLL = -log_softmax(logits)[target] (ProjectedAdaptiveLogSoftmax.forward, n_clusters == 0
 * branch, model.py:64-73) and its backward dlogits = (softmax - onehot) * dloss (bf16, Vp columns). */
int commu_nll_fwd(const float* logits, int64_t ld, int V, const int64_t* target, int64_t rows, float* nll,
                  float* lse, void* stream);
int commu_nll_bwd(const float* logits, int64_t ld, int V, int Vp, const float* lse, const int64_t* target,
                  const float* dloss, int64_t rows, void* dlogits_bf16, int64_t ldd, void* stream);

/* out[c] += sum_r x[r,c]  (bias gradients of CoreNet.0 / CoreNet.3 / logits bias). */
int commu_colsum_bf16(const void* x, int64_t ld, int ncols, int64_t rows, float* out, void* stream);

/* fp32 master weight [R,C] -> bf16 operand shadow with segment padding
 * (dst row = (r / rseg) * rseg_pad + r % rseg, same for columns), optional transpose; and the inverse
 * accumulation of a padded fp32 gradient into the reference-layout gradient. */
int commu_cast_pad(const float* src, int64_t ld_src, int R, int C, int rseg, int rseg_pad, int cseg,
                   int cseg_pad, void* dst_bf16, int64_t ld_dst, int transpose, void* stream);
int commu_unpad_accum(const float* src, int64_t ld_src, int R, int C, int rseg, int rseg_pad, int cseg,
                      int cseg_pad, float* dst, int64_t ld_dst, float scale, void* stream);

/* Optimizer tail (train.py:159-169): out_accum += sum(g^2); then
 * g' = g * grad_scale * min(1, clip / (||g * grad_scale|| + 1e-6)) (clip_grad_norm_) followed by Adam
 * (torch.optim.Adam, no amsgrad; weight_decay is its L2 term g' += weight_decay * p, train.py:442-443) on flat
 * fp32 arenas.  step >= 1.  p_bf16 (optional): bf16 copy of the updated parameters in the same flat layout - the GEMM
 * operand shadows, so that no separate cast pass runs after the optimizer step. */
int commu_sumsq(const float* g, int64_t n, float* out_accum, void* stream);
int commu_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, int step, const float* gnorm_sq, float clip, float grad_scale,
                    float weight_decay, float* gnorm_out, void* p_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused relative-position attention (RelPartialLearnableMultiHeadAttn.forward, model.py:312-345,
 * with _rel_shift :251-265 and the mask of :549-574 folded in analytically).
 *   q   : bf16 [T*B, ldq]   row = i*B + b, head h at columns h*64 .. h*64+63 (head dim padded to 64)
 *   k, v: bf16 [K*B, ldkv]  K = T + M
 *   r   : bf16 [kr, ldr]    projected sinusoid table indexed by distance, kr >= K
 *   r_w_bias, r_r_bias: fp32 [H, 64]
 *   reset: uint8 [B] or NULL (1 = this column restarted: memory keys j < M are masked)
 *   valid(i,j) = j <= i + M  &&  (!same_length || j > i - shift)  &&  (!reset[b] || j >= M)
 *   out : bf16 [T*B, ldo];  lse: fp32 [B,H,T] = log sum_j exp(scale * score)
 *   qu_save / qv_save (both or neither): bf16 [T*B, ldq] receive bf16(q + r_w_bias), bf16(q + r_r_bias)
 *   for the backward.
 * ------------------------------------------------------------------------------------------ */
int commu_relattn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                      const void* r, int64_t ldr, int kr, const float* r_w_bias, const float* r_r_bias,
                      const unsigned char* reset, int T, int M, int B, int H, int same_length, int shift,
                      float scale, void* out, int64_t ldo, float* lse, void* qu_save, void* qv_save,
                      void* stream);
/* Same contract, tcgen05 implementation (TMA-staged K/V/R tiles, S / BD / PV accumulators in TMEM,
 * P fed to the PV MMA from TMEM).  Requires ldo % 8 == 0.
 * p_save / mt_save (both or neither, sizes from commu_relattn_bwd_sizes): the forward additionally keeps, for the
 * materialised backward, P~ = bf16(exp2(score - m_tile)) [B*H, Tpad, Kp] (sign bit = dropped by the attention
 * dropout) and the per-(key tile, query row) maximum m_tile [B*H, Kp/128, Tpad] it was taken at (log2 domain);
 * P = P~ * exp2(m_tile - LSE * log2 e). */
int commu_relattn_fwd_tc(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                         const void* r, int64_t ldr, int kr, const float* r_w_bias, const float* r_r_bias,
                         const unsigned char* reset, int T, int M, int B, int H, int same_length, int shift,
                         float scale, void* out, int64_t ldo, float* lse, void* qu_save, void* qv_save,
                         void* p_save, float* mt_save, void* stream);
/* Dropout on the attention probabilities (reference: self.dropatt, commu/model/model.py:211, 337) for the
 * subsequent tcgen05 attention calls of this process (forward and the three backward passes): p in [0, 1),
 * 0 = off.  Masks are a pure function of (seed, batch, head, query, key) and are recomputed in the backward, so
 * the forward and the backward of a layer must run under the same (p, seed).  Not supported by the v1 kernels. */
int commu_relattn_set_dropout(float p, unsigned long long seed);
/* Elementwise dropout with optional residual add - the remaining nn.Dropout sites of the reference
 * (commu/model/model.py:166-168, 349, 585-586, 600): out = res + keep(x) / (1 - p) over [rows, cols].
 * x is fp32 or bf16 (x_is_bf16), res (fp32) may be NULL, out_f32 / out_bf16: at least one, in place allowed.
 * cols and all leading dimensions are multiples of 4.  The keep decision of (seed, row, column) is reproducible:
 * the backward applies the same call to the incoming gradient. */
int commu_dropout(const void* x, int x_is_bf16, int64_t ldx, const float* res, int64_t ldres, int64_t rows,
                  int cols, float p, uint64_t seed, float* out_f32, int64_t ldo, void* out_bf16, int64_t ldob,
                  void* stream);
/* Backward of the above (the reference uses torch autograd).  dq: bf16 [T*B, lddq]; dk, dv: bf16
 * [K*B, lddkv] (every key row written); dr: fp32 [kr, H*64] and du, dvb: fp32 [H,64] are accumulated
 * (+=, caller zeroes); delta_ws: fp32 [B,H,T] workspace.
 * Product path (p_save, mt_save and ws given - what commu_relattn_fwd_tc stored): the score gradient is formed ONCE
 * from the stored probabilities and written to the workspace in a coarse-sheared layout in which the relative shift
 * is a TMA stride; dq, dR and the two bias gradients are then pure TMA + tcgen05.mma streams (csrc/attn_bwd_mat.cu).
 * The first ws_zero_bytes of the workspace (commu_relattn_bwd_sizes) must be zero the first time it is used with a
 * shape (T, M, B, H); the kernels keep it valid afterwards, so one buffer serves every layer and step of that shape
 * (calls that share a workspace must be ordered on one stream: it also carries the ticket counter and the hand-over
 * flags of the merged dq launch, which every call resets; 128-byte alignment).
 * With p_save == NULL the three recompute passes run instead (no workspace). */
int commu_relattn_bwd_sizes(int T, int M, int B, int H, int64_t* p_bytes, int64_t* mt_bytes, int64_t* ws_bytes,
                            int64_t* ws_zero_bytes);
int commu_relattn_bwd(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                      int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset, int T,
                      int M, int B, int H, int same_length, int shift, float scale, const void* out,
                      int64_t ldo, const float* lse, const void* dout, int64_t lddo, float* delta_ws,
                      void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, float* dr, float* du,
                      float* dvb, const void* p_save, const float* mt_save, void* ws, int64_t ws_bytes,
                      void* stream);
/* The materialised backward on its own (delta = rowsum(dO * O) [B,H,T] already computed). */
int commu_relattn_bwd_mat(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                          int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset, int T,
                          int M, int B, int H, int same_length, int shift, float scale, const float* lse,
                          const void* dout, int64_t lddo, const float* delta, const void* p_save,
                          const float* mt_save, void* ws, int64_t ws_bytes, void* dq, int64_t lddq, void* dk,
                          void* dv, int64_t lddkv, float* dr, float* du, float* dvb, void* stream);

/* Run-time choice of the implementation of each backward pass used by commu_relattn_bwd:
 * 1 = tcgen05 kernel (default), 0 = v1 warp-MMA kernel, negative = leave unchanged. */
int commu_relattn_bwd_set_impl(int dq_tc, int dkv_tc, int dr_tc);
/* dR pass of commu_relattn_bwd on tcgen05 tensor cores (diagonal walk: one CTA per 128 distances, dR
 * accumulated in TMEM, added to dr once per CTA).  It also computes d r_r_bias from the column sums of the
 * position-term gradient: the share is added to dvb and subtracted from du, where commu_relattn_bwd_dq_tc
 * left colsum(dq) = d r_w_bias + d r_r_bias - the two tcgen05 passes run as a pair (COMMU_ATTN_BWD_DR=v1 or
 * COMMU_ATTN_BWD_DQ=v1 selects the warp-MMA passes for both). */
int commu_relattn_bwd_dr_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                            int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset,
                            int T, int M, int B, int H, int same_length, int shift, float scale,
                            const float* lse, const void* dout, int64_t lddo, const float* delta, float* dr,
                            float* du, float* dvb, void* stream);
/* dq pass of commu_relattn_bwd on tcgen05 tensor cores (dS fed to the dq MMA from TMEM, the inverse relative
 * shift scattered into 128 x 128 distance blocks in shared memory).  du += colsum(dq) = d r_w_bias + d r_r_bias;
 * dvb is left to commu_relattn_bwd_dr_tc (see there). */
int commu_relattn_bwd_dq_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                            int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset,
                            int T, int M, int B, int H, int same_length, int shift, float scale,
                            const float* lse, const void* dout, int64_t lddo, const float* delta, void* dq,
                            int64_t lddq, float* du, float* dvb, void* stream);
/* dk / dv pass of commu_relattn_bwd on tcgen05 tensor cores (TMA-staged tiles, TMEM accumulators);
 * `delta` = rowsum(dO * O) [B,H,T] must already be computed.  commu_relattn_bwd dispatches here by
 * default (COMMU_ATTN_BWD_DKV=v1 selects the warp-MMA pass instead). */
int commu_relattn_bwd_dkv_tc(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v,
                             int64_t ldkv, const void* r, int64_t ldr, int kr, const unsigned char* reset,
                             int T, int M, int B, int H, int same_length, int shift, float scale,
                             const float* lse, const void* dout, int64_t lddo, const float* delta, void* dk,
                             void* dv, int64_t lddkv, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange: one NCCL sum all-reduce of the flat fp32 gradient arena per
 * optimizer step (replaces the per-micro-batch DDP bucket all-reduce, train.py:155, 467-473).
 * NCCL is dlopen'ed (nccl_path may be NULL/"" to use the default soname).
 * ------------------------------------------------------------------------------------------ */
int commu_comm_unique_id(const char* nccl_path, void* id_out_128_bytes);
int commu_comm_init(const char* nccl_path, const void* id_128_bytes, int rank, int world);
int commu_allreduce_sum_f32(float* buf, int64_t n, void* stream);
int commu_comm_destroy(void);

/* ------------------------------------------------------------------------------------------
 * Incremental decode (generate.py hot loop: InferenceTask.calc_logits_and_mems / calc_probs /
 * apply_sampling / infer_token, commu/midi_generator/midi_inferrer.py:199-237, over
 * MemTransformerLM.forward_generate, commu/model/model.py:606-628).
 * ------------------------------------------------------------------------------------------ */
/* out[b,n] = act(sum_k x[b,k] W[n,k] + bias[n]) + res[b,n], B <= 64 rows; W fp32 or bf16 [N,K]. */
int commu_decode_linear(const float* x, int64_t ldx, const void* w, int64_t ldw, int w_bf16, const float* bias,
                        int relu, const float* res, int64_t ldr, float* out, int64_t ldo, int B, int N, int K,
                        void* stream);
/* Same contract on a register-tiled fp32 SIMT GEMM (one CTA = all rows x 32 columns x one K split; partial sums of the
 * splits are added in a fixed order by the last CTA of a column tile, so results are run-to-run identical).  splits
 * = 0 lets the library pick; scratch: fp32 [splits * ceil(N/32) * 2048], counters: int32 [ceil(N/32)] zeroed once. */
int commu_decode_linear_tiled(const float* x, int64_t ldx, const void* w, int64_t ldw, int w_bf16, const float* bias,
                              int relu, const float* res, int64_t ldr, float* out, int64_t ldo, int B, int N, int K,
                              int splits, float* scratch, int* counters, void* stream);
/* qkv_net of the fp32 engine on the tiled kernel ([q|k|v] = x W^T, W fp32 [3*H*Dh, K]; model.py:283-310 at T = 1): the
 * epilogue scatters q to the staged query [B,H,64] and k / v into ring slot `slot` (or dev_state[0]) of the fp32 caches
 * [B,H,C,64]; padding columns Dh..63 are never written (the buffers are zero-initialised once). */
int commu_decode_qkv_tiled(const float* x, int64_t ldx, const float* w, int64_t ldw, int B, int H, int Dh, int K,
                           float* q_out, float* k_cache, float* v_cache, int C, int slot, const int* dev_state,
                           float* scratch, int* counters, void* stream);
/* dst[row*row_stride + h*head_stride + offset + e] = src[row, col_off + h*Dh + e], e < 64 (zero padded):
 * stages q, appends K / V to the ring cache slot, builds the R-by-distance table. */
int commu_pad_heads(const float* src, int64_t ld_src, int col_off, int rows, int H, int Dh, void* dst,
                    int dst_bf16, int64_t row_stride, int64_t head_stride, int64_t offset, const int* dev_state,
                    void* stream);
/* Device-resident step bookkeeping (int32[4] {slot, n_vis, cached, step}, init {-1,0,0,0}) so that a whole
 * decode step can be captured once in a CUDA graph and replayed: when `dev_state` is non-NULL the ring
 * slot / visible-key count / sampler offset of commu_pad_heads, commu_decode_attn and commu_sample come
 * from it instead of the host arguments. */
int commu_decode_advance(int* state, int C, int mem_len, int extra_visible, void* stream);
/* Single-query relative attention over the projected K/V ring cache [B,H,C,64]: ages 0..n_vis-1
 * (age 0 at ring slot cur_slot) with score_a = scale*((q+r_w_bias).k_a + (q+r_r_bias).R[a]).
 * out_bf16 (optional, same leading dim) receives a bf16 copy: the operand of the o_net GEMM. */
int commu_decode_attn(const float* q, const void* kcache, const void* vcache, const void* rtab, int cache_bf16,
                      const float* r_w_bias, const float* r_r_bias, int B, int H, int C, int n_vis, int cur_slot,
                      float scale, float* out, int64_t ldo, const int* dev_state, void* out_bf16, int splits,
                      float* partial, int* counters, void* stream);
/* (splits > 1: the visible keys of every (sequence, head) are cut into `splits` ranges handled by different CTAs so
 * that the grid fills the SMs evenly; partial: fp32 [B*H*splits*66], counters: int32 [B*H] zero-initialised once; the
 * last CTA of a (sequence, head) merges the ranges in order - run-to-run identical.) */
/* ---- fused token-step kernels of the bf16 decode engine (one decoder layer = 5 launches) ----
 * commu_decode_fused_linear: out[b, n0..n0+15] per CTA = epilogue( prologue(input)[b, :K] . W[n, :K] ), B <= 64
 * rows as the M side of warp-level bf16 MMAs, fp32 accumulation, weights streamed once per call.
 *   prologue 0: x = emb[tokens[b]] * emb_scale          (AdaptiveEmbedding, commu/model/model.py:409-420)
 *            1: x = LayerNorm(z[b]; gamma, beta, eps)   (post-LN of the previous block, model.py:352 / :181)
 *            2: x = a_bf16[b, :K]                       (already a bf16 operand)
 *            for 0 / 1 the fp32 x is also written to x_out (the residual stream) when non-NULL; d_true is the
 *            embedding / LayerNorm width (<= K, columns d_true..K-1 are zero).
 *   epilogue 0: N = 3*H*64 columns in the padded head layout [q|k|v][H][64]: q -> q_out fp32 [B,H,64], k / v ->
 *               ring slot `slot` (or dev_state[0]) of the bf16 caches [B,H,C,64]   (qkv_net, model.py:283-310)
 *            1: out_f32[b,n] = acc + bias[n] + res[b,n]                              (o_net / FF output + residual)
 *            2: out_bf16[b,n] = bf16(relu(acc + bias[n]))                            (CoreNet.0 + ReLU, model.py:163-165)
 *            3: out_f32[b,n] = acc + bias[n], n < N                                  (tied logits, model.py:44-51)
 * w: bf16 [ceil16(N), K] row-major (rows beyond N zero).  K is a multiple of 64; split_k in {1,2,4,8} splits K
 * over a thread-block cluster (prologue 2 only) with a fixed-order DSMEM reduction; K / split_k <= 1024.
 * pdl != 0 launches with programmatic stream serialization (the kernel prefetches its weights before
 * griddepcontrol.wait). */
typedef struct {
  int prologue, epilogue;
  int B, K, N, d_true;
  int split_k, pdl;
  const int64_t* tokens;
  const float* emb;
  float emb_scale;
  const float* z;
  int64_t ldz;
  const float* gamma;
  const float* beta;
  float eps;
  const void* a_bf16;
  int64_t lda;
  float* x_out;
  int64_t ldx;
  const void* w;
  int64_t ldw;
  const float* bias;
  const float* res;
  int64_t ldr;
  float* out_f32;
  int64_t ldo;
  void* out_bf16;
  int64_t ldob;
  float* q_out;
  void* k_cache;
  void* v_cache;
  int H, C, slot;
  const int* dev_state;
} CommuDecLinear;
int commu_decode_fused_linear(const CommuDecLinear* args, void* stream);
/* commu_decode_attn over a bf16 cache [B,H,C,64], two implementations:
 *  impl & 8 (product path): stream-K tensor-core kernel.  C must be a multiple of 64; rtab is the reversed, doubled
 *    table [H, 2C, 64] with row j = R[C-1 - (j mod C)].  A persistent grid of `splits` CTAs per SM cuts the flat list
 *    of B*H*(visible 64-slot tiles) into equal contiguous runs; a producer warp feeds K / V / R tiles by TMA (128-byte
 *    swizzle) into an mbarrier ring (4 stages, 3 with impl & 2), four consumer warps run warp-level bf16 MMAs.
 *    partial: fp32 [B*H][C/64][66] scratch, counters: int32 [B*H] zero-initialised once (the CTA that arrives last at
 *    a (sequence, head) merges its runs in run order and re-zeroes the counter: deterministic).
 *  impl == 0 (cross-check): SIMT kernel, rtab [C,H,64], grid H x B x splits, partial fp32 [B*H*splits*66]. */
int commu_decode_attn_split(const float* q, const void* kcache, const void* vcache, const void* rtab,
                            const float* r_w_bias, const float* r_r_bias, int B, int H, int C, int n_vis, int cur_slot,
                            float scale, int splits, float* partial, int* counters, void* out_bf16, float* out_f32,
                            int64_t ldo, const int* dev_state, int pdl, int impl, void* stream);
/* Sampler over B rows of raw logits (token 0 is never sampled, midi_inferrer.py:206/:220): temperature
 * (0 = greedy one-hot, :211-213), top-k (:224-226), top-p (new), wrong-token mask (:227-229),
 * renormalise (:230-231), counter-based multinomial draw (:234-237).  tokens and/or probs_out. */
int commu_sample(const float* logits, int64_t ld, int B, int V, float temperature, int top_k, float top_p,
                 const unsigned char* wrong, uint64_t seed, uint64_t offset, int64_t* tokens, float* probs_out,
                 int64_t ldp, const int* dev_state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMMU_B200_H_ */
