/*
 * cdcmdr.h - C-ABI of the B200-native CDC-MDR hot path (libcdcmdr.so, sm_100a).
 *
 * The reference (Chrissie-Law/Causal-Domain-Clustering-for-Multi-Domain-Recommendation) has no FFI:
 * its hot path is Python calling torch aten kernels.  These entry points are what a
 * binding for that path would bind: each one replaces the aten call sites named next to it
 * (paths relative to the reference root).  Conventions:
 *   - plain pointers to DEVICE memory and sizes; no torch types; nothing here allocates or synchronises;
 *   - every launch goes to the cudaStream_t passed last (as void*; 0 = legacy default stream);
 *   - return 0 on success, non-zero on error, message via cdcmdr_last_error() (thread-local);
 *   - row-major everywhere; "ld*" = leading dimension in ELEMENTS;
 *   - activations are fp32 (exact-parity path) or bf16 (tensor-core path); `bf16` pointers are uint16_t*.
 */
#ifndef CDCMDR_H
#define CDCMDR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cdcmdr_stream_t;

const char* cdcmdr_last_error(void);
int cdcmdr_version(void);
/* number of kernels launched by this library since load / since the last reset (bench.py gpu_launches) */
int64_t cdcmdr_launch_count(void);
void cdcmdr_launch_count_reset(void);

/* ---------------------------------------------------------------------------------------------
 * a1  FeaturesEmbedding.forward                                        model/layer.py:147-157
 * out[b, f*E + e] = table[x[b,f] + offsets[f], e].  Either output may be NULL.  out_bf16 has leading
 * dimension ld_bf16 >= F*E (padding columns are left untouched).  *oob_flag (device int, may be NULL)
 * is set to 1 if any index falls outside [0, V) (such rows read as zeros).
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_embed_gather_fwd(const int32_t* x, const int64_t* offsets, const float* table,
                            float* out_f32, uint16_t* out_bf16, int64_t ld_bf16,
                            int64_t B, int F, int E, int64_t V, int* oob_flag, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a2  embedding backward = autograd of nn.Embedding (sparse=False)     model/layer.py:140, run.py:491
 * plan: stable radix sort of (row = x+offsets, position = b*F+f) -> unique rows, segment bounds,
 * fixed-size partial segments (deterministic summation order).  Workspace layout is opaque.
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_embed_plan_bytes(int64_t n_idx, int64_t V, int E_max);
int cdcmdr_embed_plan_build(const int32_t* x, const int64_t* offsets, int64_t B, int F, int64_t V, int E_max,
                            void* plan, size_t plan_bytes, cdcmdr_stream_t s);
/* dense [V,E] gradient (reference semantics, what torch.optim.Adam consumes).  grad_out: [B, ldg]. */
int cdcmdr_embed_bwd_dense(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F, int E,
                           int64_t V, float* grad_table, cdcmdr_stream_t s);

typedef struct {
  float lr, beta1, beta2, eps, weight_decay; /* torch.optim.Adam(lr, betas, eps, weight_decay)  run.py:720 */
  float l2;                                  /* l2_reg_embedding: grad += 2*l2*w              layer.py:31,96-112 */
  int step;                                  /* 1-based step count t                                          */
} cdcmdr_adam_t;

/* a2+a7+a17 fused, reference-exact ("dense_exact"): ONE sweep over all V rows:
 *   g = segment_sum(row) + 2*l2*w + wd*w ; Adam(m, v, w) ; optionally accumulates sum(w_old^2) partials
 *   (the table term of get_regularization_loss) into reg_partials[gridDim] (NULL to skip).
 * "sparse_lazy": touched rows only (documented deviation from the reference). */
int cdcmdr_embed_bwd_adam_dense_exact(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                      int E, int64_t V, float* table, float* m, float* v,
                                      const cdcmdr_adam_t* h, double* reg_sumsq /* device scalar or NULL */,
                                      cdcmdr_stream_t s);
int cdcmdr_embed_bwd_adam_sparse_lazy(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                      int E, int64_t V, float* table, float* m, float* v,
                                      const cdcmdr_adam_t* h, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a3/a4/a5/a9/a11  nn.Linear / F.linear / torch.matmul call sites      layer.py:119,185,193,275,336
 * Strided, grouped fp32 GEMM on CUDA cores (exact-parity path):
 *   C[g](m,n) = epi( sum_k A[g](m,k) * Bt[g](n,k) ),  A(m,k) = A[g*a_gs + m*a_rs + k*a_cs], etc.
 *   epi(v) = v + bias[g*bias_gs + n]; relu if act==1; then, if mask != NULL,
 *            v *= (mask[g*mask_gs + m*mask_rs + n] > 0) * mask_scale; then dropout(drop_p, seed, salt);
 *            if accumulate: v += C.
 * split_k > 1 reduces over K in split_k deterministic slices through `workspace` (>= split_k*G*M*N floats).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A; const float* Bt; float* C;
  int64_t M, N, K;
  int64_t a_rs, a_cs, b_rs, b_cs, c_rs;
  int32_t G; int64_t a_gs, b_gs, c_gs;
  const float* bias; int64_t bias_gs;
  int32_t act;
  const float* mask; int64_t mask_rs, mask_gs; float mask_scale;
  float drop_p; uint64_t seed; uint32_t salt;
  int32_t accumulate;
  int32_t split_k; float* workspace;
} cdcmdr_gemm_f32_t;
int cdcmdr_gemm_f32(const cdcmdr_gemm_f32_t* p, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 / TMEM / TMA bf16 GEMM (tensor-core path), fp32 accumulate:
 *   acc[g](m,n) = sum_k A[g](m,k) * Bt[g](n,k)
 * A is [M_total, lda] bf16 row-major; group g reads columns [g*a_gk, g*a_gk + K) and rows [0,M).
 * Bt: group g uses rows [g*b_gn, g*b_gn + N) of a [*, ldb] bf16 row-major matrix (K contiguous).
 * a_mn_major / b_mn_major != 0: the operand is stored TRANSPOSED (reduction index is the row index):
 *   A(m,k) = A[k*lda + g*a_gk + m],  Bt(n,k) = Bt[k*ldb + g*b_gn + n]  -> weight-gradient GEMMs (K = batch).
 * Columns n < n_main: main = epi(acc + bias) as bf16 at out_main[m*ld_main + g*main_gn + n]
 *     epi: relu (act==1), * (mask>0)*mask_scale (mask bf16 laid out like out_main), dropout.
 * Columns n >= n_main: aux = acc + bias as fp32 at out_aux[m*ld_aux + g*aux_gn + (n - n_main)].
 * split_k > 1 (requires n_main == 0): slice z of K is written to out_aux + z*aux_split_stride (partials).
 * K must be a multiple of 64 (pad with zeros), N a multiple of 16, N <= 256 per tile column block.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const uint16_t* A; int64_t lda; int64_t a_rows;   /* a_rows: number of valid rows of the stored A matrix */
  const uint16_t* Bt; int64_t ldb; int64_t b_rows;
  int64_t M, N, K;
  int32_t G; int64_t a_gk, b_gn;
  int32_t a_mn_major, b_mn_major;
  const float* bias; int64_t bias_gs;
  int64_t n_main;
  uint16_t* out_main; int64_t ld_main, main_gn;
  float* out_aux; int64_t ld_aux, aux_gn;
  int32_t act;
  const uint16_t* mask; float mask_scale;
  float drop_p; uint64_t seed; uint32_t salt;
  int32_t split_k; int64_t aux_split_stride;
  int32_t block_n;                                   /* 0 = auto */
} cdcmdr_gemm_bf16_t;
int cdcmdr_gemm_bf16_tc(const cdcmdr_gemm_bf16_t* p, cdcmdr_stream_t s);
/* sum split-K partials: out[i] = sum_z part[z*stride + i] (+ out[i] if accumulate), deterministic order */
int cdcmdr_splitk_reduce(const float* part, int64_t stride, int32_t splits, float* out, int64_t n,
                         int32_t accumulate, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a5/a9  gate softmax + expert-weighted sum (CGC.forward, MMoE.forward)  ple.py:106-123, mmoe.py:56-60
 * H: expert outputs [B, ldh], expert e at columns [e*h, (e+1)*h).  logits: fp32 [B, ldl]; gate j reads
 * logits[:, gate_col[j] : gate_col[j]+gate_n[j]] and mixes experts gate_sel[j*max_sel + s].
 * out gate j -> out[:, j*h : (j+1)*h].  probs (fp32 [B, n_gates*max_sel]) saved for backward.
 * is_bf16: H/out (and dOut/dH) are bf16, else fp32.  Descriptor arrays live in DEVICE memory.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_gates, n_experts, h, max_sel;
  const int32_t* gate_col; const int32_t* gate_n; const int32_t* gate_sel;   /* device */
} cdcmdr_mix_desc_t;
int cdcmdr_gate_mix_fwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* logits, int64_t ldl,
                        void* out, int64_t ldo, float* probs, int64_t B, int is_bf16, cdcmdr_stream_t s);
/* backward: dH[:, e] = (sum_{j contains e} p_j[e]*dOut_j) * (H_e > 0) * relu_scale   (relu_scale<0: no mask)
 *           dlogits gate j -> dlogits[:, gate_col[j] + s] (fp32 if dlogits_f32 else bf16 at dlogits_bf16) */
int cdcmdr_gate_mix_bwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* probs,
                        const void* dOut, int64_t ldo, void* dH, int64_t lddh, float relu_scale,
                        float* dlogits_f32, uint16_t* dlogits_bf16, int64_t lddl,
                        int64_t B, int is_bf16, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a4/a8/a14  BatchNorm1d (train: batch stats; eval: running stats) + ReLU + dropout
 *            layer.py:187,202-204,279 ; star.py:169-181 ; torch BatchNorm1d defaults (eps 1e-5, momentum .1)
 * Z fp32 [B, ldz] with C columns.  Train: computes mean / biased var per column (deterministic two-stage
 * reduction through `scratch` >= 2*C*grid doubles), updates running stats with the unbiased var, writes
 * save_mean/save_invstd (fp32 [C]) and A = dropout(relu(gamma*xhat+beta)).  gamma2/beta2 (may be NULL):
 * STAR partitioned norm, gamma*gamma2 and beta+beta2.  relu=0 skips the activation.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* gamma; const float* beta; const float* gamma2; const float* beta2;
  float* running_mean; float* running_var;
  float* save_mean; float* save_invstd;
  int32_t train, relu;
  float drop_p; uint64_t seed; uint32_t salt;
} cdcmdr_bn_t;
size_t cdcmdr_bn_scratch_bytes(int64_t C);
int cdcmdr_bn_fwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, void* A, int64_t lda_, int a_is_bf16,
                  int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s);
/* backward through dropout/relu/BN: dZ from dA.  A is the forward output (its sign gives the relu/dropout mask).
 * dgamma/dbeta fp32 [C] (+= if accumulate).  With gamma2: dgamma is the gradient of the PRODUCT gamma*gamma2. */
int cdcmdr_bn_bwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                  const float* dA, int64_t ldda, float* dZ, int64_t lddz, float* dgamma, float* dbeta,
                  int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a8/a15/a17  tower logit + FeaturesLinear + Sigmoid + tower selection + BCELoss(mean) and its backward
 *             layer.py:48-56 ; cdc.py:99-111 ; run.py:483-484,723
 * logits fp32 [B, T] (tower outputs), lin fp32 [B] (may be NULL).  y = sigmoid(logits + lin) -> pred [B,T].
 * mode 0: per-sample column sel[b] (int64 [B]; pred.gather / CDC split without domain_i)
 * mode 1: fixed column `col` (CDC split with domain_i)
 * mode 2: mean over T (CDC warmup)
 * mode 3: every column is its own sample (single-output models: T must be 1)
 * target int16/float per `target_is_f32`.  Writes psel fp32 [B], loss_sum (device double, sum over batch of
 * the clamped BCE terms; divide by B on the host or in reg_loss_finalize), dlogits fp32 [B,T] = dLoss/dlogit
 * (already divided by B * grad_scale) if dlogits != NULL.
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_sigmoid_select_bce(const float* logits, const float* lin, int64_t B, int32_t T, int32_t mode,
                              const int64_t* sel, int32_t col, const void* target, int target_is_f32,
                              float* pred, float* psel, double* loss_sum, float* dlogits, float inv_batch,
                              cdcmdr_stream_t s);

/* a7  get_regularization_loss: sum_i coef[i] * w[i]^2 over a flat parameter arena  layer.py:96-112 */
int cdcmdr_reg_l2_sum(const float* w, const float* coef, int64_t n, double* out_sum, void* scratch,
                      cdcmdr_stream_t s);
size_t cdcmdr_reduce_scratch_bytes(void);

/* a17  torch.optim.Adam over a flat arena, fused with the L2-regulariser gradient  run.py:720-721
 *   g = grad[i] + 2*l2coef[i]*w[i] + wd*w[i] ; m,v,w update.  present[i]==0 (uint8, may be NULL)
 *   marks parameters whose grad is None in the reference (skipped). */
int cdcmdr_adam_dense(float* w, const float* grad, float* m, float* v, const float* l2coef,
                      const uint8_t* present, int64_t n, const cdcmdr_adam_t* h, cdcmdr_stream_t s);

/* pack fp32 parameter rows into (zero-padded) bf16 GEMM operands: dst[r, 0:K] = src[row_off[r] : +K]
 * (row_off < 0 -> zero row); if dst_t != NULL also the transpose dst_t[k, r] with leading dim ldt. */
int cdcmdr_pack_rows_bf16(const float* src, const int64_t* row_off, int64_t rows, int64_t K,
                          uint16_t* dst, int64_t ldd, uint16_t* dst_t, int64_t ldt, cdcmdr_stream_t s);
/* scatter rows of a dense fp32 matrix back into the arena: dst[row_off[r] : +K] (+)= src[r*lds : +K] */
int cdcmdr_unpack_rows_f32(const float* src, int64_t lds, const int64_t* row_off, int64_t rows, int64_t K,
                           float* dst, int accumulate, cdcmdr_stream_t s);
/* column sums of a [B, C] matrix (bias gradients): out[c] (+)= sum_b X[b*ld + c]; deterministic */
int cdcmdr_colsum(const void* X, int64_t ld, int is_bf16, int64_t B, int64_t C, float* out, int accumulate,
                  void* scratch, cdcmdr_stream_t s);
size_t cdcmdr_colsum_scratch_bytes(int64_t C);
/* elementwise helpers on device buffers */
int cdcmdr_cast_f32_bf16(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols,
                         cdcmdr_stream_t s);
int cdcmdr_cast_bf16_f32(const uint16_t* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols,
                         int accumulate, cdcmdr_stream_t s);
/* out = a*b (op 0), a+b (op 1), out += a*b (op 2) over n contiguous floats: STAR W_d*W_s, b_d+b_s  star.py:91-92 */
int cdcmdr_ewise_f32(const float* a, const float* b, float* out, int64_t n, int op, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a10/a11/a12  cross-network elementwise stages                         layer.py:495-515, 332-343, 380-407
 * cross_fuse_fwd:  out = x0 * xw + b + x     xw is [B,1] (bcast, v1: ldxw==0 semantics via xw_cols==1) or [B,D]
 * cross_fuse_bwd:  given dout: dx0 += dout*xw ; dxw = dout*x0 (v2: [B,D]; v1: row-summed [B,1]) ; dx = dout
 *                  db (+)= colsum(dout) is done with cdcmdr_colsum.
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_cross_fuse_fwd(const float* x0, const float* x, const float* xw, int xw_cols, const float* b,
                          float* out, int64_t B, int64_t D, cdcmdr_stream_t s);
int cdcmdr_cross_fuse_bwd(const float* x0, const float* xw, int xw_cols, const float* dout,
                          float* dx0_acc, float* dxw, int64_t B, int64_t D, cdcmdr_stream_t s);
/* CrossNetMix per-layer expert combine:  x_next = x + sum_e g[b,e] * x0 * (u[e][b,:] + bias)
 *   u: [n_exp, B, D]; gate probabilities g: [B, n_exp]                    layer.py:395-404 */
int cdcmdr_crossmix_combine_fwd(const float* x0, const float* x, const float* u, const float* g,
                                const float* bias, float* out, int64_t B, int64_t D, int n_exp,
                                cdcmdr_stream_t s);
/* backward: du[e] = dout*g[:,e]*x0 ; dgate[b,e] = sum_d dout*x0*(u_e+bias) ; dx0 += sum_e dout*g_e*(u_e+bias) */
int cdcmdr_crossmix_combine_bwd(const float* x0, const float* u, const float* g, const float* bias,
                                const float* dout, float* du, float* dgate, float* dx0_acc,
                                int64_t B, int64_t D, int n_exp, cdcmdr_stream_t s);
/* y = tanh(x) in place / dy *= (1 - y^2) */
int cdcmdr_tanh_fwd(float* x, int64_t n, cdcmdr_stream_t s);
int cdcmdr_tanh_bwd(const float* y, float* dy, int64_t n, cdcmdr_stream_t s);
/* row softmax over small n (<=64) columns, fwd and bwd (dz = p*(dp - sum(dp*p))) */
int cdcmdr_softmax_rows_fwd(const float* z, int64_t ldz, float* p, int64_t ldp, int64_t B, int n, cdcmdr_stream_t s);
int cdcmdr_softmax_rows_bwd(const float* p, int64_t ldp, const float* dp, int64_t lddp, float* dz, int64_t lddz,
                            int64_t B, int n, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a14  STAR per-domain routing: stable partition of rows by group id    star.py:84-86,107,112-114
 * perm[i] = source row of output row i (groups ascending, original order inside a group);
 * counts[g] rows per group, group_start[g] exclusive prefix.  group is int64 [B] (values outside
 * [0,n_group) are dropped and *n_routed < B).  scratch >= cdcmdr_route_scratch_bytes(B, n_group).
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_route_scratch_bytes(int64_t B, int n_group);
int cdcmdr_route_partition(const int64_t* group, int64_t B, int n_group, int32_t* perm, int32_t* counts,
                           int32_t* group_start, void* scratch, cdcmdr_stream_t s);
/* dst[i, :] = src[perm[i], :] (gather) ; or dst[perm[i], :] (+)= src[i, :] (scatter, for the backward) */
int cdcmdr_permute_rows(const void* src, int64_t lds, const int32_t* perm, int64_t n, int64_t cols, int elt_bytes,
                        void* dst, int64_t ldd, int scatter, cdcmdr_stream_t s);
/* a15  groups[b] = domain2group[x[b, domain_idx]]                       cdc.py:105 */
int cdcmdr_domain_to_group(const int32_t* x, int64_t B, int F, int domain_idx, const int64_t* domain2group,
                           int n_domain, int64_t* groups, cdcmdr_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CDCMDR_H */
