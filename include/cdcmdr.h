/*
 * cdcmdr.h - C-ABI of the B200-native CDC-MDR hot path (libcdcmdr.so, sm_100a).
 *
 * The reference (Chrissie-Law/Causal-Domain-Clustering-for-Multi-Domain-Recommendation) has no FFI:
 * its hot path is Python calling torch aten kernels.  These entry points are what a
 * binding for that path would bind: each one replaces the aten call sites named next to it
 * (paths relative to the reference root).  Conventions:
 *   - plain pointers to DEVICE memory and sizes; no torch types; nothing here allocates or synchronises,
 *     so every entry point may be recorded into a CUDA graph;
 *   - every launch goes to the cudaStream_t passed last (as void*; 0 = legacy default stream);
 *   - return 0 on success, non-zero on error, message via cdcmdr_last_error() (thread-local);
 *   - row-major everywhere; "ld*" = leading dimension in ELEMENTS;
 *   - activations are fp32 (exact-parity path) or bf16 (tensor-core path); bf16 pointers are uint16_t*;
 *   - all reductions are deterministic (fixed partition + fixed-order tree; no float atomics).
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
 * Device-resident step state (so that a whole training step can live in one CUDA graph):
 *   optimizer step count t, the dropout seed of this step and the Adam scalars derived from t.
 * cdcmdr_step_tick: t += 1; seed = hash(base_seed, t); lr_t = lr/(1-beta1^t); bc2_sqrt = sqrt(1-beta2^t)
 *                   (torch.optim.Adam bias corrections, run.py:720-721).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t step;
  uint64_t seed;
  float lr_t, beta1, beta2, eps, weight_decay, bc2_sqrt;
  float pad[2];
} cdcmdr_step_state_t;                       /* 48 bytes, device memory */
int cdcmdr_step_state_init(cdcmdr_step_state_t* st, int64_t step, cdcmdr_stream_t s);
int cdcmdr_step_tick(cdcmdr_step_state_t* st, float lr, float beta1, float beta2, float eps, float weight_decay,
                     uint64_t base_seed, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a1  FeaturesEmbedding.forward                                        model/layer.py:147-157
 * out[b, f*E + e] = table[x[b,f] + offsets[f], e].  Either output may be NULL.  out_f32 has leading
 * dimension F*E; out_bf16 has leading dimension ld_bf16 >= F*E (padding columns are left untouched).
 * *oob_flag (device int, may be NULL) is set to 1 if any index falls outside [0, V) (such rows read as 0).
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_embed_gather_fwd(const int32_t* x, const int64_t* offsets, const float* table,
                            float* out_f32, uint16_t* out_bf16, int64_t ld_bf16,
                            int64_t B, int F, int E, int64_t V, int* oob_flag, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a2  embedding backward = autograd of nn.Embedding (sparse=False)     model/layer.py:140, run.py:491
 * plan: stable radix sort of (row = x+offsets, position = b*F+f) -> unique rows, segment bounds
 * (deterministic summation order: ascending position inside a row).  Workspace layout is opaque.
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_embed_plan_bytes(int64_t n_idx, int64_t V, int E_max);
int cdcmdr_embed_plan_build(const int32_t* x, const int64_t* offsets, int64_t B, int F, int64_t V, int E_max,
                            void* plan, size_t plan_bytes, cdcmdr_stream_t s);
/* dense [V,E] gradient (reference semantics, what torch.optim.Adam consumes).  grad_out: fp32 [B, ldg]. */
int cdcmdr_embed_bwd_dense(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F, int E,
                           int64_t V, float* grad_table, cdcmdr_stream_t s);
/* a2+a7+a17 fused, reference-exact ("dense_exact"): ONE sweep over all V rows:
 *   g = segment_sum(row) + 2*l2*w + wd*w ; Adam(m, v, w)            layer.py:31,96-112 ; run.py:720
 *   optionally accumulates sum(w_old^2) (the table term of get_regularization_loss) into *reg_sumsq.
 * "sparse_lazy": touched rows only (documented deviation from the reference). */
int cdcmdr_embed_bwd_adam_dense_exact(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                      int E, int64_t V, float* table, float* m, float* v, float l2,
                                      const cdcmdr_step_state_t* st, double* reg_sumsq /* device scalar or NULL */,
                                      cdcmdr_stream_t s);
/* The same update from BF16 row gradients (the replicas' row-gradient exchange delivers bf16): grad_out bf16 [B, ldg],
 * widened inside the segment sums - no separate cast pass over the inbox.  Requires E % 4 == 0, ldg % 4 == 0. */
int cdcmdr_embed_bwd_adam_dense_exact_g16(const uint16_t* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F, int E,
                                          int64_t V, float* table, float* m, float* v, float l2, const cdcmdr_step_state_t* h,
                                          double* reg_sumsq, cdcmdr_stream_t s);
int cdcmdr_embed_bwd_adam_sparse_lazy(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                      int E, int64_t V, float* table, float* m, float* v, float l2,
                                      const cdcmdr_step_state_t* st, cdcmdr_stream_t s);
/* sparse_lazy with an incrementally maintained regulariser value (a full sum of squares per step would read the whole shard:
 * 16 GB per GPU at 500 M x 64 over 8 GPUs).  reg_running (device double): sum(w^2) over the table, initialised once by the caller
 * (cdcmdr_reg_l2_sum); the call writes its value BEFORE this update to reg_before (what the step's loss uses; may be NULL) and
 * adds the change of the touched rows (fixed reduction order). */
int cdcmdr_embed_bwd_adam_sparse_lazy_reg(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                          int E, int64_t V, float* table, float* m, float* v, float l2,
                                          const cdcmdr_step_state_t* h, double* reg_running, double* reg_before,
                                          cdcmdr_stream_t s);
/* a1 over a ROW-RANGE sharded table in NVLink peer memory (BASELINE configs[4]): rank r owns rows
 * [r*rows_per, (r+1)*rows_per) of the concatenated table; shards = DEVICE array of every rank's shard base pointer (mapped into
 * this process; the host side uses torch's symmetric-memory allocator for the mapping).  One kernel reads every row where it
 * lives - no index / row exchange.  x, offsets, outputs as cdcmdr_embed_gather_fwd (global offsets, V = total rows).  Bit-exact. */
int cdcmdr_embed_gather_peer(const int32_t* x, const int64_t* offsets, const float* const* shards, int64_t rows_per,
                             float* out_f32, uint16_t* out_bf16, int64_t ld_bf16, int64_t B, int F, int E, int64_t V,
                             int* oob_flag, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a3/a4/a5/a9/a11  nn.Linear / F.linear / torch.matmul call sites      layer.py:119,185,193,275,336
 * Strided, grouped fp32 GEMM on CUDA cores (exact-parity path):
 *   C[g](m,n) = epi( sum_k A[g](m,k) * Bt[g](n,k) ),  A(m,k) = A[g*a_gs + m*a_rs + k*a_cs], etc.
 *   epi(v) = v + bias[g*bias_gs + n]; relu if act==1; then, if mask != NULL,
 *            v *= (mask[g*mask_gs + m*mask_rs + n] > 0) * mask_scale; then dropout(drop_p, *seed_dev, salt);
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
  float drop_p; const uint64_t* seed_dev; uint32_t salt;
  int32_t accumulate;
  int32_t split_k; float* workspace;
} cdcmdr_gemm_f32_t;
int cdcmdr_gemm_f32(const cdcmdr_gemm_f32_t* p, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 / TMEM / TMA bf16 GEMM (tensor-core path of the same call sites), fp32 accumulate in TMEM:
 *   acc[g](m,n) = sum_k A[g](m,k) * Bt[g](n,k)
 * Each operand is a STORED row-major bf16 matrix [rows, cols] with leading dimension ld (elements), read in place
 * through a TMA tensor map (pointer and ld*2 must be 16-byte aligned; out-of-bounds reads are zero):
 *   K-major  (x_mn_major == 0): stored [MN index, K index]   - activations / weights of forward and input-gradient GEMMs
 *   MN-major (x_mn_major != 0): stored [K index, MN index]   - weight-gradient GEMMs reduce over the batch (the row index)
 * Group g starts at (MN, K) offset (g*a_gm, g*a_gk) of A and (g*b_gn, g*b_gk) of Bt; with K offsets K % 64 must be 0.
 * Columns n < n_main: main = epi(acc + bias) as bf16 at out_main[m*ld_main + g*main_gn + n]
 *     epi: relu (act==1), * (mask>0)*mask_scale (mask bf16 at mask[m*ld_mask + g*mask_gn + n]), dropout, += if accumulate.
 * Columns n >= n_main: aux = acc + bias as fp32 at out_aux[m*ld_aux + g*aux_gn + (n - n_main)] (+= if accumulate).
 * split_k > 1 (requires n_main == 0): K slice z is written to out_aux + z*aux_split_stride; sum the slices with
 * cdcmdr_splitk_reduce.  cdcmdr_gemm_bf16_tc_splits(K, want) = number of non-empty slices to pass as split_k.
 * block_n: N tile (multiple of 16, <= 256); 0 = auto.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const uint16_t* A; int64_t lda; int64_t a_rows, a_cols;
  const uint16_t* Bt; int64_t ldb; int64_t b_rows, b_cols;
  int64_t M, N, K;
  int32_t G; int64_t a_gm, a_gk, b_gn, b_gk;
  int32_t a_mn_major, b_mn_major;
  const float* bias; int64_t bias_gs;
  int64_t n_main;
  uint16_t* out_main; int64_t ld_main, main_gn;
  float* out_aux; int64_t ld_aux, aux_gn;
  int32_t act;
  const uint16_t* mask; int64_t ld_mask, mask_gn; float mask_scale;
  float drop_p; const uint64_t* seed_dev; uint32_t salt;
  int32_t accumulate;
  int32_t split_k; int64_t aux_split_stride;
  int32_t block_n;
  /* a11 CrossNetV2 layer fused into the epilogue (layer.py:339-343: x_{l+1} = x0 * (x_l W^T) + b + x_l), cross_x0 != NULL:
   *   acc = A Bt^T (this GEMM: A = bf16 x_l, Bt = bf16 W_l) ; y = cross_x0[m, n] * acc + bias[n] + cross_x[m, n]
   *   out_main[m*ld_main + n] = bf16(y)  (the layer output = next layer's GEMM operand; n_main must equal N)
   *   out_aux[m*ld_aux + n]   = acc      (fp32, kept for the backward's dx0 += dx * acc; may be NULL)
   * cross_x0 / cross_x are bf16 [M, N] with pitch ld_cross (x is normally the A operand itself); they are read as 32 x 32 boxes by
   * TMA and both outputs leave by TMA stores.  Requires N % 32 == 0, 16-byte aligned bases and pitches, G == 1, a bias, and no
   * mask / dropout / accumulate / split-K / activation. */
  const uint16_t* cross_x0; const uint16_t* cross_x; int64_t ld_cross;
} cdcmdr_gemm_bf16_t;
int cdcmdr_gemm_bf16_tc(const cdcmdr_gemm_bf16_t* p, cdcmdr_stream_t s);
int cdcmdr_gemm_bf16_tc_splits(int64_t K, int32_t want);
/* Tuning switches of cdcmdr_gemm_bf16_tc (bit mask; mode < 0 only queries; returns the previous mask).  Every combination computes
 * the same values (same K order per output element):
 *   bit 0  single-CTA 128 x block_n tiles only
 *   bit 1  no A-resident column sweep (every tile streams its A k-blocks through the ring)
 *   bit 2  CTA pairs (tcgen05 cta_group::2: 256 x block_n tiles, B operand split between the two SMs of a pair) whenever the tile
 *          shape allows; default (neither bit): pairs only when a tile's K loop has >= 32 k-blocks
 *   bit 3  no lean epilogue: by default a split-K launch whose slices run >= 48 k-blocks uses 192 threads (one epilogue warp per
 *          TMEM lane quarter instead of four) and 24 KB less shared memory, so that small HBM-bound CTAs of a concurrent stream
 *          fit on the SM next to it (the embedding backward runs under the weight-gradient GEMM)
 * (Measured on B200: pairs +8 % at 128 k-blocks per tile, -20 % at 6; reading staged tiles back for coalesced st.global instead of
 * TMA stores was 15-25 % slower and was removed.) */
int cdcmdr_gemm_bf16_tc_mode(int mode);
/* Diagnostics: while `counters8` (device, 16 x uint64 - the first 8 documented here, [8..11] = epilogue warp 0 inside the bf16 store
 * path: accumulator load, math, staging + store issue, bias slice - zeroed by the caller) is installed, every cdcmdr_gemm_bf16_tc launch adds
 * the SM cycles its warps spent in each pipeline wait, summed over CTAs: [0] TMA producer waiting for a free stage, [1] MMA issuer
 * waiting for a drained accumulator, [2] MMA issuer waiting for operands, [3] epilogue warp 0 waiting for an accumulator,
 * [4] epilogue warp 0 waiting for its staging buffer, [5] epilogue warp 0 lifetime, [6] CTA lifetime, [7] CTAs.  NULL switches it off. */
int cdcmdr_gemm_bf16_tc_profile(uint64_t* counters8);
/* out[r*ld_out + c] (+)= sum_z part[z*stride + r*ld_part + c]   (deterministic order) */
int cdcmdr_splitk_reduce(const float* part, int64_t stride, int32_t splits, float* out, int64_t rows, int64_t cols,
                         int64_t ld_part, int64_t ld_out, int32_t accumulate, cdcmdr_stream_t s);
/* dst[c*ldd + r] = src[r*lds + c] for a bf16 matrix */
int cdcmdr_transpose_bf16(const uint16_t* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols,
                          cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a5/a9  gate softmax + expert-weighted sum (CGC.forward, MMoE.forward)  ple.py:106-123, mmoe.py:56-60
 * H: expert outputs [B, ldh], expert e at columns [e*h, (e+1)*h).  logits: fp32 [B, ldl]; gate j reads
 * logits[:, gate_col[j] : gate_col[j]+gate_n[j]] and mixes experts gate_sel[j*max_sel + s].
 * out gate j -> out[:, j*h : (j+1)*h].  probs (fp32 [B, n_gates*max_sel]) saved for backward.
 * is_bf16: H/out (and dOut/dH) are bf16, else fp32.  Descriptor arrays live in DEVICE memory.
 * Limits: n_gates <= 32, max_sel <= 32, n_experts <= 64.  With h in {64, 128}, n_pairs <= 32 and n_gates <= 8 the backward runs
 * as one warp per row with every lane owning h/32 columns of each expert block (same results, fixed summation order).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_gates, n_experts, h, max_sel;
  const int32_t* gate_col; const int32_t* gate_n; const int32_t* gate_sel;   /* device */
  int32_t n_pairs;   /* sum of gate_n (the host's copy; 0 = not given): lets the launcher pick the row-per-warp backward kernel */
} cdcmdr_mix_desc_t;
int cdcmdr_gate_mix_fwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* logits, int64_t ldl,
                        void* out, int64_t ldo, float* probs, int64_t B, int is_bf16, cdcmdr_stream_t s);
/* backward: dH[:, e] = (sum_{j contains e} p_j[e]*dOut_j) * (H_e > 0) * relu_scale   (relu_scale<=0: no mask)
 *           dlogits gate j -> dlogits[:, gate_col[j] + s] (fp32) */
int cdcmdr_gate_mix_bwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* probs,
                        const void* dOut, int64_t ldo, void* dH, int64_t lddh, float relu_scale,
                        float* dlogits, int64_t lddl, int64_t B, int is_bf16, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a4/a8/a14  BatchNorm1d (train: batch stats; eval: running stats) + ReLU + dropout
 *            layer.py:187,202-204,279 ; star.py:169-181 ; torch BatchNorm1d defaults (eps 1e-5, momentum .1)
 * Z fp32 [B, ldz] with C columns.  Train: computes mean / biased var per column (deterministic two-stage
 * reduction through `scratch`), updates running stats with the unbiased var, writes save_mean/save_invstd
 * (fp32 [C]) and A = dropout(relu(gamma*xhat+beta)).  gamma2/beta2 (may be NULL): STAR partitioned norm,
 * gamma*gamma2 and beta+beta2.  relu=0 skips the activation.  Eval: save_mean/save_invstd are filled from
 * the running statistics.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* gamma; const float* beta; const float* gamma2; const float* beta2;
  float* running_mean; float* running_var;
  float* save_mean; float* save_invstd;
  int32_t train, relu;
  float drop_p; const uint64_t* seed_dev; uint32_t salt;
} cdcmdr_bn_t;
size_t cdcmdr_bn_scratch_bytes(int64_t C);
int cdcmdr_bn_fwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, void* A, int64_t lda_, int a_is_bf16,
                  int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s);
/* backward through dropout/relu/BN: dZ from dA.  A is the forward output (its sign gives the relu/dropout mask).
 * (A, dA, dZ) dtypes: all fp32, or (bf16, fp32, bf16), or all bf16.
 * dgamma/dbeta fp32 [C] (+= if accumulate).  With gamma2: dgamma is the gradient of the PRODUCT gamma*gamma2
 * and dbeta that of the SUM. */
int cdcmdr_bn_bwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                  const void* dA, int64_t ldda, int da_is_bf16, void* dZ, int64_t lddz, int dz_is_bf16,
                  float* dgamma, float* dbeta,
                  int accumulate, int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s);

/* The same two reductions split at the point where data-parallel ranks exchange per-feature sums (train-mode
 * BatchNorm statistics are over the GLOBAL batch in the single-device reference).  `sums` is a device double[2*C]:
 *   forward : sums[c] = sum_b z, sums[C+c] = sum_b z^2 over the local B rows; the caller all-reduces it over ranks and
 *             passes n_total = total rows; cdcmdr_bn_fwd_apply derives mean / var / running statistics and normalises.
 *   backward: cdcmdr_bn_bwd_stats writes sums[c] = sum_b dy, sums[C+c] = sum_b dy*xhat (local rows) and the LOCAL parameter
 *             gradients dgamma / dbeta (the gradient all-reduce adds the ranks); after the all-reduce of `sums`,
 *             cdcmdr_bn_bwd_apply computes dZ for the local rows.
 * cdcmdr_bn_fwd / cdcmdr_bn_bwd are exactly stats + apply with n_total = B.  B may be 0 (a rank with no rows). */
int cdcmdr_bn_fwd_stats(const float* Z, int64_t ldz, int64_t B, int64_t C, double* sums, void* scratch, cdcmdr_stream_t s);
int cdcmdr_bn_fwd_apply(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, void* A, int64_t lda_, int a_is_bf16,
                        int64_t B, int64_t n_total, int64_t C, const double* sums, cdcmdr_stream_t s);
int cdcmdr_bn_bwd_stats(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                        const void* dA, int64_t ldda, int da_is_bf16, float* dgamma, float* dbeta, int accumulate,
                        int64_t B, int64_t C, double* sums, void* scratch, cdcmdr_stream_t s);
int cdcmdr_bn_bwd_apply(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                        const void* dA, int64_t ldda, int da_is_bf16, void* dZ, int64_t lddz, int dz_is_bf16,
                        int64_t B, int64_t n_total, int64_t C, const double* sums, void* scratch, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a4/a8  the Linear(d, 1) output layer of G towers                       layer.py:192-193
 * fwd:  logit[b*ldo + g] = A[b, g*d:(g+1)*d] . w[g*d:(g+1)*d] + bias[g]      (A fp32 or bf16)
 * bwd:  dA[b, g*d+k] = dlogit[b,g]*w[g,k] (fp32, may be NULL) ; dW[g,k] = sum_b dlogit[b,g]*A[b,g*d+k] ;
 *       dbias[g] = sum_b dlogit[b,g]     (deterministic; scratch >= cdcmdr_colsum_scratch_bytes(G*d))
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_rowdot_fwd(const void* A, int64_t lda_, int a_is_bf16, const float* w, const float* bias, float* out,
                      int64_t ldo, int64_t B, int G, int d, cdcmdr_stream_t s);
int cdcmdr_rowdot_bwd(const void* A, int64_t lda_, int a_is_bf16, const float* w, const float* dlogit, int64_t ldl,
                      float* dA, int64_t ldda, float* dW, float* dbias, int64_t B, int G, int d, void* scratch,
                      cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a8/a15/a17  tower logit + FeaturesLinear + Sigmoid + tower selection + BCELoss(mean) and its backward
 *             layer.py:48-56 ; cdc.py:99-111 ; run.py:483-484,723
 * logits fp32 [B, T] (tower outputs), lin fp32 (lin[b*ld_lin], may be NULL).  pred = sigmoid(logits + lin).
 * mode 0: per-sample column sel[b] (int64 [B]; pred.gather / CDC split without domain_i)
 * mode 1: fixed column `col` (CDC split with domain_i)
 * mode 2: mean over T (CDC warmup)
 * mode 3: no selection (forward only: target must be NULL)
 * target (NULL: forward only) is int16 or float per `target_is_f32`.  Writes pred [B,T], psel fp32 [B] (may be
 * NULL), *loss_sum (device double: sum over the batch of the clamped BCE terms) and
 * dlogits fp32 [B,T] = dLoss/dlogit with Loss = loss_sum * inv_batch (if dlogits != NULL); dlin (may be NULL):
 * dlin[b*ld_dlin] = sum_t dlogits[b,t], the gradient of the FeaturesLinear logit.
 * scratch >= cdcmdr_reduce_scratch_bytes().
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_sigmoid_select_bce(const float* logits, const float* lin, int64_t ld_lin, int64_t B, int32_t T, int32_t mode,
                              const int64_t* sel, int32_t col, const void* target, int target_is_f32,
                              float* pred, float* psel, double* loss_sum, float* dlogits,
                              float* dlin, int64_t ld_dlin, float inv_batch, void* scratch, cdcmdr_stream_t s);
/* autograd path: dlogits[b,t] = dpred[b,t] * pred[b,t] * (1 - pred[b,t]) ; dlin[b*ld_dlin] = sum_t dlogits[b,t] */
int cdcmdr_sigmoid_bwd(const float* pred, const float* dpred, float* dlogits, float* dlin, int64_t ld_dlin,
                       int64_t B, int32_t T, cdcmdr_stream_t s);

/* a7  get_regularization_loss: sum_i coef[i] * w[i]^2 over a flat parameter arena  layer.py:96-112
 *     coef == NULL: every element uses coef_scalar (the embedding table) */
int cdcmdr_reg_l2_sum(const float* w, const float* coef, float coef_scalar, int64_t n, double* out_sum, void* scratch,
                      cdcmdr_stream_t s);
size_t cdcmdr_reduce_scratch_bytes(void);
/* grad[i] (+)= scale * 2*coef[i]*w[i]   (backward of the regulariser for the autograd path; coef NULL -> scalar) */
int cdcmdr_reg_l2_grad(const float* w, const float* coef, float coef_scalar, float scale, float* grad, int accumulate,
                       int64_t n, cdcmdr_stream_t s);
/* out = (A > 0) ? dA * scale : 0 over a [rows, cols] block: ReLU/dropout backward on its own (batch-size-1 path
 * of MultiLayerPerceptron, where BatchNorm is skipped)                               layer.py:202-204 */
int cdcmdr_relu_mask_f32(const float* dA, int64_t ldda, const float* A, int64_t lda_, float* out, int64_t ldo,
                         int64_t rows, int64_t cols, float scale, cdcmdr_stream_t s);
/* the same with each operand fp32 (flag 0) or bf16 (flag 1): the tensor-core path on a one-row batch, where
 * activations are bf16 and the tower gradient fp32                                    layer.py:202-204 */
int cdcmdr_relu_mask(const void* dA, int64_t ldda, int da_bf16, const void* A, int64_t lda_, int a_bf16, void* out,
                     int64_t ldo, int out_bf16, int64_t rows, int64_t cols, float scale, cdcmdr_stream_t s);

/* a17  torch.optim.Adam over a flat arena, fused with the L2-regulariser gradient  run.py:720-721
 *   g = grad[i] + 2*l2coef[i]*w[i] + wd*w[i] ; m,v,w update.  present[i]==0 (uint8, may be NULL)
 *   marks parameters whose grad is None in the reference (skipped).  l2coef may be NULL. */
int cdcmdr_adam_dense(float* w, const float* grad, float* m, float* v, const float* l2coef,
                      const uint8_t* present, int64_t n, const cdcmdr_step_state_t* st, cdcmdr_stream_t s);

/* column sums of a [B, C] matrix (bias gradients): out[c] (+)= sum_b X[b*ld + c]; deterministic */
int cdcmdr_colsum(const void* X, int64_t ld, int is_bf16, int64_t B, int64_t C, float* out, int accumulate,
                  void* scratch, cdcmdr_stream_t s);
size_t cdcmdr_colsum_scratch_bytes(int64_t C);
/* elementwise helpers on device buffers */
int cdcmdr_cast_f32_bf16(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols,
                         cdcmdr_stream_t s);
int cdcmdr_cast_bf16_f32(const uint16_t* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols,
                         int accumulate, cdcmdr_stream_t s);
/* out = a*b (op 0), a+b (op 1), out += a*b (op 2), out += a (op 3; b ignored) over n contiguous floats:
 * STAR W_d*W_s, b_d+b_s and their gradients  star.py:91-92 */
int cdcmdr_ewise_f32(const float* a, const float* b, float* out, int64_t n, int op, cdcmdr_stream_t s);
/* the same factors for all G towers in one launch: a, out [G, n] blocks, b one shared [n] block (ops 0, 1) or [G, n] (op 2)
 *   op 0: out[g,i] = a[g,i]*b[i]   op 1: out[g,i] = a[g,i]+b[i]   op 2: out[i] = sum_g a[g,i]*b[g,i]   op 3: out[i] += sum_g a[g,i]
 *   op 4: out[g,i] += a[g,i]                                                                            star.py:91-92, 100-101 */
int cdcmdr_ewise_group_f32(const float* a, const float* b, float* out, int64_t n, int G, int op, cdcmdr_stream_t s);
/* out[r, c] (+)= a[r*lda + c] (op 1: * b[r*ldb + c]) for a strided 2-D block */
int cdcmdr_add2d_f32(const float* a, int64_t lda, float* out, int64_t ldo, int64_t rows, int64_t cols,
                     int accumulate, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a10/a11/a12  cross-network elementwise stages                         layer.py:495-515, 332-343, 380-407
 * cross_fuse_fwd:  out = x0 * xw + b + x     xw is [B,1] (xw_cols==1, v1) or [B,D] (v2)
 * cross_fuse_bwd:  given dout: dx0_acc += dout*xw ; dxw = dout*x0 (v2: [B,D]; v1: row-summed [B,1])
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
/* backward: du[e] = dout*g[:,e]*x0 ; dgate[b,e] = sum_d dout*x0*(u_e+bias) ; dx0_acc += sum_e dout*g_e*(u_e+bias) */
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
 * counts[g] rows per group, group_start[g] exclusive prefix (n_group+1 entries; the last is the number of
 * routed rows).  group is int64 [B] (values outside [0,n_group) are dropped).  n_group <= 256.
 * scratch >= cdcmdr_route_scratch_bytes(B, n_group).
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_route_scratch_bytes(int64_t B, int n_group);
int cdcmdr_route_partition(const int64_t* group, int64_t B, int n_group, int32_t* perm, int32_t* counts,
                           int32_t* group_start, void* scratch, cdcmdr_stream_t s);
/* gather: dst[i, :] = src[perm[i], :] ; scatter: dst[perm[i], :] = src[i, :]   (rows of cols*elt_bytes bytes).
 * perm == NULL: identity, i.e. a strided 2-D copy (packing / unpacking the per-owner blocks of the embedding exchange) */
int cdcmdr_permute_rows(const void* src, int64_t lds, const int32_t* perm, int64_t n, int64_t cols, int elt_bytes,
                        void* dst, int64_t ldd, int scatter, cdcmdr_stream_t s);
/* N1  affinity probe of CDC: per-domain BCE means of one batched evaluation      run.py:551-560, cdc.py:116-119
 * The reference evaluates the model once per domain (`cdc_test_all_domain`: n_domain forwards, each followed by a host BCE);
 * here the domains' batches are concatenated, evaluated in one pass, and this entry point reduces the selected predictions
 * per contiguous row segment: out_mean[s] = mean over rows [seg_start[s], seg_start[s+1]) of the clamped BCE term (torch
 * BCELoss: log terms clamped at -100).  pred fp32 with row stride ld_pred; target int16 or float per `target_is_f32`;
 * seg_start int64 [n_seg + 1] in device memory; empty segments give NaN like torch's mean of an empty tensor.
 * scratch >= cdcmdr_bce_segments_scratch_bytes(n_seg).  Deterministic (fixed chunking, fixed summation order). */
size_t cdcmdr_bce_segments_scratch_bytes(int n_seg);
int cdcmdr_bce_segments(const float* pred, int64_t ld_pred, const void* target, int target_is_f32, const int64_t* seg_start,
                        int n_seg, float* out_mean, void* scratch, cdcmdr_stream_t s);
/* strided batch of 2-D copies: dst[b*dst_bs + r*ldd + c] = src[b*src_bs + r*lds + c] for b < batches, r < rows, c < cols
 * (element offsets).  Packs the per-task gate weights of a CGC level (ple.py:89-94) into the block-diagonal operand of one
 * GEMM, and unpacks that operand's gradient, in one launch instead of one per gate. */
int cdcmdr_copy2d_batched(const void* src, int64_t src_bs, int64_t lds, void* dst, int64_t dst_bs, int64_t ldd, int64_t batches,
                          int64_t rows, int64_t cols, int elt_bytes, cdcmdr_stream_t s);
/* a15  groups[b] = domain2group[x[b, domain_idx]]                       cdc.py:105 */
int cdcmdr_domain_to_group(const int32_t* x, int64_t B, int F, int domain_idx, const int64_t* domain2group,
                           int n_domain, int64_t* groups, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * N3  field self-attention block   BaseModel.build_atten / atten_forward, model/layer.py:58-84
 * (config.use_atten, on in the stock config.py:24-28: PLE / MMoE add its scalar output to every tower logit)
 *
 * The Linear layers of the block (atten_embedding, nn.MultiheadAttention in_proj / out_proj, V_res_embedding) act on the
 * token matrix [B*L, .] (row b*L + l = field l of sample b) and run on cdcmdr_gemm_f32.  These entry points are the rest:
 *
 * cdcmdr_attn_fwd: the core of torch.nn.MultiheadAttention (batch_first=False semantics, no masks) for every sample and head:
 *   qkv fp32 [B*L, ld] with columns [Q | K | V], A = H*dh each, head h in columns h*dh.. of its third;
 *   P = softmax_j(scale * Q_h K_h^T) over the L tokens of the sample (scale = 1/sqrt(dh)), dropout on P in train mode
 *   (drop_p > 0: the library's stateless hash, seed *seed_dev, salt), out[b*L + i, h*dh + d] = sum_j P[i, j] V[j, d].
 *   probs (may be NULL for inference) receives the pre-dropout softmax [B, H, L, L] for the backward.  L <= 32.
 * cdcmdr_attn_bwd: dqkv [B*L, lddq] (same column layout) from dout [B*L, lddo], qkv and probs; same drop_p / seed / salt.
 * cdcmdr_attn_pool_fwd: the head  F.relu(z).view(B, L*A) -> atten_linear (no bias):  lin[b*ld_lin] (+)= sum_j relu(z[b, j]) w[j],
 *   z fp32 [B, n] contiguous (n = L*A); accumulate != 0 adds onto the FeaturesLinear logit already in lin (layer.py:52-54).
 * cdcmdr_attn_pool_bwd: dz[b, j] = dlin[b*ld_dlin] * w[j] * [z[b, j] > 0] ; dw[j] = sum_b dlin[b] * relu(z[b, j])
 *   (deterministic two-stage column sum in double; scratch >= cdcmdr_attn_pool_scratch_bytes(B, n)).
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_attn_fwd(const float* qkv, int64_t ld, float* out, int64_t ldo, float* probs, int64_t B, int L, int H, int dh,
                    float scale, float drop_p, const uint64_t* seed_dev, uint32_t salt, cdcmdr_stream_t s);
int cdcmdr_attn_bwd(const float* qkv, int64_t ld, const float* probs, const float* dout, int64_t lddo, float* dqkv, int64_t lddq,
                    int64_t B, int L, int H, int dh, float scale, float drop_p, const uint64_t* seed_dev, uint32_t salt,
                    cdcmdr_stream_t s);
int cdcmdr_attn_pool_fwd(const float* z, const float* w, float* lin, int64_t ld_lin, int accumulate, int64_t B, int64_t n,
                         cdcmdr_stream_t s);
size_t cdcmdr_attn_pool_scratch_bytes(int64_t B, int64_t n);
int cdcmdr_attn_pool_bwd(const float* z, const float* w, const float* dlin, int64_t ld_dlin, float* dz, float* dw, int64_t B,
                         int64_t n, void* scratch, cdcmdr_stream_t s);
/* The same four stages on bf16 token matrices (the tensor-core path: the block's projections run on cdcmdr_gemm_bf16_tc, so qkv,
 * the core's output, dout and dqkv are bf16; arithmetic inside stays fp32).  The forward stores NO probabilities: the backward
 * recomputes the L x L softmax of every (sample, head) from q and k (same dropout hash).  dh must be even.  w, lin, dlin, dw fp32. */
int cdcmdr_attn_fwd_bf16(const uint16_t* qkv, int64_t ld, uint16_t* out, int64_t ldo, int64_t B, int L, int H, int dh, float scale,
                         float drop_p, const uint64_t* seed_dev, uint32_t salt, cdcmdr_stream_t s);
int cdcmdr_attn_bwd_bf16(const uint16_t* qkv, int64_t ld, const uint16_t* dout, int64_t lddo, uint16_t* dqkv, int64_t lddq,
                         int64_t B, int L, int H, int dh, float scale, float drop_p, const uint64_t* seed_dev, uint32_t salt,
                         cdcmdr_stream_t s);
int cdcmdr_attn_pool_fwd_bf16(const uint16_t* z, const float* w, float* lin, int64_t ld_lin, int accumulate, int64_t B, int64_t n,
                              cdcmdr_stream_t s);
int cdcmdr_attn_pool_bwd_bf16(const uint16_t* z, const float* w, const float* dlin, int64_t ld_dlin, uint16_t* dz, float* dw,
                              int64_t B, int64_t n, void* scratch, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * a5/a6  first CGC level of PLE chained in ONE kernel (ple.py:54,96-124; layer.py:184-190): every expert's
 *   A0_e = dropout(relu(X W0_e^T + b0_e))   [B, d0]      H_e = dropout(relu(A0_e W1_e^T + b1_e))   [B, d1]
 * with A0_e handed from the first GEMM's epilogue to the second GEMM through shared memory, plus the gate / wide-linear logits
 * Lg = X Wg^T + bg (fp32) that read the same X.  bf16 operands, fp32 accumulation in TMEM (tcgen05).
 *   X  bf16 [B, K0] (pitch ldx);  W0 bf16 [nE*d0 + n_g, K0] row-major: the experts' layer-0 weights followed by the n_g gate rows;
 *   b0 fp32 [nE*d0 + n_g];  W1 bf16 [nE*d1, d0];  b1 fp32 [nE*d1].
 *   H  bf16 [B, nE*d1] (pitch ldh), always written.  A0 bf16 [B, nE*d0] (pitch lda0): written for the backward when non-NULL;
 *   NULL (inference) = the layer-0 activation never leaves the SM.  Lg fp32 [B, n_g] (pitch ldg), required when n_g > 0.
 *   drop_p > 0: nn.Dropout on both layers (the library's stateless generator, seed *seed_dev, salts salt0 / salt1).
 * Geometry (cdcmdr_ple_chain_ok): K0 % 8 == 0 and K0 <= 384, d0 in {128, 256}, d1 in {64, 128}, n_g <= 128; bases and pitches
 * 16-byte aligned.  Other shapes use the per-layer GEMM entry points.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const uint16_t* X; int64_t ldx; int64_t B; int32_t K0;
  const uint16_t* W0; const float* b0;
  const uint16_t* W1; const float* b1;
  int32_t nE, d0, d1, n_g;
  uint16_t* A0; int64_t lda0;
  uint16_t* H; int64_t ldh;
  float* Lg; int64_t ldg;
  float drop_p; const uint64_t* seed_dev; uint32_t salt0, salt1;
} cdcmdr_ple_chain_t;
int cdcmdr_ple_chain_ok(int32_t K0, int32_t d0, int32_t d1, int32_t n_g);
int cdcmdr_ple_chain_fwd(const cdcmdr_ple_chain_t* p, cdcmdr_stream_t s);
/* Diagnostics: while `counters16` (device, 16 x uint64, zeroed by the caller) is installed every cdcmdr_ple_chain_fwd launch adds
 * SM cycles spent per pipeline wait, summed over CTAs: [0] producer waiting for a free weight slot, [1] MMA issuer waiting for a
 * free accumulator slot, [2] for weights, [3] for activation blocks, [4] epilogue warp 0 waiting for an accumulator, [5] for a
 * free activation block, [6] epilogue warp 0 lifetime, [7] CTAs, [8] MMA warp lifetime, [9] producer lifetime.  NULL = off.  The counters exist only in a
 * library built with -DCDCMDR_CHAIN_PROF (they cost ~15 % of the kernel); otherwise the call is accepted and nothing is counted. */
int cdcmdr_ple_chain_profile(uint64_t* counters16);

/* ---------------------------------------------------------------------------------------------
 * N4  evaluation metrics on the device (Run.test / evaluate_multi_domain, run.py:647-711: scikit-learn roc_auc_score and
 *     log_loss over the whole set and per domain).  pred fp32 [n] probabilities; target int16 (or fp32) [n]; domain int32 (or
 *     int64) [n] in [0, n_domain), or NULL = one set.  out double [n_domain, 4] = {AUC, log loss, positives, samples} per
 *     domain: AUC = Mann-Whitney statistic with midranks for tied predictions (= the trapezoidal ROC integral), log loss with p
 *     clipped to [FLT_EPSILON, 1 - FLT_EPSILON] as sklearn does for fp32 predictions; both NaN for a set with one class only (the
 *     reference's `except ValueError` branch).  One stable sort by (domain, prediction) + fixed-order reductions: deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_auc_logloss_scratch_bytes(int64_t n, int32_t n_domain);
int cdcmdr_auc_logloss(const float* pred, const void* target, int target_is_f32, const void* domain, int domain_is_i64,
                       int64_t n, int32_t n_domain, double* out, void* scratch, cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * (e) multi-GPU: small all-reduce over NVLink peer memory (cross-replica BatchNorm statistics - the per-feature
 *     (sum, sum of squares) between cdcmdr_bn_*_stats and cdcmdr_bn_*_apply - and the per-rank loss sums; the reference is a
 *     single-device program, layer.py:187 / run.py:484).  peer_bufs: DEVICE array of `world` pointers, entry r = rank r's symmetric
 *     buffer of cdcmdr_peer_allreduce_bytes(world, max_n) bytes, zero-initialised before the first call, mapped into this process
 *     (the host side uses torch's symmetric-memory allocator for the mapping).  out[i] = sum over ranks (in rank order, identical
 *     on every rank) of in[i], i < n <= max_n.  seq: device uint64, zero-initialised, private to this rank; every rank must issue
 *     the same sequence of calls.  One kernel, no host synchronisation, CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------- */
size_t cdcmdr_peer_allreduce_bytes(int world, int64_t max_n);
int cdcmdr_peer_allreduce_f64(double* const* peer_bufs, int rank, int world, const double* in, double* out, int64_t n,
                              int64_t max_n, uint64_t* seq, cdcmdr_stream_t s);

/* The dense gradient arena between replicas (sum over replicas of d(global mean loss), a few MB of fp32): two-shot
 * all-reduce over peer memory - chunk c of `in` is pushed to rank c's inbox, rank c sums its inbox slots in rank order and stores the
 * sums into every rank's outbox, outbox -> out (in == out allowed).  Every element is summed by ONE rank: replicas stay bit identical.
 *   chunk = cdcmdr_peer_allreduce_f32_chunk(world, n) floats; inbox[r]: rank r's world*chunk floats; outbox[r]: rank r's world*chunk
 *   floats; peer_flags[r]: rank r's uint64[2][world], zero-initialised; my_inbox / my_outbox = inbox[rank] / outbox[rank];
 *   seqs2: uint64[2], zero-initialised, private to the rank.  Five launches, no host synchronisation, CUDA-graph capturable. */
int64_t cdcmdr_peer_allreduce_f32_chunk(int world, int64_t n);
int cdcmdr_peer_allreduce_f32(float* const* inbox, float* const* outbox, uint64_t* const* peer_flags, float* my_inbox,
                              const float* my_outbox, int rank, int world, const float* in, float* out, int64_t n, uint64_t* seqs2,
                              cdcmdr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * (e) multi-GPU: the embedding exchange of the replica step fused into its producers / consumers over NVLink peer memory
 *     (replaces the index / row / row-gradient all-to-alls around a1 and a2 - model/layer.py:147-157 and its autograd -
 *     when the table is sharded at field boundaries: owner o holds the rows of fields [fbound[o], fbound[o+1])).
 *     Every pointer array is a DEVICE array of `world` pointers, entry r = rank r's buffer, mapped into this process (symmetric
 *     memory).  All calls are one kernel, never synchronise with the host, and are CUDA-graph capturable.
 *
 * cdcmdr_peer_barrier     peer_flags[r]: rank r's uint64[n_slots][world], zero-initialised; seqs: uint64[n_slots] private to the rank,
 *                         zero-initialised.  Returns (on the stream) once every rank has issued its matching call on `slot`; all
 *                         device writes a rank issued on the stream before its call are visible to every rank after theirs.
 * cdcmdr_dp_push_ids      x [B, F] int32 (this rank's indices) -> recv_ids[o][(rank*B + b)*nf_o + j] = x[b, fbound[o] + j].
 * cdcmdr_dp_gather_push   owner side: for every received index (requester p, sample b, owned field j) the table row
 *                         shard[recv_ids[(p*B + b)*nf + j] + off_local[j]] -> as fp32 or bf16 (out_bf16) either straight into the
 *                         requester's matrix, xs[p][b*ldx + col0 + j*E ...] (packed = 0: runs of nf*E elements cross NVLink), or into
 *                         the requester's staging buffer where this owner's [B, nf*E] block is contiguous,
 *                         xs[p][B*col0 + (b*nf + j)*E ...] (packed = 1: whole 512-byte warp stores; the requester unpacks with
 *                         cdcmdr_copy2d_batched).  An index outside [0, Vl) reads as zeros and sets *oob_flag (as
 *                         cdcmdr_embed_gather_fwd).  Bit-exact.
 * cdcmdr_dp_push_grads    dX [B, ldg] fp32 (gradient of this rank's gathered rows) -> grad_recv[o][(rank*B + b)*nf_o*E + c] =
 *                         dX[b, fbound[o]*E + c], as fp32 or rounded to bf16 (out_bf16).
 * ------------------------------------------------------------------------------------------- */
int cdcmdr_peer_barrier(uint64_t* const* peer_flags, int rank, int world, int slot, int n_slots, uint64_t* seqs, cdcmdr_stream_t s);
int cdcmdr_dp_push_ids(const int32_t* x, int64_t B, int F, int32_t* const* recv_ids, const int32_t* fbound, int rank, int world,
                       cdcmdr_stream_t s);
int cdcmdr_dp_gather_push(const int32_t* recv_ids, const int64_t* off_local, const float* shard, int64_t Vl, void* const* xs,
                          int out_bf16, int64_t ldx, int col0, int64_t B, int nf, int E, int world, int packed, int* oob_flag,
                          cdcmdr_stream_t s);
int cdcmdr_dp_push_grads(const float* dX, int64_t ldg, int64_t B, int F, int E, void* const* grad_recv, int out_bf16,
                         const int32_t* fbound, int rank, int world, cdcmdr_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CDCMDR_H */
