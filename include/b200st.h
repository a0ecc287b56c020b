/*
 * b200st.h — C ABI of libb200st.so: hand-written sm_100a kernels for the joint speech-translation
 * hot path (forward + backward of acoustic encoder -> embedding passing -> translation decoder -> loss).
 *
 * The reference (EdieLu/speech-translation-joint-embedding-passing) has no FFI of its own: every kernel
 * it launches is a PyTorch library call made from Python nn.Modules.  Each entry
 * point below therefore names the reference *call site* (file:line under the reference root) whose
 * library kernel(s) it replaces.  The Python host layer (speech-translation-joint-embedding-passing_b200/)
 * keeps the reference's module signatures and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C, raw device pointers, int64 sizes/strides in ELEMENTS, no torch types;
 *   - `dtype`: B200ST_F32 (0) or B200ST_BF16 (1) = storage type of activation tensors; parameters
 *     named `const float*` are always fp32 (the nn.Parameter storage itself); all accumulation is fp32;
 *   - every call is asynchronous on `stream` (a cudaStream_t), allocates nothing, never syncs;
 *   - return 0 on success, non-zero on error with a message in b200st_last_error() (thread-local).
 */
#ifndef B200ST_H_
#define B200ST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ST_F32 0
#define B200ST_BF16 1

typedef void* b200st_stream_t; /* cudaStream_t */

int b200st_version(void);
const char* b200st_last_error(void);
/* Number of kernels launched by this library in the calling process since load (bench gpu_launches). */
int64_t b200st_launch_count(void);

/* ---- dense contraction -------------------------------------------------------------------------
 * C[b] = relu?( alpha * op(A[b]) * op(B[b]) + bias ) + R[b]        b = 0..batch-1
 * op(A) is M x K: trans_a=0 -> A stored [M,K] (lda), trans_a=1 -> A stored [K,M].
 * op(B) is K x N: trans_b=0 -> B stored [K,N] (ldb), trans_b=1 -> B stored [N,K] (an nn.Linear weight).
 * R (same dtype/shape as C, may alias C, may be NULL), bias fp32[N] or NULL.
 * relu: 0 none, 1 ReLU as written above, 2 "gate": C = (R > 0) ? alpha*op(A)*op(B) + bias : 0 -- R is then the
 * forward activation whose sign masks the result (ReLU backward, layers.py:247, fused into the dX GEMM).
 * Replaces every nn.Linear / matmul / bmm on the path: layers.py:131-134,158-160,192-194,238-249,
 * Seq2seq.py:124-131,180,195,207,253, Dec.py:96-98,432-436, attention.py:192-193, and the LSTM input
 * projections inside torch.nn.LSTM (Enc.py:50-66, Dec.py:104-118). */
int b200st_gemm(int dtype_ab, int dtype_c, int trans_a, int trans_b,
                int64_t M, int64_t N, int64_t K, float alpha,
                const void* A, int64_t lda, int64_t stride_a,
                const void* B, int64_t ldb, int64_t stride_b,
                void* C, int64_t ldc, int64_t stride_c,
                const void* R, int64_t ldr, int64_t stride_r,
                const float* bias, int relu, int64_t batch, b200st_stream_t stream);

/* C = relu?( alpha * (op(A) op(B) + op(A2) op(B2)) + bias ) + R   (unbatched; both pairs share trans_a / trans_b,
 * M and N; K and K2 are their depths).  One launch with a two-segment K loop into the same accumulator on the
 * tensor-core path; replaces the pair "dx = dG_f W_ih_f; dx += dG_r W_ih_r" of a bidirectional LSTM layer's input
 * gradient (autograd of torch.nn.LSTM(bidirectional=True), Enc.py:150-167). */
int b200st_gemm2(int dtype_ab, int dtype_c, int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, int64_t K2,
                 float alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
                 const void* A2, int64_t lda2, const void* B2, int64_t ldb2,
                 void* C, int64_t ldc, const void* R, int64_t ldr, const float* bias, int relu,
                 b200st_stream_t stream);

/* Persistent tensor-core GEMM (one CTA per SM walking the tile list, double-buffered TMEM accumulator) for problems
 * with >= 4 tiles per SM.  Bit mask: bit 0 = persistent kernels, bit 1 = the CTA-pair (cta_group::2, 256 x 256 tile per
 * two SMs) kernel for the largest shapes; default 3, 0 = always one tile per CTA.  Returns the previous mask (test /
 * bench hook). */
int b200st_set_gemm_persistent(int on);
/* SM budget of the PERSISTENT GEMM kernels launched from now on (0 = all SMs; returns the previous value).  Deferred
 * weight-gradient GEMMs that run on side streams under a latency-bound recurrence kernel are launched with a budget that
 * leaves that kernel's SMs free: a persistent CTA holds its SM for the whole GEMM. */
int b200st_set_gemm_sm_budget(int n);

/* GEMM kernel selection: 0 = auto (bf16 operands that TMA can address -> tcgen05 tensor-core kernel, everything
 * else -> exact CUDA-core kernel), 1 = CUDA cores only, 2 = tensor cores required (error if not eligible).
 * Returns the previous mode.  Used by tests to compare the two kernels; the product leaves it at 0. */
int b200st_set_gemm_backend(int mode);

/* GEMM with a fused residual + LayerNorm epilogue (csrc/gemm_ln.cu; bf16, N = 512 = d_model):
 *   Y[M, N]  = A[M, K] . W[N, K]^T (+ bias[N]) + R[M, N]        -- a sub-layer's output: fc + skip (layers.py:190-197),
 *                                                                  w_2 + bias + skip (layers.py:247-252)
 *   YN[M, N] = LayerNorm(Y; gamma, beta, eps), mean[M], rstd[M]  -- the NEXT sub-layer's pre-norm (layers.py:153, 245;
 *                                                                  TFEnc.py:89, TFDec.py:127 for the final norm)
 * replaces b200st_gemm(+residual) followed by b200st_layernorm_fwd.  Four CTAs (a cluster) cover a 128-row block; row
 * statistics are exchanged through distributed shared memory.  b200st_gemm_ln_eligible returns 1 when the shapes /
 * alignments are served (bf16, N == 512, K % 8 == 0, 16-byte aligned rows; Y / YN rows 32-byte aligned); callers use
 * the two separate kernels otherwise.  bias and R may be NULL.  drop_p > 0: dropout on the projection's output before the
 * skip connection (layers.py:194-195, 248-250) with the mask b200st_dropout(site, rng) draws for the dense [M, N] tensor. */
int b200st_gemm_ln_eligible(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W,
                            int64_t ldw, const void* R, int64_t ldr, const void* Y, int64_t ldy, const void* YN,
                            int64_t ldyn, const float* bias, const float* gamma, const float* beta);
int b200st_gemm_ln(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                   const float* bias, const void* R, int64_t ldr, void* Y, int64_t ldy, const float* gamma,
                   const float* beta, float eps, void* YN, int64_t ldyn, float* mean, float* rstd, float drop_p,
                   const int64_t* rng, int64_t site, b200st_stream_t stream);

/* The backward twin (csrc/gemm_ln.cu): the input-gradient GEMM that feeds a LayerNorm backward, with that backward as
 * its epilogue (replaces b200st_gemm followed by b200st_layernorm_bwd_partial):
 *   G[M, N]  = A[M, K] . W[K, N]                              -- dqn = dqp . w_qs, dy = dz . w_1 (N = 512 = d_model)
 *   DX[M, N] = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)) + ADD,  g = G * gamma, xhat = (X - mean) * rstd
 *   partials[row block, 0:N] = sum_rows G * xhat (dgamma), partials[row block, N:2N] = sum_rows G (dbeta): fp32
 *   [b200st_gemm_lnbwd_blocks(M), 2N], column-summed by the caller off the critical path.
 * X, ADD (may be NULL), DX are dense [M, N] bf16.  *_eligible as for b200st_gemm_ln. */
int b200st_gemm_lnbwd_eligible(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W,
                               int64_t ldw, const void* X, const void* ADD, const void* DX, const float* gamma);
int64_t b200st_gemm_lnbwd_blocks(int64_t M);
int b200st_gemm_lnbwd(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                      const void* X, const float* gamma, const float* mean, const float* rstd, const void* ADD, void* DX,
                      float* partials, b200st_stream_t stream);

/* ---- LayerNorm (layers.py:139,153,240,245; TFEnc.py:61,89; TFDec.py:58,127) -------------------- */
int b200st_layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y,
                         float* mean, float* rstd, int64_t rows, int64_t cols, float eps,
                         b200st_stream_t stream);
/* dgamma/dbeta are ACCUMULATED into (caller zero-fills or passes the running .grad buffer). */
int b200st_layernorm_bwd(int dtype, const void* dy, const void* x, const float* gamma,
                         const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                         int64_t rows, int64_t cols, b200st_stream_t stream);
/* Same, plus `add` (same shape/dtype as dx, may be NULL): dx = add + LayerNorm gradient.  Lets a residual block's
 * backward hand the skip-connection gradient to the kernel instead of running a separate add. */
int b200st_layernorm_bwd_add(int dtype, const void* dy, const void* x, const float* gamma,
                             const float* mean, const float* rstd, const void* add, void* dx,
                             float* dgamma, float* dbeta, int64_t rows, int64_t cols, b200st_stream_t stream);
/* Same row work, but the dgamma / dbeta contributions of each CTA are STORED as one row of
 * partials[b200st_layernorm_bwd_partial_blocks(rows, cols)][2 * cols] (fp32: cols sums for dgamma, then cols for dbeta)
 * instead of being reduced with same-address atomics; the caller column-sums the partials (b200st_colsum) -- off the
 * critical path, dgamma / dbeta being parameter gradients.  _blocks() returns 0 when the shape is not supported
 * (needs cols % 128 == 0, cols <= 1024). */
int64_t b200st_layernorm_bwd_partial_blocks(int64_t rows, int64_t cols);
int b200st_layernorm_bwd_partial(int dtype, const void* dy, const void* x, const float* gamma,
                                 const float* mean, const float* rstd, const void* add, void* dx, float* partials,
                                 int64_t rows, int64_t cols, b200st_stream_t stream);

/* ---- multi-head scaled-dot-product attention core (layers.py:162-170,213-229) -------------------
 * q,k,v are the projection outputs viewed as [B, L, H, d] with row strides ldq/ldk/ldv (elements);
 * mask uint8 [B, 1|Lq, Lk] (nonzero = keep; masked scores are SET to -1e9, layers.py:224), addressed
 * mask[b*mask_sb + i*mask_sq + j]; NULL = no mask.  o: [B, Lq, H*d] (ldo); p: [B,H,Lq,Lk] softmax
 * probabilities (the `attn` the reference returns).  Scores use (q / temperature) . k. */
int b200st_mha_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                   const void* v, int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq,
                   void* o, int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                   int64_t d, float temperature, b200st_stream_t stream);
/* Attention kernel selection (test hook), a bit mask: bit 0 set = Transformer attention core on CUDA-core tiles only
 * (default: bf16, d = 64, Lq and Lk <= 64 -> tcgen05 kernel, one (batch, head) per CTA with S / O / dP / dV / dQ / dK
 * accumulated in TMEM); bit 1 set = LAS attention step with one CTA per sequence (default for bf16: a 4-CTA cluster
 * per sequence, keys split across the CTAs and merged through distributed shared memory).  Returns the previous mask. */
int b200st_set_mha_backend(int mode);
/* ds: workspace [B,H,Lq,Lk] (same dtype as p); dq/dk/dv written (not accumulated). */
int b200st_mha_bwd(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                   const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                   void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                   int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature,
                   b200st_stream_t stream);

/* The same pair with attention dropout: attn = dropout(softmax(.), p) multiplies V (layers.py:226; hard-wired p = 0.1
 * in training mode, layers.py:207).  The keep mask of element (b, h, i, j) is Philox4x32-10(seed = rng[0], step =
 * rng[1], site, index ((b*H + h)*Lq + i)*Lk + j) -- see b200st_dropout; backward recomputes it.  `p` still receives
 * the UN-dropped probabilities (backward needs them).  Implemented by the shared-memory tile kernels. */
int b200st_mha_fwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq, void* o,
                           int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                           float temperature, float drop_p, const int64_t* rng, int64_t site,
                           b200st_stream_t stream);
int b200st_mha_bwd_dropout(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                           const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                           void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B,
                           int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature, float drop_p,
                           const int64_t* rng, int64_t site, b200st_stream_t stream);

/* Single-query attention over a key/value cache: the decoder step of incremental decoding (forward_translate /
 * forward_eval, Seq2seq.py:260-393, which upstream re-runs the whole decoder on the prefix every step).
 * q: [n_hyp, H*d] (row stride ldq) = the newest position's queries; key t of hypothesis b is read at
 * k_cache + slot*stride_b + t*stride_t + h*d with slot = anc[t*n_hyp + b] (int32 ancestry table: the cache slot that
 * holds position t of b's history after beam re-ordering, Seq2seq.py:381-384) or, when anc is NULL, slot = b / bdiv
 * (cross-attention: the bdiv beams of one utterance share its encoder keys).  mask: uint8 [*, Lk] rows of stride mask_sb,
 * row b / mask_bdiv, nonzero = keep, NULL = none; masked scores are SET to -1e9 like b200st_mha_fwd.
 * o: [n_hyp, H*d] (ldo).  Scores use (q / temperature) . k. */
int b200st_mha_decode(int dtype, const void* q, int64_t ldq, const void* k_cache, const void* v_cache,
                      int64_t stride_b, int64_t stride_t, const int32_t* anc, int64_t n_hyp, int64_t bdiv,
                      const uint8_t* mask, int64_t mask_sb, int64_t mask_bdiv, void* o, int64_t ldo, int64_t H,
                      int64_t Lk, int64_t d, float temperature, b200st_stream_t stream);



/* ---- LSTM cell pointwise (one step of torch.nn.LSTM, Dec.py:393-419) ----------------------------
 * gates (+ gates_b + gates_c, NULLs skipped) [B,4H] = pre-activations x W_ih^T + h W_hh^T + b_ih + b_hh, possibly
 * delivered as partial products that were computed concurrently; PyTorch order i,f,g,o.
 * acts [B,4H] fp32 = post-activation gates, c/c_prev fp32 [B,H] (c_prev NULL = zeros),
 * h [B,H] dtype; if residual != NULL, out_res = h + residual (Dec.py:417-418). */
int b200st_lstm_cell_fwd(int dtype, const void* gates, const void* gates_b, const void* gates_c,
                         const float* c_prev, void* h, float* c, float* acts, const void* residual,
                         void* out_res, int64_t B, int64_t H, b200st_stream_t stream);
/* dh = dh_a + dh_b + dh_c (NULLs skipped); dc_next NULL = zeros; c_prev NULL = zeros. */
int b200st_lstm_cell_bwd(int dtype, const void* dh_a, const void* dh_b, const void* dh_c,
                         const float* dc_next, const float* acts, const float* c_prev, const float* c,
                         void* dgates, float* dc_prev, int64_t B, int64_t H, b200st_stream_t stream);

/* ---- bidirectional packed-sequence LSTM recurrence (torch.nn.LSTM(bidirectional) on a
 * PackedSequence, Enc.py:150-157,172-177,189-194,206-211), one persistent thread-block-cluster kernel.
 * xproj [2][T][B][4H] dtype: per direction, time-major input projections INCLUDING both biases.
 * w_hh_f / w_hh_r: [4H,H] fp32 (weight_hh_l0 / weight_hh_l0_reverse).  lens int32[B]: valid frames.
 * out: h written at  out[(t/pair)*out_ld_t + b*out_ld_b + (t%pair)*2H + dir*H + u]  (zeros for t>=len):
 *      pair=2 folds the pyramid frame-pair concat (Enc.py:166-167) into the store.
 * hs   [2][T+1][B][H] dtype: forward dir stores h_t at [0][t+1] ([0][0] = 0); reverse at [1][t] ([1][T] = 0).
 * acts [2][T][B][4H] fp32, cs [2][T][B][H] fp32: saved for backward (NULL to skip, inference). */
int b200st_blstm_fwd(int dtype, const void* xproj, const float* w_hh_f, const float* w_hh_r,
                     const int32_t* lens, void* out, int64_t out_ld_t, int64_t out_ld_b, int pair,
                     void* hs, float* acts, float* cs, int64_t T, int64_t B, int64_t H,
                     b200st_stream_t stream);
/* Debug aid: register (or clear with NULL) a device buffer of >= 128 int64; the tensor-core recurrence kernels then
 * record clock64() at fixed points of time steps 64..71 (16 slots per step) for the first CTA. */
int b200st_debug_timeline(void* buf);
/* Debug aid: a one-thread kernel that writes the device's %globaltimer (ns) to *slot at this point of `stream` -- captured into a
 * CUDA graph it time-stamps section boundaries of the replayed step (scripts/graph_timeline.py). */
int b200st_debug_stamp(int64_t* slot, b200st_stream_t stream);
/* Recurrence kernel selection: 0 = auto (bf16 activations with H = 256 -> tcgen05 kernels, lstm_tc.cu), 1 = CUDA cores
 * only, 2 = same as 0, 3 = register-resident warp-MMA kernels (lstm_rg.cu: measured on a par with the tcgen05 kernels,
 * profiles/r02_blstm_experiments.txt; kept selectable, not the default).  Returns the previous mode (test / profiling hook). */
int b200st_set_blstm_backend(int mode);
/* Layout of the saved state (`acts`, `cs`) that b200st_blstm_fwd writes and b200st_blstm_bwd reads for this dtype / H
 * under the current backend: 0 = plain acts [2, T, B, 4H], cs [2, T, B, H] (fp32);  1 = kernel-private blocked layout
 * of the register-resident kernels, sized for B rounded up to a multiple of 16:
 *   acts [2][T][B/16][rank 8][thread 256][gate 4][seq 2], cs [2][T][B/16][rank 8][thread 256][seq 2]   (fp32)
 * thread = (n-tile nt * 4 + unit block ub) * 32 + r * 4 + c  <->  unit 32 rank + 8 ub + r, sequence 16 grp + 8 nt + 2 c + seq.
 * The saved state is opaque to callers (Enc.py keeps nothing comparable); the layout is documented for the tests. */
int b200st_blstm_saved_layout(int dtype, int64_t H);
/* dout in the same layout as out; dgates [2][T][B][4H] dtype written (zeros for t>=len). */
int b200st_blstm_bwd(int dtype, const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair,
                     const float* acts, const float* cs, const float* w_hh_f, const float* w_hh_r,
                     const int32_t* lens, void* dgates, int64_t T, int64_t B, int64_t H,
                     b200st_stream_t stream);

/* ---- On-device beam search of Seq2seq._step_translate (Seq2seq.py:337-393) --------------------------------------------
 * topk_logsoftmax: score[r, 0..k) = the k largest log_softmax(x[r, :]) values (fp32, descending; ties: lower index first),
 * pred[r, 0..k) their column indices; one pass for the log-sum-exp, k <= 8.  Replaces log_softmax + topk (Seq2seq.py:254-257). */
int b200st_topk_logsoftmax(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t k, float* score,
                           int64_t* pred, b200st_stream_t stream);
/* beam_select: one decode position of the beam bookkeeping for n_utt utterances x k hypotheses (hypothesis h = u * k + m):
 *   first != 0 (position 1): scores[h] += cand_score[u*k, m]; token = cand_pred[u*k, m]           (Seq2seq.py:349-356)
 *   else: candidates (r, j): (scores[u*k+r] + (eos[u*k+r] ? (j == 0 ? 0 : -1e9) : cand_score[u*k+r, j])) / len_map[u*k+r]^penalty;
 *         the k best (ties: lower r*k+j first) become the new hypotheses: scores[h] = value * len_map[h]^penalty (slot h's own
 *         length: the reference's rule), token = cand_pred[u*k+r, j], and positions [0, pos) of preds / anc / tokmask of slot h
 *         are copied from slot u*k+r                                                               (Seq2seq.py:358-383)
 *   then preds[h, pos] = token; eos[h] |= token == EOS; len_map[h] += !eos[h]; *n_done = number of finished hypotheses.
 * preds int64 [n_hyp, ld_preds]; anc int32 [>= pos, n_hyp] KV-cache ancestry table or NULL; tokmask uint8 [n_hyp, ld_tok];
 * done_u int32 [n_utt] and ticket u32 [1] (zero once) are scratch. */
int b200st_beam_select(float* scores, const float* cand_score, const int64_t* cand_pred, uint8_t* eos, float* len_map,
                       float penalty, int64_t pos, int first, int64_t* preds, int64_t ld_preds, int32_t* anc,
                       uint8_t* tokmask, int64_t ld_tok, int64_t k, int64_t n_utt, int32_t* done_u, void* ticket,
                       int64_t* n_done, b200st_stream_t stream);

/* ---- Persistent LAS decoder loop, forward (Dec.forward / forward_step / decode, Dec.py:130-233,320-438) ------------
 * ONE launch runs all S decode steps (3 uni-LSTM layers -> bilinear attention -> acous_ffn -> vocabulary projection ->
 * arg-max feedback and the EOS/PAD length rule) for bf16 activations with decoder width 512 and 512-wide keys/values;
 * replaces S x ~9 dependent launches of b200st_gemm / lstm_cell_fwd / las_attn_fwd / argmax_rows.  `args` is a HOST array
 * of `n_args` = 42 int64 slots holding device pointers and sizes, in this order:
 *   0 wk bf16[B,Tk,512] projected keys   1 enc bf16[B,Tk,512] values   2 klens int32[B]|0
 *   3 gx0 bf16: free running [V,2048] = E W_ih0[:, :E]^T + b0 (row gathered by the fed-back token), teacher forced [S,B,2048]
 *   4..15 per layer i = 0,1,2: wx_i bf16 (layer 0: W_ih0[:, E:]), row stride of wx_i, whh_i bf16[2048,512], bias_i f32[2048]|0
 *   16 w_ffn bf16[512,1024]   17 w_out bf16[V,512]   18 b_out f32[V]
 *   19 CV bf16[S+1,B,512] (row 0 zero)   20..28 per layer: H_i bf16[S+1,B,512] (row 0 zero), C_i f32[S+1,B,512] (row 0 zero),
 *   ACT_i f32[S,B,2048]   29 RES1 bf16[S,B,512]   30 CTX bf16[S,B,512]   31 PROBS f32[S,B,Tk]   32 LOGITS bf16[S,B,V]|0
 *   33 SYM int64[S,B]   34 lengths int32[B] (pre-set to S+1)   35 best u64[2,B] scratch   36 barrier u32[1] scratch
 *   37 B   38 Tk   39 S   40 V   41 teacher forcing (0 | 1)
 * Needs 128 co-resident CTAs (one per SM). */
int b200st_las_decoder_fwd(const int64_t* args, int64_t n_args, b200st_stream_t stream);
/* Debug aid: register an int64 [S * 16 + 1] device buffer; CTA 0 then records %globaltimer (ns) at the phase boundaries of
 * every decode step ([s][0] start, [1..3] after LSTM layer 0..2 + barrier, [7] end of attention, [4] after its barrier,
 * [5] after acous_ffn + barrier, [6] after the vocabulary phase + barrier; [8..11] inside the layer-1/2 hand-over: after the fresh-half GEMM, the cell, the barrier arrive, the recurrent-half GEMM; [S*16] kernel start).  NULL switches it off. */
int b200st_las_decoder_timeline(void* buf);

/* ---- LAS bilinear attention step (attention.py:190-193,250-273; Dec.py:423-425) ------------------
 * score[b,j] = q[b] . wk[b,j]; j >= klens[b] -> -1e12; softmax; ctx[b] = sum_j p[b,j] vals[b,j]. */
int b200st_las_attn_fwd(int dtype, const void* q, const void* wk, const void* vals,
                        const int32_t* klens, void* ctx, float* probs, int64_t B, int64_t Tk,
                        int64_t D, int64_t Dv, b200st_stream_t stream);
/* dscore [B,Tk] fp32 out; dq [B,D] dtype out.  (d wk / d vals are batched GEMMs over saved stacks.) */
int b200st_las_attn_bwd(int dtype, const void* dctx, const void* wk, const void* vals,
                        const float* probs, float* dscore, void* dq, int64_t B, int64_t Tk,
                        int64_t D, int64_t Dv, b200st_stream_t stream);
/* Key / value gradients of the attention over all S decode steps in one launch (attention.py:203-289 in reverse):
 * out[b, t, :] = sum_s w[s, b, t] * x[s, b, :];  w fp32 [S, B, Tk] (dscore -> d(W k), probs -> d values), x [S, B, D]
 * (decoder outputs / context gradients), out [B, Tk, D] in `dtype`.  Replaces two batched thin-K GEMMs (+ two casts). */
int b200st_las_stack_grad(int dtype, const float* w, const void* x, void* out, int64_t S, int64_t B, int64_t Tk, int64_t D,
                          b200st_stream_t stream);
/* argmax over the last dim (Dec.py:331 topk(1), Seq2seq.py:255 topk(1)); first index wins ties. */
int b200st_argmax_rows(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols,
                       int64_t* idx, int64_t idx_stride, b200st_stream_t stream);
/* arg-max + the Dec.decode lengths rule below in one launch (lengths may be NULL). */
int b200st_argmax_rows_lengths(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                               int64_t idx_stride, int32_t* lengths, int step, b200st_stream_t stream);
/* Same, and (table != NULL) the chosen token's embedding row table[idx] (fp32 [cols, dim]) is written to
 * emb[r * ld_emb ..] in `dtype`: the free-running LAS decoder's "feed the arg-max back" (Dec.py:331-341) in one launch. */
int b200st_argmax_rows_embed(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                             int64_t idx_stride, int32_t* lengths, int step, const float* table, void* emb,
                             int64_t ld_emb, int64_t dim, b200st_stream_t stream);
/* Same, plus a second gather in the activation dtype: out2[r * ld_out2 ..] = table2[idx[r]] ([cols, dim2], `dtype`).  The
 * LAS decoder passes table2 = E W_ih0[:, :E]^T + b (made once per forward): the next step's first-layer gate
 * contribution of the fed-back token (Dec.py:383,393-401) becomes a row gather instead of a GEMM on the step's chain. */
int b200st_argmax_rows_embed2(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                              int64_t idx_stride, int32_t* lengths, int step, const float* table, void* emb,
                              int64_t ld_emb, int64_t dim, const void* table2, void* out2, int64_t ld_out2,
                              int64_t dim2, b200st_stream_t stream);
/* Dec.decode lengths rule (Dec.py:334-340) kept on device: if sym in {EOS,PAD} and lengths[b] > step
 * then lengths[b] = step + 1. */
int b200st_las_update_lengths(const int64_t* sym, int64_t sym_stride, int32_t* lengths, int step,
                              int64_t B, b200st_stream_t stream);

/* ---- embeddings and the embedding-passing mix (Seq2seq.py:183-211; Dec.py:166,223) ---------------- */
int b200st_embedding_fwd(int dtype, const int64_t* ids, const float* table, void* out, int64_t ld_out,
                         int64_t n, int64_t dim, int64_t vocab, b200st_stream_t stream);
/* dtable[ids[i]] += dout[i] for ids[i] != padding_idx (nn.Embedding(padding_idx=PAD), Seq2seq.py:106). */
int b200st_embedding_bwd(int dtype, const int64_t* ids, const void* dout, int64_t ld_dout,
                         float* dtable, int64_t n, int64_t dim, int64_t vocab, int64_t padding_idx,
                         b200st_stream_t stream);
/* cat[i] = [ table[ids[i]] , dyn[i] ]  (Seq2seq.py:188-191) written in `dtype`, row width E + D. */
int b200st_mix_gather_concat(int dtype, const int64_t* ids, const float* table, const void* dyn,
                             int64_t ld_dyn, void* cat, int64_t n, int64_t E, int64_t D, int64_t vocab,
                             b200st_stream_t stream);

/* ---- softmax / loss over the vocabulary (Seq2seq.py:254-255; Dec.py:436; loss.py:130-132) --------- */
int b200st_log_softmax_fwd(int dtype, const void* x, void* y, int64_t rows, int64_t cols,
                           int64_t* argmax, b200st_stream_t stream);
int b200st_log_softmax_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t rows,
                           int64_t cols, b200st_stream_t stream);
/* loss_sum += sum_{mask} -logp[r, target[r]];  ld = row stride of logp. */
int b200st_masked_nll_fwd(int dtype, const void* logp, int64_t ld, const int64_t* target,
                          const uint8_t* mask, float* loss_sum, int64_t rows, int64_t cols,
                          b200st_stream_t stream);
/* dlogp (dense [rows, cols], ld) = 0 except dlogp[r, target[r]] = -gscale[0] where mask. */
int b200st_masked_nll_bwd(int dtype, const float* gscale, const int64_t* target, const uint8_t* mask,
                          void* dlogp, int64_t ld, int64_t rows, int64_t cols, b200st_stream_t stream);
/* Fused softmax + masked NLL + gradient (K17): one read of logits, one write of dlogits.
 * loss_sum += sum_mask (lse - logit[target]); dlogits = (softmax - onehot) * mask * scale[0];
 * eps = label smoothing (0 for parity with the reference, which has none).  dlogits may alias logits. */
int b200st_softmax_nll_fused(int dtype, const void* logits, int64_t ld, const int64_t* target,
                             const uint8_t* mask, const float* scale, float eps, float* loss_sum,
                             void* dlogits, int64_t ld_d, int64_t rows, int64_t cols,
                             b200st_stream_t stream);

/* ---- small glue kernels ------------------------------------------------------------------------ */
int b200st_add(int dtype, const void* a, const void* b, void* out, int64_t n, b200st_stream_t stream);
/* out[b,l,:] = x[b,l,:] + pe[l,:]  (TFEnc.py:82-83, TFDec.py:85-86) */
int b200st_add_posenc(int dtype, const void* x, const float* pe, void* out, int64_t B, int64_t L,
                      int64_t D, b200st_stream_t stream);
/* in [A,Bd,C] -> out [Bd,A,C] */
int b200st_transpose01(int dtype_in, int dtype_out, const void* in, void* out, int64_t A, int64_t Bd,
                       int64_t C, b200st_stream_t stream);
int b200st_cast(int dtype_in, int dtype_out, const void* in, void* out, int64_t n,
                b200st_stream_t stream);
/* out[c] (+)= sum_r x[r, c] */
int b200st_colsum(int dtype, const void* x, int64_t ld, float* out, int64_t rows, int64_t cols,
                  int accumulate, b200st_stream_t stream);
int b200st_relu_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t n,
                    b200st_stream_t stream);
/* mask[b, i, j] = (ids[b, j] != pad) && (!causal || j <= i), uint8 [B, Lq, L]; Lq = causal ? L : 1
 * (Seq2seq.py:204-205 / layers.py:269-289). */
int b200st_token_mask(const int64_t* ids, uint8_t* mask, int64_t B, int64_t L, int64_t pad, int causal,
                      b200st_stream_t stream);
/* mask[b, 0, j] = j < lengths[b]   (Seq2seq.py:494-497) */
int b200st_length_mask(const int32_t* lengths, uint8_t* mask, int64_t B, int64_t L,
                       b200st_stream_t stream);

/* ---- dropout (nn.Dropout call sites: Enc.py:159-212, Dec.py:166,386-429, Seq2seq.py:195-209, layers.py:182-249) --
 * y[r, c] = x[r, c] * keep / (1 - p) (+ residual[r, c]) over a rows x cols view (row strides ldx / ldr / ldy).
 * keep is a pure function of (rng[0] = seed, rng[1] = step, site, element index r * ld_mask + col_off + c):
 * Philox4x32-10 with key {seed_lo, seed_hi ^ step_hi}, counter {index/4 (64 bit), site, step_lo}, word index%4,
 * keep iff word >= p * 2^32.  rng is a 2 x int64 DEVICE array; b200st_rng_advance adds 1 to rng[1] (launched once per
 * forward pass, so a replayed CUDA graph draws new masks).  Backward = the same call on the gradient (same site). */
int b200st_dropout(int dtype, const void* x, int64_t ldx, const void* residual, int64_t ldr, void* y, int64_t ldy,
                   int64_t rows, int64_t cols, int64_t ld_mask, int64_t col_off, float p, const int64_t* rng,
                   int64_t site, b200st_stream_t stream);
int b200st_rng_advance(int64_t* rng, b200st_stream_t stream);

/* ---- input stage: Dataset.load_acous_from_flis (utils/dataset.py:155-184) on the device -----------------------------
 * packed: fp32 [sum(lens), F], the batch's utterances back to back (no padding crosses PCIe); offsets int64 [B] = first
 * frame of utterance i in `packed`; lens int32 [B]; mu, sd: fp32 [B, F] per-UTTERANCE rows of the speaker statistics
 * (dataset.py:134-153) or both NULL (acous_norm off).  out fp32 [B, T_pad, F]:
 *   out[i, t, :] = (packed[offsets[i] + t, :] - mu[i]) / sd[i]   for t < lens[i]     (IEEE division, dataset.py:173)
 *   out[i, t, :] = 0                                             otherwise            (pad_sequence, dataset.py:178-182)
 * with T_pad = max(lens) + 8 - max(lens) % 8 chosen by the caller (dataset.py:179). */
int b200st_fbank_norm_pad(const float* packed, const int64_t* offsets, const int32_t* lens, const float* mu,
                          const float* sd, float* out, int64_t B, int64_t T_pad, int64_t F, b200st_stream_t stream);

/* ---- fused gradient-norm clip + Adam over all parameter tensors -----------------------
 * Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step() behind Optimizer.step()
 * (modules/optim.py:31-36; constructed at trainer/trainer_base.py:422-426).  `table` is a device array
 * int64 [n_tensors][b200st_opt_table_cols() = 7] = {param, grad, exp_avg, exp_avg_sq (fp32 device pointers), numel, and up
 * to two bf16 destinations (or 0) that receive the updated parameter -- the operand copies the tensor-core GEMMs read, so
 * no cast pass follows the step};
 * `blockmap` is int32 [n_blocks][2] = {tensor index, chunk index}, one CTA per chunk of b200st_opt_chunk() elements.
 * multi_sqnorm writes one partial sum of squares per block; adam_prepare (one CTA, fixed summation order) turns them
 * into scal[4] = {clip coefficient = min(1, max_grad_norm / (||g|| + 1e-6)), lr[0] / (1 - beta1^t), sqrt(1 - beta2^t),
 * ||g||} and advances the device step counter t (fp32 scalar, the optimizer state's 'step'); partials == NULL or
 * max_grad_norm <= 0 disables clipping.  multi_adam applies torch.optim.Adam's update (amsgrad off, L2 weight decay)
 * with the gradient scaled by the clip coefficient on the fly.  No host synchronisation: graph-capturable. */
int b200st_opt_chunk(void);
int b200st_opt_table_cols(void);
int b200st_multi_sqnorm(const int64_t* table, const int32_t* blockmap, int64_t n_blocks, float* partials,
                        b200st_stream_t stream);
int b200st_adam_prepare(const float* partials, int64_t n_blocks, float max_grad_norm, const float* lr, double beta1,
                        double beta2, float* step, float* scal, b200st_stream_t stream);
int b200st_multi_adam(const int64_t* table, const int32_t* blockmap, int64_t n_blocks, const float* scal, double beta1,
                      double beta2, double eps, double weight_decay, b200st_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ST_H_ */
