/*
 * v2m_b200 -- C ABI of the B200-native (sm_100a) kernels for the Affective Multimodal
 * Transformer hot path of khangklj/Video2Music.
 *
 * The reference has no FFI boundary: its hot path is a set of PyTorch nn.Modules calling ATen
 * (SURVEY.md section 8b).  This header is therefore the boundary a maintainer of the reference
 * would bind (ctypes / cffi / a 20-line torch extension) to replace the ATen op sequences listed
 * next to each entry point; INTEGRATION.md shows that binding.  Conventions:
 *   - plain C, raw DEVICE pointers, explicit sizes/strides, caller-owned outputs and workspaces;
 *   - no allocation, no global state, re-entrant per stream; `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success or a V2M_* status; v2m_last_error() gives the text
 *     (thread-local); the Python host turns it into RuntimeError like the reference's asserts /
 *     raises (rpr.py:242-247, grouped_query_attention.py:58-66);
 *   - dtype codes: 0 = float32, 1 = bfloat16.
 * Paths relative to the reference root are cited as file:line.
 */
#ifndef V2M_B200_H_
#define V2M_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2M_OK 0
#define V2M_BAD_ARG 1
#define V2M_CUDA_ERROR 2
#define V2M_UNSUPPORTED 3
#define V2M_NO_DEVICE 4

#define V2M_F32 0
#define V2M_BF16 1

int v2m_abi_version(void);
const char* v2m_last_error(void);
/* sizeof() of the ABI structs as compiled: 0 v2m_epilogue, 1 v2m_attn, 2 v2m_dec_layer, 3 v2m_decode, 4 v2m_attn_bwd_t (binding self-check). */
int64_t v2m_struct_size(int32_t which);
/* 1 if the current CUDA device is compute capability 10.x (the only target), else 0. */
int v2m_device_ok(void);

/* ---- dense linear layers ---------------------------------------------------------------------
 * C[M,N] = epilogue(A[M,K] * W[N,K]^T); replaces torch.nn.functional.linear at
 * model/rpr.py:253,263,277,417, model/rpr.py:67 (FFN), model/video_music_transformer.py:1001,1022,1042
 * and the stock nn.TransformerEncoderLayer linears created at video_music_transformer.py:967-971.
 * Epilogue, in order: +bias[n]; *alpha for n < alpha_cols (q scaling, rpr.py:328); ReLU (rpr.py:67);
 * + row_scale[m]*col_vec[n] (key column of Linear_chord, video_music_transformer.py:991-1001);
 * + residual[(m % res_mod)*ldr + n] (residual add rpr.py:59-68 or positional encoding
 * positional_encoding.py:22 when res_mod = sequence length).
 * head_scatter != 0 writes element (m,n) into the head-major KV cache layout
 * [part][batch][head][cap][dh] (see kernels.h) instead of C[m*ldc+n].                              */
typedef struct v2m_epilogue {
  const float* bias;
  const void* residual;     /* fp32, or bf16 when residual_bf16 */
  int32_t ldr, res_mod;
  const float* row_scale;
  const float* col_vec;
  float alpha;
  int32_t alpha_cols;
  int32_t relu;
  int32_t residual_bf16;
  int32_t head_scatter, S, H, dh, cap, pos0;
  int64_t part_stride;
  /* Training-time inverted dropout fused into the epilogue (nn.Dropout after a linear layer, rpr.py:58-69 and the stock
   * encoder layer; PositionalEncoding dropout, positional_encoding.py:21-23).  drop_scale = 1/(1-p), 0 = off; element (m, n)
   * is kept iff byte (n & 3) of mix(drop_seed, m, n / 4) >= drop_thresh = round(p * 256), drop_scale = 256 / (256 - drop_thresh)
   * (stateless: v2m_dy_prep recomputes the same mask);
   * drop_after_res: y = drop(acc + residual) instead of drop(acc) + residual.  Both GEMM paths (fp32 SIMT, bf16 tcgen05). */
  float drop_scale;
  uint32_t drop_thresh, drop_seed;
  int32_t drop_after_res;
  const uint32_t* drop_seed_dev;   /* optional device counter added to drop_seed: fresh masks when a CUDA graph is replayed */
  /* bf16 GEMMs with an fp32 output and no other epilogue operation: C += A W^T instead of C = A W^T (coalesced vector
   * reductions into the caller's buffer) -- weight gradients accumulate straight into the optimiser's gradient buffer. */
  int32_t accumulate;
  /* bf16 GEMMs: `residual` (bf16) is a GATE instead of an addend: C = residual[m][n] > 0 ? acc * gate_scale : 0 -- the backward of
   * relu(linear1(x)) followed by dropout (rpr.py:67, nn.TransformerEncoderLayer) applied by the dX GEMM of linear2: the gate is the
   * saved hidden activation, gate_scale the dropout scale 1 / (1 - p) (1 without dropout). */
  int32_t residual_gate;
  float gate_scale;
  int32_t pad_;
} v2m_epilogue;

/* fp32 SIMT GEMM (exact path, batch-invariant summation order). */
int v2m_gemm_f32(const float* A, int32_t lda, const float* W, int32_t ldw, float* C, int32_t ldc,
                 int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream);
/* bf16 tcgen05/TMEM/TMA GEMM; C is bf16 (out_dtype 1) or fp32 (0); lda, ldw multiples of 8. */
int v2m_gemm_bf16(const void* A, int32_t lda, const void* W, int32_t ldw, void* C, int32_t ldc, int32_t out_dtype,
                  int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream);

/* General tcgen05 GEMM: a_mn / b_mn != 0 mean the operand is stored transposed ([K, M] resp. [K, N] row-major) and is fed
 * to the tensor core as an MN-major operand -- backward pass: dX = dY W (b_mn), dW = dY^T X (a_mn, b_mn).           */
int v2m_gemm_bf16_general(const void* A, int32_t lda, int32_t a_mn, const void* W, int32_t ldw, int32_t b_mn, void* C, int32_t ldc,
                          int32_t out_dtype, int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream);

/* ---- attention ---------------------------------------------------------------------------------
 * out = softmax(q k^T + skew(q Er^T) + causal mask) v for every (batch, head); replaces
 * model/rpr.py:387-414 (bmm, einsum with Er, _skew :439-455, mask, softmax, bmm), the stock
 * nn.MultiheadAttention cores (rpr.py:62, encoder layers) when Er == NULL, and
 * scaled_dot_product_gqa (model/grouped_query_attention.py:122-156) when Hkv < Hq.
 * Element (b,l,h,d) of q is q[b*q_sb + l*q_sl + h*dh + d] (same for k, v, o).
 * lse (optional) [B*Hq, Lq]; p_out (optional) [B*Hq, Lq, Lk] = the probabilities (need_weights).  */
typedef struct v2m_attn {
  const void* q; const void* k; const void* v; void* o;
  int64_t q_sb, q_sl, k_sb, k_sl, v_sb, v_sl, o_sb, o_sl;
  int32_t B, Hq, Hkv, Lq, Lk, dh;
  int32_t causal;
  const void* Er; int32_t er_len;
  float q_scale;
  float* lse; float* p_out;
  float drop_scale; uint32_t drop_thresh, drop_seed;   /* dropout of the probabilities (training, fp32 and bf16 kernels): see v2m_epilogue */
  const uint32_t* drop_seed_dev;
  const int32_t* lk_dev;   /* optional device word (fp32 kernel, no Er / p_out): only the first min(Lk, *lk_dev) keys exist -- one captured
                              CUDA graph then serves every position of a KV-cached generation (cached_decode.py)             */
} v2m_attn;
int v2m_attn_fwd(const v2m_attn* p, int32_t dtype, void* stream);

/* fp32 GEMM with explicit element strides: A(m,k) = A[m*a_rs + k*a_cs], W(n,k) = W[n*w_rs + k*w_cs]; used by the
 * backward pass (dX = dY W, dW = dY^T X) without materialising transposes.                                        */
int v2m_gemm_f32_strided(const float* A, int32_t a_rs, int32_t a_cs, const float* W, int32_t w_rs, int32_t w_cs, float* C,
                         int32_t ldc, int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream);

/* ---- backward pass of the training step (run_model_vevo.py:84-124; the reference relies on torch autograd) ------
 * v2m_attn_bwd: dQ, dK, dV, dEr of v2m_attn_fwd (needs its lse).  dk / dv / dEr are caller-zeroed fp32 accumulators
 * (dk, dv addressed as [b*dkv_sb + j*dkv_sl + hkv*dh + d]); dq has the dtype of q.                                */
typedef struct v2m_attn_bwd_t {
  const void* q; const void* k; const void* v; const void* o; const void* dO; const float* lse; const void* Er;
  void* dq; float* dk; float* dv; float* dEr;
  int64_t q_sb, q_sl, k_sb, k_sl, v_sb, v_sl, o_sb, o_sl, do_sb, do_sl, dq_sb, dq_sl, dkv_sb, dkv_sl;
  int32_t B, Hq, Hkv, Lq, Lk, dh, causal, er_len, dtype;
  float q_scale;
  float drop_scale; uint32_t drop_thresh, drop_seed;   /* the forward's probability dropout (v2m_attn_bwd and v2m_attn_bwd_tc) */
  const uint32_t* drop_seed_dev;
  float dq_scale;   /* dq is stored times dq_scale (0 = 1): backward of a query scaling applied by the projection's epilogue */
} v2m_attn_bwd_t;
int v2m_attn_bwd(const v2m_attn_bwd_t* p, void* stream);
/* Tensor-core attention backward (bf16, head_dim 64, q pre-scaled: q_scale must be 1): same gradients as v2m_attn_bwd from
 * mma.sync kernels (csrc/attn_bwd_tc.cu).  Here a->dk / a->dv are BF16 outputs (strides dkv_sb / dkv_sl in elements) that
 * are written once -- no zeroing, no accumulation; a->dEr (fp32 [er_len, 64]) must be zeroed by the caller.
 * ws: v2m_attn_bwd_tc_workspace(B, Hq, Lq, Lk, has_er) bytes of device scratch (P, dS and skewed dS tiles, bf16). */
int64_t v2m_attn_bwd_tc_workspace(int32_t B, int32_t Hq, int32_t Lq, int32_t Lk, int32_t has_er);
int v2m_attn_bwd_tc(const v2m_attn_bwd_t* a, void* ws, int64_t ws_bytes, void* stream);
/* dz = dy * relu'(y) * (n < alpha_cols ? alpha : 1) * dropmask(m, n), db[n] += sum_m dz[m][n]: gradient of the fused linear
 * epilogue (drop_scale = 0: no dropout; otherwise the mask of v2m_epilogue with the same seed / threshold). */
int v2m_dy_prep(const void* dy, int32_t dy_dtype, int64_t ld_dy, const void* y, int32_t y_dtype, int64_t ld_y, int32_t relu,
                float alpha, int32_t alpha_cols, void* dz, int32_t dz_dtype, int64_t ld_dz, float* db, int32_t M, int32_t N,
                float drop_scale, uint32_t drop_thresh, uint32_t drop_seed, const uint32_t* drop_seed_dev, void* stream);
int v2m_layernorm_bwd(const void* x, int32_t x_dtype, const float* gamma, const void* dy, int32_t dy_dtype, void* dx,
                      int32_t dx_dtype, float* dgamma, float* dbeta, int32_t M, int32_t D, float eps, void* stream);
int v2m_embed_bwd(const int64_t* idx, const void* d, int32_t d_dtype, int64_t ld_d, float* dtable, int32_t rows, int32_t D,
                  void* stream);
/* 0.4*CrossEntropy(label_smoothing, ignore_index) + 0.6*BCEWithLogits (run_model_vevo.py:101-119, train.py:222,233):
 * scratch3 = [sum CE over valid rows, sum BCE over elements, #valid rows]; dlogits = d(total loss)/d(logits).      */
/* Evaluation metrics, dataset/vevo_dataset.py:653-701 (compute_vevo_accuracy, compute_hits_k) for a whole batch in one
 * launch: counters[5] (int32, device) = {rows with tgt != pad, argmax hits, hits@k0, hits@k1, hits@k2}; ties go to the
 * lower class index. */
int v2m_amt_metrics(const float* logits, const int64_t* tgt, int32_t R, int32_t Cn, int64_t pad, int32_t k0, int32_t k1, int32_t k2,
                    int32_t* counters, void* stream);
/* Emotion correspondence of the evaluation loop (compute_vevo_correspondence, dataset/vevo_dataset.py:747-810; called at
 * utilities/run_model_vevo.py:307,415 with EMOTION_THRESHOLD = 0.8): logits [R, Cn], tgt_emotion [R, Ce] (columns 0..13 = chord
 * qualities, last column = padding flag), tgt_emotion_prob [R]; counters (int32[2], zeroed by the call) = {pt, right}:
 * correspondence = right / pt, -1 when pt == 0.  chord_end = 157: END / PAD predictions are never right.                      */
int v2m_amt_correspondence(const float* logits, const float* tgt_emotion, const float* tgt_emotion_prob, int32_t R, int32_t Cn, int32_t Ce,
                           float threshold, int32_t chord_end, int32_t* counters, void* stream);
int v2m_amt_loss(const float* logits, const int64_t* tgt, const float* tgt_emotion, int32_t R, int32_t Cn, int64_t ignore,
                 float smooth, float w_ce, float w_bce, float* scratch3, float* dlogits, const float* norm_in, void* stream);
/* out1[0] = number of targets != ignore (the CE normaliser, nn.CrossEntropyLoss(ignore_index), train.py:222).  A data-parallel
 * trainer all-reduces it and hands {count / world, global rows / world} back to v2m_amt_loss as norm_in (2 device floats), so
 * that the mean of the ranks' gradients is the gradient of the reference's single-process global-batch loss. */
int v2m_count_valid(const int64_t* tgt, int32_t R, int64_t ignore, float* out1, void* stream);
/* torch.optim.Adam step on flat fp32 buffers (train.py:237-238; eps ADAM_EPSILON = 10e-9 = 1e-8, utilities/constants.py:91).
 * weight_decay > 0: torch.optim.AdamW's decoupled decay p *= 1 - lr * weight_decay (train.py:239-240, the CLI default optimiser,
 * utilities/argument_funcs.py:17).  grad_scale multiplies g first (1/world_size after all-reduce). */
/* dyn (optional, device {lr, 1-b1^t, 1-b2^t}) overrides the scalar arguments at run time (CUDA-graph replay with a schedule);
 * p16 (optional) receives the bf16 mirror of the updated parameters; zero_grad clears g; ctr (optional) += 1 per call. */
int v2m_adam_step(float* p, float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float weight_decay, int32_t step,
                  float grad_scale, const float* dyn, void* p16, int32_t zero_grad, uint32_t* ctr, void* stream);

/* ---- one generation step of the generic decoder stacks (BASELINE config 4: grouped-query attention + MoE feed-forward) ----------
 * The reference generates with one FULL forward per token (video_music_transformer.py:227-315); the KV-cached step
 * (video2music_b200/cached_decode.py) pushes ONE new row per video through the layers, so its linear layers have M <= 64 rows.
 * v2m_step_linear_f32: y[M,N] = act(x[M,K] W[N,K]^T + bias[n] + row_scale[m] * col_vec[n]) -- F.linear of
 *   grouped_query_attention.py:306-309,350-356, custom_transformer.py (in_proj / out_proj), video_music_transformer.py:240-262
 *   (Linear_chord with its key column, Wout); fixed k order, result independent of M.
 * v2m_step_attn_f32: o[b, hq*64:] = softmax(q_scale * q[b, hq] . K[b, :n, hkv]) V[b, :n, hkv], hkv = hq / (Hq / Hkv), one query row
 *   per (video, query head) over the first n = min(n_max, *n_dev) cached rows (n_dev optional, device word: the launch is the same
 *   for every position, so a generation replays ONE captured CUDA graph); element (b, j, h, d) of a cache is k[b*kv_sb + j*kv_sl + h*64 + d]
 *   -- scaled_dot_product_gqa (grouped_query_attention.py:93-159) and F.multi_head_attention_forward with a single query.     */
int v2m_step_linear_f32(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias, const float* row_scale,
                        const float* col_vec, float* y, int64_t ldy, int32_t M, int32_t N, int32_t K, int32_t relu, void* stream);
int v2m_step_attn_f32(const float* q, int64_t q_sb, const float* k, const float* v, int64_t kv_sb, int64_t kv_sl, float* o, int64_t o_sb,
                      int32_t B, int32_t Hq, int32_t Hkv, int32_t dh, int32_t n_max, const int32_t* n_dev, float q_scale, void* stream);

/* ---- residual + LayerNorm (rpr.py:59-69; nn.LayerNorm eps) ------------------------------------ */
int v2m_layernorm(const void* x, int32_t x_dtype, const void* res, int32_t res_dtype, const float* gamma,
                  const float* beta, void* y, int32_t y_dtype, void* y2, int32_t y2_dtype, int32_t M, int32_t D,
                  float eps, void* stream);

/* ---- input staging (video_music_transformer.py:984-1018) -------------------------------------- */
int v2m_embed_sum(const int64_t* idx_a, const float* table_a, const int64_t* idx_b, const float* table_b, void* out,
                  int32_t out_dtype, int32_t ld_out, int32_t rows, int32_t D, void* stream);
int v2m_concat_features(const float* sem, int32_t sem_dim, const float* scene, const float* motion, int32_t motion_dim,
                        const float* emotion, int32_t emo_dim, void* out, int32_t out_dtype, int32_t ld_out,
                        int32_t rows, void* stream);
int v2m_cast_2d(const void* src, int32_t src_dtype, int64_t ld_src, void* dst, int32_t dst_dtype, int64_t ld_dst,
                int32_t rows, int32_t cols, int32_t zero_pad, void* stream);

/* mode 0: out = a * silu(b) (GLUExpert gating, moe.py:47); mode 1: out = a + alpha*b (shared expert, moe.py:301);
 * mode 2: out = sigmoid(a) (instrument classifier head, video_regression.py:197-200; b is read but unused);
 * mode 3: out = a * b * (1 - b) (gradient of mode 2: a = upstream gradient, b = the sigmoid output) */
int v2m_binary_f32(const float* a, const float* b, float* out, int64_t n, int32_t mode, float alpha, void* stream);

/* ---- KV-cached greedy decode (replaces the re-forward loop of VideoMusicTransformer.generate,
 * video_music_transformer.py:1069-1084 with beam=1).  All pointers are device pointers owned by the
 * caller; caches are head-major [B][H][cap|S][64] in `dtype`.                                      */
#define V2M_MAX_DEC_LAYERS 8
typedef struct v2m_dec_layer {
  const void* w_qkv; const float* b_qkv;
  const void* w_so;  const float* b_so;
  const void* w_cq;  const float* b_cq;
  const void* w_co;  const float* b_co;
  const void* w_f1;  const float* b_f1;
  const void* w_f2;  const float* b_f2;
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;
  const float* ln3_g; const float* ln3_b;
  const void* er;
  const void* er_sw;   /* v2m_decode_run_stream only: 8 swizzled copies of Er, see below */
  void* self_k; void* self_v;
  const void* cross_k; const void* cross_v;
} v2m_dec_layer;
typedef struct v2m_decode {
  int32_t dtype;
  int32_t B, H, E, FF, S, cap, n_layers, er_len, vocab, vocab_limit, primer_len, chord_embed;
  v2m_dec_layer layer[V2M_MAX_DEC_LAYERS];
  const float* lnf_g; const float* lnf_b;
  const void* w_out; const float* b_out;
  const float* emb_root; const float* emb_attr; const float* emb_chord;
  const void* w_chord; const float* wc_key; const float* b_chord;
  const float* pe;
  const float* key;
  int64_t* gen; int64_t* gen_root; int64_t* gen_attr;
  int32_t* step;
  float* h; void* r; float* qbuf; void* ctx; void* ff; float* logits; float* logits_all;
  void* xn;
  /* next-token rule: sample 0 = greedy arg-max (beam=1, video_music_transformer.py:1078-1084); 1 = the sampling branch
   * (:1085-1128): P(N)=0 if max_conseq_N==0, P(prev)=0 after max_conseq_chord equal tokens, inverse-CDF draw with
   * uniforms[b*cap + t+1]; gen_root / gen_attr of generated positions follow the chord id (closed form of the JSON maps). */
  int32_t sample, max_conseq_N, max_conseq_chord, pad_;
  const float* uniforms;
} v2m_decode;
int v2m_decode_run(const v2m_decode* p, int32_t n_steps, int32_t use_graph, void* stream);
/* Same loop as ONE persistent kernel launch (csrc/decode_stream.cu): a cluster of 8 CTAs (one per head) owns up to 8
 * videos for all n_steps positions starting at t0 (the host copy of *p->step).  *p is a SEPARATE parameter block:
 *  - the matrix pointers (w_qkv, w_so, w_cq, w_co, w_f1, w_f2, w_chord, w_out) reference FRAGMENT-PACKED bf16 copies:
 *    tiles of 16 output rows x 16 k in mma.m16n8k16 A-fragment order, [N/16][K/16][32 lanes][8], rows zero-padded to 16;
 *  - self_k / cross_k reference caches of 256-byte rows [K(64) | V(64)] per (video, head, position) whose 16-byte chunks
 *    sit at chunk position c ^ (position & 7) (v2m_kv_interleave builds the cross cache); self_v / cross_v are unused;
 *  - er_sw references 8 copies of Er, copy s storing row r with its chunks at c ^ ((r - s) & 7).
 * bf16, d_model 512, 8 heads, dim_feedforward 1024; V2M_UNSUPPORTED otherwise (callers then use v2m_decode_run).
 * timestamps (optional, device, ts_cap entries): globaltimer (ns) at every phase boundary of CTA 0 (measurement aid). */
int v2m_decode_run_stream(const v2m_decode* p, int32_t t0, int32_t n_steps, uint64_t* timestamps, int32_t ts_cap, void* stream);
/* bf16 tensor-core TRAINING path of the MoE experts (gradients of moe.py:44-49,191-199; torch autograd in the reference).
 * kgrouped: C[g] (M x N fp32, stride c_gstride) = A_g^T B_g over the rows [k_off[g], k_off[g+1]) of the row-major bf16 matrices A
 * [R, M] and B [R, N] -- the ragged dW_e = dY_e^T X_e with the group bounds (multiples of 64: moe_permute's 128-row aligned groups)
 * read on the device.  swiglu_pair_bwd: gradient of h = a1 * silu(g) w.r.t. the pair matrix a = [a1 | g] (M x 2 ff).
 * group_colsum: out[g][n] = column sums over group g (expert bias gradients). */
int v2m_gemm_bf16_kgrouped(const void* A, int32_t lda, const void* B, int32_t ldb, float* C, int32_t ldc, int64_t c_gstride, int32_t M,
                           int32_t N, int32_t R, int32_t n_groups, const int32_t* k_off, void* stream);
int v2m_swiglu_pair_bwd_bf16(const void* a, const void* dh, void* dag, int64_t M, int32_t ff, void* stream);
int v2m_moe_group_colsum_bf16(const void* x, int64_t ldx, const int32_t* off, int32_t n_groups, float* out, int32_t N, int32_t rows_hint,
                              void* stream);

/* out[row] = [k[row] | v[row]] (2 x 64 bf16) with the XOR swizzle above, row = (video*H + head)*S + position. */
int v2m_kv_interleave(const void* k, const void* v, void* out, int64_t rows, int32_t S, void* stream);
int64_t v2m_decode_launches_per_step(const v2m_decode* p);
/* Measurement aid: launches one decode kernel kind (0 self-attention, 1 cross-attention, 2 QKV GEMM, 3 FFN1 GEMM)
 * `reps` rounds over all layers on `stream`, reading the current step from *p->step. */
int v2m_decode_probe(const v2m_decode* p, int32_t kind, int32_t reps, void* stream);

/* ---- Mamba block, model/mamba.py:259-351 (fp32).  Row-major (B, L, ld) views with explicit leading dimensions. ----
 * conv: y = silu(conv1d(x, w, bias, groups=ED, padding=KW-1)[:, :, :L])  (mamba.py:270-276).
 * scan: delta = softplus(delta_raw + dt_bias); h = exp(delta A) h + delta B x; y = h.C + D x; out = y silu(z)
 *       (+ x (1 - sigmoid(silu z)) when plus != 0, mamba.py:283-287) -- the (B, L, ED, N) tensors of mamba.py:333-351 are
 *       never materialised.  rmsnorm: mamba.py:483-489. */
int v2m_mamba_conv_silu(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int32_t B, int32_t L,
                        int32_t ED, int32_t KW, void* stream);
/* Recurrent single-token step, MambaBlock.step / ssm_step (mamba.py:407-470): one new token per video against the cache
 * (h (B, ED, N) or NULL = zeros, inputs (B, ED, KW-1)).  step_conv: xs = silu(conv1d([in_old | x_new])[KW-1]) with x_new =
 * xz[b*ldxz + e], in_new = the window shifted by one (mamba.py:419-425,435).  step_ssm: delta = softplus(dt_proj(dbc[:, :R])),
 * h_new = exp(delta A) h_old + delta B x, y = h_new C + D x, out = y * silu(z) with dbc (B, R + 2N) = x_proj(xs) (mamba.py:437-465,
 * 427-430).  Neither kernel writes to its inputs. */
int v2m_mamba_step_conv(const float* xz, int64_t ldxz, const float* in_old, const float* w, const float* bias, float* xs, float* in_new,
                        int32_t B, int32_t ED, int32_t KW, void* stream);
int v2m_mamba_step_ssm(const float* xs, const float* dbc, int64_t lddbc, const float* dtw, const float* dtb, const float* A_log,
                       const float* D, const float* z, int64_t ldz, const float* h_old, float* h_new, float* out, int32_t B, int32_t ED,
                       int32_t N, int32_t R, void* stream);
int64_t v2m_selective_scan_workspace(int32_t B, int32_t L, int32_t ED, int32_t N);   /* bytes of `ws` (0 when L <= 64) */
int v2m_selective_scan_fwd(const float* x, int64_t ldx, const float* delta_raw, int64_t ldd, const float* dt_bias, const float* A_log,
                           const float* Bm, const float* Cm, int64_t ldbc, const float* D, const float* z, int64_t ldz, float* out,
                           int64_t ldo, int32_t B, int32_t L, int32_t ED, int32_t N, int32_t plus, float* ws, int64_t ws_bytes,
                           void* stream);
/* Backward of the two kernels above (what torch autograd derives from mamba.py:270-276 and from the materialised
 * deltaA / BX / pscan graph of mamba.py:333-351, pscan.py:196-226).
 * scan_bwd: `hs` = state workspace of v2m_selective_scan_bwd_workspace() bytes ([B][L][N][ED] fp32, recomputed inside);
 *   writes dx, ddelta_raw, dz (dz may be NULL when z is NULL); ADDS into the caller-zeroed dBm / dCm rows ([B*L][N] with
 *   leading dimension lddbc), dA_log [ED][N], dD [ED], ddt_bias [ED] (may be NULL).
 * conv_bwd: writes dx (the (B, L, lddx) view of the conv input gradient), ADDS into caller-zeroed dw [ED][KW], dbias [ED]. */
int64_t v2m_selective_scan_bwd_workspace(int32_t B, int32_t L, int32_t ED, int32_t N);
int v2m_selective_scan_bwd(const float* x, int64_t ldx, const float* delta_raw, int64_t ldd, const float* dt_bias, const float* A_log,
                           const float* Bm, const float* Cm, int64_t ldbc, const float* D, const float* z, int64_t ldz,
                           const float* dout, int64_t ldo, float* hs, int64_t hs_bytes, float* dx, int64_t lddx, float* ddelta_raw,
                           int64_t lddd, float* dBm, float* dCm, int64_t lddbc, float* dz, int64_t lddz, float* dA_log, float* dD,
                           float* ddt_bias, int32_t B, int32_t L, int32_t ED, int32_t N, int32_t plus, void* stream);
int v2m_mamba_conv_silu_bwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* dy, int64_t ldy, float* dx,
                            int64_t lddx, float* dw, float* dbias, int32_t B, int32_t L, int32_t ED, int32_t KW, void* stream);
int v2m_rmsnorm(const float* x, const float* w, float* y, int32_t M, int32_t D, float eps, void* stream);
/* gradient of v2m_rmsnorm (autograd over custom_transformer.py:27-47 / mamba.py:483-489): dx [M][D]; dw [D] is accumulated
 * into a caller-zeroed buffer (may be NULL). */
int v2m_rmsnorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int32_t M, int32_t D, float eps, void* stream);

/* ---- selective scan, model/pscan.py:154-226: H[t] = A[t]*H[t-1] + X[t] over (B,L,D,N) fp32 ---- */
int v2m_pscan_fwd(const float* A, const float* X, float* H, int32_t B, int32_t L, int32_t D, int32_t N, void* stream);
int v2m_pscan_bwd(const float* A, const float* H, const float* gH, float* gA, float* gX, int32_t B, int32_t L,
                  int32_t D, int32_t N, void* stream);

/* ---- MoE router, model/moe.py:180-190 / 244-288: gate GEMV + top-k + softmax + histogram ------
 * sel = (x Wg^T + bg) * inv_t_pre (+ sel_bias for selection only, moe.py:260-268);
 * weights = softmax(gathered logits * inv_t_post) in fp32 (moe.py:190,288).                        */
int v2m_moe_route(const float* x, const float* wg, const float* bg, const float* sel_bias, float inv_t_pre,
                  float inv_t_post, int32_t tokens, int32_t d, int32_t n_experts, int32_t k, int64_t* idx_out,
                  float* w_out, float* logits_out, int32_t* hist_out, void* stream);

/* ---- MoE expert dispatch, model/moe.py:191-199 without the per-expert Python loop / torch.where host syncs (fp32).
 * permute: off[E+1] = exclusive scan of hist rounded up to `align` rows per group; copy (t, r) of token t lands in row
 *   perm[t*k+r] of xp (expert-contiguous; fp32 or bf16); tile_group may be NULL.
 * grouped_gemm: C[row] = x_row . W1_e^T + b1_e for the rows of group e (bounds read from off[] on the device; max_rows =
 *   tokens * k sizes the grid); with Wg != NULL: C = (x W1^T + b1) * silu(x Wg^T + bg)  (GLUExpert, moe.py:44-49).
 *   W1 / Wg: [E][N][K] stacks (w_gstride elements apart), b1 / bg: [E][N] (b_gstride).
 * combine: out[t] = sum_r w[t*k+r] * yp[perm[t*k+r]], rank order. */
int v2m_moe_permute(const float* x, const int64_t* idx, const int32_t* hist, int32_t tokens, int32_t k, int32_t d, int32_t n_experts,
                    int32_t align, int32_t* off, int32_t* cursor, void* xp, int32_t xp_dtype, int32_t* perm, int32_t* tile_group,
                    int32_t n_tiles, void* stream);
/* Tensor-core expert path (bf16): permute with align = 128 and xp_dtype = V2M_BF16 (groups start on 128-row boundaries,
 * tile_group[t] = expert of row tile t or -1), then the grouped tcgen05 GEMM: rows of tile t use W rows
 * [g*N, (g+1)*N) of the stacked weights and bias + g*N.  swiglu_pair: h[m][j] = a[m][j] * silu(a[m][ff+j]) on the
 * [M, 2 ff] output of the stacked (linear1 | gate) GEMM (moe.py:46-47). */
int v2m_gemm_bf16_grouped(const void* A, int32_t lda, const void* W, int32_t ldw, void* C, int32_t ldc, int32_t out_dtype,
                          int32_t M_cap, int32_t N, int32_t K, int32_t n_groups, const int32_t* tile_group, const float* bias,
                          int32_t relu, void* stream);
int v2m_swiglu_pair_bf16(const void* a, void* h, int64_t M, int32_t ff, void* stream);
int v2m_moe_grouped_gemm(const float* A, int32_t lda, const float* W1, const float* b1, const float* Wg, const float* bg,
                         int64_t w_gstride, int64_t b_gstride, const int32_t* off, int32_t n_experts, int32_t max_rows, float* C,
                         int32_t ldc, int32_t N, int32_t K, void* stream);
int v2m_moe_combine(const float* yp, const int32_t* perm, const float* w, float* out, int32_t tokens, int32_t k, int32_t d, void* stream);

/* ---- MoE backward (fp32): what torch autograd derives from model/moe.py:180-199 (topk values -> softmax -> weighted
 * index_put of the expert outputs) and GLUExpert.forward (moe.py:44-49).
 * combine_bwd: dyp[perm[t*k+r]] = w[t,r] * dout[t];  dw[t,r] = <dout[t], yp[perm[t*k+r]]>;  dlogits [tokens][n_experts] =
 *   scale * softmax-backward of dw scattered to the selected experts (zero elsewhere); scale = inv_t_pre * inv_t_post.
 * swiglu_bwd: h = a * silu(g): dag[m][0:ff] = dh * silu(g), dag[m][ff:2ff] = dh * a * silu'(g)   (dag is [M][2 ff]).
 * grouped_dw: dW[e][n][j] = sum over the rows m of group e (off[e] <= m < off[e+1]) of dY[m][n] * X[m][j]; db[e][n] = sum_m dY[m][n]
 *   (db may be NULL).  dX of a grouped linear is v2m_moe_grouped_gemm over the transposed weight stack. */
int v2m_moe_combine_bwd(const float* dout, const float* yp, const int32_t* perm, const float* w, const int64_t* idx, float scale,
                        int32_t tokens, int32_t k, int32_t d, int32_t n_experts, float* dyp, float* dlogits, void* stream);
int v2m_swiglu_bwd(const float* a, const float* g, const float* dh, float* dag, int64_t M, int32_t ff, void* stream);
int v2m_moe_grouped_dw(const float* dY, int32_t ldy, const float* X, int32_t ldx, const int32_t* off, int32_t n_experts, float* dW,
                       float* db, int32_t N, int32_t K, void* stream);
/* Weight gradient of one dense fp32 layer (autograd of F.linear): dW [N][K] = dY[rows][N]^T X[rows][K], db [N] = column sums of
 * dY (may be NULL).  Rows are split over the GPU and summed with fp32 atomics (the outputs are cleared inside). */
int v2m_dw_f32(const float* dY, int32_t ldy, const float* X, int32_t ldx, int32_t rows, float* dW, float* db, int32_t N, int32_t K,
               void* stream);

/* ---- RoPE of the V2/V3 attention, custom_transformer.py:1044-1053 + rotate_operation.py:117-165, fp32 --------------
 * x, y: the (len, B, H*dh) projection as stored; cache: RotaryPositionalEmbeddings.cache[:len] ([len][H*dh/2][2] cos, sin).
 * Literal semantics: x is viewed as [H][len][B][dh] and the cache as [H][len][dh/2][2] without any transposition. */
int v2m_rope_quirk(const float* x, const float* cache, float* y, int32_t len, int32_t B, int32_t H, int32_t dh, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2M_B200_H_ */
