// Internal C++ interface between the kernel translation units and the C ABI (c_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace v2m {

// Epilogue shared by the fp32 SIMT GEMM and the bf16 tcgen05 GEMM.  Applied in this order:
//   v = acc + bias[n];  if (n < alpha_cols) v *= alpha;  if (relu) v = max(v, 0);
//   v += row_scale[m] * col_vec[n];  v += residual[(m % res_mod) * ldr + n]
struct GemmEpilogue {
  const float* bias = nullptr;
  const float* residual = nullptr;   // fp32 (or bf16 when residual_bf16 != 0)
  int ldr = 0;
  int res_mod = 0;                    // 0: residual row = m
  const float* row_scale = nullptr;  // [M]
  const float* col_vec = nullptr;    // [N]
  float alpha = 1.f;
  int alpha_cols = 0;
  int relu = 0;
  int residual_bf16 = 0;
  // Output addressing.  head_scatter == 0: C[m*ldc + n].  Otherwise the (m, n) element goes to
  // the head-major KV-cache layout  C + (n / (H*dh)) * part_stride + (((m / S)*H + h)*cap + (m % S) + pos0)*dh + d
  // with h = (n % (H*dh)) / dh, d = n % dh  (used for the cross-attention K|V projection of memory).
  int head_scatter = 0;
  int S = 0, H = 0, dh = 0, cap = 0, pos0 = 0;
  long long part_stride = 0;
  // Inverted dropout fused into the epilogue (training): element (m, n) is kept iff drop_keep(drop_seed, m, n, drop_thresh)
  // (thresh = round(p * 256), one hash byte per element) and scaled by drop_scale = 256 / (256 - thresh); 0 disables it.  drop_after_res: y = drop(acc + residual) (the
  // positional-encoding dropout) instead of y = drop(acc) + residual (sub-layer dropout before the residual add).
  float drop_scale = 0.f;
  unsigned int drop_thresh = 0, drop_seed = 0;
  int drop_after_res = 0;
  const unsigned int* drop_seed_dev = nullptr;   // optional device counter added to drop_seed (CUDA-graph replays need fresh masks)
  // bf16 GEMM, fp32 output only: C += A W^T (vector reductions into C, which the caller owns and has initialised) instead of
  // C = A W^T: weight gradients accumulate straight into the optimiser's gradient buffer.  No other epilogue operation.
  int accumulate = 0;
  // bf16 GEMM: the (bf16) `residual` operand is a GATE instead of an addend: C = residual[m][n] > 0 ? acc * gate_scale : 0.  The dX
  // GEMM of the layer that FOLLOWS a ReLU (+ fused dropout) applies that activation's backward in its epilogue: the gate is the
  // saved activation output (positive exactly where the ReLU was active and the element was kept), gate_scale the dropout scale.
  int residual_gate = 0;
  float gate_scale = 1.f;
};

#ifdef __CUDACC__
__device__ __forceinline__ long long epi_out_index(const GemmEpilogue& ep, int m, int n, int ldc) {
  if (!ep.head_scatter) return (long long)m * ldc + n;
  const int hd = ep.H * ep.dh;
  const int part = n / hd, rem = n - part * hd;
  const int h = rem / ep.dh, d = rem - h * ep.dh;
  const int b = m / ep.S, s = m - b * ep.S;
  return (long long)part * ep.part_stride + (((long long)b * ep.H + h) * ep.cap + s + ep.pos0) * ep.dh + d;
}
#endif

// a_cs / w_cs: element stride along K (1 = row-major operands; lda = 1 with a_cs = ld reads a transposed operand).
int gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
             const GemmEpilogue& ep, cudaStream_t stream, int a_cs = 1, int w_cs = 1);

// bf16 tensor-core GEMM (tcgen05/TMEM/TMA).  A [M,K] bf16 row-major (lda), W [N,K] bf16 row-major (ldw).
// out_bf16: C is bf16, else fp32.
int gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int out_bf16, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream);

int swiglu_pair_bwd_bf16(const void* a, const void* dh, void* dag, long long M, int ff, cudaStream_t stream);
int moe_group_colsum_bf16(const void* x, long long ldx, const int* off, int n_groups, float* out, int N, int rows_hint, cudaStream_t stream);
int gemm_bf16_tc_kgrouped(const void* A, int lda, const void* B, int ldb, float* C, int ldc, long long c_gstride, int M, int N,
                          int R, int n_groups, const int* k_off, cudaStream_t stream);
// General form: a_mn / b_mn != 0 means the operand is stored transposed ([K, M] resp. [K, N] row-major).
int gemm_bf16_tc_general(const void* A, int lda, int a_mn, const void* W, int ldw, int b_mn, void* C, int ldc, int out_bf16,
                         int M, int N, int K, const GemmEpilogue& ep, cudaStream_t stream);

// Attention, one (batch, head) problem per blockIdx.y.  Element (b, l, h, d) of q lives at
// q + b*q_sb + l*q_sl + h*dh + d (same for k, v, o).  kv head of query head hq is hq / (Hq / Hkv).
struct AttnParams {
  const void* q; const void* k; const void* v; void* o;
  long long q_sb, q_sl, k_sb, k_sl, v_sb, v_sl, o_sb, o_sl;
  int B, Hq, Hkv, Lq, Lk, dh;
  int causal;              // keys j <= i only
  const void* Er;          // [er_len, dh] or null; adds q_i . Er[er_len-1-(i-j)] for j <= i (rpr.py:391-395,439-455)
  int er_len;
  float q_scale;           // multiplies q on load (1.0 when the projection epilogue already scaled it)
  float* lse;              // optional [B*Hq, Lq] log-sum-exp (for backward)
  float* p_out;            // optional [B*Hq, Lq, Lk] probabilities (need_weights=True)
  // training: inverted dropout of the probabilities (rpr.py:407, F.multi_head_attention_forward): element (b*Hq+h, i, j) is kept
  // iff drop_keep(drop_seed, (b*Hq+h)*Lq + i, j, drop_thresh), scaled by drop_scale = 1/(1-p); 0 = off.  bf16 path only.
  float drop_scale = 0.f;
  unsigned int drop_thresh = 0, drop_seed = 0;
  const unsigned int* drop_seed_dev = nullptr;   // optional device counter added to drop_seed
  // optional device word: only the first min(Lk, *lk_dev) keys exist (fp32 kernel).  Lets ONE captured CUDA graph be replayed
  // for every position of a KV-cached generation: the cache grows, the launch parameters do not.
  const int* lk_dev = nullptr;
};
int attn_fwd_f32(const AttnParams& p, cudaStream_t stream);
int attn_fwd_bf16_tc(const AttnParams& p, cudaStream_t stream);

struct AttnBwdParams {
  const void* q; const void* k; const void* v; const void* o; const void* dO; const float* lse; const void* Er;
  void* dq; float* dk; float* dv; float* dEr;
  long long q_sb, q_sl, k_sb, k_sl, v_sb, v_sl, o_sb, o_sl, do_sb, do_sl, dq_sb, dq_sl, dkv_sb, dkv_sl;
  int B, Hq, Hkv, Lq, Lk, dh, causal, er_len, dtype;
  float q_scale;
  float drop_scale;                 // the forward's probability dropout (same mask function), 0 = off; tensor-core path only
  unsigned int drop_thresh, drop_seed;
  const unsigned int* drop_seed_dev;   // optional device counter added to drop_seed
  // dq is stored times dq_scale (0 = 1): the backward of the 1 / sqrt(d) that the query projection's epilogue applied in the forward,
  // so that the projection's own backward needs no pass that rescales its incoming gradient
  float dq_scale;
};
int attn_bwd(const AttnBwdParams& p, cudaStream_t stream);
// Tensor-core variant (bf16, head_dim 64, q pre-scaled): p.dk / p.dv are BF16 outputs written once (no accumulation),
// p.dEr fp32 zeroed by the caller; ws = attn_bwd_tc_workspace() bytes of device scratch (P, dS and skewed-dS tiles).
long long attn_bwd_tc_workspace(int B, int Hq, int Lq, int Lk, int has_er);
int attn_bwd_tc(const AttnBwdParams& p, void* ws, long long ws_bytes, cudaStream_t stream);
// tcgen05 / TMEM single-kernel backward of the stock attentions (no Er, no mask, Lq, Lk <= 320): csrc/attn_bwd_tc5.cu
bool attn_bwd_tc5_supported(const AttnBwdParams& p);
int attn_bwd_tc5(const AttnBwdParams& p, void* ws, long long ws_bytes, cudaStream_t stream);

int dy_prep(const void* dy, int dy_dtype, long long ld_dy, const void* y, int y_dtype, long long ld_y, int relu, float alpha,
            int alpha_cols, void* dz, int dz_dtype, long long ld_dz, float* db, int M, int N, float drop_scale, unsigned int drop_thresh,
            unsigned int drop_seed, const unsigned int* drop_seed_dev, cudaStream_t stream);
int layernorm_bwd(const void* x, int x_dtype, const float* gamma, const void* dy, int dy_dtype, void* dx, int dx_dtype,
                  float* dgamma, float* dbeta, int M, int D, float eps, cudaStream_t stream);
int embed_bwd(const long long* idx, const void* d, int d_dtype, long long ld_d, float* dtable, int rows, int D,
              cudaStream_t stream);
int amt_loss(const float* logits, const long long* tgt, const float* tgt_emotion, int R, int Cn, long long ignore, float smooth,
             float w_ce, float w_bce, float* scratch3, float* dlogits, const float* norm_in, cudaStream_t stream);
int count_valid(const long long* tgt, int R, long long ignore, float* out1, cudaStream_t stream);
int adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float weight_decay, int step,
              float grad_scale, const float* dyn, void* p16, int zero_grad, unsigned int* ctr, cudaStream_t stream);

// y = LayerNorm(x (+ res)) * gamma + beta over the last dim D (eps 1e-5 like nn.LayerNorm).
// dtype codes: 0 = fp32, 1 = bf16.  y2 (optional) receives a second copy in dtype y2_dtype.
int layernorm(const void* x, int x_dtype, const void* res, int res_dtype, const float* gamma, const float* beta,
              void* y, int y_dtype, void* y2, int y2_dtype, int M, int D, float eps, cudaStream_t stream);

// out[row, :] = table_a[idx_a[row]] (+ table_b[idx_b[row]]);  tables fp32 [*, D]; out dtype 0/1, ld_out elements.
int embed_sum(const long long* idx_a, const float* table_a, const long long* idx_b, const float* table_b,
              void* out, int out_dtype, int ld_out, int rows, int D, cudaStream_t stream);

// Video feature concat (video_music_transformer.py:1003-1018): out[row] = [sem(768) | scene | motion(md) | emotion(6) | 0-pad]
int concat_features(const float* sem, int sem_dim, const float* scene, const float* motion, int motion_dim,
                    const float* emotion, int emo_dim, void* out, int out_dtype, int ld_out, int rows, cudaStream_t stream);

int cast_copy(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, cudaStream_t stream);
// dst[r, 0:cols] = src[r, 0:cols] with independent leading dims and dtypes, zero-filling dst columns [cols, ld_dst).
int cast_copy_2d(const void* src, int src_dtype, long long ld_src, void* dst, int dst_dtype, long long ld_dst,
                 int rows, int cols, int zero_pad, cudaStream_t stream);

// out[row] = [K | V] (2 x 64 bf16) with 16-byte chunks XOR-swizzled by (row % S) & 7: cache layout of decode_stream.cu
int kv_interleave(const void* k, const void* v, void* out, long long rows, int S, cudaStream_t stream);

int binary_op(const float* a, const float* b, float* out, long long n, int mode, float alpha, cudaStream_t stream);

// ------------------------------------------------------------------ KV-cached decode (decode.cu)
constexpr int kMaxDecLayers = 8;
struct DecLayer {
  const void* w_qkv; const float* b_qkv;
  const void* w_so;  const float* b_so;
  const void* w_cq;  const float* b_cq;
  const void* w_co;  const float* b_co;
  const void* w_f1;  const float* b_f1;
  const void* w_f2;  const float* b_f2;
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;
  const float* ln3_g; const float* ln3_b;
  const void* er;                 // [er_len, dh] in the compute dtype
  const void* er_sw;              // stream path only: [8][er_len, dh] bf16, copy s holds row r with its 16-byte chunks at
                                  // position c ^ ((r - s) & 7)  (conflict-free ldmatrix for slices starting at rows = s mod 8)
  void* self_k; void* self_v;     // [B, H, cap, dh]
  const void* cross_k; const void* cross_v;  // [B, H, S, dh]
};
struct DecodeParams {
  int dtype;                      // 0 fp32, 1 bf16: type of weights, KV caches and the T scratch buffers
  int B, H, E, FF, S, cap, n_layers, er_len, vocab, vocab_limit, primer_len, chord_embed;
  DecLayer layer[kMaxDecLayers];
  const float* lnf_g; const float* lnf_b;
  const void* w_out; const float* b_out;        // [vocab, E]
  const float* emb_root; const float* emb_attr; const float* emb_chord;  // fp32 tables
  const void* w_chord; const float* wc_key; const float* b_chord;        // Linear_chord split: [E,E] | [E] | [E]
  const float* pe;                // [cap, E]
  const float* key;               // [B]
  long long* gen; long long* gen_root; long long* gen_attr;  // [B, cap] token buffers (int64 like the reference)
  int* step;                      // device scalar: position t being processed
  float* h;                       // [B, E]  residual stream (post-LN), fp32
  void* r;                        // [B, E]  pre-LN sum, dtype T
  float* qbuf;                    // [B, 3E] fp32 (q only is used)
  void* ctx;                      // [B, E]  attention output, dtype T
  void* ff;                       // [B, FF] dtype T
  float* logits;                  // [B, vocab]
  float* logits_all;              // optional [B, cap, vocab] (tests): row t receives the step-t logits
  void* xn;                       // [B, E]  normalised activations (GEMM operand), dtype T
  // next-token rule: 0 = greedy arg-max (beam=1), 1 = sampling branch of generate() with its constraints
  int sample, max_conseq_N, max_conseq_chord, pad_;
  const float* uniforms;          // [B, cap] uniforms in [0, 1) consumed at position t + 1 (sampling only)
};
// Runs `n_steps` decode steps starting at *step (device).  use_graph: capture one step into a CUDA graph and replay.
int decode_run(const DecodeParams& p, int n_steps, int use_graph, cudaStream_t stream);
long long decode_kernel_launches_per_step(const DecodeParams& p);
int decode_probe(const DecodeParams& p, int kind, int reps, cudaStream_t stream);
// One persistent cluster kernel for n_steps positions starting at t0 (bf16, d_model 512, 8 heads, FF 1024); the weight
// pointers of p reference fragment-packed matrices (decode_stream.cu).  kUnsupported otherwise.
int decode_run_stream(const DecodeParams& p, int t0, int n_steps, unsigned long long* ts, int ts_cap, cudaStream_t stream);

// ------------------------------------------------------------------ selective scan (pscan.cu)
int pscan_fwd(const float* A, const float* X, float* H, int B, int L, int D, int N, cudaStream_t stream);
int pscan_bwd(const float* A, const float* H, const float* gH, float* gA, float* gX, int B, int L, int D, int N,
              cudaStream_t stream);

// ------------------------------------------------------------------ Mamba block (mamba.cu), fp32
int mamba_step_conv(const float* xz, long long ldxz, const float* in_old, const float* w, const float* bias, float* xs, float* in_new,
                    int B, int ED, int KW, cudaStream_t stream);
int mamba_step_ssm(const float* xs, const float* dbc, long long lddbc, const float* dtw, const float* dtb, const float* A_log,
                   const float* D, const float* z, long long ldz, const float* h_old, float* h_new, float* out, int B, int ED, int N,
                   int R, cudaStream_t stream);
int mamba_conv_silu(const float* x, long long ldx, const float* w, const float* bias, float* y, long long ldy, int B, int L, int ED,
                    int KW, cudaStream_t stream);
long long selective_scan_workspace(int B, int L, int ED, int N);
int selective_scan_fwd(const float* x, long long ldx, const float* delta_raw, long long ldd, const float* dt_bias,
                       const float* A_log, const float* Bm, const float* Cm, long long ldbc, const float* Dp, const float* z,
                       long long ldz, float* out, long long ldo, int B, int L, int ED, int N, int plus, float* ws,
                       long long ws_bytes, cudaStream_t stream);
// one generation step of the generic decoder stacks (csrc/step_f32.cu): M <= 64 new rows per launch, one query row per (video, head)
int step_linear_f32(const float* x, long long ldx, const float* W, long long ldw, const float* bias, const float* row_scale,
                    const float* col_vec, float* y, long long ldy, int M, int N, int K, int relu, cudaStream_t stream);
int step_moe_gemm_f32(const float* A, int lda, const float* W1, const float* b1, const float* Wg, const float* bg, long long w_gstride,
                      long long b_gstride, const int* off, int n_experts, float* C, int ldc, int N, int K, cudaStream_t stream);
int step_attn_f32(const float* q, long long q_sb, const float* k, const float* v, long long kv_sb, long long kv_sl, float* o, long long o_sb,
                  int B, int Hq, int Hkv, int dh, int n_max, const int* n_dev, float q_scale, cudaStream_t stream);
int rmsnorm(const float* x, const float* w, float* y, int M, int D, float eps, cudaStream_t stream);
long long selective_scan_bwd_workspace(int B, int L, int ED, int N);
int selective_scan_bwd(const float* x, long long ldx, const float* delta_raw, long long ldd, const float* dt_bias, const float* A_log,
                       const float* Bm, const float* Cm, long long ldbc, const float* Dp, const float* z, long long ldz,
                       const float* dout, long long ldo, float* hs, long long hs_bytes, float* dx, long long lddx, float* ddraw,
                       long long lddd, float* dBm, float* dCm, long long lddbc, float* dz, long long lddz, float* dA_log, float* dD,
                       float* ddt_bias, int B, int L, int ED, int N, int plus, cudaStream_t stream);
int mamba_conv_silu_bwd(const float* x, long long ldx, const float* w, const float* bias, const float* dy, long long ldy, float* dx,
                        long long lddx, float* dw, float* dbias, int B, int L, int ED, int KW, cudaStream_t stream);
int rmsnorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int M, int D, float eps, cudaStream_t stream);

// ------------------------------------------------------------------ MoE (moe.cu)
int moe_route(const float* x, const float* wg, const float* bg, const float* sel_bias, float inv_t_pre, float inv_t_post,
              int tokens, int d, int n_experts, int k, long long* idx_out, float* w_out, float* logits_out,
              int* hist_out, cudaStream_t stream);

// Expert dispatch (moe.cu): expert-contiguous token permutation, ragged grouped GEMM with device-side group bounds
// (optionally the fused SwiGLU pair of GLUExpert), weighted combine in rank order.
int moe_permute(const float* x, const long long* idx, const int* hist, int tokens, int k, int d, int n_experts, int align, int* off,
                int* cursor, void* xp, int xp_bf16, int* perm, int* tile_group, int n_tiles, cudaStream_t stream);
int swiglu_pair_bf16(const void* a, void* h, long long M, int ff, cudaStream_t stream);
int gemm_bf16_tc_grouped(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int out_bf16, int M_cap, int N, int K,
                         int n_groups, const int* tile_group, const float* bias, int relu, cudaStream_t stream);
int moe_grouped_gemm(const float* A, int lda, const float* W1, const float* b1, const float* Wg, const float* bg, long long w_gstride,
                     long long b_gstride, const int* off, int n_experts, int max_rows, float* C, int ldc, int N, int K,
                     cudaStream_t stream);
int moe_combine(const float* yp, const int* perm, const float* w, float* out, int tokens, int k, int d, cudaStream_t stream);
// backward of the dispatch: combine / softmax-over-top-k, SwiGLU, per-expert weight gradients over the ragged groups
int moe_combine_bwd(const float* dout, const float* yp, const int* perm, const float* w, const long long* idx, float scale, int tokens,
                    int k, int d, int n_experts, float* dyp, float* dlogits, cudaStream_t stream);
int swiglu_bwd(const float* a, const float* g, const float* dh, float* dag, long long M, int ff, cudaStream_t stream);
int moe_grouped_dw(const float* dY, int ldy, const float* X, int ldx, const int* off, int n_experts, float* dW, float* db, int N, int K,
                   cudaStream_t stream);
int dw_f32(const float* dY, int ldy, const float* X, int ldx, int rows, float* dW, float* db, int N, int K, cudaStream_t stream);

// accuracy / hits@k counters of the evaluation loop (train.cu)
int amt_correspondence(const float* logits, const float* emo, const float* prob, int R, int Cn, int Ce, float thr, int chord_end,
                       int* counters, cudaStream_t stream);
int amt_metrics(const float* logits, const long long* tgt, int R, int Cn, long long pad, int k0, int k1, int k2, int* counters,
                cudaStream_t stream);

// RoPE with the reference's reinterpretations (elementwise.cu)
int rope_quirk(const float* x, const float* cache, float* y, int len, int B, int H, int dh, cudaStream_t stream);

}  // namespace v2m
