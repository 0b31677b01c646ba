// extern "C" entry points declared in include/v2m_b200.h + error plumbing.
#include "common.cuh"
#include "kernels.h"
#include "../../include/v2m_b200.h"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

namespace v2m {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool dep_launch_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("V2M_PDL"); on = (e && e[0] == '1') ? 1 : 0; }   // opt-in: see the note in common.cuh
  return on != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return kCudaError;
  }
  return kOk;
}

static GemmEpilogue to_ep(const v2m_epilogue* e) {
  GemmEpilogue g;
  if (!e) return g;
  g.bias = e->bias;
  g.residual = static_cast<const float*>(e->residual);
  g.ldr = e->ldr; g.res_mod = e->res_mod;
  g.row_scale = e->row_scale; g.col_vec = e->col_vec;
  g.alpha = e->alpha; g.alpha_cols = e->alpha_cols; g.relu = e->relu; g.residual_bf16 = e->residual_bf16;
  g.head_scatter = e->head_scatter; g.S = e->S; g.H = e->H; g.dh = e->dh; g.cap = e->cap; g.pos0 = e->pos0;
  g.part_stride = e->part_stride;
  g.drop_scale = e->drop_scale; g.drop_thresh = e->drop_thresh; g.drop_seed = e->drop_seed; g.drop_after_res = e->drop_after_res;
  g.drop_seed_dev = e->drop_seed_dev;
  g.accumulate = e->accumulate;
  g.residual_gate = e->residual_gate; g.gate_scale = e->gate_scale;
  return g;
}

static_assert(sizeof(v2m_dec_layer) == sizeof(DecLayer), "v2m_dec_layer must mirror DecLayer");
static_assert(sizeof(v2m_decode) == sizeof(DecodeParams), "v2m_decode must mirror DecodeParams");

}  // namespace v2m

using namespace v2m;

extern "C" {

int v2m_abi_version(void) { return 1; }
const char* v2m_last_error(void) { return g_err; }

int64_t v2m_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(v2m_epilogue);
    case 1: return sizeof(v2m_attn);
    case 2: return sizeof(v2m_dec_layer);
    case 3: return sizeof(v2m_decode);
    case 4: return sizeof(v2m_attn_bwd_t);
    default: return -1;
  }
}

int v2m_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int v2m_gemm_f32(const float* A, int32_t lda, const float* W, int32_t ldw, float* C, int32_t ldc, int32_t M, int32_t N,
                 int32_t K, const v2m_epilogue* ep, void* stream) {
  GemmEpilogue g = to_ep(ep);
  V2M_REQUIRE(!g.residual_bf16, "v2m_gemm_f32: bf16 residual not supported on the fp32 path");
  V2M_REQUIRE(!g.accumulate, "v2m_gemm_f32: accumulate exists on the bf16 tensor-core path only");
  return gemm_f32(A, lda, W, ldw, C, ldc, M, N, K, g, static_cast<cudaStream_t>(stream));
}

int v2m_gemm_f32_strided(const float* A, int32_t a_rs, int32_t a_cs, const float* W, int32_t w_rs, int32_t w_cs, float* C,
                         int32_t ldc, int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream) {
  GemmEpilogue g = to_ep(ep);
  V2M_REQUIRE(!g.residual_bf16, "v2m_gemm_f32_strided: bf16 residual not supported on the fp32 path");
  return gemm_f32(A, a_rs, W, w_rs, C, ldc, M, N, K, g, static_cast<cudaStream_t>(stream), a_cs, w_cs);
}

int v2m_gemm_bf16(const void* A, int32_t lda, const void* W, int32_t ldw, void* C, int32_t ldc, int32_t out_dtype,
                  int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream) {
  return gemm_bf16_tc(A, lda, W, ldw, C, ldc, out_dtype == V2M_BF16, M, N, K, to_ep(ep), static_cast<cudaStream_t>(stream));
}

int v2m_gemm_bf16_general(const void* A, int32_t lda, int32_t a_mn, const void* W, int32_t ldw, int32_t b_mn, void* C, int32_t ldc,
                          int32_t out_dtype, int32_t M, int32_t N, int32_t K, const v2m_epilogue* ep, void* stream) {
  return gemm_bf16_tc_general(A, lda, a_mn, W, ldw, b_mn, C, ldc, out_dtype == V2M_BF16, M, N, K, to_ep(ep),
                              static_cast<cudaStream_t>(stream));
}

int v2m_attn_fwd(const v2m_attn* a, int32_t dtype, void* stream) {
  V2M_REQUIRE(a != nullptr, "v2m_attn_fwd: null params");
  AttnParams p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.o = a->o;
  p.q_sb = a->q_sb; p.q_sl = a->q_sl; p.k_sb = a->k_sb; p.k_sl = a->k_sl;
  p.v_sb = a->v_sb; p.v_sl = a->v_sl; p.o_sb = a->o_sb; p.o_sl = a->o_sl;
  p.B = a->B; p.Hq = a->Hq; p.Hkv = a->Hkv; p.Lq = a->Lq; p.Lk = a->Lk; p.dh = a->dh;
  p.causal = a->causal; p.Er = a->Er; p.er_len = a->er_len; p.q_scale = a->q_scale;
  p.lse = a->lse; p.p_out = a->p_out;
  p.drop_scale = a->drop_scale; p.drop_thresh = a->drop_thresh; p.drop_seed = a->drop_seed; p.drop_seed_dev = a->drop_seed_dev;
  p.lk_dev = a->lk_dev;
  if (dtype == V2M_F32) return attn_fwd_f32(p, static_cast<cudaStream_t>(stream));
  if (dtype == V2M_BF16) return attn_fwd_bf16_tc(p, static_cast<cudaStream_t>(stream));
  set_last_error("v2m_attn_fwd: dtype %d unsupported", dtype);
  return kUnsupported;
}

int v2m_attn_bwd(const v2m_attn_bwd_t* a, void* stream) {
  V2M_REQUIRE(a != nullptr, "v2m_attn_bwd: null params");
  static_assert(sizeof(v2m_attn_bwd_t) == sizeof(AttnBwdParams), "v2m_attn_bwd_t must mirror AttnBwdParams");
  AttnBwdParams p;
  memcpy(&p, a, sizeof(p));
  return attn_bwd(p, static_cast<cudaStream_t>(stream));
}

int64_t v2m_attn_bwd_tc_workspace(int32_t B, int32_t Hq, int32_t Lq, int32_t Lk, int32_t has_er) {
  return attn_bwd_tc_workspace(B, Hq, Lq, Lk, has_er);
}

int v2m_attn_bwd_tc(const v2m_attn_bwd_t* a, void* ws, int64_t ws_bytes, void* stream) {
  V2M_REQUIRE(a != nullptr, "v2m_attn_bwd_tc: null params");
  AttnBwdParams p;
  memcpy(&p, a, sizeof(p));
  return attn_bwd_tc(p, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int v2m_dy_prep(const void* dy, int32_t dy_dtype, int64_t ld_dy, const void* y, int32_t y_dtype, int64_t ld_y, int32_t relu,
                float alpha, int32_t alpha_cols, void* dz, int32_t dz_dtype, int64_t ld_dz, float* db, int32_t M, int32_t N,
                float drop_scale, uint32_t drop_thresh, uint32_t drop_seed, const uint32_t* drop_seed_dev, void* stream) {
  return dy_prep(dy, dy_dtype, ld_dy, y, y_dtype, ld_y, relu, alpha, alpha_cols, dz, dz_dtype, ld_dz, db, M, N, drop_scale, drop_thresh,
                 drop_seed, drop_seed_dev, static_cast<cudaStream_t>(stream));
}

int v2m_layernorm_bwd(const void* x, int32_t x_dtype, const float* gamma, const void* dy, int32_t dy_dtype, void* dx,
                      int32_t dx_dtype, float* dgamma, float* dbeta, int32_t M, int32_t D, float eps, void* stream) {
  return layernorm_bwd(x, x_dtype, gamma, dy, dy_dtype, dx, dx_dtype, dgamma, dbeta, M, D, eps, static_cast<cudaStream_t>(stream));
}

int v2m_embed_bwd(const int64_t* idx, const void* d, int32_t d_dtype, int64_t ld_d, float* dtable, int32_t rows, int32_t D,
                  void* stream) {
  return embed_bwd(reinterpret_cast<const long long*>(idx), d, d_dtype, ld_d, dtable, rows, D, static_cast<cudaStream_t>(stream));
}

int v2m_amt_correspondence(const float* logits, const float* tgt_emotion, const float* tgt_emotion_prob, int32_t R, int32_t Cn, int32_t Ce,
                           float threshold, int32_t chord_end, int32_t* counters, void* stream) {
  return amt_correspondence(logits, tgt_emotion, tgt_emotion_prob, R, Cn, Ce, threshold, chord_end, counters, static_cast<cudaStream_t>(stream));
}

int v2m_amt_metrics(const float* logits, const int64_t* tgt, int32_t R, int32_t Cn, int64_t pad, int32_t k0, int32_t k1, int32_t k2,
                    int32_t* counters, void* stream) {
  return amt_metrics(logits, reinterpret_cast<const long long*>(tgt), R, Cn, pad, k0, k1, k2, counters, static_cast<cudaStream_t>(stream));
}

int v2m_amt_loss(const float* logits, const int64_t* tgt, const float* tgt_emotion, int32_t R, int32_t Cn, int64_t ignore,
                 float smooth, float w_ce, float w_bce, float* scratch3, float* dlogits, const float* norm_in, void* stream) {
  return amt_loss(logits, reinterpret_cast<const long long*>(tgt), tgt_emotion, R, Cn, ignore, smooth, w_ce, w_bce, scratch3,
                  dlogits, norm_in, static_cast<cudaStream_t>(stream));
}

int v2m_count_valid(const int64_t* tgt, int32_t R, int64_t ignore, float* out1, void* stream) {
  return count_valid(reinterpret_cast<const long long*>(tgt), R, ignore, out1, static_cast<cudaStream_t>(stream));
}

int v2m_adam_step(float* p, float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float weight_decay, int32_t step,
                  float grad_scale, const float* dyn, void* p16, int32_t zero_grad, uint32_t* ctr, void* stream) {
  return adam_step(p, g, m, v, n, lr, b1, b2, eps, weight_decay, step, grad_scale, dyn, p16, zero_grad, ctr, static_cast<cudaStream_t>(stream));
}

int v2m_layernorm(const void* x, int32_t x_dtype, const void* res, int32_t res_dtype, const float* gamma, const float* beta,
                  void* y, int32_t y_dtype, void* y2, int32_t y2_dtype, int32_t M, int32_t D, float eps, void* stream) {
  return layernorm(x, x_dtype, res, res_dtype, gamma, beta, y, y_dtype, y2, y2_dtype, M, D, eps,
                   static_cast<cudaStream_t>(stream));
}

int v2m_step_linear_f32(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias, const float* row_scale,
                        const float* col_vec, float* y, int64_t ldy, int32_t M, int32_t N, int32_t K, int32_t relu, void* stream) {
  return step_linear_f32(x, ldx, W, ldw, bias, row_scale, col_vec, y, ldy, M, N, K, relu, static_cast<cudaStream_t>(stream));
}

int v2m_step_attn_f32(const float* q, int64_t q_sb, const float* k, const float* v, int64_t kv_sb, int64_t kv_sl, float* o, int64_t o_sb,
                      int32_t B, int32_t Hq, int32_t Hkv, int32_t dh, int32_t n_max, const int32_t* n_dev, float q_scale, void* stream) {
  return step_attn_f32(q, q_sb, k, v, kv_sb, kv_sl, o, o_sb, B, Hq, Hkv, dh, n_max, n_dev, q_scale, static_cast<cudaStream_t>(stream));
}

int v2m_embed_sum(const int64_t* idx_a, const float* table_a, const int64_t* idx_b, const float* table_b, void* out,
                  int32_t out_dtype, int32_t ld_out, int32_t rows, int32_t D, void* stream) {
  return embed_sum(reinterpret_cast<const long long*>(idx_a), table_a, reinterpret_cast<const long long*>(idx_b), table_b,
                   out, out_dtype, ld_out, rows, D, static_cast<cudaStream_t>(stream));
}

int v2m_concat_features(const float* sem, int32_t sem_dim, const float* scene, const float* motion, int32_t motion_dim,
                        const float* emotion, int32_t emo_dim, void* out, int32_t out_dtype, int32_t ld_out, int32_t rows,
                        void* stream) {
  return concat_features(sem, sem_dim, scene, motion, motion_dim, emotion, emo_dim, out, out_dtype, ld_out, rows,
                         static_cast<cudaStream_t>(stream));
}

int v2m_cast_2d(const void* src, int32_t src_dtype, int64_t ld_src, void* dst, int32_t dst_dtype, int64_t ld_dst,
                int32_t rows, int32_t cols, int32_t zero_pad, void* stream) {
  return cast_copy_2d(src, src_dtype, ld_src, dst, dst_dtype, ld_dst, rows, cols, zero_pad,
                      static_cast<cudaStream_t>(stream));
}

int v2m_rope_quirk(const float* x, const float* cache, float* y, int32_t len, int32_t B, int32_t H, int32_t dh, void* stream) {
  return rope_quirk(x, cache, y, len, B, H, dh, static_cast<cudaStream_t>(stream));
}

int v2m_binary_f32(const float* a, const float* b, float* out, int64_t n, int32_t mode, float alpha, void* stream) {
  return binary_op(a, b, out, n, mode, alpha, static_cast<cudaStream_t>(stream));
}

int v2m_decode_run(const v2m_decode* p, int32_t n_steps, int32_t use_graph, void* stream) {
  V2M_REQUIRE(p != nullptr, "v2m_decode_run: null params");
  DecodeParams d;
  memcpy(&d, p, sizeof(d));
  return decode_run(d, n_steps, use_graph, static_cast<cudaStream_t>(stream));
}

int v2m_decode_run_stream(const v2m_decode* p, int32_t t0, int32_t n_steps, uint64_t* timestamps, int32_t ts_cap, void* stream) {
  V2M_REQUIRE(p != nullptr, "v2m_decode_run_stream: null params");
  DecodeParams d;
  memcpy(&d, p, sizeof(d));
  int rc = decode_run_stream(d, t0, n_steps, reinterpret_cast<unsigned long long*>(timestamps), ts_cap, static_cast<cudaStream_t>(stream));
  if (rc == kUnsupported)
    set_last_error("decode_run_stream: configuration not covered by the streamed cluster kernel (bf16, d_model 512, 8 heads, "
                   "dim_feedforward 1024, all clusters co-resident)");
  return rc;
}

int v2m_gemm_bf16_kgrouped(const void* A, int32_t lda, const void* B, int32_t ldb, float* C, int32_t ldc, int64_t c_gstride, int32_t M,
                           int32_t N, int32_t R, int32_t n_groups, const int32_t* k_off, void* stream) {
  V2M_REQUIRE(A && B && C && k_off, "v2m_gemm_bf16_kgrouped: null pointer");
  return gemm_bf16_tc_kgrouped(A, lda, B, ldb, C, ldc, c_gstride, M, N, R, n_groups, k_off, static_cast<cudaStream_t>(stream));
}

int v2m_swiglu_pair_bwd_bf16(const void* a, const void* dh, void* dag, int64_t M, int32_t ff, void* stream) {
  V2M_REQUIRE(a && dh && dag, "v2m_swiglu_pair_bwd_bf16: null pointer");
  return swiglu_pair_bwd_bf16(a, dh, dag, M, ff, static_cast<cudaStream_t>(stream));
}

int v2m_moe_group_colsum_bf16(const void* x, int64_t ldx, const int32_t* off, int32_t n_groups, float* out, int32_t N, int32_t rows_hint,
                              void* stream) {
  return moe_group_colsum_bf16(x, ldx, off, n_groups, out, N, rows_hint, static_cast<cudaStream_t>(stream));
}

int v2m_kv_interleave(const void* k, const void* v, void* out, int64_t rows, int32_t S, void* stream) {
  V2M_REQUIRE(k && v && out, "v2m_kv_interleave: null pointer");
  return kv_interleave(k, v, out, rows, S, static_cast<cudaStream_t>(stream));
}

int v2m_decode_probe(const v2m_decode* p, int32_t kind, int32_t reps, void* stream) {
  V2M_REQUIRE(p != nullptr, "v2m_decode_probe: null params");
  DecodeParams d;
  memcpy(&d, p, sizeof(d));
  return decode_probe(d, kind, reps, static_cast<cudaStream_t>(stream));
}

int64_t v2m_decode_launches_per_step(const v2m_decode* p) {
  DecodeParams d;
  memcpy(&d, p, sizeof(d));
  return decode_kernel_launches_per_step(d);
}

int v2m_mamba_conv_silu(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int32_t B, int32_t L,
                        int32_t ED, int32_t KW, void* stream) {
  return mamba_conv_silu(x, ldx, w, bias, y, ldy, B, L, ED, KW, static_cast<cudaStream_t>(stream));
}

int v2m_mamba_step_conv(const float* xz, int64_t ldxz, const float* in_old, const float* w, const float* bias, float* xs, float* in_new,
                        int32_t B, int32_t ED, int32_t KW, void* stream) {
  V2M_REQUIRE(xz && in_old && w && xs && in_new, "v2m_mamba_step_conv: null pointer");
  return mamba_step_conv(xz, ldxz, in_old, w, bias, xs, in_new, B, ED, KW, static_cast<cudaStream_t>(stream));
}

int v2m_mamba_step_ssm(const float* xs, const float* dbc, int64_t lddbc, const float* dtw, const float* dtb, const float* A_log,
                       const float* D, const float* z, int64_t ldz, const float* h_old, float* h_new, float* out, int32_t B, int32_t ED,
                       int32_t N, int32_t R, void* stream) {
  V2M_REQUIRE(xs && dbc && dtw && dtb && A_log && D && z && h_new && out, "v2m_mamba_step_ssm: null pointer");
  return mamba_step_ssm(xs, dbc, lddbc, dtw, dtb, A_log, D, z, ldz, h_old, h_new, out, B, ED, N, R, static_cast<cudaStream_t>(stream));
}

int64_t v2m_selective_scan_workspace(int32_t B, int32_t L, int32_t ED, int32_t N) { return selective_scan_workspace(B, L, ED, N); }

int v2m_selective_scan_fwd(const float* x, int64_t ldx, const float* delta_raw, int64_t ldd, const float* dt_bias, const float* A_log,
                           const float* Bm, const float* Cm, int64_t ldbc, const float* D, const float* z, int64_t ldz, float* out,
                           int64_t ldo, int32_t B, int32_t L, int32_t ED, int32_t N, int32_t plus, float* ws, int64_t ws_bytes,
                           void* stream) {
  return selective_scan_fwd(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, D, z, ldz, out, ldo, B, L, ED, N, plus, ws, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

int v2m_rmsnorm(const float* x, const float* w, float* y, int32_t M, int32_t D, float eps, void* stream) {
  return rmsnorm(x, w, y, M, D, eps, static_cast<cudaStream_t>(stream));
}

int64_t v2m_selective_scan_bwd_workspace(int32_t B, int32_t L, int32_t ED, int32_t N) { return selective_scan_bwd_workspace(B, L, ED, N); }

int v2m_selective_scan_bwd(const float* x, int64_t ldx, const float* delta_raw, int64_t ldd, const float* dt_bias, const float* A_log,
                           const float* Bm, const float* Cm, int64_t ldbc, const float* D, const float* z, int64_t ldz,
                           const float* dout, int64_t ldo, float* hs, int64_t hs_bytes, float* dx, int64_t lddx, float* ddelta_raw,
                           int64_t lddd, float* dBm, float* dCm, int64_t lddbc, float* dz, int64_t lddz, float* dA_log, float* dD,
                           float* ddt_bias, int32_t B, int32_t L, int32_t ED, int32_t N, int32_t plus, void* stream) {
  return selective_scan_bwd(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, D, z, ldz, dout, ldo, hs, hs_bytes, dx, lddx,
                            ddelta_raw, lddd, dBm, dCm, lddbc, dz, lddz, dA_log, dD, ddt_bias, B, L, ED, N, plus,
                            static_cast<cudaStream_t>(stream));
}

int v2m_mamba_conv_silu_bwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* dy, int64_t ldy, float* dx,
                            int64_t lddx, float* dw, float* dbias, int32_t B, int32_t L, int32_t ED, int32_t KW, void* stream) {
  return mamba_conv_silu_bwd(x, ldx, w, bias, dy, ldy, dx, lddx, dw, dbias, B, L, ED, KW, static_cast<cudaStream_t>(stream));
}

int v2m_rmsnorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int32_t M, int32_t D, float eps, void* stream) {
  return rmsnorm_bwd(x, w, dy, dx, dw, M, D, eps, static_cast<cudaStream_t>(stream));
}

int v2m_moe_permute(const float* x, const int64_t* idx, const int32_t* hist, int32_t tokens, int32_t k, int32_t d, int32_t n_experts,
                    int32_t align, int32_t* off, int32_t* cursor, void* xp, int32_t xp_dtype, int32_t* perm, int32_t* tile_group,
                    int32_t n_tiles, void* stream) {
  return moe_permute(x, reinterpret_cast<const long long*>(idx), hist, tokens, k, d, n_experts, align, off, cursor, xp,
                     xp_dtype == V2M_BF16, perm, tile_group, n_tiles, static_cast<cudaStream_t>(stream));
}

int v2m_gemm_bf16_grouped(const void* A, int32_t lda, const void* W, int32_t ldw, void* C, int32_t ldc, int32_t out_dtype,
                          int32_t M_cap, int32_t N, int32_t K, int32_t n_groups, const int32_t* tile_group, const float* bias,
                          int32_t relu, void* stream) {
  return gemm_bf16_tc_grouped(A, lda, W, ldw, C, ldc, out_dtype == V2M_BF16, M_cap, N, K, n_groups, tile_group, bias, relu,
                              static_cast<cudaStream_t>(stream));
}

int v2m_swiglu_pair_bf16(const void* a, void* h, int64_t M, int32_t ff, void* stream) {
  return swiglu_pair_bf16(a, h, M, ff, static_cast<cudaStream_t>(stream));
}

int v2m_moe_grouped_gemm(const float* A, int32_t lda, const float* W1, const float* b1, const float* Wg, const float* bg,
                         int64_t w_gstride, int64_t b_gstride, const int32_t* off, int32_t n_experts, int32_t max_rows, float* C,
                         int32_t ldc, int32_t N, int32_t K, void* stream) {
  return moe_grouped_gemm(A, lda, W1, b1, Wg, bg, w_gstride, b_gstride, off, n_experts, max_rows, C, ldc, N, K,
                          static_cast<cudaStream_t>(stream));
}

int v2m_moe_combine(const float* yp, const int32_t* perm, const float* w, float* out, int32_t tokens, int32_t k, int32_t d, void* stream) {
  return moe_combine(yp, perm, w, out, tokens, k, d, static_cast<cudaStream_t>(stream));
}

int v2m_moe_combine_bwd(const float* dout, const float* yp, const int32_t* perm, const float* w, const int64_t* idx, float scale,
                        int32_t tokens, int32_t k, int32_t d, int32_t n_experts, float* dyp, float* dlogits, void* stream) {
  return moe_combine_bwd(dout, yp, perm, w, reinterpret_cast<const long long*>(idx), scale, tokens, k, d, n_experts, dyp, dlogits,
                         static_cast<cudaStream_t>(stream));
}

int v2m_swiglu_bwd(const float* a, const float* g, const float* dh, float* dag, int64_t M, int32_t ff, void* stream) {
  return swiglu_bwd(a, g, dh, dag, M, ff, static_cast<cudaStream_t>(stream));
}

int v2m_moe_grouped_dw(const float* dY, int32_t ldy, const float* X, int32_t ldx, const int32_t* off, int32_t n_experts, float* dW,
                       float* db, int32_t N, int32_t K, void* stream) {
  return moe_grouped_dw(dY, ldy, X, ldx, off, n_experts, dW, db, N, K, static_cast<cudaStream_t>(stream));
}

int v2m_dw_f32(const float* dY, int32_t ldy, const float* X, int32_t ldx, int32_t rows, float* dW, float* db, int32_t N, int32_t K,
               void* stream) {
  return dw_f32(dY, ldy, X, ldx, rows, dW, db, N, K, static_cast<cudaStream_t>(stream));
}

int v2m_pscan_fwd(const float* A, const float* X, float* H, int32_t B, int32_t L, int32_t D, int32_t N, void* stream) {
  return pscan_fwd(A, X, H, B, L, D, N, static_cast<cudaStream_t>(stream));
}

int v2m_pscan_bwd(const float* A, const float* H, const float* gH, float* gA, float* gX, int32_t B, int32_t L, int32_t D,
                  int32_t N, void* stream) {
  return pscan_bwd(A, H, gH, gA, gX, B, L, D, N, static_cast<cudaStream_t>(stream));
}

int v2m_moe_route(const float* x, const float* wg, const float* bg, const float* sel_bias, float inv_t_pre, float inv_t_post,
                  int32_t tokens, int32_t d, int32_t n_experts, int32_t k, int64_t* idx_out, float* w_out, float* logits_out,
                  int32_t* hist_out, void* stream) {
  return moe_route(x, wg, bg, sel_bias, inv_t_pre, inv_t_post, tokens, d, n_experts, k,
                   reinterpret_cast<long long*>(idx_out), w_out, logits_out, hist_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
