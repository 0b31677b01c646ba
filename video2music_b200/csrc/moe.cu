// MoE router: gate GEMV + top-k + softmax + expert histogram in one pass (model/moe.py:180-190,
// 244-288).  One warp per token; the gate logits are accumulated in fp32 in a fixed order
// (lane-strided partial sums, then an xor-shuffle tree) so routing never depends on the batch.
// top-k = k rounds of arg-max, lowest index first on ties, values in descending order exactly like
// torch.topk's sorted output.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

constexpr int kMaxExperts = 16;

__global__ void __launch_bounds__(256) moe_route_kernel(const float* __restrict__ x, const float* __restrict__ wg,
                                                        const float* __restrict__ bg, const float* __restrict__ sel_bias,
                                                        float inv_t_pre, float inv_t_post, int tokens, int d, int n_experts,
                                                        int k, long long* __restrict__ idx_out, float* __restrict__ w_out,
                                                        float* __restrict__ logits_out, int* __restrict__ hist_out) {
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  const float* xr = x + (size_t)tok * d;
  float logit[kMaxExperts];
#pragma unroll
  for (int e = 0; e < kMaxExperts; ++e) {
    if (e < n_experts) {
      const float* w = wg + (size_t)e * d;
      float acc = 0.f;
      for (int i = lane; i < d; i += 32) acc = fmaf(xr[i], __ldg(w + i), acc);
      logit[e] = (warp_sum(acc) + bg[e]) * inv_t_pre;
    } else {
      logit[e] = -INFINITY;
    }
  }
  if (lane != 0) return;
  if (logits_out)
    for (int e = 0; e < n_experts; ++e) logits_out[(size_t)tok * n_experts + e] = logit[e];
  float sel[kMaxExperts];
#pragma unroll
  for (int e = 0; e < kMaxExperts; ++e) sel[e] = (e < n_experts) ? logit[e] + (sel_bias ? sel_bias[e] : 0.f) : -INFINITY;
  float wv[kMaxExperts];
  int wi[kMaxExperts];
  float mx = -INFINITY;
  for (int r = 0; r < k; ++r) {
    int best = 0;
    float bv = -INFINITY;
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e)
      if (e < n_experts && sel[e] > bv) { bv = sel[e]; best = e; }
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e)
      if (e == best) { wv[r] = logit[e] * inv_t_post; sel[e] = -INFINITY; }   // weights come from the un-biased logits
    wi[r] = best;
    mx = fmaxf(mx, wv[r]);
  }
  float sum = 0.f;
  for (int r = 0; r < k; ++r) { wv[r] = expf(wv[r] - mx); sum += wv[r]; }
  for (int r = 0; r < k; ++r) {
    idx_out[(size_t)tok * k + r] = wi[r];
    w_out[(size_t)tok * k + r] = wv[r] / sum;
    if (hist_out) atomicAdd(hist_out + wi[r], 1);
  }
}

int moe_route(const float* x, const float* wg, const float* bg, const float* sel_bias, float inv_t_pre, float inv_t_post,
              int tokens, int d, int n_experts, int k, long long* idx_out, float* w_out, float* logits_out,
              int* hist_out, cudaStream_t stream) {
  V2M_REQUIRE(n_experts >= 1 && n_experts <= kMaxExperts, "moe_route: n_experts=%d (max %d)", n_experts, kMaxExperts);
  V2M_REQUIRE(k >= 1 && k <= n_experts, "moe_route: k=%d out of range", k);
  if (tokens == 0) return kOk;
  const int wpb = 8;
  moe_route_kernel<<<(tokens + wpb - 1) / wpb, wpb * 32, 0, stream>>>(x, wg, bg, sel_bias, inv_t_pre, inv_t_post, tokens, d,
                                                                      n_experts, k, idx_out, w_out, logits_out, hist_out);
  return check_launch("moe_route");
}

}  // namespace v2m
