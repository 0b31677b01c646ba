// Next-token choice of the decode kernels: one warp per video.
//   sample == 0: greedy arg-max over [:vocab_limit], first index wins on ties (generate(beam=1), video_music_transformer.py:1078-1084)
//   sample != 0: the sampling branch (:1085-1104): probabilities softmax(logits)[:vocab_limit] with P(N) = 0 when
//                max_conseq_N == 0 and P(previous chord) = 0 when the last max_conseq_chord tokens are equal, then one
//                draw from Categorical(p) -- here by inverse CDF with the caller's uniform u in [0, 1)
//                (torch.multinomial uses another algorithm, so streams differ from the reference's for the same seed).
#pragma once
#include "common.cuh"

namespace v2m {

// hist(k): token at position t - k (k = 0 is the current input token); only called for k < max_conseq_chord.
template <class Hist>
__device__ __forceinline__ int pick_token(const float* logits, int vocab, int vocab_limit, int sample, int max_conseq_N,
                                          int max_conseq_chord, int t, float u, Hist hist, int lane) {
  if (!sample) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = lane; n < vocab_limit; n += 32) {
      const float v = __ldcg(logits + n);
      if (v > best) { best = v; bi = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    return bi;
  }
  constexpr int PER = 8;                                   // lane l owns indices [8 l, 8 l + 8): prefix order == index order
  float mx = -INFINITY;
  float v[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int n = lane * PER + e;
    v[e] = n < vocab ? __ldcg(logits + n) : -INFINITY;
    mx = fmaxf(mx, v[e]);
  }
  mx = warp_max(mx);                                       // softmax runs over the whole vocabulary (:1069), then [:vocab_limit]
  // repeat constraint (:1093-1104): cur_i = t + 1 >= max_conseq_chord and the last max_conseq_chord tokens are equal
  int banned = -1;
  if (t + 1 >= max_conseq_chord) {
    const int prev = hist(0);
    bool same = true;
    for (int k = 1; k < max_conseq_chord; ++k) same = same && (hist(k) == prev);
    if (same) banned = prev;
  }
  float loc = 0.f;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int n = lane * PER + e;
    float pj = (n < vocab_limit) ? __expf(v[e] - mx) : 0.f;
    if ((max_conseq_N == 0 && n == 0) || n == banned) pj = 0.f;
    v[e] = pj;
    loc += pj;
  }
  float incl = loc;                                        // inclusive scan of the lane totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  const float target = u * total;
  float run = incl - loc;
  int pick = 0x7fffffff, last = -1;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    run += v[e];
    if (v[e] > 0.f) {
      last = lane * PER + e;
      if (run > target && pick == 0x7fffffff) pick = lane * PER + e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  return pick != 0x7fffffff ? pick : last;                 // rounding at the top end: the last admissible token
}

// chord id -> (root id, attribute id): dataset/vevo_meta/chord_inv.json + chord_root.json + chord_attr.json of the reference
// are the closed form below (checked against those JSON tables in the CPU tests); "N" and plain major names get attribute 1
// (video_music_transformer.py:1109-1116).
__device__ __forceinline__ void chord_root_attr(int c, long long& root, long long& attr) {
  if (c <= 0) { root = 0; attr = 1; return; }
  root = (c - 1) / 13 + 1;
  attr = (c - 1) % 13 + 1;
}

}  // namespace v2m
