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


// ---- bf16 tensor-core training path of the experts (gradients of moe.py:44-49,191-199 on the grouped tcgen05 GEMMs) ----------
// swiglu_pair_bwd: a [M, 2 ff] = (x W1^T + b1 | x Wg^T + bg) as written by the grouped GEMM, dh [M, ff] -> dag [M, 2 ff]:
//   d a1 = dh * silu(g),  d g = dh * a1 * sigmoid(g) * (1 + g * (1 - sigmoid(g)))
__global__ void __launch_bounds__(256) swiglu_pair_bwd_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ dh,
                                                                   bf16* __restrict__ dag, long long M, int ff) {
  const long long n2 = M * (long long)(ff / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (ff / 2);
    const int c = (int)(i - r * (ff / 2)) * 2;
    const float2 a1 = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(a + r * 2 * ff + c));
    const float2 g = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(a + r * 2 * ff + ff + c));
    const float2 d = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(dh + r * ff + c));
    const float s0 = 1.f / (1.f + expf(-g.x)), s1 = 1.f / (1.f + expf(-g.y));
    *reinterpret_cast<uint32_t*>(dag + r * 2 * ff + c) = f2_to_bf16x2(d.x * g.x * s0, d.y * g.y * s1);
    *reinterpret_cast<uint32_t*>(dag + r * 2 * ff + ff + c) =
        f2_to_bf16x2(d.x * a1.x * s0 * (1.f + g.x * (1.f - s0)), d.y * a1.y * s1 * (1.f + g.y * (1.f - s1)));
  }
}

int swiglu_pair_bwd_bf16(const void* a, const void* dh, void* dag, long long M, int ff, cudaStream_t stream) {
  V2M_REQUIRE(ff > 0 && ff % 2 == 0, "swiglu_pair_bwd_bf16: ff=%d must be even", ff);
  if (M == 0) return kOk;
  const long long want = (M * (ff / 2) + 255) / 256;
  swiglu_pair_bwd_bf16_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(static_cast<const bf16*>(a), static_cast<const bf16*>(dh),
                                                                                            static_cast<bf16*>(dag), M, ff);
  return check_launch("swiglu_pair_bwd_bf16");
}

// out[g][n] = sum over the rows of group g of x[row][n] (bias gradients of the experts): grid (column blocks of 64, row
// splits, groups), group bounds read on the device, partial sums added to the zero-initialised output.
__global__ void __launch_bounds__(256) moe_group_colsum_bf16_kernel(const bf16* __restrict__ x, long long ldx, const int* __restrict__ off,
                                                                    float* __restrict__ out, int N) {
  const int g = blockIdx.z;
  const int r0 = off[g], r1 = off[g + 1];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;           // 32 column pairs x 8 row lanes
  const int n = blockIdx.x * 64 + 2 * cx;
  float s0 = 0.f, s1 = 0.f;
  for (int r = r0 + blockIdx.y * 8 + ry; r < r1 && n < N; r += gridDim.y * 8) {
    const float2 v = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(x + (long long)r * ldx + n));
    s0 += v.x; s1 += v.y;
  }
  __shared__ float part[8][66];
  part[ry][2 * cx] = s0; part[ry][2 * cx + 1] = s1;
  __syncthreads();
  if (ry == 0) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0 += part[i][2 * cx]; a1 += part[i][2 * cx + 1]; }
    if (n < N && a0 != 0.f) atomicAdd(out + (long long)g * N + n, a0);
    if (n + 1 < N && a1 != 0.f) atomicAdd(out + (long long)g * N + n + 1, a1);
  }
}

int moe_group_colsum_bf16(const void* x, long long ldx, const int* off, int n_groups, float* out, int N, int rows_hint, cudaStream_t stream) {
  V2M_REQUIRE(N > 0 && N % 2 == 0 && n_groups > 0 && off && out, "moe_group_colsum_bf16: bad arguments");
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)n_groups * N * 4, stream);
  if (e != cudaSuccess) { set_last_error("moe_group_colsum_bf16: memset: %s", cudaGetErrorString(e)); return kCudaError; }
  int splits = (rows_hint / n_groups + 255) / 256;
  splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
  moe_group_colsum_bf16_kernel<<<dim3((N + 63) / 64, splits, n_groups), 256, 0, stream>>>(static_cast<const bf16*>(x), ldx, off, out, N);
  return check_launch("moe_group_colsum_bf16");
}

// ------------------------------------------------------------------------------------------------------------------
// Expert dispatch without host round trips (the reference loops over experts in Python with torch.where per expert,
// model/moe.py:192-199):
//   moe_offsets   : exclusive scan of the router's histogram -> off[E+1]; clears the per-expert cursors
//   moe_permute   : token copy (t, r) goes to row off[e] + cursor[e]++ of the expert-contiguous matrix xp; perm[t*k+r] = row
//                   (row order inside a group is arbitrary; every row's result is independent of its position)
//   moe_grouped_* : C[row] = epi(xp[row] . W_e^T) for the rows of group e = blockIdx.z, with the group bounds read from
//                   off[] ON THE DEVICE (ragged groups, grid sized for the worst case, surplus CTAs exit at once).
//                   glu != 0: two weight stacks, C = (x W1^T + b1) * silu(x Wg^T + bg)   (GLUExpert, moe.py:44-49)
//   moe_combine   : out[t] = sum_r w[t][r] * yp[perm[t*k+r]]  in rank order (deterministic)        (moe.py:196-199)
// align > 1 (tensor-core path): every group starts on an `align`-row boundary and tile_group[t] names the group that owns
// row tile t of `align` rows (-1 for the unused tail), which is all the grouped tcgen05 GEMM needs to pick its weight tile.
__global__ void moe_offsets_kernel(const int* __restrict__ hist, int n_experts, int align, int* __restrict__ off,
                                   int* __restrict__ cursor, int* __restrict__ tile_group, int n_tiles) {
  __shared__ int s_off[kMaxExperts + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < n_experts; ++e) {
      off[e] = acc; s_off[e] = acc; cursor[e] = 0;
      acc += (hist[e] + align - 1) / align * align;
    }
    off[n_experts] = acc; s_off[n_experts] = acc;
  }
  __syncthreads();
  if (tile_group)
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
      const int r = t * align;
      int g = -1;
      for (int e = 0; e < n_experts; ++e)
        if (r >= s_off[e] && r < s_off[e] + hist[e]) g = e;
      tile_group[t] = g;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) moe_permute_kernel(const float* __restrict__ x, const long long* __restrict__ idx,
                                                          const int* __restrict__ off, int* __restrict__ cursor,
                                                          T* __restrict__ xp, int* __restrict__ perm, int tokens, int k, int d) {
  const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (item >= tokens * k) return;
  const int t = item / k;
  int row = 0;
  if (lane == 0) {
    const int e = (int)idx[item];
    row = off[e] + atomicAdd(cursor + e, 1);
    perm[item] = row;
  }
  row = __shfl_sync(0xffffffffu, row, 0);
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)t * d);
  if (sizeof(T) == 4) {
    float4* dst = reinterpret_cast<float4*>(xp + (size_t)row * d);
    for (int i = lane; i < d / 4; i += 32) dst[i] = src[i];
  } else {
    uint2* dst = reinterpret_cast<uint2*>(xp + (size_t)row * d);
    for (int i = lane; i < d / 4; i += 32) {
      const float4 v = src[i];
      dst[i] = make_uint2(f2_to_bf16x2(v.x, v.y), f2_to_bf16x2(v.z, v.w));
    }
  }
}

// h[m][j] = a[m][j] * silu(a[m][ff + j]) over the [M, 2 ff] bf16 output of the stacked (linear1 | gate) GEMM  (moe.py:46-47)
__global__ void __launch_bounds__(256) swiglu_pair_bf16_kernel(const bf16* __restrict__ a, bf16* __restrict__ h, long long M, int ff) {
  const long long n8 = M * (ff / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (ff / 8);
    const int j = (int)(i - m * (ff / 8)) * 8;
    const uint4 va = *reinterpret_cast<const uint4*>(a + m * 2 * ff + j);
    const uint4 vg = *reinterpret_cast<const uint4*>(a + m * 2 * ff + ff + j);
    const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, ug[4] = {vg.x, vg.y, vg.z, vg.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 fa = bf16x2_to_f2(ua[q]), fg = bf16x2_to_f2(ug[q]);
      o[q] = f2_to_bf16x2(fa.x * (fg.x / (1.f + expf(-fg.x))), fa.y * (fg.y / (1.f + expf(-fg.y))));
    }
    *reinterpret_cast<uint4*>(h + m * ff + j) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

constexpr int GB = 64, GK = 16;       // 64 x 64 output tile, 256 threads, 4 x 4 outputs each
template <bool GLU, bool VEC>
__global__ void __launch_bounds__(256) moe_grouped_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W1,
                                                               const float* __restrict__ b1, const float* __restrict__ Wg,
                                                               const float* __restrict__ bg, long long w_gstride,
                                                               long long b_gstride, const int* __restrict__ off,
                                                               float* __restrict__ C, int ldc, int N, int K) {
  __shared__ __align__(16) float As[GK][GB + 4];
  __shared__ __align__(16) float Bs[GK][GB + 4];
  __shared__ __align__(16) float Gs[GLU ? GK : 1][GB + 4];
  const int e = blockIdx.z;
  const int m_begin = off[e], m_end = off[e + 1];
  const int m0 = m_begin + blockIdx.y * GB, n0 = blockIdx.x * GB;
  if (m0 >= m_end) return;
  const float* W1e = W1 + (size_t)e * w_gstride;
  const float* Wge = GLU ? Wg + (size_t)e * w_gstride : nullptr;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4], accg[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; accg[i][j] = 0.f; }
  for (int k0 = 0; k0 < K; k0 += GK) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va, vg = va;
    const int gm = m0 + lrow, gn = n0 + lrow, gk = k0 + lk;
    if (VEC) {
      if (gm < m_end) va = __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk));
      if (gn < N) {
        vb = __ldg(reinterpret_cast<const float4*>(W1e + (size_t)gn * K + gk));
        if (GLU) vg = __ldg(reinterpret_cast<const float4*>(Wge + (size_t)gn * K + gk));
      }
    } else {                                        // odd K / leading dimensions (d_ff = 2 d_model + 1 experts): guarded scalars
      float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f}, tg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (gk + i < K) {
          if (gm < m_end) ta[i] = __ldg(A + (size_t)gm * lda + gk + i);
          if (gn < N) {
            tb[i] = __ldg(W1e + (size_t)gn * K + gk + i);
            if (GLU) tg[i] = __ldg(Wge + (size_t)gn * K + gk + i);
          }
        }
      }
      va = make_float4(ta[0], ta[1], ta[2], ta[3]); vb = make_float4(tb[0], tb[1], tb[2], tb[3]); vg = make_float4(tg[0], tg[1], tg[2], tg[3]);
    }
    As[lk + 0][lrow] = va.x; As[lk + 1][lrow] = va.y; As[lk + 2][lrow] = va.z; As[lk + 3][lrow] = va.w;
    Bs[lk + 0][lrow] = vb.x; Bs[lk + 1][lrow] = vb.y; Bs[lk + 2][lrow] = vb.z; Bs[lk + 3][lrow] = vb.w;
    if (GLU) { Gs[lk + 0][lrow] = vg.x; Gs[lk + 1][lrow] = vg.y; Gs[lk + 2][lrow] = vg.z; Gs[lk + 3][lrow] = vg.w; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (GLU) {
        const float4 g4 = *reinterpret_cast<const float4*>(&Gs[k][tx * 4]);
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) accg[i][j] = fmaf(a[i], gg[j], accg[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= m_end) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (b1 ? b1[(size_t)e * b_gstride + n] : 0.f);
      if (GLU) {
        const float gv = accg[i][j] + (bg ? bg[(size_t)e * b_gstride + n] : 0.f);
        v = v * (gv / (1.f + expf(-gv)));
      }
      C[(size_t)m * ldc + n] = v;
    }
  }
}

__global__ void __launch_bounds__(256) moe_combine_kernel(const float* __restrict__ yp, const int* __restrict__ perm,
                                                          const float* __restrict__ w, float* __restrict__ out, int tokens, int k, int d) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= tokens) return;
  for (int i = lane; i < d / 4; i += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
      const float wr = w[(size_t)t * k + r];
      const float4 v = reinterpret_cast<const float4*>(yp + (size_t)perm[(size_t)t * k + r] * d)[i];
      acc.x = fmaf(wr, v.x, acc.x); acc.y = fmaf(wr, v.y, acc.y); acc.z = fmaf(wr, v.z, acc.z); acc.w = fmaf(wr, v.w, acc.w);
    }
    reinterpret_cast<float4*>(out + (size_t)t * d)[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the expert dispatch (the reference gets these from autograd over the per-expert index_put / where loop,
// model/moe.py:190-199):
//   moe_combine_bwd : dyp[perm[t,r]] = w[t,r] * dout[t];  dw[t,r] = <dout[t], yp[perm[t,r]]>;  softmax-over-top-k backward
//                     scattered into the dense gate-logit gradient: dlogits[t, idx[t,r]] = scale * w_r (dw_r - sum_s w_s dw_s)
//   swiglu_bwd      : h = a * silu(g)  ->  dag[:, :ff] = dh * silu(g),  dag[:, ff:] = dh * a * sig(g) (1 + g (1 - sig(g)))
//   moe_grouped_dw  : dW_e[n][j] = sum_{rows m of group e} dY[m][n] * X[m][j],  db_e[n] = sum_m dY[m][n]   (ragged groups, bounds
//                     read from off[] on the device; an empty group writes zeros)
__global__ void __launch_bounds__(256) moe_combine_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ yp,
                                                              const int* __restrict__ perm, const float* __restrict__ w,
                                                              const long long* __restrict__ idx, float scale, int tokens, int k, int d,
                                                              int n_experts, float* __restrict__ dyp, float* __restrict__ dlogits) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= tokens) return;
  const float4* g4 = reinterpret_cast<const float4*>(dout + (size_t)t * d);
  float dw[kMaxExperts];
#pragma unroll
  for (int r = 0; r < kMaxExperts; ++r) {
    dw[r] = 0.f;
    if (r < k) {
      const int row = perm[(size_t)t * k + r];
      const float wr = w[(size_t)t * k + r];
      const float4* y4 = reinterpret_cast<const float4*>(yp + (size_t)row * d);
      float4* o4 = reinterpret_cast<float4*>(dyp + (size_t)row * d);
      float acc = 0.f;
      for (int i = lane; i < d / 4; i += 32) {
        const float4 g = g4[i], y = y4[i];
        acc = fmaf(g.x, y.x, acc); acc = fmaf(g.y, y.y, acc); acc = fmaf(g.z, y.z, acc); acc = fmaf(g.w, y.w, acc);
        o4[i] = make_float4(wr * g.x, wr * g.y, wr * g.z, wr * g.w);
      }
      dw[r] = warp_sum(acc);
    }
  }
  if (lane != 0) return;
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < kMaxExperts; ++r)
    if (r < k) s = fmaf(w[(size_t)t * k + r], dw[r], s);
  for (int e = 0; e < n_experts; ++e) dlogits[(size_t)t * n_experts + e] = 0.f;
#pragma unroll
  for (int r = 0; r < kMaxExperts; ++r)
    if (r < k) dlogits[(size_t)t * n_experts + (int)idx[(size_t)t * k + r]] = scale * w[(size_t)t * k + r] * (dw[r] - s);
}

__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const float* __restrict__ a, const float* __restrict__ g,
                                                         const float* __restrict__ dh, float* __restrict__ dag, long long M, int ff) {
  const long long n = M * ff;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / ff;
    const int j = (int)(i - m * ff);
    const float gv = g[i], av = a[i], dv = dh[i];
    const float sg = 1.f / (1.f + expf(-gv));
    dag[m * 2 * ff + j] = dv * gv * sg;
    dag[m * 2 * ff + ff + j] = dv * av * sg * (1.f + gv * (1.f - sg));
  }
}

__global__ void __launch_bounds__(256) moe_grouped_dw_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx,
                                                             const int* __restrict__ off, int total_rows, int splits,
                                                             float* __restrict__ dW, float* __restrict__ db, int N, int K) {
  // off == NULL: one group of total_rows rows.  splits > 1: the rows of a group are cut into `splits` ranges (blockIdx.z = e *
  // splits + s) whose partial sums are ADDED into the caller-zeroed outputs (split-K for tall, narrow weight gradients).
  __shared__ __align__(16) float Ys[GK][GB + 4];
  __shared__ __align__(16) float Xs[GK][GB + 4];
  const int e = blockIdx.z / splits;
  int m_begin = off ? off[e] : 0, m_end = off ? off[e + 1] : total_rows;
  if (splits > 1) {
    const int per = ((m_end - m_begin + splits - 1) / splits + GK - 1) / GK * GK;
    m_begin += (blockIdx.z % splits) * per;
    m_end = min(m_end, m_begin + per);
  }
  const int n0 = blockIdx.y * GB, k0 = blockIdx.x * GB;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lm = tid >> 4, lc = (tid & 15) * 4;          // loader: row lm of the 16-row chunk, 4 consecutive columns
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;
  for (int m0 = m_begin; m0 < m_end; m0 += GK) {
    const int m = m0 + lm;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + lc + c, kk = k0 + lc + c;
      Ys[lm][lc + c] = (m < m_end && n < N) ? __ldg(dY + (size_t)m * ldy + n) : 0.f;
      Xs[lm][lc + c] = (m < m_end && kk < K) ? __ldg(X + (size_t)m * ldx + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < GK; ++r) {
      const float4 y4 = *reinterpret_cast<const float4*>(&Ys[r][ty * 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
      const float y[4] = {y4.x, y4.y, y4.z, y4.w}, x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(y[i], x[j], acc[i][j]);
    }
    if (db && blockIdx.x == 0 && tid < GB) {
#pragma unroll
      for (int r = 0; r < GK; ++r) bacc += Ys[r][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + tx * 4 + j;
      if (kk < K) {
        if (splits > 1) atomicAdd(dW + ((size_t)e * N + n) * K + kk, acc[i][j]);
        else dW[((size_t)e * N + n) * K + kk] = acc[i][j];
      }
    }
  }
  if (db && blockIdx.x == 0 && tid < GB && n0 + tid < N) {
    if (splits > 1) atomicAdd(db + (size_t)e * N + n0 + tid, bacc);
    else db[(size_t)e * N + n0 + tid] = bacc;
  }
}

int moe_permute(const float* x, const long long* idx, const int* hist, int tokens, int k, int d, int n_experts, int align, int* off,
                int* cursor, void* xp, int xp_bf16, int* perm, int* tile_group, int n_tiles, cudaStream_t stream) {
  V2M_REQUIRE(d % 4 == 0 && n_experts >= 1 && n_experts <= kMaxExperts && k >= 1 && align >= 1, "moe_permute: bad dims d=%d E=%d k=%d",
              d, n_experts, k);
  moe_offsets_kernel<<<1, 256, 0, stream>>>(hist, n_experts, align, off, cursor, tile_group, n_tiles);
  if (tokens > 0) {
    const long long warps = (long long)tokens * k;
    const int grid = (int)((warps + 7) / 8);
    if (xp_bf16) moe_permute_kernel<bf16><<<grid, 256, 0, stream>>>(x, idx, off, cursor, static_cast<bf16*>(xp), perm, tokens, k, d);
    else moe_permute_kernel<float><<<grid, 256, 0, stream>>>(x, idx, off, cursor, static_cast<float*>(xp), perm, tokens, k, d);
  }
  return check_launch("moe_permute");
}

int swiglu_pair_bf16(const void* a, void* h, long long M, int ff, cudaStream_t stream) {
  V2M_REQUIRE(ff % 8 == 0, "swiglu_pair: ff=%d must be a multiple of 8", ff);
  if (M == 0) return kOk;
  const long long want = (M * (ff / 8) + 255) / 256;
  swiglu_pair_bf16_kernel<<<(int)(want < 148 * 8 ? want : 148 * 8), 256, 0, stream>>>(static_cast<const bf16*>(a), static_cast<bf16*>(h), M, ff);
  return check_launch("swiglu_pair_bf16");
}

int moe_grouped_gemm(const float* A, int lda, const float* W1, const float* b1, const float* Wg, const float* bg, long long w_gstride,
                     long long b_gstride, const int* off, int n_experts, int max_rows, float* C, int ldc, int N, int K,
                     cudaStream_t stream) {
  V2M_REQUIRE(K > 0 && N > 0, "moe_grouped_gemm: bad dims N=%d K=%d", N, K);
  if (max_rows == 0) return kOk;
  dim3 grid((N + GB - 1) / GB, (max_rows + GB - 1) / GB, n_experts);
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const bool vec = K % 16 == 0 && lda % 4 == 0 && al16(A) && al16(W1) && al16(Wg) && w_gstride % 4 == 0;
  // a generation step (one token per video): weight-streaming kernel over the whole chip instead of mostly empty 64 x 64 tiles
  if (vec && max_rows <= 256) return step_moe_gemm_f32(A, lda, W1, b1, Wg, bg, w_gstride, b_gstride, off, n_experts, C, ldc, N, K, stream);
#define V2M_GG(GLU, VEC) moe_grouped_gemm_kernel<GLU, VEC><<<grid, 256, 0, stream>>>(A, lda, W1, b1, Wg, bg, w_gstride, b_gstride, off, C, ldc, N, K)
  if (Wg) { if (vec) V2M_GG(true, true); else V2M_GG(true, false); }
  else { if (vec) V2M_GG(false, true); else V2M_GG(false, false); }
#undef V2M_GG
  return check_launch("moe_grouped_gemm");
}

int moe_combine(const float* yp, const int* perm, const float* w, float* out, int tokens, int k, int d, cudaStream_t stream) {
  V2M_REQUIRE(d % 4 == 0, "moe_combine: d=%d must be a multiple of 4", d);
  if (tokens == 0) return kOk;
  moe_combine_kernel<<<(tokens + 7) / 8, 256, 0, stream>>>(yp, perm, w, out, tokens, k, d);
  return check_launch("moe_combine");
}

int moe_combine_bwd(const float* dout, const float* yp, const int* perm, const float* w, const long long* idx, float scale, int tokens,
                    int k, int d, int n_experts, float* dyp, float* dlogits, cudaStream_t stream) {
  V2M_REQUIRE(d % 4 == 0 && k >= 1 && k <= kMaxExperts && n_experts >= k && n_experts <= kMaxExperts,
              "moe_combine_bwd: bad dims d=%d k=%d E=%d", d, k, n_experts);
  if (tokens == 0) return kOk;
  moe_combine_bwd_kernel<<<(tokens + 7) / 8, 256, 0, stream>>>(dout, yp, perm, w, idx, scale, tokens, k, d, n_experts, dyp, dlogits);
  return check_launch("moe_combine_bwd");
}

int swiglu_bwd(const float* a, const float* g, const float* dh, float* dag, long long M, int ff, cudaStream_t stream) {
  V2M_REQUIRE(ff > 0, "swiglu_bwd: ff=%d", ff);
  if (M == 0) return kOk;
  const long long want = (M * ff + 255) / 256;
  swiglu_bwd_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(a, g, dh, dag, M, ff);
  return check_launch("swiglu_bwd");
}

int moe_grouped_dw(const float* dY, int ldy, const float* X, int ldx, const int* off, int n_experts, float* dW, float* db, int N, int K,
                   cudaStream_t stream) {
  V2M_REQUIRE(K > 0 && N > 0 && n_experts >= 1, "moe_grouped_dw: bad dims N=%d K=%d E=%d", N, K, n_experts);
  dim3 grid((K + GB - 1) / GB, (N + GB - 1) / GB, n_experts);
  moe_grouped_dw_kernel<<<grid, 256, 0, stream>>>(dY, ldy, X, ldx, off, 0, 1, dW, db, N, K);
  return check_launch("moe_grouped_dw");
}

// dW[N][K] = dY[rows][N]^T X[rows][K] (+ db[N] = column sums of dY) for ONE dense layer on the fp32 path: the same tile kernel
// (operands read along their contiguous dimension), rows split over enough CTAs to fill the GPU, partial sums added atomically.
int dw_f32(const float* dY, int ldy, const float* X, int ldx, int rows, float* dW, float* db, int N, int K, cudaStream_t stream) {
  V2M_REQUIRE(K > 0 && N > 0 && rows >= 0, "dw_f32: bad dims N=%d K=%d rows=%d", N, K, rows);
  const int tiles = ((K + GB - 1) / GB) * ((N + GB - 1) / GB);
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int max_splits = (rows + 255) / 256;                       // at least 256 rows per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 1) {
    cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, stream);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, stream);
  }
  dim3 grid((K + GB - 1) / GB, (N + GB - 1) / GB, splits);
  moe_grouped_dw_kernel<<<grid, 256, 0, stream>>>(dY, ldy, X, ldx, nullptr, rows, splits, dW, db, N, K);
  return check_launch("dw_f32");
}

}  // namespace v2m
