// fp32 SIMT GEMM  C[M,N] = epi(A[M,K] * W[N,K]^T)  -- the exact-precision path.
//
// Replaces the F.linear calls of the reference on the fp32 parity path
// (model/rpr.py:253,277,417; model/video_music_transformer.py:1001,1022,1042; stock
// nn.TransformerEncoderLayer linears).  Plain FFMA with a fixed k order, so each output row is
// bit-identical whatever the batch size (batch invariance, needed for greedy-token parity).
#include "common.cuh"
#include "kernels.h"

namespace v2m {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool ALIGNED>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ W, int ldw,
                                                       float* __restrict__ C, int ldc, int M, int N, int K,
                                                       GemmEpilogue ep, int a_cs, int w_cs) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int lrow = tid >> 2;         // 0..63
  const int lk = (tid & 3) * 4;      // 0,4,8,12
  // global -> register loads of one k tile (two rows of A and of W per thread); issued for tile k0 + BK BEFORE the FMAs of
  // tile k0, so the memory latency hides behind the arithmetic (same k order and same products as a plain loop: the results
  // are bit-identical, only the waiting moved)
  float4 va[2], vb[2];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + 64 * i;
      va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      vb[i] = va[i];
      const int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      if (ALIGNED) {
        if (gm < M && gk < K) va[i] = __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk));
        if (gn < N && gk < K) vb[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk));
      } else {
        if (gm < M) {
          const float* p = A + (size_t)gm * lda + (size_t)gk * a_cs;      // element (m, k) at m*lda + k*a_cs
          if (gk + 0 < K) va[i].x = __ldg(p);
          if (gk + 1 < K) va[i].y = __ldg(p + (size_t)a_cs);
          if (gk + 2 < K) va[i].z = __ldg(p + 2 * (size_t)a_cs);
          if (gk + 3 < K) va[i].w = __ldg(p + 3 * (size_t)a_cs);
        }
        if (gn < N) {
          const float* p = W + (size_t)gn * ldw + (size_t)gk * w_cs;
          if (gk + 0 < K) vb[i].x = __ldg(p);
          if (gk + 1 < K) vb[i].y = __ldg(p + (size_t)w_cs);
          if (gk + 2 < K) vb[i].z = __ldg(p + 2 * (size_t)w_cs);
          if (gk + 3 < K) vb[i].w = __ldg(p + 3 * (size_t)w_cs);
        }
      }
    }
  };
  load_tile(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + 64 * i;
      As[lk + 0][r] = va[i].x; As[lk + 1][r] = va[i].y; As[lk + 2][r] = va[i].z; As[lk + 3][r] = va[i].w;
      Bs[lk + 0][r] = vb[i].x; Bs[lk + 1][r] = vb[i].y; Bs[lk + 2][r] = vb[i].z; Bs[lk + 3][r] = vb[i].w;
    }
    __syncthreads();
    if (k0 + BK < K) load_tile(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const uint32_t dseed = ep.drop_seed + (ep.drop_seed_dev ? *ep.drop_seed_dev : 0u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float rs = ep.row_scale ? ep.row_scale[m] : 0.f;
    const float* res = ep.residual ? ep.residual + (size_t)(ep.res_mod > 0 ? m % ep.res_mod : m) * ep.ldr : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[n];
      if (n < ep.alpha_cols) v *= ep.alpha;
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.row_scale) v = fmaf(rs, ep.col_vec[n], v);
      // training: inverted dropout with the same stateless mask as the bf16 epilogue (dy_prep recomputes it in backward)
      if (ep.drop_scale != 0.f && !ep.drop_after_res) v = drop_keep(dseed, m, n, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
      if (res) v += res[n];  // fp32 path: residual is always fp32
      if (ep.drop_scale != 0.f && ep.drop_after_res) v = drop_keep(dseed, m, n, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
      C[epi_out_index(ep, m, n, ldc)] = v;
    }
  }
}

int gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
             const GemmEpilogue& ep, cudaStream_t stream, int a_cs, int w_cs) {
  V2M_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_f32: bad dims M=%d N=%d K=%d", M, N, K);
  if (M == 0) return kOk;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  const bool aligned = (a_cs == 1) && (w_cs == 1) && (lda % 4 == 0) && (ldw % 4 == 0) && (K % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(W) % 16 == 0);
  if (aligned)
    gemm_f32_kernel<true><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep, 1, 1);
  else
    gemm_f32_kernel<false><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, ep, a_cs, w_cs);
  return check_launch("gemm_f32");
}

}  // namespace v2m
