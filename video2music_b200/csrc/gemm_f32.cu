// fp32 SIMT GEMM  C[M,N] = epi(A[M,K] * W[N,K]^T)  -- the exact-precision path.
//
// Replaces the F.linear calls of the reference on the fp32 parity path
// (model/rpr.py:253,277,417; model/video_music_transformer.py:1001,1022,1042; stock
// nn.TransformerEncoderLayer linears).  Plain FFMA with a fixed k order, so each output row is
// bit-identical whatever the batch size (batch invariance, needed for greedy-token parity).
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------------------------------
// Skinny shapes: one of N, K is tiny (Mamba's x_proj 256 -> 40 and dt_proj 8 -> 256 and their input gradients; M = videos x
// steps rows).  The 128 x 128 tiles above turn them into 150 CTAs that mostly multiply padding (85-112 us per launch at
// M = 19200).  Here a CTA owns 64 rows and ALL columns: W^T sits in shared memory as [k][n] for the whole launch and is read as
// 16-byte broadcasts; a thread = (row, one of four column groups) with NB accumulators; the rows' inputs are staged through
// shared memory in 64-wide k chunks (coalesced global reads whichever of m / k is the contiguous direction, then conflict-free
// [k][row] reads) and the results leave through a shared-memory tile so that global writes are contiguous along n.  Fixed k order
// per output, independent of M: batch-invariant like the tiled kernel.  (A first version with a row's inputs read straight from
// global memory by its thread was slower than the tiles: 32 different lines per load instruction.)
constexpr int RK_ROWS = 64, RK_KC = 64, RK_GROUPS = 4, RK_THREADS = RK_ROWS * RK_GROUPS;
template <int NB>
__global__ void __launch_bounds__(RK_THREADS) gemm_f32_rows_kernel(const float* __restrict__ A, int lda, int a_cs,
                                                                  const float* __restrict__ W, int ldw, int w_cs, float* __restrict__ C,
                                                                  int ldc, int M, int N, int K, GemmEpilogue ep, int npad) {
  extern __shared__ __align__(16) float rk_smem[];
  float* Ws = rk_smem;                                               // [K][npad]: Ws[k][n] = W[n][k], zero for n >= N
  float* xs = Ws + (size_t)K * npad;                                 // [RK_KC][RK_ROWS + 1]
  float* ys = xs + RK_KC * (RK_ROWS + 1);                            // [RK_GROUPS][RK_ROWS][NB + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * RK_ROWS;
  if (w_cs == 1) {                                                   // k contiguous in W: consecutive threads walk k (coalesced reads)
    for (int i = tid; i < K * npad; i += RK_THREADS) {
      const int n = i / K, k = i - n * K;
      Ws[(size_t)k * npad + n] = n < N ? __ldg(W + (size_t)n * ldw + k) : 0.f;
    }
  } else {                                                           // transposed view of W: n contiguous
    for (int i = tid; i < K * npad; i += RK_THREADS) {
      const int k = i / npad, n = i - k * npad;
      Ws[i] = n < N ? __ldg(W + (size_t)n * ldw + (size_t)k * w_cs) : 0.f;
    }
  }
  const int row = tid & (RK_ROWS - 1), grp = tid / RK_ROWS;          // column group of this thread
  const int m = m0 + row;
  const uint32_t dseed = ep.drop_seed + (ep.drop_seed_dev ? *ep.drop_seed_dev : 0u);
  const int n_passes = (npad + RK_GROUPS * NB - 1) / (RK_GROUPS * NB);   // > 1 only when K <= RK_KC (one chunk, staged once)
  float acc[NB];
  for (int kc0 = 0; kc0 < K; kc0 += RK_KC) {
    const int kc = min(RK_KC, K - kc0);
    __syncthreads();                                                 // previous chunk consumed (and Ws complete)
    if (a_cs == 1) {                                                 // k contiguous: a warp reads 32 k of one row
      for (int r = warp; r < RK_ROWS; r += RK_THREADS / 32)
        for (int k = lane; k < kc; k += 32)
          xs[k * (RK_ROWS + 1) + r] = (m0 + r < M) ? __ldg(A + (size_t)(m0 + r) * lda + kc0 + k) : 0.f;
    } else {                                                         // transposed view: rows contiguous, a warp reads 32 rows of one k
      for (int k = warp; k < kc; k += RK_THREADS / 32)
        for (int r = lane; r < RK_ROWS; r += 32)
          xs[k * (RK_ROWS + 1) + r] = (m0 + r < M) ? __ldg(A + (size_t)(m0 + r) * lda + (size_t)(kc0 + k) * a_cs) : 0.f;
    }
    __syncthreads();
    for (int pass = 0; pass < n_passes; ++pass) {
      const int n0 = (pass * RK_GROUPS + grp) * NB;
      if (kc0 == 0 || n_passes > 1) {
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.f;
      }
      if (n0 < npad) {
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
          const float xv = xs[k * (RK_ROWS + 1) + row];
          const float4* wr = reinterpret_cast<const float4*>(Ws + (size_t)(kc0 + k) * npad + n0);
#pragma unroll
          for (int j = 0; j < NB / 4; ++j) {
            const float4 w = wr[j];
            acc[4 * j] = fmaf(xv, w.x, acc[4 * j]);         acc[4 * j + 1] = fmaf(xv, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(xv, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(xv, w.w, acc[4 * j + 3]);
          }
        }
      }
      if (kc0 + RK_KC < K) continue;                                 // more k chunks to go (single pass in that case)
      // ---- epilogue by the owning thread, then a contiguous copy-out of both column groups
      if (n0 < npad && m < M) {
        const float rs = ep.row_scale ? ep.row_scale[m] : 0.f;
        const float* res = ep.residual ? ep.residual + (size_t)(ep.res_mod > 0 ? m % ep.res_mod : m) * ep.ldr : nullptr;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int n = n0 + j;
          float v = acc[j];
          if (n < N) {
            if (ep.bias) v += ep.bias[n];
            if (n < ep.alpha_cols) v *= ep.alpha;
            if (ep.relu) v = fmaxf(v, 0.f);
            if (ep.row_scale) v = fmaf(rs, ep.col_vec[n], v);
            if (ep.drop_scale != 0.f && !ep.drop_after_res) v = drop_keep(dseed, m, n, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
            if (res) v += res[n];
            if (ep.drop_scale != 0.f && ep.drop_after_res) v = drop_keep(dseed, m, n, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
          }
          ys[((size_t)grp * RK_ROWS + row) * (NB + 1) + j] = v;
        }
      }
      __syncthreads();
      for (int g2 = 0; g2 < RK_GROUPS; ++g2) {
        const int nb0 = (pass * RK_GROUPS + g2) * NB;
        const int nbv = min(NB, N - nb0);
        if (nbv <= 0) continue;
        for (int idx = tid; idx < RK_ROWS * nbv; idx += RK_THREADS) {
          const int r = idx / nbv, j = idx - r * nbv;
          if (m0 + r < M) C[epi_out_index(ep, m0 + r, nb0 + j, ldc)] = ys[((size_t)g2 * RK_ROWS + r) * (NB + 1) + j];
        }
      }
      if (pass + 1 < n_passes) __syncthreads();                      // ys is rewritten by the next pass
    }
  }
}

static size_t rows_smem(int K, int npad, int nb) {
  return ((size_t)K * npad + RK_KC * (RK_ROWS + 1) + RK_GROUPS * RK_ROWS * (nb + 1)) * sizeof(float);
}

template <int NB>
static int launch_rows(const float* A, int lda, int a_cs, const float* W, int ldw, int w_cs, float* C, int ldc, int M, int N, int K,
                       const GemmEpilogue& ep, int npad, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f32_rows_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) { set_last_error("gemm_f32(rows): smem attribute: %s", cudaGetErrorString(e)); return kCudaError; }
    attr = true;
  }
  gemm_f32_rows_kernel<NB><<<(M + RK_ROWS - 1) / RK_ROWS, RK_THREADS, rows_smem(K, npad, NB), stream>>>(A, lda, a_cs, W, ldw, w_cs, C, ldc,
                                                                                                    M, N, K, ep, npad);
  return check_launch("gemm_f32(rows)");
}

int gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
             const GemmEpilogue& ep, cudaStream_t stream, int a_cs, int w_cs) {
  V2M_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_f32: bad dims M=%d N=%d K=%d", M, N, K);
  if (M == 0) return kOk;
  {
    // skinny shapes (see gemm_f32_rows_kernel): tiny N or tiny K, many rows, W^T fits in shared memory.  V2M_GEMM_ROWS=0: off.
    static int rows_on = -1;
    if (rows_on < 0) { const char* e = getenv("V2M_GEMM_ROWS"); rows_on = (e && e[0] == '0') ? 0 : 1; }
    const int nb = N <= 32 ? 8 : 16;                                 // accumulators per thread; four column groups per CTA
    const int npad = ((N + nb - 1) / nb) * nb;
    // more than one pass over the columns needs the whole K staged at once (K <= 64)
    const bool shape_ok = (N <= RK_GROUPS * nb && K <= 1024) || K <= RK_KC;
    if (rows_on && shape_ok && M >= 2048 && N <= 1024 && rows_smem(K, npad, nb) <= 100 * 1024) {
      if (nb == 8) return launch_rows<8>(A, lda, a_cs, W, ldw, w_cs, C, ldc, M, N, K, ep, npad, stream);
      return launch_rows<16>(A, lda, a_cs, W, ldw, w_cs, C, ldc, M, N, K, ep, npad, stream);
    }
  }
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
