// Mamba block kernels (fp32): depthwise causal conv + SiLU, fused selective scan, RMSNorm.
//
// The reference materialises deltaA = exp(delta * A) and BX = delta * B * x as (B, L, ED, N) tensors, scans them with
// pscan and contracts the result with C (model/mamba.py:333-351: 3 x 16 KB per token and direction through HBM).
// selective_scan_fwd_kernel keeps the N = 16 states of one channel in registers instead: it reads x, delta, z (B, L, ED)
// and B, C (B, L, N) once and writes y (B, L, ED) -- the fusion the reference itself points at through its optional
// mamba_ssm hook (mamba.py:308-317).  Thread <-> (batch, channel); a warp reads 128 contiguous bytes per tensor and step.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

// y[b][l][c] = silu(bias[c] + sum_k w[c][k] * x[b][l - (KW-1) + k][c])      (nn.Conv1d(groups=ED, padding=KW-1)[:, :, :L],
// mamba.py:270-274).  x / y: (B, L, ld) row-major views.
__global__ void __launch_bounds__(256) mamba_conv_silu_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y, long long ldy,
                                                              int B, int L, int ED, int KW) {
  const long long total = (long long)B * L * ED;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ED);
    const long long bl = i / ED;
    const int l = (int)(bl % L);
    float acc = bias ? bias[c] : 0.f;
    for (int k = 0; k < KW; ++k) {
      const int ls = l - (KW - 1) + k;
      if (ls >= 0) acc = fmaf(w[c * KW + k], x[(bl - l + ls) * ldx + c], acc);
    }
    y[bl * ldy + c] = acc / (1.f + expf(-acc));
  }
}

int mamba_conv_silu(const float* x, long long ldx, const float* w, const float* bias, float* y, long long ldy, int B, int L, int ED,
                    int KW, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && L > 0 && ED > 0 && KW >= 1 && KW <= 8, "mamba_conv_silu: bad dims B=%d L=%d ED=%d KW=%d", B, L, ED, KW);
  const long long total = (long long)B * L * ED;
  const long long want = (total + 255) / 256;
  mamba_conv_silu_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(x, ldx, w, bias, y, ldy, B, L, ED, KW);
  return check_launch("mamba_conv_silu");
}

// softplus with torch's threshold (F.softplus: x for x > 20)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// delta = softplus(delta_raw + dt_bias); h_n = exp(delta * A_n) h_n + delta * B_n * x; y = sum_n h_n C_n + D x;
// out = y * silu(z) (+ x * (1 - sigmoid(z)) for the "mamba+" variant, mamba.py:283-287).  A_n = -exp(A_log[c][n]).
template <int N>
__global__ void __launch_bounds__(128) selective_scan_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ delta_raw, long long ldd,
                                                                 const float* __restrict__ dt_bias, const float* __restrict__ A_log,
                                                                 const float* __restrict__ Bm, const float* __restrict__ Cm,
                                                                 long long ldbc, const float* __restrict__ Dp,
                                                                 const float* __restrict__ z, long long ldz, float* __restrict__ out,
                                                                 long long ldo, int L, int ED, int plus) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= ED) return;
  float A[N], h[N];
#pragma unroll
  for (int n = 0; n < N; ++n) { A[n] = -expf(A_log[(long long)c * N + n]); h[n] = 0.f; }
  const float Dc = Dp[c], db = dt_bias ? dt_bias[c] : 0.f;
  for (int l = 0; l < L; ++l) {
    const long long row = (long long)b * L + l;
    const float xv = x[row * ldx + c];
    const float dl = softplus_f(delta_raw[row * ldd + c] + db);
    const float dx = dl * xv;
    const float* Br = Bm + row * ldbc;
    const float* Cr = Cm + row * ldbc;
    float y = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      h[n] = fmaf(expf(dl * A[n]), h[n], dx * __ldg(Br + n));
      y = fmaf(h[n], __ldg(Cr + n), y);
    }
    y = fmaf(Dc, xv, y);
    float o = y;
    if (z) {
      const float zv = z[row * ldz + c];
      const float sg = 1.f / (1.f + expf(-zv));
      o = y * (zv * sg);
      if (plus) o = fmaf(xv, 1.f - 1.f / (1.f + expf(-(zv * sg))), o);     // x * (1 - sigmoid(silu(z)))
    }
    out[row * ldo + c] = o;
  }
}

int selective_scan_fwd(const float* x, long long ldx, const float* delta_raw, long long ldd, const float* dt_bias,
                       const float* A_log, const float* Bm, const float* Cm, long long ldbc, const float* Dp, const float* z,
                       long long ldz, float* out, long long ldo, int B, int L, int ED, int N, int plus, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && L > 0 && ED > 0, "selective_scan: bad dims B=%d L=%d ED=%d", B, L, ED);
  V2M_REQUIRE(N == 16, "selective_scan: d_state %d unsupported (16)", N);
  dim3 grid((ED + 127) / 128, B);
  selective_scan_fwd_kernel<16><<<grid, 128, 0, stream>>>(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, Dp, z, ldz, out, ldo,
                                                          L, ED, plus);
  return check_launch("selective_scan_fwd");
}

// y = x * rsqrt(mean(x^2) + eps) * w   (mamba.py:483-489); one warp per row
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                      int M, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * D;
  float sq = 0.f;
  for (int d = lane; d < D; d += 32) sq = fmaf(xr[d], xr[d], sq);
  const float r = rsqrtf(warp_sum(sq) / (float)D + eps);
  for (int d = lane; d < D; d += 32) y[(long long)row * D + d] = xr[d] * r * (w ? w[d] : 1.f);
}

int rmsnorm(const float* x, const float* w, float* y, int M, int D, float eps, cudaStream_t stream) {
  if (M == 0) return kOk;
  rmsnorm_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, w, y, M, D, eps);
  return check_launch("rmsnorm");
}

}  // namespace v2m
