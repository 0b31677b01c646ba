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
//
// The recurrence is sequential along L only; (video, channel) pairs alone give too few threads (8 x 256 at the L = 4096
// stress shape), so L is cut into chunks of kScanChunk steps and scanned in three passes:
//   pass 0  thread = (channel, chunk, video): chunk-local end state from h = 0, and sum(delta) of the chunk    -> ws
//   carry   thread = (channel, state, video): h_end[ch] = exp(A_n * sum_delta[ch]) * h_end[ch-1] + local[ch]  (in place;
//           the decay of a whole chunk has the closed form prod_t exp(delta_t A_n) = exp(A_n * sum_t delta_t))
//   pass 1  thread = (channel, chunk, video): rescan the chunk from its true start state and emit the outputs
// ws layout [B][n_chunks-1][N+1][ED] fp32 (channel fastest: coalesced), row N = sum(delta).  The inputs are read twice
// (x, delta, B, C: ~2 KB per step for ED = 256) instead of streaming three (B,L,ED,N) tensors (48 KB per step).
constexpr int kScanChunk = 64;

template <int N, int PASS, bool VEC>
__global__ void __launch_bounds__(128) selective_scan_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ delta_raw, long long ldd,
                                                                 const float* __restrict__ dt_bias, const float* __restrict__ A_log,
                                                                 const float* __restrict__ Bm, const float* __restrict__ Cm,
                                                                 long long ldbc, const float* __restrict__ Dp,
                                                                 const float* __restrict__ z, long long ldz, float* __restrict__ out,
                                                                 long long ldo, float* __restrict__ ws, int L, int ED, int n_chunks,
                                                                 int plus) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y, b = blockIdx.z;
  if (c >= ED) return;
  const int l0 = ch * kScanChunk, l1 = min(L, l0 + kScanChunk);
  float A2[N], h[N];                       // A2 = A * log2(e): exp(delta * A) = 2^(delta * A2)
#pragma unroll
  for (int n = 0; n < N; ++n) { A2[n] = -expf(A_log[(long long)c * N + n]) * 1.4426950408889634f; h[n] = 0.f; }
  if (PASS == 1 && ch > 0) {
    const float* w = ws + ((long long)b * (n_chunks - 1) + (ch - 1)) * (N + 1) * ED + c;
#pragma unroll
    for (int n = 0; n < N; ++n) h[n] = w[(long long)n * ED];
  }
  const float Dc = Dp[c], db = dt_bias ? dt_bias[c] : 0.f;
  float dsum = 0.f;
  for (int l = l0; l < l1; ++l) {
    const long long row = (long long)b * L + l;
    const float xv = x[row * ldx + c];
    const float dl = softplus_f(delta_raw[row * ldd + c] + db);
    const float dx = dl * xv;
    float Bv[N], Cv[N];                    // the step's B and C rows are the same for every channel: broadcast loads
    if (VEC) {
      const float4* Br = reinterpret_cast<const float4*>(Bm + row * ldbc);
      const float4* Cr = reinterpret_cast<const float4*>(Cm + row * ldbc);
#pragma unroll
      for (int q = 0; q < N / 4; ++q) {
        const float4 bq = __ldg(Br + q);
        Bv[4 * q] = bq.x; Bv[4 * q + 1] = bq.y; Bv[4 * q + 2] = bq.z; Bv[4 * q + 3] = bq.w;
        if (PASS == 1) {
          const float4 cq = __ldg(Cr + q);
          Cv[4 * q] = cq.x; Cv[4 * q + 1] = cq.y; Cv[4 * q + 2] = cq.z; Cv[4 * q + 3] = cq.w;
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < N; ++n) { Bv[n] = __ldg(Bm + row * ldbc + n); Cv[n] = PASS == 1 ? __ldg(Cm + row * ldbc + n) : 0.f; }
    }
    float y = 0.f;
    dsum += dl;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      h[n] = fmaf(ex2_approx(dl * A2[n]), h[n], dx * Bv[n]);
      if (PASS == 1) y = fmaf(h[n], Cv[n], y);
    }
    if (PASS == 1) {
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
  if (PASS == 0) {
    float* w = ws + ((long long)b * (n_chunks - 1) + ch) * (N + 1) * ED + c;
#pragma unroll
    for (int n = 0; n < N; ++n) w[(long long)n * ED] = h[n];
    w[(long long)N * ED] = dsum;
  }
}

template <int N>
__global__ void __launch_bounds__(128) selective_scan_carry_kernel(const float* __restrict__ A_log, float* __restrict__ ws, int ED,
                                                                   int n_chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y, b = blockIdx.z;
  if (c >= ED) return;
  const float A2 = -expf(A_log[(long long)c * N + n]) * 1.4426950408889634f;
  float h = 0.f;
  for (int ch = 0; ch < n_chunks - 1; ++ch) {
    float* w = ws + ((long long)b * (n_chunks - 1) + ch) * (N + 1) * ED + c;
    h = fmaf(ex2_approx(A2 * w[(long long)N * ED]), h, w[(long long)n * ED]);
    w[(long long)n * ED] = h;
  }
}

long long selective_scan_workspace(int B, int L, int ED, int N) {
  const int n_chunks = (L + kScanChunk - 1) / kScanChunk;
  return (long long)B * (n_chunks - 1) * (N + 1) * ED * (long long)sizeof(float);
}

int selective_scan_fwd(const float* x, long long ldx, const float* delta_raw, long long ldd, const float* dt_bias,
                       const float* A_log, const float* Bm, const float* Cm, long long ldbc, const float* Dp, const float* z,
                       long long ldz, float* out, long long ldo, int B, int L, int ED, int N, int plus, float* ws,
                       long long ws_bytes, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && L > 0 && ED > 0, "selective_scan: bad dims B=%d L=%d ED=%d", B, L, ED);
  V2M_REQUIRE(N == 16, "selective_scan: d_state %d unsupported (16)", N);
  const int n_chunks = (L + kScanChunk - 1) / kScanChunk;
  V2M_REQUIRE(n_chunks == 1 || (ws && ws_bytes >= selective_scan_workspace(B, L, ED, N)),
              "selective_scan: workspace of %lld bytes needed (v2m_selective_scan_workspace)", selective_scan_workspace(B, L, ED, N));
  V2M_REQUIRE(n_chunks <= 65535 && B <= 65535, "selective_scan: L=%d / B=%d too large for the grid", L, B);
  const int gx = (ED + 127) / 128;
  const bool vec = ldbc % 4 == 0 && reinterpret_cast<uintptr_t>(Bm) % 16 == 0 && reinterpret_cast<uintptr_t>(Cm) % 16 == 0;
#define V2M_SCAN(PASS, VEC, GY)                                                                                                  \
  selective_scan_fwd_kernel<16, PASS, VEC><<<dim3(gx, GY, B), 128, 0, stream>>>(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, \
                                                                               Dp, z, ldz, out, ldo, ws, L, ED, n_chunks, plus)
  if (n_chunks > 1) {
    if (vec) V2M_SCAN(0, true, n_chunks - 1); else V2M_SCAN(0, false, n_chunks - 1);
    selective_scan_carry_kernel<16><<<dim3(gx, 16, B), 128, 0, stream>>>(A_log, ws, ED, n_chunks);
  }
  if (vec) V2M_SCAN(1, true, n_chunks); else V2M_SCAN(1, false, n_chunks);
#undef V2M_SCAN
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

// backward of rmsnorm: with r = rsqrt(mean(x^2) + eps), g = w o dy:  dx = r g - x r^3 mean(g o x);  dw += dy o x r  (fp32 atomics
// into a caller-zeroed buffer, rows folded per CTA first).  One warp per row, 8 rows per CTA.
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dw,
                                                          int M, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * D;
  const float* gr = dy + (long long)row * D;
  float sq = 0.f, dot = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float xv = xr[d];
    sq = fmaf(xv, xv, sq);
    dot = fmaf(gr[d] * (w ? w[d] : 1.f), xv, dot);
  }
  const float r = rsqrtf(warp_sum(sq) / (float)D + eps);
  const float c = warp_sum(dot) * r * r * r / (float)D;
  for (int d = lane; d < D; d += 32) {
    const float xv = xr[d], gv = gr[d];
    dx[(long long)row * D + d] = r * gv * (w ? w[d] : 1.f) - xv * c;
    if (dw) atomicAdd(dw + d, gv * xv * r);
  }
}

int rmsnorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int M, int D, float eps, cudaStream_t stream) {
  if (M == 0) return kOk;
  rmsnorm_bwd_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, w, dy, dx, dw, M, D, eps);
  return check_launch("rmsnorm_bwd");
}

}  // namespace v2m
