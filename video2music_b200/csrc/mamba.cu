// Mamba block kernels (fp32): depthwise causal conv + SiLU, fused selective scan, RMSNorm.
//
// The reference materialises deltaA = exp(delta * A) and BX = delta * B * x as (B, L, ED, N) tensors, scans them with
// pscan and contracts the result with C (model/mamba.py:333-351: 3 x 16 KB per token and direction through HBM).
// selective_scan_fwd_kernel keeps the N = 16 states of one channel in registers instead: it reads x, delta, z (B, L, ED)
// and B, C (B, L, N) once and writes y (B, L, ED) -- the fusion the reference itself points at through its optional
// mamba_ssm hook (mamba.py:308-317).  Thread <-> (batch, channel); a warp reads 128 contiguous bytes per tensor and step.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

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
  // the steps are sequential, their inputs are not: the loads of 4 steps are issued together (one memory latency per 4 steps)
  for (int lg = l0; lg < l1; lg += 4) {
  float xq[4], dq[4], zq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long row = (long long)b * L + min(lg + j, l1 - 1);
    xq[j] = x[row * ldx + c];
    dq[j] = delta_raw[row * ldd + c];
    zq[j] = (PASS == 1 && z) ? z[row * ldz + c] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int l = lg + j;
    if (l >= l1) break;
    const long long row = (long long)b * L + l;
    const float xv = xq[j];
    const float dl = softplus_f(dq[j] + db);
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
        const float zv = zq[j];
        const float sg = 1.f / (1.f + expf(-zv));
        o = y * (zv * sg);
        if (plus) o = fmaf(xv, 1.f - 1.f / (1.f + expf(-(zv * sg))), o);     // x * (1 - sigmoid(silu(z)))
      }
      out[row * ldo + c] = o;
    }
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

// ---------------------------------------------------------------------------------------------------------------------------
// Backward of the fused selective scan (exact fp32 training path; the reference differentiates the materialised
// deltaA / BX tensors and the pscan through torch autograd, mamba.py:333-351 + pscan.py:196-226).
// Like the forward, L is cut into chunks of kScanChunk steps so that (video, channel, chunk) triples fill the GPU:
//   forward pass 0 + carry (the kernels above)            -> true state at every chunk start                          (wsf)
//   bwd pass 0   thread = (channel, chunk >= 1, video)    reverse sweep of the chunk from gh = 0: the gradient its own outputs
//                send to the state before the chunk, and sum(delta) of the chunk                                      (wsg)
//   bwd carry    thread = (channel, state, video)         G[ch-1] = exp(A_n sum_delta[ch]) G[ch] + local[ch], last chunk first:
//                G[ch] = dL/dh arriving at the END of chunk ch from all later chunks (in place)
//   bwd pass 1   thread = (channel, chunk, video), one CTA = 128 channels of one (video, chunk) walking in lockstep:
//        sweep 1 (l0 .. l1-1)   recompute the states from the chunk-start state, store them, Hs[b][l][n][c]
//        sweep 2 (l1-1 .. l0)   gh_n = dL/dh_l[n] carried in registers, starting from G[ch]; per step
//            y  = sum_n h_n C_n + D x;  out = y silu(z) (+ x (1 - sigmoid(silu(z))) for mamba+)
//            dC_n = dy h_n;  gh_n += dy C_n;  with a_n = exp(delta A_n):  d delta += gh_n (A_n a_n h_{l-1,n} + B_n x),
//            dA_n += gh_n h_{l-1,n} delta a_n,  dB_n = gh_n delta x,  dx += gh_n delta B_n,  gh_n *= a_n
//            d delta_raw = d delta * softplus'(delta_raw + dt_bias)
//   dB / dC are sums over the channels: every warp folds its 32 channels with a transpose-reduce butterfly (2 N = 32 values per
//   step) and adds them to the caller-zeroed gradient rows; dA_log (= dA * A), dD, d dt_bias are accumulated per thread and added once.
// wsf / wsg layout [B][n_chunks-1][N+1][ED] (row N = sum(delta)); wsg entry ch-1 belongs to chunk ch before the carry and holds
// G[ch-1] after it.
template <int N, int PASS>
__global__ void __launch_bounds__(128) selective_scan_bwd_kernel(const float* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ delta_raw, long long ldd,
                                                                 const float* __restrict__ dt_bias, const float* __restrict__ A_log,
                                                                 const float* __restrict__ Bm, const float* __restrict__ Cm,
                                                                 long long ldbc, const float* __restrict__ Dp,
                                                                 const float* __restrict__ z, long long ldz,
                                                                 const float* __restrict__ dout, long long ldo,
                                                                 const float* __restrict__ wsf, float* __restrict__ wsg,
                                                                 float* __restrict__ Hs, float* __restrict__ dx, long long lddx,
                                                                 float* __restrict__ ddraw, long long lddd, float* __restrict__ dBm,
                                                                 float* __restrict__ dCm, long long lddbc, float* __restrict__ dz,
                                                                 long long lddz, float* __restrict__ dA_log, float* __restrict__ dD,
                                                                 float* __restrict__ ddt_bias, int L, int ED, int n_chunks, int plus) {
  static_assert(2 * N == 32, "the channel fold maps 2 N values onto the 32 lanes of a warp");
  // B | C rows of the chunk (shared by every channel): staged once per CTA and read as shared-memory broadcasts -- per-lane
  // broadcast loads from global memory cost the load/store unit one request per value and step
  __shared__ __align__(16) float sBC[PASS == 1 ? kScanChunk : 1][2 * N];
  const int tid = threadIdx.x, lane = tid & 31, c = blockIdx.x * 128 + tid, b = blockIdx.z;
  const int ch = PASS == 0 ? blockIdx.y + 1 : blockIdx.y;            // pass 0 skips chunk 0 (nobody needs its start gradient)
  const bool act = c < ED;
  const int cc = act ? c : ED - 1;                                   // inactive lanes read a valid channel and contribute nothing
  const int l0 = ch * kScanChunk, l1 = min(L, l0 + kScanChunk);
  float A[N], A2[N], h[N], gh[N], gA[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    A[n] = -expf(A_log[(long long)cc * N + n]);
    A2[n] = A[n] * 1.4426950408889634f;
    h[n] = 0.f; gh[n] = 0.f; gA[n] = 0.f;
  }
  const float Dc = Dp[cc], db = dt_bias ? dt_bias[cc] : 0.f;
  if (PASS == 0) {
    float dsum = 0.f;
    for (int lg = l1 - 1; lg >= l0; lg -= 4) {                         // loads of 4 steps issued together
      float dq[4], gq[4], zq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long row = (long long)b * L + max(lg - j, l0);
        dq[j] = delta_raw[row * ldd + cc];
        gq[j] = dout[row * ldo + cc];
        zq[j] = z ? z[row * ldz + cc] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = lg - j;
        if (l < l0) break;
        const long long row = (long long)b * L + l;
        const float dl = softplus_f(dq[j] + db);
        float dy = gq[j];
        if (z) {
          const float zv = zq[j];
          dy *= zv / (1.f + expf(-zv));
        }
        dsum += dl;
#pragma unroll
        for (int n = 0; n < N; ++n) gh[n] = fmaf(dy, __ldg(Cm + row * ldbc + n), gh[n]) * ex2_approx(dl * A2[n]);
      }
    }
    if (act) {
      float* w = wsg + ((long long)b * (n_chunks - 1) + (ch - 1)) * (N + 1) * ED + c;
#pragma unroll
      for (int n = 0; n < N; ++n) w[(long long)n * ED] = gh[n];
      w[(long long)N * ED] = dsum;
    }
    return;
  }
  for (int i = tid; i < (l1 - l0) * 2 * N; i += 128) {
    const int li = i / (2 * N), v = i % (2 * N);
    const long long row = (long long)b * L + l0 + li;
    sBC[li][v] = v < N ? Bm[row * ldbc + v] : Cm[row * ldbc + (v - N)];
  }
  __syncthreads();
  const float* h0 = ch > 0 ? wsf + ((long long)b * (n_chunks - 1) + (ch - 1)) * (N + 1) * ED + cc : nullptr;   // chunk-start state
  if (ch > 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) h[n] = h0[(long long)n * ED];
  }
  if (ch < n_chunks - 1) {
    const float* w = wsg + ((long long)b * (n_chunks - 1) + ch) * (N + 1) * ED + cc;
#pragma unroll
    for (int n = 0; n < N; ++n) gh[n] = w[(long long)n * ED];
  }
  for (int l = l0; l < l1; ++l) {
    const long long row = (long long)b * L + l;
    const float xv = x[row * ldx + cc];
    const float dl = softplus_f(delta_raw[row * ldd + cc] + db);
    const float dxv = dl * xv;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      h[n] = fmaf(ex2_approx(dl * A2[n]), h[n], dxv * sBC[l - l0][n]);
      if (act) Hs[(row * N + n) * ED + c] = h[n];
    }
  }
  float gD = 0.f, gdb = 0.f;
  for (int l = l1 - 1; l >= l0; --l) {
    const long long row = (long long)b * L + l;
    const float xv = x[row * ldx + cc];
    const float raw = delta_raw[row * ldd + cc] + db;
    const float dl = softplus_f(raw);
    const float g = act ? dout[row * ldo + c] : 0.f;
    float y = Dc * xv;
#pragma unroll
    for (int n = 0; n < N; ++n) y = fmaf(h[n], sBC[l - l0][N + n], y);
    float dy = g, dzv = 0.f, dxv = 0.f;
    if (z) {
      const float zv = z[row * ldz + cc];
      const float sg = 1.f / (1.f + expf(-zv));
      const float sl = zv * sg, dsl = sg * (1.f + zv * (1.f - sg));
      dy = g * sl;
      dzv = g * y * dsl;
      if (plus) {
        const float s2 = 1.f / (1.f + expf(-sl));
        dxv = g * (1.f - s2);
        dzv -= g * xv * s2 * (1.f - s2) * dsl;
      }
    }
    dxv = fmaf(dy, Dc, dxv);
    gD = fmaf(dy, xv, gD);
    float ddl = 0.f;
    float fold[2 * N];                                             // this channel's dB_n | dC_n of the step
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float Bv = sBC[l - l0][n], Cv = sBC[l - l0][N + n];
      const float hp = (l > l0) ? Hs[((row - 1) * N + n) * ED + cc] : (ch > 0 ? h0[(long long)n * ED] : 0.f);
      const float a = ex2_approx(dl * A2[n]);
      fold[N + n] = act ? dy * h[n] : 0.f;                         // dC_n
      gh[n] = fmaf(dy, Cv, gh[n]);
      const float t = gh[n] * hp;
      ddl = fmaf(t, A[n] * a, ddl);
      ddl = fmaf(gh[n] * Bv, xv, ddl);
      gA[n] = fmaf(t, dl * a, gA[n]);
      fold[n] = act ? gh[n] * dl * xv : 0.f;                       // dB_n
      dxv = fmaf(gh[n] * dl, Bv, dxv);
      gh[n] *= a;
      h[n] = hp;
    }
    const float ddr = ddl * (raw > 20.f ? 1.f : 1.f / (1.f + expf(-raw)));
    gdb += ddr;
    if (act) {
      dx[row * lddx + c] = dxv;
      ddraw[row * lddd + c] = ddr;
      if (dz) dz[row * lddz + c] = dzv;
    }
    // dB / dC are sums over the channels: the 32 channels of a warp meet in a transpose-reduce butterfly (lane v ends with sum
    // #v), one reduction per warp and value -- no shared memory and no CTA barrier inside the sweep, warps run independently
    const float sum = warp_transpose_reduce32(fold, lane);
    atomicAdd((lane < N ? dBm + row * lddbc + lane : dCm + row * lddbc + (lane - N)), sum);
  }
  if (act) {
#pragma unroll
    for (int n = 0; n < N; ++n) atomicAdd(dA_log + (long long)c * N + n, gA[n] * A[n]);
    atomicAdd(dD + c, gD);
    if (ddt_bias) atomicAdd(ddt_bias + c, gdb);
  }
}

// bwd pass 1, second version (the one that runs; the kernel above with PASS == 1 is kept for A/B runs, V2M_SCAN_BWD_OLD=1):
// thread = (channel, STATE), CTA = 16 channels x 16 states of one (video, chunk).  One state per thread means the whole history of
// the chunk (64 values) fits in REGISTERS: the recomputed states never go to memory (the first version wrote and re-read
// 2 x 64 B per (token, channel) through HBM and waited for a dependent global load in every step of both sweeps).  Everything a
// step needs that depends on (step, channel) only -- delta = softplus(.), x, dy = g silu(z), softplus', the z / "mamba+" terms -- is
// computed ONCE per (step, channel) by the staging phase (not once per state) and parked in shared memory together with the
// chunk's B | C rows, so the two sweeps touch shared memory and registers only.  Sums over the states (d delta, dx, y) are
// 16-lane butterflies; dB / dC (sums over channels) are folded over the warp's two channels by one shuffle, over the CTA's 16
// channels by shared-memory reductions, and leave as one global reduction per (step, state) and CTA.
template <int N>
__global__ void __launch_bounds__(256, 2) selective_scan_bwd_state_kernel(
    const float* __restrict__ x, long long ldx, const float* __restrict__ delta_raw, long long ldd, const float* __restrict__ dt_bias,
    const float* __restrict__ A_log, const float* __restrict__ Bm, const float* __restrict__ Cm, long long ldbc,
    const float* __restrict__ Dp, const float* __restrict__ z, long long ldz, const float* __restrict__ dout, long long ldo,
    const float* __restrict__ wsf, const float* __restrict__ wsg, float* __restrict__ dx, long long lddx, float* __restrict__ ddraw,
    long long lddd, float* __restrict__ dBm, float* __restrict__ dCm, long long lddbc, float* __restrict__ dz, long long lddz,
    float* __restrict__ dA_log, float* __restrict__ dD, float* __restrict__ ddt_bias, int L, int ED, int n_chunks, int plus) {
  static_assert(N == 16, "lane = 16 x (channel & 1) + state");
  constexpr int CH = kScanChunk, CC = 16;                            // steps per chunk, channels per CTA
  // [array][step][channel | state].  The three per-(step, channel) outputs have the form  out = M * (sum over the states) + Q  with
  // M and Q known at staging time:  d delta_raw = softplus' * sum,   dx = 1 * sum + dx0,   dz = gz * sum_y + (gz D x + dz0);
  // the lane that ends up with the sum (states 0 / 4 / 8 of the channel) applies its (M, Q) pair and parks the result over M.
  __shared__ __align__(16) float sm[11][CH][CC];
  float (*sB)[CC] = sm[0], (*sC)[CC] = sm[1], (*s_dl)[CC] = sm[2], (*s_x)[CC] = sm[3], (*s_dy)[CC] = sm[4];
  float (*sM)[CH][CC] = &sm[5], (*sQ)[CH][CC] = &sm[8];              // role 0: d delta_raw, 1: dx, 2: dz
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = tid & 15, cl = tid >> 4, ch = blockIdx.y, b = blockIdx.z;
  const int c = blockIdx.x * CC + cl;
  const bool act = c < ED;
  const int cc = act ? c : ED - 1;
  const int l0 = ch * CH, len = min(L, l0 + CH) - l0;
  // ---- staging: one (step, channel) element per thread and round (all loads first); B | C rows
  {
    const int scl = tid & 15, sc = blockIdx.x * CC + scl;
    const bool sact = sc < ED;
    const float Dc = sact ? Dp[sc] : 0.f, db = (dt_bias && sact) ? dt_bias[sc] : 0.f;
    float vx[CH / 16], vr[CH / 16], vg[CH / 16], vz[CH / 16];
#pragma unroll
    for (int k = 0; k < CH / 16; ++k) {
      const int li = (tid >> 4) + 16 * k;
      const bool ok = li < len && sact;
      const long long row = (long long)b * L + l0 + (ok ? li : 0);
      const int scc = sact ? sc : 0;
      vx[k] = ok ? x[row * ldx + scc] : 0.f;
      vr[k] = ok ? delta_raw[row * ldd + scc] : 0.f;
      vg[k] = ok ? dout[row * ldo + scc] : 0.f;
      vz[k] = (ok && z) ? z[row * ldz + scc] : 0.f;
    }
    for (int i = tid; i < CH * N; i += 256) {
      const int li = i / N, v = i % N;
      const long long row = (long long)b * L + l0 + li;
      sB[li][v] = li < len ? Bm[row * ldbc + v] : 0.f;
      sC[li][v] = li < len ? Cm[row * ldbc + v] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < CH / 16; ++k) {
      const int li = (tid >> 4) + 16 * k;
      float dl = 0.f, xv = 0.f, dy = 0.f, spd = 0.f, gz = 0.f, dx0 = 0.f, dz0 = 0.f;
      if (li < len && sact) {
        xv = vx[k];
        const float raw = vr[k] + db;
        dl = softplus_f(raw);
        spd = raw > 20.f ? 1.f : 1.f / (1.f + expf(-raw));
        const float g = vg[k];
        dy = g;
        if (z) {
          const float zv = vz[k];
          const float sg = 1.f / (1.f + expf(-zv));
          const float sl = zv * sg, dsl = sg * (1.f + zv * (1.f - sg));
          dy = g * sl;
          gz = g * dsl;
          if (plus) {
            const float s2 = 1.f / (1.f + expf(-sl));
            dx0 = g * (1.f - s2);
            dz0 = -g * xv * s2 * (1.f - s2) * dsl;
          }
        }
        dx0 = fmaf(dy, Dc, dx0);
      }
      s_dl[li][scl] = dl; s_x[li][scl] = xv; s_dy[li][scl] = dy;
      sM[0][li][scl] = spd; sQ[0][li][scl] = 0.f;
      sM[1][li][scl] = 1.f; sQ[1][li][scl] = dx0;
      sM[2][li][scl] = gz;  sQ[2][li][scl] = fmaf(gz, Dc * xv, dz0);
    }
  }
  const float A = -expf(A_log[(long long)cc * N + n]), A2 = A * 1.4426950408889634f;
  const float h0 = (ch > 0) ? wsf[(((long long)b * (n_chunks - 1) + (ch - 1)) * (N + 1) + n) * ED + cc] : 0.f;
  float gh = (ch < n_chunks - 1) ? wsg[(((long long)b * (n_chunks - 1) + ch) * (N + 1) + n) * ED + cc] : 0.f;
  if (!act) gh = 0.f;
  __syncthreads();
  // ---- sweep 1: the states of the chunk, kept in registers
  float hist[CH];
  {
    float h = h0;
#pragma unroll
    for (int l = 0; l < CH; ++l) {
      const float dl = s_dl[l][cl];
      h = fmaf(ex2_approx(dl * A2), h, dl * s_x[l][cl] * sB[l][n]);   // steps past the chunk's end: dl = 0, x = 0 -> h unchanged
      hist[l] = h;
    }
  }
  // ---- sweep 2: last step first.  s_acc[warp][l][lane]: this warp's two-channel sum of dB_n (lanes 0..15) | dC_n (lanes 16..31)
  // of step l (shared memory: 64 more live registers per thread would halve the occupancy)
  extern __shared__ float s_acc[];                                   // [8 warps][CH steps][32 lanes]
  float* my_acc = s_acc + (size_t)warp * CH * 32 + lane;
  float gA = 0.f, gD = 0.f, gdb = 0.f;
  const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, up = (lane & 16) != 0;
  const int role = n >> 2;                                           // lanes with state 0 / 4 / 8 end up with the three sums
  const bool writer = (n & 3) == 0 && n < 12;
  float* pM = &sM[writer ? role : 0][0][cl];
  const float* pQ = &sQ[writer ? role : 0][0][cl];
#pragma unroll
  for (int l = CH - 1; l >= 0; --l) {
    if (l < len) {                                                   // uniform over the CTA
      if ((l & 7) == 7) asm volatile("" ::: "memory");              // keeps the unrolled steps from being interleaved too far (registers)
      const float dl = s_dl[l][cl], xv = s_x[l][cl], dy = s_dy[l][cl];
      const float Bv = sB[l][n], Cv = sC[l][n];
      const float hl = hist[l], hp = l > 0 ? hist[l > 0 ? l - 1 : 0] : h0;
      const float a = ex2_approx(dl * A2);
      gh = fmaf(dy, Cv, gh);
      const float t = gh * hp;
      const float p_ddl = fmaf(t, A * a, gh * Bv * xv);
      gA = fmaf(t, dl * a, gA);
      const float p_dx = gh * dl * Bv;
      const float p_y = hl * Cv;
      const float dBc = gh * dl * xv, dCc = dy * hl;
      gh *= a;
      gD = fmaf(dy, xv, gD);
      // sums over the 16 states of a channel as a transpose-reduce: 5 shuffles for the three sums (instead of 12); they end in
      // the lanes with state bits (b3, b2) = (0, 0): d delta, (0, 1): dx, (1, x): y
      const float g0 = __shfl_xor_sync(0xffffffffu, b3 ? p_ddl : p_y, 8);
      const float g1 = __shfl_xor_sync(0xffffffffu, p_dx, 8);
      const float k0 = (b3 ? p_y : p_ddl) + g0;                      // b3: y          else: d delta
      const float k1 = p_dx + g1;                                    //                else: dx
      float r = (b3 ? k0 : (b2 ? k1 : k0)) + __shfl_xor_sync(0xffffffffu, b3 ? k0 : (b2 ? k0 : k1), 4);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      // dB | dC over the two channels of the warp: one exchange, lanes 0..15 keep dB, lanes 16..31 keep dC
      my_acc[l * 32] = (up ? dCc : dBc) + __shfl_xor_sync(0xffffffffu, up ? dBc : dCc, 16);
      if (writer) {
        const float o = fmaf(pM[l * CC], r, pQ[l * CC]);
        gdb += o;                                                    // meaningful in the role-0 lane only
        pM[l * CC] = o;                                              // outputs parked in place, written coalesced below
      }
    }
  }
  if (act) {
    atomicAdd(dA_log + (long long)c * N + n, gA * A);
    if (n == 0) {
      atomicAdd(dD + c, gD);
      if (ddt_bias) atomicAdd(ddt_bias + c, gdb);
    }
  }
  __syncthreads();
  {
    const int scl = tid & 15, sc = blockIdx.x * CC + scl;
    if (sc < ED) {
#pragma unroll
      for (int k = 0; k < CH / 16; ++k) {
        const int li = (tid >> 4) + 16 * k;
        if (li < len) {
          const long long row = (long long)b * L + l0 + li;
          ddraw[row * lddd + sc] = sM[0][li][scl];
          dx[row * lddx + sc] = sM[1][li][scl];
          if (dz) dz[row * lddz + sc] = sM[2][li][scl];
        }
      }
    }
  }
  // ---- dB | dC over the CTA's 16 channels (8 warps), then one global reduction per (step, state) and CTA
  for (int i = tid; i < len * 32; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_acc[(size_t)w * CH * 32 + i];
    const int li = i >> 5, v = i & 31;
    const long long row = (long long)b * L + l0 + li;
    atomicAdd(v < N ? dBm + row * lddbc + v : dCm + row * lddbc + (v - N), s);
  }
}

// G[ch-1] = exp(A_n sum_delta[ch]) G[ch] + local[ch] for ch = n_chunks-1 .. 1, in place (entry ch-1 of wsg), G[n_chunks-1] = 0
template <int N>
__global__ void __launch_bounds__(128) selective_scan_bwd_carry_kernel(const float* __restrict__ A_log, float* __restrict__ wsg, int ED,
                                                                       int n_chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y, b = blockIdx.z;
  if (c >= ED) return;
  const float A2 = -expf(A_log[(long long)c * N + n]) * 1.4426950408889634f;
  float G = 0.f;
  for (int ch = n_chunks - 1; ch >= 1; --ch) {
    float* w = wsg + ((long long)b * (n_chunks - 1) + (ch - 1)) * (N + 1) * ED + c;
    G = fmaf(ex2_approx(A2 * w[(long long)N * ED]), G, w[(long long)n * ED]);
    w[(long long)n * ED] = G;
  }
}

long long selective_scan_bwd_workspace(int B, int L, int ED, int N) {
  return (long long)B * L * ED * N * (long long)sizeof(float) + 2 * selective_scan_workspace(B, L, ED, N);
}

int selective_scan_bwd(const float* x, long long ldx, const float* delta_raw, long long ldd, const float* dt_bias, const float* A_log,
                       const float* Bm, const float* Cm, long long ldbc, const float* Dp, const float* z, long long ldz,
                       const float* dout, long long ldo, float* hs, long long hs_bytes, float* dx, long long lddx, float* ddraw,
                       long long lddd, float* dBm, float* dCm, long long lddbc, float* dz, long long lddz, float* dA_log, float* dD,
                       float* ddt_bias, int B, int L, int ED, int N, int plus, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && L > 0 && ED > 0, "selective_scan_bwd: bad dims B=%d L=%d ED=%d", B, L, ED);
  V2M_REQUIRE(N == 16, "selective_scan_bwd: d_state %d unsupported (16)", N);
  V2M_REQUIRE(hs && hs_bytes >= selective_scan_bwd_workspace(B, L, ED, N), "selective_scan_bwd: workspace of %lld bytes needed",
              selective_scan_bwd_workspace(B, L, ED, N));
  const int n_chunks = (L + kScanChunk - 1) / kScanChunk;
  V2M_REQUIRE(n_chunks <= 65535 && B <= 65535, "selective_scan_bwd: L=%d / B=%d too large for the grid", L, B);
  const long long ws_floats = selective_scan_workspace(B, L, ED, N) / (long long)sizeof(float);
  float* wsf = hs;                         // chunk-start states (forward pass 0 + carry)
  float* wsg = hs + ws_floats;             // chunk-end gradients
  float* Hs = hs + 2 * ws_floats;          // all states
  const int gx = (ED + 127) / 128;
  if (n_chunks > 1) {
    const bool vec = ldbc % 4 == 0 && reinterpret_cast<uintptr_t>(Bm) % 16 == 0 && reinterpret_cast<uintptr_t>(Cm) % 16 == 0;
    if (vec)
      selective_scan_fwd_kernel<16, 0, true><<<dim3(gx, n_chunks - 1, B), 128, 0, stream>>>(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm,
                                                                                            ldbc, Dp, z, ldz, nullptr, 0, wsf, L, ED,
                                                                                            n_chunks, plus);
    else
      selective_scan_fwd_kernel<16, 0, false><<<dim3(gx, n_chunks - 1, B), 128, 0, stream>>>(x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm,
                                                                                             ldbc, Dp, z, ldz, nullptr, 0, wsf, L, ED,
                                                                                             n_chunks, plus);
    selective_scan_carry_kernel<16><<<dim3(gx, 16, B), 128, 0, stream>>>(A_log, wsf, ED, n_chunks);
    selective_scan_bwd_kernel<16, 0><<<dim3(gx, n_chunks - 1, B), 128, 0, stream>>>(
        x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, Dp, z, ldz, dout, ldo, wsf, wsg, Hs, dx, lddx, ddraw, lddd, dBm, dCm, lddbc,
        dz, lddz, dA_log, dD, ddt_bias, L, ED, n_chunks, plus);
    selective_scan_bwd_carry_kernel<16><<<dim3(gx, 16, B), 128, 0, stream>>>(A_log, wsg, ED, n_chunks);
  }
  static int use_old = -1;
  if (use_old < 0) { const char* e = getenv("V2M_SCAN_BWD_OLD"); use_old = (e && atoi(e) == 1) ? 1 : 0; }
  if (use_old)
    selective_scan_bwd_kernel<16, 1><<<dim3(gx, n_chunks, B), 128, 0, stream>>>(
        x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, Dp, z, ldz, dout, ldo, wsf, wsg, Hs, dx, lddx, ddraw, lddd, dBm, dCm, lddbc, dz,
        lddz, dA_log, dD, ddt_bias, L, ED, n_chunks, plus);
  {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(selective_scan_bwd_state_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           8 * kScanChunk * 32 * 4);
      if (e != cudaSuccess) { set_last_error("selective_scan_bwd: smem attribute: %s", cudaGetErrorString(e)); return kCudaError; }
      attr = true;
    }
  }
  if (!use_old)
    selective_scan_bwd_state_kernel<16><<<dim3((ED + 15) / 16, n_chunks, B), 256, 8 * kScanChunk * 32 * 4, stream>>>(
        x, ldx, delta_raw, ldd, dt_bias, A_log, Bm, Cm, ldbc, Dp, z, ldz, dout, ldo, wsf, wsg, dx, lddx, ddraw, lddd, dBm, dCm, lddbc, dz,
        lddz, dA_log, dD, ddt_bias, L, ED, n_chunks, plus);
  return check_launch("selective_scan_bwd");
}

// Backward of mamba_conv_silu: thread <-> (channel, chunk of kScanChunk steps, video).  The thread walks its chunk plus a
// (KW-1)-step halo on the right with a KW-wide sliding window of inputs (primed with the KW-1 inputs before the chunk) and of
// pending input gradients: position p collects w[k] * dpre from the outputs p .. p+KW-1, so every position of the chunk is
// complete once the halo has been visited; contributions that fall before the chunk belong to the previous chunk's thread
// (which recomputes them in its own halo).  dw / dbias are summed over the chunk's own outputs per thread and added once
// (caller-zeroed buffers).
template <int KW>
__global__ void __launch_bounds__(128) mamba_conv_silu_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, const float* __restrict__ dy,
                                                                  long long ldy, float* __restrict__ dx, long long lddx,
                                                                  float* __restrict__ dw, float* __restrict__ dbias, int L, int ED) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y, b = blockIdx.z;
  if (c >= ED) return;
  const int l0 = ch * kScanChunk, l1 = min(L, l0 + kScanChunk), lh = min(L, l1 + KW - 1);
  float wk[KW], win[KW], acc[KW], gw[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) { wk[k] = w[c * KW + k]; acc[k] = 0.f; gw[k] = 0.f; }
  // window before the first step: win[k] will hold x[l - (KW-1) + k] after the shift at step l
#pragma unroll
  for (int k = 1; k < KW; ++k) {
    const int ls = l0 - KW + k;                                      // becomes win[k-1] at step l0
    win[k] = ls >= 0 ? x[((long long)b * L + ls) * ldx + c] : 0.f;
  }
  win[0] = 0.f;
  const float bs = bias ? bias[c] : 0.f;
  float gb = 0.f;
  for (int l = l0; l < lh; ++l) {
    const long long row = (long long)b * L + l;
#pragma unroll
    for (int k = 0; k < KW - 1; ++k) { win[k] = win[k + 1]; acc[k] = acc[k + 1]; }
    win[KW - 1] = x[row * ldx + c];
    acc[KW - 1] = 0.f;
    float pre = bs;
#pragma unroll
    for (int k = 0; k < KW; ++k) pre = fmaf(wk[k], win[k], pre);
    const float sg = 1.f / (1.f + expf(-pre));
    const float dpre = dy[row * ldy + c] * sg * (1.f + pre * (1.f - sg));
    if (l < l1) {                                                    // own outputs only: the halo belongs to the next chunk
      gb += dpre;
#pragma unroll
      for (int k = 0; k < KW; ++k) gw[k] = fmaf(dpre, win[k], gw[k]);
    }
#pragma unroll
    for (int k = 0; k < KW; ++k) acc[k] = fmaf(wk[k], dpre, acc[k]);
    const int p = l - (KW - 1);                                      // acc[0] is complete now
    if (p >= l0 && p < l1) dx[(row - (KW - 1)) * lddx + c] = acc[0];
  }
  // positions whose later outputs do not exist (end of the sequence): acc[k] belongs to position lh - 1 - (KW-1) + k
#pragma unroll
  for (int k = 1; k < KW; ++k) {
    const int p = lh - 1 - (KW - 1) + k;
    if (p >= l0 && p < l1) dx[((long long)b * L + p) * lddx + c] = acc[k];
  }
#pragma unroll
  for (int k = 0; k < KW; ++k) atomicAdd(dw + c * KW + k, gw[k]);
  if (dbias) atomicAdd(dbias + c, gb);
}

int mamba_conv_silu_bwd(const float* x, long long ldx, const float* w, const float* bias, const float* dy, long long ldy, float* dx,
                        long long lddx, float* dw, float* dbias, int B, int L, int ED, int KW, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && L > 0 && ED > 0 && B <= 65535, "mamba_conv_silu_bwd: bad dims B=%d L=%d ED=%d", B, L, ED);
  V2M_REQUIRE(KW == 4, "mamba_conv_silu_bwd: d_conv %d unsupported (4)", KW);
  const int n_chunks = (L + kScanChunk - 1) / kScanChunk;
  V2M_REQUIRE(n_chunks <= 65535, "mamba_conv_silu_bwd: L=%d too large for the grid", L);
  mamba_conv_silu_bwd_kernel<4><<<dim3((ED + 127) / 128, n_chunks, B), 128, 0, stream>>>(x, ldx, w, bias, dy, ldy, dx, lddx, dw, dbias, L, ED);
  return check_launch("mamba_conv_silu_bwd");
}

// ---- recurrent single-token step (MambaBlock.step / ssm_step, mamba.py:407-470) -------------------------------------------
// One new token per video: the cache holds the last KW-1 inputs of the depthwise convolution (B, ED, KW-1) and the SSM
// state h (B, ED, N).  Two kernels around the three small GEMMs (in_proj, x_proj, out_proj) of the step:
//   mamba_step_conv : xs = silu(conv1d([inputs | x_new])[KW-1]) and the shifted input window          (mamba.py:419-425,435)
//   mamba_step_ssm  : delta = softplus(dt_proj(delta_raw)); h = exp(delta A) h + delta B x; y = h C + D x; out = y silu(z)
// Nothing is modified in place: the caller's cache tensors stay valid, as in the reference (which builds new tensors).
__global__ void __launch_bounds__(256) mamba_step_conv_kernel(const float* __restrict__ xz, long long ldxz, const float* __restrict__ in_old,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ xs, float* __restrict__ in_new, int B, int ED, int KW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * ED) return;
  const int b = i / ED, e = i - b * ED;
  const float xn = xz[(long long)b * ldxz + e];
  const float* io = in_old + (long long)i * (KW - 1);
  float* in = in_new + (long long)i * (KW - 1);
  float acc = bias ? bias[e] : 0.f;
  for (int j = 0; j < KW - 1; ++j) {
    const float v = io[j];
    acc = fmaf(w[e * KW + j], v, acc);
    if (j > 0) in[j - 1] = v;
  }
  acc = fmaf(w[e * KW + KW - 1], xn, acc);
  if (KW > 1) in[KW - 2] = xn;
  xs[i] = acc / (1.f + expf(-acc));
}

template <int N>
__global__ void __launch_bounds__(128) mamba_step_ssm_kernel(const float* __restrict__ xs, const float* __restrict__ dbc, long long lddbc,
                                                             const float* __restrict__ dtw, const float* __restrict__ dtb,
                                                             const float* __restrict__ A_log, const float* __restrict__ D,
                                                             const float* __restrict__ z, long long ldz, const float* __restrict__ h_old,
                                                             float* __restrict__ h_new, float* __restrict__ out, int B, int ED, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * ED) return;
  const int b = i / ED, e = i - b * ED;
  const float* row = dbc + (long long)b * lddbc;            // [delta_raw (R) | B (N) | C (N)]
  float dr = dtb[e];
  for (int r = 0; r < R; ++r) dr = fmaf(dtw[e * R + r], row[r], dr);
  const float delta = softplus_f(dr);
  const float x = xs[i];
  float y = 0.f;
#pragma unroll
  for (int n4 = 0; n4 < N; n4 += 4) {
    float4 hv = h_old ? *reinterpret_cast<const float4*>(h_old + (long long)i * N + n4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 al = *reinterpret_cast<const float4*>(A_log + (long long)e * N + n4);
    float hh[4] = {hv.x, hv.y, hv.z, hv.w};
    const float aa[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dA = expf(delta * -expf(aa[k]));
      hh[k] = fmaf(dA, hh[k], delta * row[R + n4 + k] * x);
      y = fmaf(hh[k], row[R + N + n4 + k], y);
    }
    *reinterpret_cast<float4*>(h_new + (long long)i * N + n4) = make_float4(hh[0], hh[1], hh[2], hh[3]);
  }
  y = fmaf(D[e], x, y);
  const float zz = z[(long long)b * ldz + e];
  out[i] = y * (zz / (1.f + expf(-zz)));
}

int mamba_step_conv(const float* xz, long long ldxz, const float* in_old, const float* w, const float* bias, float* xs, float* in_new,
                    int B, int ED, int KW, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && ED > 0 && KW >= 1 && KW <= 8, "mamba_step_conv: bad dims B=%d ED=%d KW=%d", B, ED, KW);
  mamba_step_conv_kernel<<<(B * ED + 255) / 256, 256, 0, stream>>>(xz, ldxz, in_old, w, bias, xs, in_new, B, ED, KW);
  return check_launch("mamba_step_conv");
}

int mamba_step_ssm(const float* xs, const float* dbc, long long lddbc, const float* dtw, const float* dtb, const float* A_log,
                   const float* D, const float* z, long long ldz, const float* h_old, float* h_new, float* out, int B, int ED, int N,
                   int R, cudaStream_t stream) {
  V2M_REQUIRE(B > 0 && ED > 0 && R >= 1, "mamba_step_ssm: bad dims B=%d ED=%d R=%d", B, ED, R);
  V2M_REQUIRE(N == 16, "mamba_step_ssm: d_state %d unsupported (16)", N);
  mamba_step_ssm_kernel<16><<<(B * ED + 127) / 128, 128, 0, stream>>>(xs, dbc, lddbc, dtw, dtb, A_log, D, z, ldz, h_old, h_new, out, B, ED, R);
  return check_launch("mamba_step_ssm");
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
