// Selective-scan recurrence  H[t] = A[t] * H[t-1] + X[t]  along L of (B, L, D, N) fp32 tensors,
// forward and backward -- replaces PScan.forward / PScan.backward (model/pscan.py:154-226), i.e. the
// Blelloch up/down sweep the reference runs as ~2*log2(L) strided in-place PyTorch kernels over
// tensors padded to the next power of two (pscan.py:15-35,170-176).  No padding is needed here.
//
// HBM-bound (12 B per element forward, 20 B backward), so the layout is all about coalescing:
// the (D, N) plane is contiguous, so a lane owns VEC consecutive channels (a warp touches
// 128*VEC contiguous bytes per time step) and the L axis is split into one chunk per warp:
//   1. every warp scans its chunk with a zero carry-in and records the chunk aggregate
//      (prod A, local H_end);
//   2. the per-chunk aggregates are combined by an inclusive warp-shuffle scan with the operator
//      (P2,h2) o (P1,h1) = (P1*P2, P2*h1 + h2), lane <-> chunk, giving every chunk's carry-in;
//   3. the chunk is replayed with its carry-in and H is written.
// Short sequences (chunk <= 32 steps, e.g. L = 300 with 10 warps) keep the chunk's A and X in
// registers between 1 and 3, so each input byte is read from HBM exactly once; long sequences
// (L = 4096) re-read the chunk, which is served by L2 (a CTA's working set is L * 256 * VEC bytes).
//
// Backward (pscan.py:191-226): gX[t] = gH[t] + A[t+1] * gX[t+1] is the same recurrence run from the
// end with A shifted by one step (:218), and gA[t] = H[t-1] * gX[t], gA[0] = 0 (:223-224).
#include "common.cuh"
#include "kernels.h"

namespace v2m {

constexpr int PS_MAXREG = 32;   // chunk length kept in registers

template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float type; };
template <> struct VecT<4> { typedef float4 type; };

template <int VEC> __device__ __forceinline__ void ldv(const float* p, float* o);
template <> __device__ __forceinline__ void ldv<1>(const float* p, float* o) { o[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<4>(const float* p, float* o) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <int VEC> __device__ __forceinline__ void stv(float* p, const float* o);
template <> __device__ __forceinline__ void stv<1>(float* p, const float* o) { *p = o[0]; }
template <> __device__ __forceinline__ void stv<4>(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

// BWD == false: a(t) = A[t],   x(t) = X[t],  out H[t]
// BWD == true : runs over u = L-1-t; a(u) = A[t+1] (0 at t = L-1), x(u) = gH[t]; out gX[t], gA[t] = H[t-1]*gX[t]
template <int VEC, bool REG, bool BWD>
__global__ void __launch_bounds__(REG ? 640 : 1024) pscan_kernel(const float* __restrict__ A, const float* __restrict__ X,
                                                     const float* __restrict__ Hprev, float* __restrict__ out,
                                                     float* __restrict__ out_gA, int L, int DN, int chunk) {
  extern __shared__ float sm[];                 // [2][nw][32*VEC]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cpb = 32 * VEC;                     // channels per block
  const int blocks_per_batch = DN / cpb;
  const int b = blockIdx.x / blocks_per_batch;
  const int c0 = (blockIdx.x % blocks_per_batch) * cpb + lane * VEC;
  const size_t base = (size_t)b * L * DN + c0;
  const int u0 = warp * chunk, u1 = min(L, u0 + chunk);

  float ra[REG ? PS_MAXREG : 1][VEC], rx[REG ? PS_MAXREG : 1][VEC];
  float P[VEC], h[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { P[e] = 1.f; h[e] = 0.f; }

  auto load_ax = [&](int u, float* a, float* x) {
    if (!BWD) {
      ldv<VEC>(A + base + (size_t)u * DN, a);
      ldv<VEC>(X + base + (size_t)u * DN, x);
    } else {
      const int t = L - 1 - u;
      if (t + 1 < L) ldv<VEC>(A + base + (size_t)(t + 1) * DN, a);
      else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = 0.f;
      }
      ldv<VEC>(X + base + (size_t)t * DN, x);
    }
  };

  // ---- 1. local scan of the chunk (zero carry-in)
  if (REG) {
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) load_ax(u0 + i, ra[i], rx[i]);
    }
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { h[e] = fmaf(ra[i][e], h[e], rx[i][e]); P[e] *= ra[i][e]; }
      }
    }
  } else {
#pragma unroll 4
    for (int u = u0; u < u1; ++u) {
      float a[VEC], x[VEC];
      load_ax(u, a, x);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { h[e] = fmaf(a[e], h[e], x[e]); P[e] *= a[e]; }
    }
  }
  float* sP = sm;
  float* sH = sm + nw * cpb;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    sP[warp * cpb + lane * VEC + e] = P[e];
    sH[warp * cpb + lane * VEC + e] = h[e];
  }
  __syncthreads();
  // ---- 2. scan of the chunk aggregates: lane <-> chunk, one channel at a time per warp
  for (int ch = warp; ch < cpb; ch += nw) {
    float p = lane < nw ? sP[lane * cpb + ch] : 1.f;
    float g = lane < nw ? sH[lane * cpb + ch] : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float pp = __shfl_up_sync(0xffffffffu, p, o);
      const float gp = __shfl_up_sync(0xffffffffu, g, o);
      if (lane >= o) { g = fmaf(p, gp, g); p *= pp; }
    }
    // exclusive carry-in of chunk `lane` = inclusive result of chunk lane-1
    const float carry = __shfl_up_sync(0xffffffffu, g, 1);
    if (lane < nw) sH[lane * cpb + ch] = lane == 0 ? 0.f : carry;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < VEC; ++e) h[e] = sH[warp * cpb + lane * VEC + e];
  // ---- 3. replay with the carry-in and write
  auto emit = [&](int u, const float* hv) {
    if (!BWD) {
      stv<VEC>(out + base + (size_t)u * DN, hv);
    } else {
      const int t = L - 1 - u;
      stv<VEC>(out + base + (size_t)t * DN, hv);
      float ga[VEC];
      if (t >= 1) {
        float hp[VEC];
        ldv<VEC>(Hprev + base + (size_t)(t - 1) * DN, hp);
#pragma unroll
        for (int e = 0; e < VEC; ++e) ga[e] = hp[e] * hv[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) ga[e] = 0.f;
      }
      stv<VEC>(out_gA + base + (size_t)t * DN, ga);
    }
  };
  if (REG) {
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) h[e] = fmaf(ra[i][e], h[e], rx[i][e]);
        emit(u0 + i, h);
      }
    }
  } else {
#pragma unroll 4
    for (int u = u0; u < u1; ++u) {
      float a[VEC], x[VEC];
      load_ax(u, a, x);
#pragma unroll
      for (int e = 0; e < VEC; ++e) h[e] = fmaf(a[e], h[e], x[e]);
      emit(u, h);
    }
  }
}

template <bool BWD>
static int pscan_launch(const float* A, const float* X, const float* Hprev, float* out, float* out_gA, int B, int L, int D,
                        int N, cudaStream_t stream) {
  V2M_REQUIRE(B >= 0 && L >= 1 && D >= 1 && N >= 1, "pscan: bad shape (%d,%d,%d,%d)", B, L, D, N);
  if (B == 0) return kOk;
  const int DN = D * N;
  V2M_REQUIRE(DN % 32 == 0, "pscan: D*N=%d must be a multiple of 32", DN);
  // Short sequences (L <= 640): one channel per lane, chunk (<= 32 steps) resident in registers, <= 20 warps.
  // Long sequences: two-pass chunks; float4 channels per lane when that still fills 148 SMs twice over.
  int nw = (L + PS_MAXREG - 1) / PS_MAXREG;
  const bool reg = nw <= 20;
  const bool vec4 = !reg && (DN % 128 == 0) && ((long long)B * (DN / 128) >= 2 * 148) &&
                    (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                    (!BWD || (reinterpret_cast<uintptr_t>(Hprev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_gA) % 16 == 0));
  const int vec = vec4 ? 4 : 1;
  if (!reg) nw = 32;
  if (nw < 1) nw = 1;
  const int chunk = (L + nw - 1) / nw;
  const int grid = B * (DN / (32 * vec));
  const size_t smem = sizeof(float) * 2 * nw * 32 * vec;
#define PS_GO(V, R) pscan_kernel<V, R, BWD><<<grid, nw * 32, smem, stream>>>(A, X, Hprev, out, out_gA, L, DN, chunk)
  if (vec4) PS_GO(4, false);
  else if (reg) PS_GO(1, true);
  else PS_GO(1, false);
#undef PS_GO
  return check_launch("pscan");
}

int pscan_fwd(const float* A, const float* X, float* H, int B, int L, int D, int N, cudaStream_t stream) {
  return pscan_launch<false>(A, X, nullptr, H, nullptr, B, L, D, N, stream);
}

int pscan_bwd(const float* A, const float* H, const float* gH, float* gA, float* gX, int B, int L, int D, int N,
              cudaStream_t stream) {
  return pscan_launch<true>(A, gH, H, gX, gA, B, L, D, N, stream);
}

}  // namespace v2m
