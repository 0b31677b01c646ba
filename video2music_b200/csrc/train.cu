// Backward-pass and optimiser kernels of the AMT training step (run_model_vevo.py:84-124 of the reference:
// forward, 0.4*CE(label_smoothing 0.1, ignore PAD) + 0.6*BCEWithLogits, backward, Adam).
// The heavy contractions reuse the GEMM kernels (dX = dY W, dW = dY^T X) and attn_bwd.cu; this file
// holds the bandwidth-bound pieces.  dtype codes: 0 = fp32, 1 = bf16.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

__device__ __forceinline__ float ld_any(const void* p, int dtype, size_t i) {
  return dtype == 0 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, size_t i, float v) {
  if (dtype == 0) static_cast<float*>(p)[i] = v;
  else static_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

// ---- gradient of the fused linear epilogue --------------------------------------------------------------------
// dz = dy * (relu ? y > 0 : 1) * (n < alpha_cols ? alpha : 1);  db[n] = sum_m dz[m][n]
// (the epilogue computed y = relu((x W^T + b) * alpha_n); residual terms pass dy through unchanged.)
// Block = 32 columns x 8 row-lanes; rows strided over the grid; db via shared partial sums + one atomic per column.
__global__ void __launch_bounds__(256) dy_prep_kernel(const void* __restrict__ dy, int dy_dtype, long long ld_dy,
                                                      const void* __restrict__ y, int y_dtype, long long ld_y, int relu,
                                                      float alpha, int alpha_cols, void* __restrict__ dz, int dz_dtype,
                                                      long long ld_dz, float* __restrict__ db, int M, int N, float drop_scale,
                                                      unsigned int drop_thresh, unsigned int drop_seed,
                                                      const unsigned int* __restrict__ drop_seed_dev) {
  if (drop_seed_dev) drop_seed += *drop_seed_dev;
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (n < N) {
    const float sc = n < alpha_cols ? alpha : 1.f;
    for (int m = blockIdx.y * 8 + ty; m < M; m += gridDim.y * 8) {
      float g = ld_any(dy, dy_dtype, (size_t)m * ld_dy + n);
      if (relu && !(ld_any(y, y_dtype, (size_t)m * ld_y + n) > 0.f)) g = 0.f;
      g *= sc;
      if (drop_scale != 0.f) g = drop_keep(drop_seed, m, n, drop_thresh) ? g * drop_scale : 0.f;   // mask of the forward epilogue
      if (dz) st_any(dz, dz_dtype, (size_t)m * ld_dz + n, g);
      acc += g;
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N && db) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][tx];
    atomicAdd(db + n, s);
  }
}

// bf16 fast path: a thread owns 8 consecutive columns (one 16-byte access per tensor and row), 32 row-lanes per block
__global__ void __launch_bounds__(256) dy_prep_bf16x8_kernel(const bf16* __restrict__ dy, long long ld_dy, const bf16* __restrict__ y,
                                                             long long ld_y, int relu, float alpha, int alpha_cols,
                                                             bf16* __restrict__ dz, long long ld_dz, float* __restrict__ db, int M,
                                                             int N, float drop_scale, unsigned int drop_thresh,
                                                             unsigned int drop_seed, const unsigned int* __restrict__ drop_seed_dev) {
  grid_dep_launch();
  grid_dep_wait();
  if (drop_seed_dev) drop_seed += *drop_seed_dev;
  __shared__ float part[32][8 * 8 + 1];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;            // 8 column groups x 32 row lanes
  const int n0 = (blockIdx.x * 8 + tx) * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (n0 < N) {
    float sc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sc[e] = (n0 + e < alpha_cols) ? alpha : 1.f;
    // db-only calls (no dz, no ReLU mask, no dropout: the plain linear layers, two thirds of all calls) only sum columns:
    // four independent 16-byte loads per thread and iteration keep the memory pipe busy
    int m = blockIdx.y * 32 + ty;
    if (!dz && !relu && drop_scale == 0.f) {
      const int stride = gridDim.y * 32;
      for (; m + 3 * stride < M; m += 4 * stride) {
        uint4 q4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q4[u] = *reinterpret_cast<const uint4*>(dy + (size_t)(m + u * stride) * ld_dy + n0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w[4] = {q4[u].x, q4[u].y, q4[u].z, q4[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f2 = bf16x2_to_f2(w[e]);
            acc[2 * e] = fmaf(f2.x, sc[2 * e], acc[2 * e]);
            acc[2 * e + 1] = fmaf(f2.y, sc[2 * e + 1], acc[2 * e + 1]);
          }
        }
      }
    }
    for (; m < M; m += gridDim.y * 32) {
      const uint4 gq = *reinterpret_cast<const uint4*>(dy + (size_t)m * ld_dy + n0);
      float g[8];
      float2 f;
      f = bf16x2_to_f2(gq.x); g[0] = f.x; g[1] = f.y;   f = bf16x2_to_f2(gq.y); g[2] = f.x; g[3] = f.y;
      f = bf16x2_to_f2(gq.z); g[4] = f.x; g[5] = f.y;   f = bf16x2_to_f2(gq.w); g[6] = f.x; g[7] = f.y;
      if (relu) {
        const uint4 yq = *reinterpret_cast<const uint4*>(y + (size_t)m * ld_y + n0);
        float yv[8];
        f = bf16x2_to_f2(yq.x); yv[0] = f.x; yv[1] = f.y;   f = bf16x2_to_f2(yq.y); yv[2] = f.x; yv[3] = f.y;
        f = bf16x2_to_f2(yq.z); yv[4] = f.x; yv[5] = f.y;   f = bf16x2_to_f2(yq.w); yv[6] = f.x; yv[7] = f.y;
#pragma unroll
        for (int e = 0; e < 8; ++e) if (!(yv[e] > 0.f)) g[e] = 0.f;
      }
      uint32_t hd[2] = {0u, 0u};
      if (drop_scale != 0.f) { hd[0] = drop_hash4(drop_seed, m, n0 >> 2); hd[1] = drop_hash4(drop_seed, m, (n0 >> 2) + 1); }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[e] *= sc[e];
        if (drop_scale != 0.f) g[e] = drop_keep_byte(hd[e >> 2], e, drop_thresh) ? g[e] * drop_scale : 0.f;
        acc[e] += g[e];
      }
      if (dz) {
        uint4 o;
        o.x = f2_to_bf16x2(g[0], g[1]); o.y = f2_to_bf16x2(g[2], g[3]); o.z = f2_to_bf16x2(g[4], g[5]); o.w = f2_to_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(dz + (size_t)m * ld_dz + n0) = o;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[ty][tx * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64 && db) {
    const int n = blockIdx.x * 64 + threadIdx.x;
    if (n < N) {
      float s = 0.f;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) s += part[i][threadIdx.x];
      atomicAdd(db + n, s);
    }
  }
}

int dy_prep(const void* dy, int dy_dtype, long long ld_dy, const void* y, int y_dtype, long long ld_y, int relu, float alpha,
            int alpha_cols, void* dz, int dz_dtype, long long ld_dz, float* db, int M, int N, float drop_scale, unsigned int drop_thresh,
            unsigned int drop_seed, const unsigned int* drop_seed_dev, cudaStream_t stream) {
  if (M == 0 || N == 0) return kOk;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (dy_dtype == 1 && (!dz || dz_dtype == 1) && (!relu || y_dtype == 1) && N % 8 == 0 && ld_dy % 8 == 0 && (!dz || ld_dz % 8 == 0) &&
      (!relu || ld_y % 8 == 0) && al16(dy) && al16(dz) && (!relu || al16(y))) {
    int gy2 = (M + 31) / 32;
    if (gy2 > 148) gy2 = 148;
    dim3 grid2((N + 63) / 64, gy2);
    launch_dep(dy_prep_bf16x8_kernel, grid2, dim3(256), 0, stream, static_cast<const bf16*>(dy), ld_dy, static_cast<const bf16*>(y), ld_y, relu, alpha,
               alpha_cols, static_cast<bf16*>(dz), ld_dz, db, M, N, drop_scale, drop_thresh, drop_seed, drop_seed_dev);
    return check_launch("dy_prep_bf16x8");
  }
  int gy = (M + 7) / 8;
  if (gy > 148 * 2) gy = 148 * 2;
  dim3 grid((N + 31) / 32, gy);
  dy_prep_kernel<<<grid, 256, 0, stream>>>(dy, dy_dtype, ld_dy, y, y_dtype, ld_y, relu, alpha, alpha_cols, dz, dz_dtype, ld_dz,
                                          db, M, N, drop_scale, drop_thresh, drop_seed, drop_seed_dev);
  return check_launch("dy_prep");
}

// ---- LayerNorm backward ---------------------------------------------------------------------------------------
// y = LN(x) * g + b with x the pre-norm sum.  dx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)), dxh = dy * g;
// dg[d] += dy * xh, db[d] += dy.  One warp per row; a CTA accumulates its 8 rows' dg/db in shared memory first.
constexpr int LNB_MAX = 32;
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ gamma,
                                                            const void* __restrict__ dy, int dy_dtype, void* __restrict__ dx,
                                                            int dx_dtype, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int M, int D, float eps) {
  extern __shared__ float sacc[];                 // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const size_t base = (size_t)row * D;
    float xv[LNB_MAX], gv[LNB_MAX];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < LNB_MAX; ++c) {
      const int d = lane + 32 * c;
      xv[c] = d < D ? ld_any(x, x_dtype, base + d) : 0.f;
      sum += xv[c];
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < LNB_MAX; ++c) {
      const int d = lane + 32 * c;
      if (d < D) { const float t = xv[c] - mean; sq = fmaf(t, t, sq); }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < LNB_MAX; ++c) {
      const int d = lane + 32 * c;
      if (d < D) {
        const float xh = (xv[c] - mean) * rstd;
        const float g = ld_any(dy, dy_dtype, base + d);
        atomicAdd(&sacc[d], g * xh);
        atomicAdd(&sacc[D + d], g);
        const float dxh = g * gamma[d];
        xv[c] = xh;
        gv[c] = dxh;
        s1 += dxh;
        s2 = fmaf(dxh, xh, s2);
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int c = 0; c < LNB_MAX; ++c) {
      const int d = lane + 32 * c;
      if (d < D) st_any(dx, dx_dtype, base + d, rstd * (gv[c] - s1 - xv[c] * s2));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sacc[i]);
    atomicAdd(dbeta + i, sacc[D + i]);
  }
}

// D == 512 fast path: a lane owns 16 contiguous columns (128-bit accesses), dgamma / dbeta partials live in registers over
// all rows a warp visits, are combined per CTA in shared memory and leave with one global atomic per column and CTA.
__device__ __forceinline__ void lnb_ld16(const bf16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  float2 f;
  f = bf16x2_to_f2(a.x); v[0] = f.x; v[1] = f.y;   f = bf16x2_to_f2(a.y); v[2] = f.x; v[3] = f.y;
  f = bf16x2_to_f2(a.z); v[4] = f.x; v[5] = f.y;   f = bf16x2_to_f2(a.w); v[6] = f.x; v[7] = f.y;
  f = bf16x2_to_f2(b.x); v[8] = f.x; v[9] = f.y;   f = bf16x2_to_f2(b.y); v[10] = f.x; v[11] = f.y;
  f = bf16x2_to_f2(b.z); v[12] = f.x; v[13] = f.y; f = bf16x2_to_f2(b.w); v[14] = f.x; v[15] = f.y;
}
__device__ __forceinline__ void lnb_ld16(const float* p, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = reinterpret_cast<const float4*>(p)[q];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void lnb_st16(bf16* p, const float* v) {
  uint4 a, b;
  a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
  b.x = f2_to_bf16x2(v[8], v[9]);   b.y = f2_to_bf16x2(v[10], v[11]); b.z = f2_to_bf16x2(v[12], v[13]); b.w = f2_to_bf16x2(v[14], v[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 8) = b;
}
__device__ __forceinline__ void lnb_st16(float* p, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// raw (unconverted) 16-element slices: what the prefetch of the next row holds in registers
template <typename T> struct LnbRaw;
template <> struct LnbRaw<bf16> { uint4 a, b; };
template <> struct LnbRaw<float> { float4 q[4]; };
__device__ __forceinline__ void lnb_raw_ld(const bf16* p, LnbRaw<bf16>& r) {
  r.a = *reinterpret_cast<const uint4*>(p); r.b = *reinterpret_cast<const uint4*>(p + 8);
}
__device__ __forceinline__ void lnb_raw_ld(const float* p, LnbRaw<float>& r) {
#pragma unroll
  for (int q = 0; q < 4; ++q) r.q[q] = reinterpret_cast<const float4*>(p)[q];
}
__device__ __forceinline__ void lnb_cvt(const LnbRaw<bf16>& r, float* v) {
  float2 f;
  f = bf16x2_to_f2(r.a.x); v[0] = f.x; v[1] = f.y;   f = bf16x2_to_f2(r.a.y); v[2] = f.x; v[3] = f.y;
  f = bf16x2_to_f2(r.a.z); v[4] = f.x; v[5] = f.y;   f = bf16x2_to_f2(r.a.w); v[6] = f.x; v[7] = f.y;
  f = bf16x2_to_f2(r.b.x); v[8] = f.x; v[9] = f.y;   f = bf16x2_to_f2(r.b.y); v[10] = f.x; v[11] = f.y;
  f = bf16x2_to_f2(r.b.z); v[12] = f.x; v[13] = f.y; f = bf16x2_to_f2(r.b.w); v[14] = f.x; v[15] = f.y;
}
__device__ __forceinline__ void lnb_cvt(const LnbRaw<float>& r, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) { v[4 * q] = r.q[q].x; v[4 * q + 1] = r.q[q].y; v[4 * q + 2] = r.q[q].z; v[4 * q + 3] = r.q[q].w; }
}

// The rows of a warp are independent: the NEXT row's x / dy slices are fetched (raw, 8 registers each in bf16) before the
// current row's reductions, so the four dependent warp sums of a row overlap the memory latency of the next one.  The
// per-CTA fold of dgamma / dbeta goes through per-warp shared-memory rows (shared-memory float atomics are CAS loops).
template <typename T>
__global__ void __launch_bounds__(256) layernorm512_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                               const T* __restrict__ dy, T* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, int M, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float sacc[8][2 * 512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = gridDim.x * 8;
  float gm[16], dg[16], db[16];
  lnb_ld16(gamma + lane * 16, gm);
#pragma unroll
  for (int e = 0; e < 16; ++e) { dg[e] = 0.f; db[e] = 0.f; }
  int row = blockIdx.x * 8 + warp;
  LnbRaw<T> rx, rg;
  if (row < M) {
    lnb_raw_ld(x + (size_t)row * 512 + lane * 16, rx);
    lnb_raw_ld(dy + (size_t)row * 512 + lane * 16, rg);
  }
  while (row < M) {
    float xv[16], gv[16];
    lnb_cvt(rx, xv);
    lnb_cvt(rg, gv);
    const int nrow = row + nw;
    if (nrow < M) {
      lnb_raw_ld(x + (size_t)nrow * 512 + lane * 16, rx);
      lnb_raw_ld(dy + (size_t)nrow * 512 + lane * 16, rg);
    }
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) sum += xv[e];
    const float mean = warp_sum(sum) * (1.f / 512.f);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { const float t = xv[e] - mean; sq = fmaf(t, t, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / 512.f) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float xh = (xv[e] - mean) * rstd;
      dg[e] = fmaf(gv[e], xh, dg[e]);
      db[e] += gv[e];
      const float dxh = gv[e] * gm[e];
      xv[e] = xh;
      gv[e] = dxh;
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
    s1 = warp_sum(s1) * (1.f / 512.f);
    s2 = warp_sum(s2) * (1.f / 512.f);
#pragma unroll
    for (int e = 0; e < 16; ++e) gv[e] = rstd * (gv[e] - s1 - xv[e] * s2);
    lnb_st16(dx + (size_t)row * 512 + lane * 16, gv);
    row = nrow;
  }
  lnb_st16(&sacc[warp][lane * 16], dg);
  lnb_st16(&sacc[warp][512 + lane * 16], db);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 512; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sacc[w][i];
    atomicAdd(i < 512 ? dgamma + i : dbeta + (i - 512), t);
  }
}

int layernorm_bwd(const void* x, int x_dtype, const float* gamma, const void* dy, int dy_dtype, void* dx, int dx_dtype,
                  float* dgamma, float* dbeta, int M, int D, float eps, cudaStream_t stream) {
  V2M_REQUIRE(D > 0 && D <= 32 * LNB_MAX, "layernorm_bwd: D=%d unsupported", D);
  if (M == 0) return kOk;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (D == 512 && x_dtype == dy_dtype && x_dtype == dx_dtype && al16(x) && al16(dy) && al16(dx) && al16(gamma)) {
    int g2 = (M + 7) / 8;
    if (g2 > 148 * 2) g2 = 148 * 2;                       // 2 CTAs per SM, the rest of a warp's rows in its loop
    if (x_dtype == 0)
      launch_dep(layernorm512_bwd_kernel<float>, dim3(g2), dim3(256), 0, stream, static_cast<const float*>(x), gamma, static_cast<const float*>(dy),
                 static_cast<float*>(dx), dgamma, dbeta, M, eps);
    else
      launch_dep(layernorm512_bwd_kernel<bf16>, dim3(g2), dim3(256), 0, stream, static_cast<const bf16*>(x), gamma, static_cast<const bf16*>(dy),
                 static_cast<bf16*>(dx), dgamma, dbeta, M, eps);
    return check_launch("layernorm512_bwd");
  }
  int grid = (M + 7) / 8;
  if (grid > 148 * 4) grid = 148 * 4;
  layernorm_bwd_kernel<<<grid, 256, 2 * D * sizeof(float), stream>>>(x, x_dtype, gamma, dy, dy_dtype, dx, dx_dtype, dgamma,
                                                                     dbeta, M, D, eps);
  return check_launch("layernorm_bwd");
}

// ---- embedding backward: dtable[idx[row]] += d[row] ----------------------------------------------------------
__global__ void embed_bwd_kernel(const long long* __restrict__ idx, const void* __restrict__ d, int d_dtype, long long ld_d,
                                 float* __restrict__ dtable, int rows, int D) {
  const int row = blockIdx.x;
  const long long i = idx[row];
  for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(dtable + i * D + c, ld_any(d, d_dtype, (size_t)row * ld_d + c));
}

int embed_bwd(const long long* idx, const void* d, int d_dtype, long long ld_d, float* dtable, int rows, int D,
              cudaStream_t stream) {
  if (rows == 0) return kOk;
  embed_bwd_kernel<<<rows, 128, 0, stream>>>(idx, d, d_dtype, ld_d, dtable, rows, D);
  return check_launch("embed_bwd");
}

// ---- loss: 0.4 * CrossEntropy(label_smoothing eps, ignore_index) + 0.6 * BCEWithLogits, and d(loss)/d(logits) ----
// run_model_vevo.py:101-119 with train.py:222,233.  logits [R, C] fp32 (R = B*T), tgt [R] int64, tgt_emotion [R, C] fp32.
// One warp per row.  out[0] += CE row sum (valid rows), out[1] += BCE element sum; n_valid is a device scalar
// (count of tgt != ignore) produced by count_valid_kernel.  dlogits gets the already weighted and normalised gradient.
__global__ void count_valid_kernel(const long long* __restrict__ tgt, int R, long long ignore, float* __restrict__ n_valid) {
  grid_dep_launch();
  grid_dep_wait();
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) c += (tgt[i] != ignore);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_valid, (float)c);
}

__global__ void __launch_bounds__(256) amt_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ tgt,
                                                       const float* __restrict__ tgt_emotion, int R, int Cn, long long ignore,
                                                       float smooth, float w_ce, float w_bce, const float* __restrict__ n_valid,
                                                       const float* __restrict__ bce_rows_p, float* __restrict__ out,
                                                       float* __restrict__ dlogits) {
  grid_dep_launch();
  grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= R) return;
  const float* x = logits + (size_t)row * Cn;
  const float* e = tgt_emotion + (size_t)row * Cn;
  const long long t = tgt[row];
  const bool in_range = t >= 0 && t < Cn;
  const bool valid = t != ignore && in_range;
  // a label outside [0, Cn) that is not the ignore index: torch's cross_entropy raises; here the loss becomes NaN (loud)
  // instead of reading x[t] out of bounds
  if (lane == 0 && t != ignore && !in_range) atomicAdd(out + 0, __int_as_float(0x7fc00000));
  float mx = -INFINITY;
  for (int c = lane; c < Cn; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float se = 0.f, sx = 0.f, bce = 0.f;
  for (int c = lane; c < Cn; c += 32) {
    const float v = x[c];
    se += expf(v - mx);
    sx += v;
    bce += fmaxf(v, 0.f) - v * e[c] + log1pf(expf(-fabsf(v)));        // BCEWithLogits, numerically stable form
  }
  se = warp_sum(se);
  sx = warp_sum(sx);
  bce = warp_sum(bce);
  const float lse = mx + logf(se);
  const float nv = fmaxf(*n_valid, 1.f);
  const float inv_bce = 1.f / ((bce_rows_p ? *bce_rows_p : (float)R) * (float)Cn);
  if (lane == 0) {
    if (valid) {
      const float nll = lse - x[t];
      const float smooth_term = lse - sx / (float)Cn;                  // mean_c(-log p_c)
      atomicAdd(out + 0, (1.f - smooth) * nll + smooth * smooth_term);
    }
    atomicAdd(out + 1, bce);
  }
  if (dlogits) {
    for (int c = lane; c < Cn; c += 32) {
      const float v = x[c];
      float g = w_bce * inv_bce * (1.f / (1.f + expf(-v)) - e[c]);
      if (valid) {
        const float p = expf(v - lse);
        const float q = (1.f - smooth) * (c == t ? 1.f : 0.f) + smooth / (float)Cn;
        g += w_ce * (p - q) / nv;
      }
      dlogits[(size_t)row * Cn + c] = g;
    }
  }
}

int amt_loss(const float* logits, const long long* tgt, const float* tgt_emotion, int R, int Cn, long long ignore, float smooth,
             float w_ce, float w_bce, float* scratch3, float* dlogits, const float* norm_in, cudaStream_t stream) {
  // scratch3: [ce_sum, bce_sum, n_valid] (device, zeroed here).  norm_in (optional, device, 2 floats): [the CE normaliser to
  // use instead of this call's own count of non-ignored targets, the row count the BCE mean divides by instead of R] --
  // a data-parallel rank passes global_count / world and global_rows / world so that the 1/world mean of the ranks'
  // gradients is the gradient of the single-process global-batch loss (run_model_vevo.py:101-119).
  if (R == 0) return kOk;
  cudaError_t e = cudaMemsetAsync(scratch3, 0, 3 * sizeof(float), stream);
  if (e != cudaSuccess) { set_last_error("amt_loss: memset: %s", cudaGetErrorString(e)); return kCudaError; }
  count_valid_kernel<<<(R + 255) / 256 < 296 ? (R + 255) / 256 : 296, 256, 0, stream>>>(tgt, R, ignore, scratch3 + 2);
  amt_loss_kernel<<<(R + 7) / 8, 256, 0, stream>>>(logits, tgt, tgt_emotion, R, Cn, ignore, smooth, w_ce, w_bce,
                                                   norm_in ? norm_in : scratch3 + 2, norm_in ? norm_in + 1 : nullptr, scratch3, dlogits);
  return check_launch("amt_loss");
}

int count_valid(const long long* tgt, int R, long long ignore, float* out1, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out1, 0, sizeof(float), stream);
  if (e != cudaSuccess) { set_last_error("count_valid: memset: %s", cudaGetErrorString(e)); return kCudaError; }
  if (R == 0) return kOk;
  count_valid_kernel<<<(R + 255) / 256 < 296 ? (R + 255) / 256 : 296, 256, 0, stream>>>(tgt, R, ignore, out1);
  return check_launch("count_valid");
}

// ---- Adam / AdamW (train.py:237-240: betas (0.9, 0.98), eps ADAM_EPSILON = 10e-9 = 1e-8, constants.py:89-91) ----------
// torch.optim.Adam semantics; weight_decay > 0 adds torch.optim.AdamW's decoupled decay p *= 1 - lr * weight_decay before the
// update (the reference CLI's default optimiser, argument_funcs.py:17, uses AdamW's default 0.01).
// dyn (optional, device): {lr, 1 - b1^t, 1 - b2^t} read at run time, so that a CUDA graph of the whole training step can be
// replayed with a changing learning-rate schedule.  Fused extras: p16 (optional) receives the bf16 mirror of the updated
// parameters, zero_grad clears the gradient buffer for the next step, ctr (optional) is a device step counter that the
// dropout kernels add to their seeds.
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float weight_decay, float bc1, float bc2,
                            float grad_scale, const float* __restrict__ dyn, bf16* __restrict__ p16, int zero_grad, unsigned int* __restrict__ ctr) {
  grid_dep_launch();
  grid_dep_wait();
  if (dyn) { lr = dyn[0]; bc1 = dyn[1]; bc2 = dyn[2]; }
  const float decay = 1.f - lr * weight_decay;
  // 16-byte path (every buffer 16-byte aligned: the trainer's flat buffers): four elements per thread and access -- the kernel is a
  // pure stream over ~34 bytes per parameter and is bound by requests in flight, not by arithmetic
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!p16 || (reinterpret_cast<uintptr_t>(p16) & 7) == 0);
  const long long n4 = vec ? n / 4 : 0;
  const float inv1 = 1.f / bc1, inv2 = 1.f / bc2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i], m4 = reinterpret_cast<const float4*>(m)[i];
    const float4 v4 = reinterpret_cast<const float4*>(v)[i], p4 = reinterpret_cast<const float4*>(p)[i];
    const float gg[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float mo[4], vo[4], po[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mo[e] = b1 * mm[e] + (1.f - b1) * gg[e];
      vo[e] = b2 * vv[e] + (1.f - b2) * gg[e] * gg[e];
      po[e] = pp[e] * decay - lr * (mo[e] / bc1) / (sqrtf(vo[e] / bc2) + eps);
    }
    reinterpret_cast<float4*>(m)[i] = make_float4(mo[0], mo[1], mo[2], mo[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vo[0], vo[1], vo[2], vo[3]);
    reinterpret_cast<float4*>(p)[i] = make_float4(po[0], po[1], po[2], po[3]);
    if (p16) reinterpret_cast<uint2*>(p16)[i] = make_uint2(f2_to_bf16x2(po[0], po[1]), f2_to_bf16x2(po[2], po[3]));
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  (void)inv1; (void)inv2;
  for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pi = p[i] * decay - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
    if (p16) p16[i] = __float2bfloat16_rn(pi);
    if (zero_grad) g[i] = 0.f;
  }
  if (ctr && blockIdx.x == 0 && threadIdx.x == 0) *ctr += 1u;
}

int adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float weight_decay, int step,
              float grad_scale, const float* dyn, void* p16, int zero_grad, unsigned int* ctr, cudaStream_t stream) {
  if (n == 0) return kOk;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  const long long want = (n + 255) / 256;
  adam_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, weight_decay, bc1, bc2, grad_scale, dyn,
                                                                            static_cast<bf16*>(p16), zero_grad, ctr);
  return check_launch("adam_step");
}

// Evaluation metrics of the run loop (dataset/vevo_dataset.py:653-701: compute_vevo_accuracy, compute_hits_k), which the
// reference evaluates with a Python loop and one .item() per token: one warp per (video, position) row ranks the target's
// logit (softmax is monotonic, so logits are ranked directly).  rank = number of classes that beat the target, with the
// lower index winning ties (the order torch.argmax / torch.topk return on equal values).  Padded targets are skipped.
// counters: [0] valid rows, [1] rank == 0 (accuracy), [2 + i] rank < ks[i] (hits@k), int32, zeroed here.
__global__ void __launch_bounds__(256) amt_metrics_kernel(const float* __restrict__ logits, const long long* __restrict__ tgt, int R,
                                                          int Cn, long long pad, int k0, int k1, int k2, int* __restrict__ counters) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const long long t = tgt[row];
  if (t == pad || t < 0 || t >= Cn) return;
  const float* lr = logits + (long long)row * Cn;
  const float lt = lr[t];
  int beat = 0;
  for (int c = lane; c < Cn; c += 32) {
    const float v = lr[c];
    beat += (v > lt || (v == lt && c < (int)t)) ? 1 : 0;
  }
  beat = warp_sum_int(beat);
  if (lane == 0) {
    atomicAdd(counters + 0, 1);
    if (beat == 0) atomicAdd(counters + 1, 1);
    if (beat < k0) atomicAdd(counters + 2, 1);
    if (beat < k1) atomicAdd(counters + 3, 1);
    if (beat < k2) atomicAdd(counters + 4, 1);
  }
}

int amt_metrics(const float* logits, const long long* tgt, int R, int Cn, long long pad, int k0, int k1, int k2, int* counters,
                cudaStream_t stream) {
  V2M_REQUIRE(R >= 0 && Cn > 0 && counters, "amt_metrics: bad arguments");
  cudaError_t e = cudaMemsetAsync(counters, 0, 5 * sizeof(int), stream);
  if (e != cudaSuccess) { set_last_error("amt_metrics: %s", cudaGetErrorString(e)); return kCudaError; }
  if (R == 0) return kOk;
  amt_metrics_kernel<<<(R + 7) / 8, 256, 0, stream>>>(logits, tgt, R, Cn, pad, k0, k1, k2, counters);
  return check_launch("amt_metrics");
}

// Emotion-correspondence metric of the evaluation loop (dataset/vevo_dataset.py:747-810: compute_vevo_correspondence), which the
// reference evaluates with a Python loop, one .item() per position and five JSON files per call: a position counts (pt) when its
// emotion row is not padding (last column != 1), has at least one of the 14 chord qualities set and comes with probability >=
// threshold; it is right when the quality of the arg-max chord is one of the set qualities.  Chord id c -> quality: the
// chord_inv / chord_attr tables reduce to 1 ("maj") for c == 0 ("N" has no ':' either, literal) and (c - 1) % 13 + 1 otherwise;
// END (157) / PAD (158) predictions count in pt but are never right.  One warp per position; lower index wins arg-max ties.
// counters: [0] pt, [1] right (int32, zeroed here).
__global__ void __launch_bounds__(256) amt_correspondence_kernel(const float* __restrict__ logits, const float* __restrict__ emo,
                                                                 const float* __restrict__ prob, int R, int Cn, int Ce, float thr,
                                                                 int chord_end, int* __restrict__ counters) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* er = emo + (long long)row * Ce;
  const float q = lane < 14 ? er[lane] : 0.f;
  const bool any_q = __any_sync(0xffffffffu, lane < 14 && q != 0.f);
  if (er[Ce - 1] == 1.f || !any_q || prob[row] < thr) return;           // warp-uniform
  const float* lr = logits + (long long)row * Cn;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < Cn; c += 32) {
    const float v = lr[c];
    if (v > best) { best = v; arg = c; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  const int quality = arg == 0 ? 1 : (arg - 1) % 13 + 1;
  const float hit = __shfl_sync(0xffffffffu, q, quality);
  if (lane == 0) {
    atomicAdd(counters + 0, 1);
    if (arg < chord_end && hit == 1.f) atomicAdd(counters + 1, 1);
  }
}

int amt_correspondence(const float* logits, const float* emo, const float* prob, int R, int Cn, int Ce, float thr, int chord_end,
                       int* counters, cudaStream_t stream) {
  V2M_REQUIRE(R >= 0 && Cn > 0 && Ce >= 14 && counters, "amt_correspondence: bad arguments (Ce = %d needs the 14 quality columns)", Ce);
  cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream);
  if (e != cudaSuccess) { set_last_error("amt_correspondence: %s", cudaGetErrorString(e)); return kCudaError; }
  if (R == 0) return kOk;
  amt_correspondence_kernel<<<(R + 7) / 8, 256, 0, stream>>>(logits, emo, prob, R, Cn, Ce, thr, chord_end, counters);
  return check_launch("amt_correspondence");
}

}  // namespace v2m
