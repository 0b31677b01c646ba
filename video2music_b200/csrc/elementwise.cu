// Bandwidth-bound helpers: residual + LayerNorm, embedding gather, feature concat, casts.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

__device__ __forceinline__ float load_any(const void* p, int dtype, size_t i) {
  return dtype == 0 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void store_any(void* p, int dtype, size_t i, float v) {
  if (dtype == 0) static_cast<float*>(p)[i] = v;
  else static_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

// One warp per row; the row lives in registers (D <= 32 * kMaxPerLane).  Two-pass mean / variance in fp32,
// the arithmetic of nn.LayerNorm (rpr.py:59-69: tgt = norm(tgt + dropout(tgt2))).
constexpr int kMaxPerLane = 32;
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x, int x_dtype,
                                                        const void* __restrict__ res, int res_dtype,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        void* __restrict__ y, int y_dtype, void* __restrict__ y2, int y2_dtype,
                                                        int M, int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const size_t base = (size_t)warp * D;
  float vals[kMaxPerLane];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxPerLane; ++c) {
    const int d = lane + 32 * c;
    float v = 0.f;
    if (d < D) {
      v = load_any(x, x_dtype, base + d);
      if (res) v += load_any(res, res_dtype, base + d);
    }
    vals[c] = v;
    sum += v;
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxPerLane; ++c) {
    const int d = lane + 32 * c;
    if (d < D) {
      const float t = vals[c] - mean;
      sq = fmaf(t, t, sq);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int c = 0; c < kMaxPerLane; ++c) {
    const int d = lane + 32 * c;
    if (d < D) {
      const float o = (vals[c] - mean) * rstd * gamma[d] + beta[d];
      store_any(y, y_dtype, base + d, o);
      if (y2) store_any(y2, y2_dtype, base + d, o);
    }
  }
}

// D == 512 fast path: a lane owns 16 contiguous columns (128-bit loads / stores), a warp walks rows grid-stride with
// gamma / beta in registers.  Same two-pass arithmetic as the generic kernel.
__device__ __forceinline__ void ld16(const bf16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  float2 f;
  f = bf16x2_to_f2(a.x); v[0] = f.x; v[1] = f.y;   f = bf16x2_to_f2(a.y); v[2] = f.x; v[3] = f.y;
  f = bf16x2_to_f2(a.z); v[4] = f.x; v[5] = f.y;   f = bf16x2_to_f2(a.w); v[6] = f.x; v[7] = f.y;
  f = bf16x2_to_f2(b.x); v[8] = f.x; v[9] = f.y;   f = bf16x2_to_f2(b.y); v[10] = f.x; v[11] = f.y;
  f = bf16x2_to_f2(b.z); v[12] = f.x; v[13] = f.y; f = bf16x2_to_f2(b.w); v[14] = f.x; v[15] = f.y;
}
__device__ __forceinline__ void ld16(const float* p, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = reinterpret_cast<const float4*>(p)[q];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void st16(bf16* p, const float* v) {
  uint4 a, b;
  a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
  b.x = f2_to_bf16x2(v[8], v[9]);   b.y = f2_to_bf16x2(v[10], v[11]); b.z = f2_to_bf16x2(v[12], v[13]); b.w = f2_to_bf16x2(v[14], v[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 8) = b;
}
__device__ __forceinline__ void st16(float* p, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm512_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           T* __restrict__ y, int M, float eps) {
  grid_dep_launch();
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = gridDim.x * 8;
  float g[16], b[16];
  ld16(gamma + lane * 16, g);
  ld16(beta + lane * 16, b);
  for (int row = blockIdx.x * 8 + warp; row < M; row += nw) {
    float v[16];
    ld16(x + (size_t)row * 512 + lane * 16, v);
    if (res) {
      float r[16];
      ld16(res + (size_t)row * 512 + lane * 16, r);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += r[e];
    }
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) sum += v[e];
    const float mean = warp_sum(sum) * (1.f / 512.f);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { const float t = v[e] - mean; sq = fmaf(t, t, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / 512.f) + eps);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = (v[e] - mean) * rstd * g[e] + b[e];
    st16(y + (size_t)row * 512 + lane * 16, v);
  }
}

int layernorm(const void* x, int x_dtype, const void* res, int res_dtype, const float* gamma, const float* beta,
              void* y, int y_dtype, void* y2, int y2_dtype, int M, int D, float eps, cudaStream_t stream) {
  V2M_REQUIRE(D > 0 && D <= 32 * kMaxPerLane, "layernorm: D=%d unsupported (<= %d)", D, 32 * kMaxPerLane);
  if (M == 0) return kOk;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (D == 512 && !y2 && x_dtype == y_dtype && (!res || res_dtype == x_dtype) && al16(x) && al16(res) && al16(y) && al16(gamma) &&
      al16(beta)) {
    int grid = (M + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    if (x_dtype == 0)
      launch_dep(layernorm512_kernel<float>, dim3(grid), dim3(256), 0, stream, static_cast<const float*>(x), static_cast<const float*>(res), gamma, beta,
                 static_cast<float*>(y), M, eps);
    else
      launch_dep(layernorm512_kernel<bf16>, dim3(grid), dim3(256), 0, stream, static_cast<const bf16*>(x), static_cast<const bf16*>(res), gamma, beta,
                 static_cast<bf16*>(y), M, eps);
    return check_launch("layernorm512");
  }
  const int rows_per_block = 8;
  layernorm_kernel<<<(M + rows_per_block - 1) / rows_per_block, rows_per_block * 32, 0, stream>>>(
      x, x_dtype, res, res_dtype, gamma, beta, y, y_dtype, y2, y2_dtype, M, D, eps);
  return check_launch("layernorm");
}

__global__ void embed_sum_kernel(const long long* __restrict__ idx_a, const float* __restrict__ table_a,
                                 const long long* __restrict__ idx_b, const float* __restrict__ table_b,
                                 void* __restrict__ out, int out_dtype, int ld_out, int rows, int D) {
  const int row = blockIdx.x;
  const long long ia = idx_a[row];
  const long long ib = idx_b ? idx_b[row] : 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = table_a[ia * D + d];
    if (idx_b) v += table_b[ib * D + d];
    store_any(out, out_dtype, (size_t)row * ld_out + d, v);
  }
}

int embed_sum(const long long* idx_a, const float* table_a, const long long* idx_b, const float* table_b,
              void* out, int out_dtype, int ld_out, int rows, int D, cudaStream_t stream) {
  if (rows == 0) return kOk;
  embed_sum_kernel<<<rows, 128, 0, stream>>>(idx_a, table_a, idx_b, table_b, out, out_dtype, ld_out, rows, D);
  return check_launch("embed_sum");
}

__global__ void concat_features_kernel(const float* __restrict__ sem, int sem_dim, const float* __restrict__ scene,
                                       const float* __restrict__ motion, int motion_dim,
                                       const float* __restrict__ emotion, int emo_dim, void* __restrict__ out,
                                       int out_dtype, int ld_out, int rows) {
  const int row = blockIdx.x;
  const int o_scene = sem_dim, o_motion = sem_dim + 1, o_emo = o_motion + motion_dim, total = o_emo + emo_dim;
  for (int c = threadIdx.x; c < ld_out; c += blockDim.x) {
    float v = 0.f;
    if (c < o_scene) v = sem[(size_t)row * sem_dim + c];
    else if (c < o_motion) v = scene[row];
    else if (c < o_emo) v = motion[(size_t)row * motion_dim + (c - o_motion)];
    else if (c < total) v = emotion[(size_t)row * emo_dim + (c - o_emo)];
    store_any(out, out_dtype, (size_t)row * ld_out + c, v);
  }
}

int concat_features(const float* sem, int sem_dim, const float* scene, const float* motion, int motion_dim,
                    const float* emotion, int emo_dim, void* out, int out_dtype, int ld_out, int rows, cudaStream_t stream) {
  V2M_REQUIRE(ld_out >= sem_dim + 1 + motion_dim + emo_dim, "concat_features: ld_out %d too small", ld_out);
  if (rows == 0) return kOk;
  concat_features_kernel<<<rows, 256, 0, stream>>>(sem, sem_dim, scene, motion, motion_dim, emotion, emo_dim, out,
                                                   out_dtype, ld_out, rows);
  return check_launch("concat_features");
}

__global__ void cast_copy_2d_kernel(const void* __restrict__ src, int src_dtype, long long ld_src, void* __restrict__ dst,
                                    int dst_dtype, long long ld_dst, int rows, int cols, int width) {
  const long long total = (long long)rows * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), c = (int)(i % width);
    const float v = c < cols ? load_any(src, src_dtype, (size_t)(r * ld_src + c)) : 0.f;
    store_any(dst, dst_dtype, (size_t)(r * ld_dst + c), v);
  }
}

int cast_copy_2d(const void* src, int src_dtype, long long ld_src, void* dst, int dst_dtype, long long ld_dst,
                 int rows, int cols, int zero_pad, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return kOk;
  const int width = zero_pad ? (int)ld_dst : cols;
  const long long total = (long long)rows * width;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  cast_copy_2d_kernel<<<blocks, 256, 0, stream>>>(src, src_dtype, ld_src, dst, dst_dtype, ld_dst, rows, cols, width);
  return check_launch("cast_copy_2d");
}

int cast_copy(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, cudaStream_t stream) {
  if (n == 0) return kOk;
  return cast_copy_2d(src, src_dtype, n, dst, dst_dtype, n, 1, (int)n, 0, stream);
}

// Cache layout of the streamed decode kernel: out[row] = [K(64) | V(64)] bf16 with the eight 16-byte chunks of each half
// stored at chunk position (c ^ (pos & 7)), pos = row % S  (row = (video * H + head) * S + pos).
__global__ void kv_interleave_kernel(const uint4* __restrict__ k, const uint4* __restrict__ v, uint4* __restrict__ out,
                                     long long rows, int S) {
  const long long total = rows * 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 4;
    const int part = (int)(i >> 3) & 1, c = (int)i & 7, pos = (int)(row % S);
    out[row * 16 + part * 8 + (c ^ (pos & 7))] = (part ? v : k)[row * 8 + c];
  }
}

int kv_interleave(const void* k, const void* v, void* out, long long rows, int S, cudaStream_t stream) {
  if (rows == 0) return kOk;
  V2M_REQUIRE(S > 0, "kv_interleave: S must be positive");
  const long long want = (rows * 16 + 255) / 256;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  kv_interleave_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(k), static_cast<const uint4*>(v),
                                                    static_cast<uint4*>(out), rows, S);
  return check_launch("kv_interleave");
}

// mode 0: out = a * silu(b)   (GLUExpert, moe.py:47);  mode 1: out = a + alpha * b  (shared expert, moe.py:301)
__global__ void binary_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                              long long n, int mode, float alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = a[i], y = b[i];
    out[i] = mode == 0 ? x * (y / (1.f + expf(-y))) : mode == 1 ? fmaf(alpha, y, x) : mode == 2 ? 1.f / (1.f + expf(-x))   // 2: sigmoid(a)
                                                                                 : x * y * (1.f - y);       // 3: sigmoid backward, a = dy, b = sigmoid
  }
}

int binary_op(const float* a, const float* b, float* out, long long n, int mode, float alpha, cudaStream_t stream) {
  if (n == 0) return kOk;
  const long long want = (n + 255) / 256;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  binary_kernel<<<blocks, 256, 0, stream>>>(a, b, out, n, mode, alpha);
  return check_launch("binary_op");
}

// RoPE exactly as the reference's V2 attention applies it (custom_transformer.py:1044-1053 + rotate_operation.py:117-165):
// the (len, B, E) projection is REINTERPRETED (.view, no transpose) as [H][len][B][dh], the rotation cache [len][E/2][2] is
// reinterpreted as [H][len][dh/2][2], and pair j of element (h', l', b', .) is rotated by cache entry (h'*len + l')*(dh/2) + j.
// One thread per (even, odd) pair.
__global__ void __launch_bounds__(256) rope_quirk_kernel(const float* __restrict__ x, const float2* __restrict__ cache,
                                                         float* __restrict__ y, long long n_pairs, int len, int B, int dh2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % dh2);
    const long long r = i / dh2;                    // (h' * len + l') * B + b'
    const long long hl = r / B;
    const float2 cs = cache[hl * dh2 + j];
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    reinterpret_cast<float2*>(y)[i] = make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
  }
}

int rope_quirk(const float* x, const float* cache, float* y, int len, int B, int H, int dh, cudaStream_t stream) {
  V2M_REQUIRE(len > 0 && B > 0 && H > 0 && dh > 0 && dh % 2 == 0, "rope_quirk: bad dims len=%d B=%d H=%d dh=%d", len, B, H, dh);
  const long long n_pairs = (long long)len * B * H * (dh / 2);
  const long long want = (n_pairs + 255) / 256;
  rope_quirk_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(x, reinterpret_cast<const float2*>(cache), y, n_pairs,
                                                                                 len, B, dh / 2);
  return check_launch("rope_quirk");
}

}  // namespace v2m
