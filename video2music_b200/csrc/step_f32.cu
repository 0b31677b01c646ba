// fp32 kernels of ONE generation step of the generic decoder stacks (BASELINE config 4: grouped-query attention + MoE feed-forward,
// cached_decode.py): the batch is one new token per video (M <= 64 rows), so a linear layer is a weight stream and the attention is one
// query row per (video, head) over the cached keys / values.  The tiled GEMM / attention kernels of the full forward run these shapes
// on 4..12 CTAs (93 us per launch measured); here every launch covers the chip and is bound by its weight / cache bytes.
//
//   step_gemm_f32_kernel   : y[M,N] = act(x[M,K] W[N,K]^T + bias[n] + row_scale[m] col_vec[n])   (F.linear at
//                            custom_transformer.py:1044-1053 / grouped_query_attention.py:306-356 / video_music_transformer.py:240-262)
//                            and, per expert group, the GLU expert GEMMs of the MoE feed-forward (moe.py:44-49);
//                            CTA = 8 output features x 64 rows, activations in registers, weights as broadcast loads, fixed k order.
//   step_attn_f32_kernel   : o = softmax(q K^T * scale) V for ONE query row per (video, query head) over the first n cached rows,
//                            n read from a device word (the launch is identical for every position -> one CUDA graph per
//                            generation); kv head = query head / (Hq / Hkv)  (grouped_query_attention.py:99-159 with one query).
// Both use a fixed summation order: results do not depend on the batch size or on n_max.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

namespace stp {

constexpr int THREADS = 256;
constexpr int NT = 8;          // output features per CTA
constexpr int MT = 64;         // rows (videos) per CTA

// warp = 8 rows, lane = a k slice (16 of every 512 k: four 16-byte chunks, 128 apart): activation and weight loads are coalesced
// 512-byte requests (a broadcast or row-strided 16-byte load costs the load/store unit the same as a full request and was the
// bound of the first two versions: 12.5 us per 64 x 512 x 512 launch), 64 accumulators per lane (8 rows x 8 feature columns), and
// the 32 lane partials of each accumulator meet in a transpose-reduce butterfly (31 shuffles per 32 values, fixed order) that
// leaves lane l with the total of (row l / 8, column l % 8).  No shared memory, no barrier.
struct StepGemm {
  const float* A; long long lda;
  const float* W1; const float* b1; const float* Wg; const float* bg;
  long long ldw, w_gstride, b_gstride;
  const int* off;            // row ranges of the groups (experts); null: one group = rows [0, M)
  const float* row_scale; const float* col_vec;
  float* C; long long ldc;
  int M, N, K, relu;
};

template <bool GLU>
__global__ void __launch_bounds__(THREADS) step_gemm_f32_kernel(const __grid_constant__ StepGemm p) {
  constexpr int NF = GLU ? 4 : 8;                                // features per CTA; 8 accumulator columns either way
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.y;
  const int r_begin = p.off ? p.off[e] : 0, r_end = p.off ? p.off[e + 1] : p.M;
  const int n0 = blockIdx.x * NF;
  const float* W1e = p.W1 + (size_t)e * p.w_gstride;
  const float* Wge = GLU ? p.Wg + (size_t)e * p.w_gstride : nullptr;
  for (int rw = r_begin + 8 * warp; rw < r_end; rw += MT) {      // this warp's 8 rows of every 64-row block
    float acc[8][8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[rr][c] = 0.f;
    for (int kb = 0; kb < p.K; kb += 512) {
      float4 xr[8][4];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const bool ok = rw + rr < r_end;
        const float* xs = p.A + (size_t)(ok ? rw + rr : r_begin) * p.lda + kb + 4 * lane;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          xr[rr][j] = (ok && kb + 128 * j + 4 * lane < p.K) ? __ldg(reinterpret_cast<const float4*>(xs + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int n = min(n0 + (GLU ? (c & 3) : c), p.N - 1);
        const float* ws = ((GLU && c >= 4) ? Wge : W1e) + (size_t)n * p.ldw + kb + 4 * lane;
        float4 wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          wv[j] = (kb + 128 * j + 4 * lane < p.K) ? __ldg(reinterpret_cast<const float4*>(ws + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          float a = acc[rr][c];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a = fmaf(xr[rr][j].x, wv[j].x, a); a = fmaf(xr[rr][j].y, wv[j].y, a);
            a = fmaf(xr[rr][j].z, wv[j].z, a); a = fmaf(xr[rr][j].w, wv[j].w, a);
          }
          acc[rr][c] = a;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {                                  // rows 4g .. 4g + 3 of the warp: 32 sums -> one per lane
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = acc[4 * g + (i >> 3)][i & 7];
      float tot = warp_transpose_reduce32(v, lane);
      const int m = rw + 4 * g + (lane >> 3), c = lane & 7;
      float gate = 0.f;
      if (GLU) gate = __shfl_down_sync(0xffffffffu, tot, 4);       // column c + 4 = the gate sum of feature c
      const int n = n0 + c;
      if (m < r_end && c < NF && n < p.N) {
        if (p.b1) tot += p.b1[(size_t)e * p.b_gstride + n];
        if (GLU) {
          const float gv = gate + p.bg[(size_t)e * p.b_gstride + n];
          tot = tot * (gv / (1.f + expf(-gv)));
        }
        if (p.row_scale) tot = fmaf(p.row_scale[m], p.col_vec[n], tot);
        if (p.relu) tot = fmaxf(tot, 0.f);
        p.C[(size_t)m * p.ldc + n] = tot;
      }
    }
  }
}

constexpr int ATH = 256;       // threads of the attention CTA (8 warps)

// grid = B * Hq.  Shared memory: q[64] | sc[n_max] | part[16][64] | red[16]
__global__ void __launch_bounds__(ATH) step_attn_f32_kernel(const float* __restrict__ q, long long q_sb, const float* __restrict__ Kc,
                                                            const float* __restrict__ Vc, long long kv_sb, long long kv_sl,
                                                            float* __restrict__ o, long long o_sb, int Hq, int Hkv, int n_max,
                                                            const int* __restrict__ n_dev, float q_scale) {
  extern __shared__ __align__(16) float sa_smem[];
  float* qs = sa_smem;
  float* sc = qs + 64;
  float* part = sc + ((n_max + 3) & ~3);
  float* red = part + 16 * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / Hq, hq = blockIdx.x % Hq, hkv = hq / (Hq / Hkv);
  const int n = n_dev ? min(n_max, __ldg(n_dev)) : n_max;
  const float* kb = Kc + (size_t)b * kv_sb + (size_t)hkv * 64;
  const float* vb = Vc + (size_t)b * kv_sb + (size_t)hkv * 64;
  if (tid < 64) qs[tid] = q[(size_t)b * q_sb + (size_t)hq * 64 + tid] * q_scale;
  __syncthreads();
  // ---- scores: warp w takes keys w, w + 8, ...; a lane holds two dims of the row (256 contiguous bytes per warp and key);
  // eight keys of the warp in flight per iteration
  const float2 qq = *reinterpret_cast<const float2*>(qs + 2 * lane);
  float mx = -INFINITY;
  for (int j0 = warp; j0 < n; j0 += 64) {
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + 8 * u;
      float2 kk = make_float2(0.f, 0.f);
      if (j < n) kk = __ldg(reinterpret_cast<const float2*>(kb + (size_t)j * kv_sl) + lane);
      s[u] = fmaf(qq.x, kk.x, qq.y * kk.y);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], off);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + 8 * u;
      if (j < n) { if (lane == 0) sc[j] = s[u]; mx = fmaxf(mx, s[u]); }
    }
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
  // ---- p = exp(s - max) (kept un-normalised), row sum in a fixed order: lane-strided partials, shuffle tree, warps 0..7
  float sum = 0.f;
  for (int j = tid; j < n; j += ATH) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  const float total = ((red[8] + red[9]) + (red[10] + red[11])) + ((red[12] + red[13]) + (red[14] + red[15]));
  // ---- o[d] = sum_j p_j V[j][d]: thread = (key group g of 16, four dims d4); keys g, g + 16, ... with four 16-byte loads in
  // flight, the sixteen partial rows meet in shared memory and are added in a fixed order
  const int d4 = tid & 15, g = tid >> 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = g; j0 < n; j0 += 64) {
    float4 vv[4];
    float pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 16 * u;
      vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      pp[u] = 0.f;
      if (j < n) { vv[u] = __ldg(reinterpret_cast<const float4*>(vb + (size_t)j * kv_sl) + d4); pp[u] = sc[j]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc.x = fmaf(pp[u], vv[u].x, acc.x); acc.y = fmaf(pp[u], vv[u].y, acc.y);
      acc.z = fmaf(pp[u], vv[u].z, acc.z); acc.w = fmaf(pp[u], vv[u].w, acc.w);
    }
  }
  reinterpret_cast<float4*>(part)[g * 16 + d4] = acc;
  __syncthreads();
  if (tid < 64) {
    float v = 0.f;
#pragma unroll
    for (int gg = 0; gg < 16; ++gg) v += part[gg * 64 + tid];
    o[(size_t)b * o_sb + (size_t)hq * 64 + tid] = v / total;
  }
}

}  // namespace stp

int step_linear_f32(const float* x, long long ldx, const float* W, long long ldw, const float* bias, const float* row_scale,
                    const float* col_vec, float* y, long long ldy, int M, int N, int K, int relu, cudaStream_t stream) {
  V2M_REQUIRE(M > 0 && N > 0 && K > 0, "step_linear_f32: empty problem %d x %d x %d", M, N, K);
  V2M_REQUIRE(K % 4 == 0 && ldw % 4 == 0 && ldx % 4 == 0 && reinterpret_cast<uintptr_t>(W) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0,
              "step_linear_f32: rows must be 16-byte aligned (K %d, ldx %lld, ldw %lld)", K, ldx, ldw);
  V2M_REQUIRE((row_scale == nullptr) == (col_vec == nullptr), "step_linear_f32: row_scale and col_vec come together");
  stp::StepGemm p{x, ldx, W, bias, nullptr, nullptr, ldw, 0, 0, nullptr, row_scale, col_vec, y, ldy, M, N, K, relu};
  stp::step_gemm_f32_kernel<false><<<dim3((N + stp::NT - 1) / stp::NT, 1), stp::THREADS, 0, stream>>>(p);
  return check_launch("step_linear_f32");
}

int step_attn_f32(const float* q, long long q_sb, const float* k, const float* v, long long kv_sb, long long kv_sl, float* o, long long o_sb,
                  int B, int Hq, int Hkv, int dh, int n_max, const int* n_dev, float q_scale, cudaStream_t stream) {
  V2M_REQUIRE(dh == 64, "step_attn_f32: head_dim %d unsupported (64)", dh);
  V2M_REQUIRE(B > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0 && n_max > 0, "step_attn_f32: bad shape B=%d Hq=%d Hkv=%d n=%d", B, Hq, Hkv, n_max);
  V2M_REQUIRE(kv_sl % 4 == 0 && kv_sb % 4 == 0 && reinterpret_cast<uintptr_t>(k) % 16 == 0 && reinterpret_cast<uintptr_t>(v) % 16 == 0,
              "step_attn_f32: cache rows must be 16-byte aligned");
  const size_t smem = sizeof(float) * (64 + ((n_max + 3) & ~3) + 16 * 64 + 16);
  V2M_REQUIRE(smem <= 48 * 1024, "step_attn_f32: n_max %d too large", n_max);
  stp::step_attn_f32_kernel<<<B * Hq, stp::ATH, smem, stream>>>(q, q_sb, k, v, kv_sb, kv_sl, o, o_sb, Hq, Hkv, n_max, n_dev, q_scale);
  return check_launch("step_attn_f32");
}

// Small-row variant of moe_grouped_gemm (moe.cu calls it when a launch carries at most a few hundred token copies): rows
// [off[e], off[e+1]) of A belong to expert e, C = A W_e^T + b_e, or with the gate stack C = (A W1_e^T + b1_e) * silu(A Wg_e^T + bg_e)
// (GLUExpert, moe.py:44-49); K % 4 == 0 and 16-byte aligned rows are the caller's `vec` condition.  The expert weights (226 MB per
// position for 6 layers x 6 experts) are read once per launch.
int step_moe_gemm_f32(const float* A, int lda, const float* W1, const float* b1, const float* Wg, const float* bg, long long w_gstride,
                      long long b_gstride, const int* off, int n_experts, float* C, int ldc, int N, int K, cudaStream_t stream) {
  stp::StepGemm p{A, lda, W1, b1, Wg, bg, K, w_gstride, b_gstride, off, nullptr, nullptr, C, ldc, 0, N, K, 0};
  const int nt = Wg ? stp::NT / 2 : stp::NT;
  dim3 grid((N + nt - 1) / nt, n_experts);
  if (Wg) stp::step_gemm_f32_kernel<true><<<grid, stp::THREADS, 0, stream>>>(p);
  else stp::step_gemm_f32_kernel<false><<<grid, stp::THREADS, 0, stream>>>(p);
  return check_launch("step_moe_gemm_f32");
}

}  // namespace v2m
