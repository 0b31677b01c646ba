// KV-cached single-token decode for chord generation (BASELINE config 2).
//
// The reference has no such path: VideoMusicTransformer.generate re-runs the whole model on
// the growing prefix every step (model/video_music_transformer.py:1069-1071).  Because the
// decoder is causal, row t of that forward depends only on positions <= t, so one step here
// computes exactly that row for all B videos at once from cached keys/values:
//   x_t   = Linear_chord([emb | key]) + pe[t]                         (:984-1001,1029)
//   6 x   TransformerDecoderLayerRPR on the single row                  (model/rpr.py:55-70)
//         self-attention logits  q.k_j + q.Er[er_len-1-(t-j)], j <= t   (rpr.py:387-395,439-455)
//   y_t   = Wout(LayerNorm(x))  ->  argmax over [:CHORD_END]            (:1042,1070-1084, beam=1)
//
// Two kinds of kernels, both HBM/L2-bandwidth bound (arithmetic intensity < 64 flop/B):
//   * skinny GEMM  [B<=64 rows] x [N,K]^T : every CTA owns 8 output features, streams its weight
//     rows once with 16-byte loads issued before the activation prologue, keeps the activation
//     tile in shared memory (LayerNorm of the previous sub-layer is applied while staging it) and
//     splits K over its 8 warps (mma.sync m16n8k16 for bf16, FFMA for fp32; tcgen05 needs M>=64
//     *per tile* and buys nothing at 16 flop/B).
//   * attention over the cache: one CTA per (video, head); K and V blocks are contiguous in the
//     head-major cache layout [B][H][pos][64] and are pulled into shared memory with one bulk
//     async copy each (TMA engine, mbarrier completion).
#include "common.cuh"
#include "kernels.h"
#include "decode_pick.cuh"
#include <type_traits>
#include <stdlib.h>

namespace v2m {

// Programmatic dependent launch: a kernel may start while its predecessor in the stream is still running; everything
// before pdl_wait() must only touch data that no kernel of the decode chain writes (weights, static cross K/V).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

enum EMode { E_QKV = 0, E_RESID = 1, E_Q = 2, E_RELU = 3, E_EMBED = 4, E_LOGITS = 5 };

struct SkinnyArgs {
  const void* W; const float* bias; int N, K;
  const void* a_src;            // [B, K] activations in the compute dtype (p.xn / p.ctx / p.ff)
  int emode; int layer;
};

constexpr int MT = 64;        // rows (videos) per tile
constexpr int NT = 8;         // output features per CTA
constexpr int SK_THREADS = 256;
constexpr int F32_KCHUNK = 512;

template <typename T> struct SkinnyCfg;
template <> struct SkinnyCfg<bf16> {
  static __host__ __device__ int a_stride(int K) { return K + 32; }                 // elements; 64 B pad -> conflict-free LDS.128
  static size_t smem(int K) { return (size_t)MT * a_stride(K) * 2 + 8 * MT * NT * 4 + 16; }
};
template <> struct SkinnyCfg<float> {
  static __host__ __device__ int a_stride(int) { return 0; }
  static size_t smem(int) { return 16; }                                           // registers only
};

__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16 consecutive elements of a row (lane-contiguous ownership: lane l holds elements [16 l, 16 l + 16) of a 512 segment)
__device__ __forceinline__ void load16(const bf16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
  float2 f;
  f = bf16x2_to_f2(a.x); v[0] = f.x; v[1] = f.y;   f = bf16x2_to_f2(a.y); v[2] = f.x; v[3] = f.y;
  f = bf16x2_to_f2(a.z); v[4] = f.x; v[5] = f.y;   f = bf16x2_to_f2(a.w); v[6] = f.x; v[7] = f.y;
  f = bf16x2_to_f2(b.x); v[8] = f.x; v[9] = f.y;   f = bf16x2_to_f2(b.y); v[10] = f.x; v[11] = f.y;
  f = bf16x2_to_f2(b.z); v[12] = f.x; v[13] = f.y; f = bf16x2_to_f2(b.w); v[14] = f.x; v[15] = f.y;
}
__device__ __forceinline__ void load16(const float* p, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = reinterpret_cast<const float4*>(p)[q];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void store16(bf16* p, const float* v) {
  uint4 a, b;
  a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
  b.x = f2_to_bf16x2(v[8], v[9]);   b.y = f2_to_bf16x2(v[10], v[11]); b.z = f2_to_bf16x2(v[12], v[13]); b.w = f2_to_bf16x2(v[14], v[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 8) = b;
}
__device__ __forceinline__ void store16(float* p, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

__device__ __forceinline__ void ln16(float* v, const float* g, const float* b, int E) {
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) sum += v[e];
  const float mean = warp_sum(sum) / (float)E;
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) { const float d = v[e] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(warp_sum(sq) / (float)E + 1e-5f);
  float gg[16], bb[16];
  load16(g, gg);
  load16(b, bb);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
}

// Row preparation, done ONCE per row (one warp per video) instead of redundantly inside every GEMM CTA:
//   PREP_LN   : xn = LayerNorm(r) -> p.xn (dtype T, GEMM operand) and p.h (fp32 residual)      (rpr.py:59,66,69)
//   PREP_LN2  : xn = LayerNorm_f(LayerNorm(r))  -> p.xn                                          (rpr.py:32-33)
//   PREP_EMBED: emb_root[.] + emb_attr[.] (or the chord embedding) of position t -> p.ctx       (video_music_transformer.py:984-989)
enum PrepMode { PREP_LN = 0, PREP_LN2 = 1, PREP_EMBED = 2 };
template <typename T>
__global__ void __launch_bounds__(256) rows_prep_kernel(const __grid_constant__ DecodeParams p, int mode, const float* g1,
                                                        const float* b1, const float* g2, const float* b2) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (row >= p.B) return;
  const int t = *p.step;
  const int k0 = lane * 16;                       // E == 512
  float v[16];
  if (mode == PREP_EMBED) {
    if (p.chord_embed) {
      load16(p.emb_chord + (size_t)p.gen[(size_t)row * p.cap + t] * p.E + k0, v);
    } else {
      float w[16];
      load16(p.emb_root + (size_t)p.gen_root[(size_t)row * p.cap + t] * p.E + k0, v);
      load16(p.emb_attr + (size_t)p.gen_attr[(size_t)row * p.cap + t] * p.E + k0, w);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += w[e];
    }
    store16(static_cast<T*>(p.ctx) + (size_t)row * p.E + k0, v);
    return;
  }
  load16(static_cast<const T*>(p.r) + (size_t)row * p.E + k0, v);
  ln16(v, g1 + k0, b1 + k0, p.E);
  if (mode == PREP_LN) store16(p.h + (size_t)row * p.E + k0, v);
  else ln16(v, g2 + k0, b2 + k0, p.E);
  store16(static_cast<T*>(p.xn) + (size_t)row * p.E + k0, v);
}

template <typename T>
__device__ __forceinline__ void skinny_epilogue(const DecodeParams& p, const SkinnyArgs& a, int row, int n, float v, int t) {
  if (row >= p.B || n >= a.N) return;
  v += a.bias[n];
  const float scaling = 1.0f / sqrtf((float)(p.E / p.H));   // float(head_dim) ** -0.5, rpr.py:251
  switch (a.emode) {
    case E_QKV: {
      const DecLayer& L = p.layer[a.layer];
      const int dh = p.E / p.H;
      if (n < p.E) {
        p.qbuf[(size_t)row * 3 * p.E + n] = v * scaling;               // rpr.py:328
      } else {
        const int m = (n - p.E) % p.E, hh = m / dh, d = m % dh;
        T* dst = static_cast<T*>(n < 2 * p.E ? L.self_k : L.self_v);
        dst[(((size_t)row * p.H + hh) * p.cap + t) * dh + d] = from_f32<T>(v);
      }
      break;
    }
    case E_Q:
      p.qbuf[(size_t)row * 3 * p.E + n] = v * scaling;
      break;
    case E_RESID:
      static_cast<T*>(p.r)[(size_t)row * p.E + n] = from_f32<T>(v + p.h[(size_t)row * p.E + n]);
      break;
    case E_RELU:
      static_cast<T*>(p.ff)[(size_t)row * p.FF + n] = from_f32<T>(fmaxf(v, 0.f));
      break;
    case E_EMBED:
      v += p.key[row] * p.wc_key[n] + p.pe[(size_t)t * p.E + n];
      p.h[(size_t)row * p.E + n] = v;                                  // residual of the first block
      static_cast<T*>(p.xn)[(size_t)row * p.E + n] = from_f32<T>(v);   // operand of layer 0's QKV projection
      break;
    case E_LOGITS:
      p.logits[(size_t)row * p.vocab + n] = v;
      if (p.logits_all) p.logits_all[((size_t)row * p.cap + t) * p.vocab + n] = v;
      break;
  }
}

template <typename T>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const __grid_constant__ DecodeParams p,
                                                                 const __grid_constant__ SkinnyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * NT, row0 = blockIdx.y * MT;
  const int K = a.K;
  pdl_trigger();

  if constexpr (std::is_same<T, bf16>::value) {
    bf16* As = reinterpret_cast<bf16*>(smem_raw);
    const int AS = SkinnyCfg<bf16>::a_stride(K);
    float* red = reinterpret_cast<float*>(smem_raw + (size_t)MT * AS * 2);
    const int g = lane >> 2, q = lane & 3;
    const int kw = K / 8, kbase = warp * kw, nch = kw / 32;          // nch in {2, 4}
    // weight stream first: its HBM latency overlaps the activation staging below
    const bf16* wrow = static_cast<const bf16*>(a.W) + (size_t)min(n0 + g, a.N - 1) * K + kbase + q * 8;
    uint4 wv[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      if (ch < nch) wv[ch] = ld_nc_v4(wrow + ch * 32);
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 8 * MT * NT);
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    pdl_wait();                                              // predecessor's outputs (and the position counter) are final
    const int t = *p.step;
    // activation rows: already in the compute dtype in global memory -> one bulk async copy per row, no instructions
    const int nrows = min(MT, p.B - row0);
    if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nrows * K * 2));
    if (tid < nrows) {
      bulk_g2s(As + (size_t)tid * AS, static_cast<const bf16*>(a.a_src) + (size_t)(row0 + tid) * K, (uint32_t)(K * 2), bar);
    } else if (tid < MT) {
      for (int k = 0; k < K; k += 8) *reinterpret_cast<uint4*>(As + (size_t)tid * AS + k) = make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(bar, 0);
    __syncthreads();
    float c[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][i] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < nch) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const bf16* ar = As + (size_t)(mt * 16 + g) * AS + kbase + ch * 32 + q * 8;
          const uint4 lo = *reinterpret_cast<const uint4*>(ar);
          const uint4 hi = *reinterpret_cast<const uint4*>(ar + 8 * AS);
          // k-permutation: a thread's 8 contiguous k values feed two MMAs (see decode notes in DESIGN.md)
          mma_bf16_16816(c[mt], lo.x, hi.x, lo.y, hi.y, wv[ch].x, wv[ch].y);
          mma_bf16_16816(c[mt], lo.z, hi.z, lo.w, hi.w, wv[ch].z, wv[ch].w);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      float* r0 = red + ((size_t)warp * MT + mt * 16 + g) * NT + 2 * q;
      r0[0] = c[mt][0]; r0[1] = c[mt][1];
      r0[8 * NT] = c[mt][2]; r0[8 * NT + 1] = c[mt][3];
    }
    __syncthreads();
#pragma unroll
    for (int o = tid; o < MT * NT; o += SK_THREADS) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[(size_t)w * MT * NT + o];
      skinny_epilogue<T>(p, a, row0 + o / NT, n0 + o % NT, s, t);
    }
  } else {
    // fp32 (exact path): warp = 8 video rows, lane = a k slice (16 of every 512 k: four 16-byte chunks, 128 apart), so that the
    // activation and weight loads are coalesced 512-byte requests (staging 64 rows through shared memory and reading the
    // weights as broadcast loads kept the load/store unit busy for ~40 us per launch); 8 rows x 8 features = 64 accumulators
    // per lane, whose 32 lane partials meet in a transpose-reduce butterfly (fixed order) that leaves lane l with the sum of
    // (row l / 8, feature l % 8).  No shared memory, no barrier.
    float acc[8][NT];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr)
#pragma unroll
      for (int c = 0; c < NT; ++c) acc[rr][c] = 0.f;
    const int rw = row0 + 8 * warp;
    pdl_wait();
    const int t = *p.step;
    if (rw < p.B) {
      for (int kb = 0; kb < K; kb += 512) {
        float4 xr[8][4];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const bool ok = rw + rr < p.B;
          const float* xs = static_cast<const float*>(a.a_src) + (size_t)(ok ? rw + rr : 0) * K + kb + 4 * lane;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            xr[rr][j] = (ok && kb + 128 * j + 4 * lane < K) ? __ldg(reinterpret_cast<const float4*>(xs + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < NT; ++c) {
          const float* ws = static_cast<const float*>(a.W) + (size_t)min(n0 + c, a.N - 1) * K + kb + 4 * lane;
          float4 wv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            wv[j] = (kb + 128 * j + 4 * lane < K) ? __ldg(reinterpret_cast<const float4*>(ws + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            float s = acc[rr][c];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s = fmaf(xr[rr][j].x, wv[j].x, s); s = fmaf(xr[rr][j].y, wv[j].y, s);
              s = fmaf(xr[rr][j].z, wv[j].z, s); s = fmaf(xr[rr][j].w, wv[j].w, s);
            }
            acc[rr][c] = s;
          }
        }
      }
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {                           // rows 4 g2 .. 4 g2 + 3 of the warp: 32 sums -> one per lane
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc[4 * g2 + (i >> 3)][i & 7];
#pragma unroll
        for (int h = 16; h >= 1; h >>= 1) {
          const bool up = (lane & h) != 0;
#pragma unroll
          for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
          }
        }
        skinny_epilogue<T>(p, a, rw + 4 * g2 + (lane >> 3), n0 + (lane & 7), v[0], t);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Attention of the single query row over the cached keys/values of one (video, head).
constexpr int DA_THREADS = 256;

template <typename T> struct DecAttnCfg;
template <> struct DecAttnCfg<bf16> { static constexpr int kVec = 8; };   // elements per 16 B
template <> struct DecAttnCfg<float> { static constexpr int kVec = 4; };

template <typename T>
__device__ __forceinline__ void load_vec_f32(const T* p, float* out);
template <>
__device__ __forceinline__ void load_vec_f32<bf16>(const bf16* p, float* out) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = bf16x2_to_f2(u.x); out[0] = f.x; out[1] = f.y;
  f = bf16x2_to_f2(u.y); out[2] = f.x; out[3] = f.y;
  f = bf16x2_to_f2(u.z); out[4] = f.x; out[5] = f.y;
  f = bf16x2_to_f2(u.w); out[6] = f.x; out[7] = f.y;
}
template <>
__device__ __forceinline__ void load_vec_f32<float>(const float* p, float* out) {
  const float4 u = *reinterpret_cast<const float4*>(p);
  out[0] = u.x; out[1] = u.y; out[2] = u.z; out[3] = u.w;
}

template <typename T, int DH>
__global__ void __launch_bounds__(DA_THREADS) dec_attn_kernel(const __grid_constant__ DecodeParams p, int layer,
                                                              int is_cross, int n_max, int e_rows) {
  constexpr int VEC = DecAttnCfg<T>::kVec;
  constexpr int NCH = DH / VEC;                 // 16-byte chunks per row
  constexpr int JG = DA_THREADS / NCH;          // row groups in the PV pass
  extern __shared__ __align__(128) unsigned char dec_attn_smem[];
  T* Ks = reinterpret_cast<T*>(dec_attn_smem);
  T* Vs = Ks + (size_t)n_max * DH;
  T* Es = Vs + (size_t)n_max * DH;              // bf16 self-attention only
  const bool es_in_smem = e_rows > 0;           // Er rows staged in shared memory (bf16 self-attention)
  float* qs = reinterpret_cast<float*>(Es + (size_t)e_rows * DH);
  float* sc = qs + DH;                          // [n_max]
  float* red = sc + ((n_max + 3) & ~3);         // [JG][DH]
  float* stat = red + JG * DH;                  // [16]
  uint64_t* bar = reinterpret_cast<uint64_t*>(stat + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const DecLayer& L = p.layer[layer];
  const int kv_cap = is_cross ? p.S : p.cap;
  const T* kg = static_cast<const T*>(is_cross ? L.cross_k : L.self_k) + ((size_t)b * p.H + h) * kv_cap * DH;
  const T* vg = static_cast<const T*>(is_cross ? L.cross_v : L.self_v) + ((size_t)b * p.H + h) * kv_cap * DH;
  pdl_trigger();
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    if (is_cross) {
      // the video K/V cache is written once before the decode loop: stream it while the predecessor still runs
      const uint32_t bytes = (uint32_t)(p.S * DH * sizeof(T));
      mbar_arrive_expect_tx(bar, bytes * 2);
      bulk_g2s(Ks, kg, bytes, bar);
      bulk_g2s(Vs, vg, bytes, bar);
    }
  }
  pdl_wait();
  const int t = *p.step;
  const int n = is_cross ? p.S : t + 1;
  const T* eg = static_cast<const T*>(L.er) + (size_t)(p.er_len - 1 - t) * DH;   // rows for j = 0..t
  if (tid == 0 && !is_cross) {
    const uint32_t bytes = (uint32_t)(n * DH * sizeof(T));
    const bool want_e = es_in_smem;
    mbar_arrive_expect_tx(bar, bytes * (want_e ? 3 : 2));
    bulk_g2s(Ks, kg, bytes, bar);
    bulk_g2s(Vs, vg, bytes, bar);
    if (want_e) bulk_g2s(Es, eg, bytes, bar);
  }
  if (tid < DH) qs[tid] = p.qbuf[(size_t)b * 3 * p.E + h * DH + tid];
  __syncthreads();
  mbar_wait(bar, 0);

  // ---- scores: thread <-> key j; 16-byte chunks visited in an order rotated by j (conflict-free)
  float lmax = -INFINITY;
  for (int j = tid; j < n; j += DA_THREADS) {
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int cc = (c + j) & (NCH - 1);
      float kv[VEC];
      load_vec_f32<T>(Ks + (size_t)j * DH + cc * VEC, kv);
      if (!is_cross) {
        float ev[VEC];
        if (es_in_smem) load_vec_f32<T>(Es + (size_t)j * DH + cc * VEC, ev);
        else load_vec_f32<T>(eg + (size_t)j * DH + cc * VEC, ev);
#pragma unroll
        for (int e = 0; e < VEC; ++e) kv[e] += ev[e];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) dot = fmaf(qs[cc * VEC + e], kv[e], dot);
    }
    sc[j] = dot;
    lmax = fmaxf(lmax, dot);
  }
  lmax = warp_max(lmax);
  if (lane == 0) stat[warp] = lmax;
  __syncthreads();
  float mx = stat[0];
#pragma unroll
  for (int w = 1; w < DA_THREADS / 32; ++w) mx = fmaxf(mx, stat[w]);
  float lsum = 0.f;
  for (int j = tid; j < n; j += DA_THREADS) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) stat[8 + warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < DA_THREADS / 32; ++w) tot += stat[8 + w];

  // ---- out = P V : thread <-> (row group jg, 16-byte chunk dc)
  const int jg = tid / NCH, dc = tid % NCH;
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  for (int j = jg; j < n; j += JG) {
    float vv[VEC];
    load_vec_f32<T>(Vs + (size_t)j * DH + dc * VEC, vv);
    const float pj = sc[j];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = fmaf(pj, vv[e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) red[jg * DH + dc * VEC + e] = acc[e];
  __syncthreads();
  if (tid < DH) {
    float o = 0.f;
#pragma unroll 8
    for (int g2 = 0; g2 < JG; ++g2) o += red[g2 * DH + tid];
    static_cast<T*>(p.ctx)[(size_t)b * p.E + h * DH + tid] = from_f32<T>(o / tot);
  }
}

// Next token of every video (decode_pick.cuh: greedy arg-max or the sampling branch with its constraints); writes gen[:, t+1]
// (and gen_root / gen_attr when sampling, video_music_transformer.py:1105-1123) and advances the step counter.
__global__ void __launch_bounds__(256) argmax_advance_kernel(const __grid_constant__ DecodeParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_trigger();
  pdl_wait();
  const int t = *p.step;
  for (int b = warp; b < p.B; b += 8) {
    const long long* g = p.gen + (size_t)b * p.cap;
    const float u = p.sample ? p.uniforms[(size_t)b * p.cap + min(t + 1, p.cap - 1)] : 0.f;
    const int bi = pick_token(p.logits + (size_t)b * p.vocab, p.vocab, p.vocab_limit, p.sample, p.max_conseq_N, p.max_conseq_chord,
                              t, u, [&](int k) { return (int)g[t - k]; }, lane);
    if (lane == 0 && t + 1 >= p.primer_len && t + 1 < p.cap) {
      p.gen[(size_t)b * p.cap + t + 1] = bi;
      if (p.sample) {
        long long root, attr;
        chord_root_attr(bi, root, attr);
        p.gen_root[(size_t)b * p.cap + t + 1] = root;
        p.gen_attr[(size_t)b * p.cap + t + 1] = attr;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.step = t + 1;
}

// ----------------------------------------------------------------------------------------------
static bool g_use_pdl = true;

template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const char* what,
                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return kCudaError;
  }
  return kOk;
}

template <typename T>
static int launch_skinny(const DecodeParams& p, const SkinnyArgs& a, cudaStream_t s) {
  const size_t smem = SkinnyCfg<T>::smem(a.K);
  dim3 grid((a.N + NT - 1) / NT, (p.B + MT - 1) / MT);
  return launch_pdl(skinny_gemm_kernel<T>, grid, dim3(SK_THREADS), smem, s, "decode skinny gemm", p, a);
}

template <typename T>
static int launch_attn(const DecodeParams& p, int layer, int is_cross, cudaStream_t s) {
  const int n_max = p.S > p.cap ? p.S : p.cap;
  const int dh = p.E / p.H;
  const int e_rows = 0;       // Er rows are read through L1 (38 KB shared by every CTA); K+V alone keep 2 CTAs per SM
  constexpr int VEC = DecAttnCfg<T>::kVec;
  const size_t smem = (size_t)(2 * n_max + e_rows) * dh * sizeof(T) +
                      sizeof(float) * (dh + ((n_max + 3) & ~3) + (DA_THREADS / (dh / VEC)) * dh + 16) + 16;
  return launch_pdl(dec_attn_kernel<T, 64>, dim3(p.B * p.H), dim3(DA_THREADS), smem, s, "decode attention", p, layer, is_cross,
                    n_max, e_rows);
}

template <typename T>
static int enqueue_step(const DecodeParams& p, cudaStream_t s) {
  int rc;
  SkinnyArgs a{};
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
  auto prep = [&](int mode, const float* g1, const float* b1, const float* g2, const float* b2) {
    return launch_pdl(rows_prep_kernel<T>, dim3((p.B + 7) / 8), dim3(256), 0, s, "decode rows_prep", p, mode, g1, b1, g2, b2);
  };
  // x_t = Linear_chord([emb | key]) + pe[t]
  RUN(prep(PREP_EMBED, nullptr, nullptr, nullptr, nullptr));
  a = SkinnyArgs{p.w_chord, p.b_chord, p.E, p.E, p.ctx, E_EMBED, 0};
  RUN(launch_skinny<T>(p, a, s));
  for (int l = 0; l < p.n_layers; ++l) {
    const DecLayer& L = p.layer[l];
    // self-attention block (rpr.py:56-59)
    a = SkinnyArgs{L.w_qkv, L.b_qkv, 3 * p.E, p.E, p.xn, E_QKV, l};
    RUN(launch_skinny<T>(p, a, s));
    RUN(launch_attn<T>(p, l, 0, s));
    a = SkinnyArgs{L.w_so, L.b_so, p.E, p.E, p.ctx, E_RESID, l};
    RUN(launch_skinny<T>(p, a, s));
    RUN(prep(PREP_LN, L.ln1_g, L.ln1_b, nullptr, nullptr));
    // cross-attention block (rpr.py:62-66)
    a = SkinnyArgs{L.w_cq, L.b_cq, p.E, p.E, p.xn, E_Q, l};
    RUN(launch_skinny<T>(p, a, s));
    RUN(launch_attn<T>(p, l, 1, s));
    a = SkinnyArgs{L.w_co, L.b_co, p.E, p.E, p.ctx, E_RESID, l};
    RUN(launch_skinny<T>(p, a, s));
    RUN(prep(PREP_LN, L.ln2_g, L.ln2_b, nullptr, nullptr));
    // feed-forward block (rpr.py:67-69)
    a = SkinnyArgs{L.w_f1, L.b_f1, p.FF, p.E, p.xn, E_RELU, l};
    RUN(launch_skinny<T>(p, a, s));
    a = SkinnyArgs{L.w_f2, L.b_f2, p.E, p.FF, p.ff, E_RESID, l};
    RUN(launch_skinny<T>(p, a, s));
    if (l + 1 < p.n_layers) RUN(prep(PREP_LN, L.ln3_g, L.ln3_b, nullptr, nullptr));
    else RUN(prep(PREP_LN2, L.ln3_g, L.ln3_b, p.lnf_g, p.lnf_b));       // + decoder final norm (rpr.py:32-33)
  }
  a = SkinnyArgs{p.w_out, p.b_out, p.vocab, p.E, p.xn, E_LOGITS, 0};
  RUN(launch_skinny<T>(p, a, s));
  RUN(launch_pdl(argmax_advance_kernel, dim3(1), dim3(256), 0, s, "decode argmax", p));
#undef RUN
  return kOk;
}

long long decode_kernel_launches_per_step(const DecodeParams& p) { return 2 + 11LL * p.n_layers + 2; }

template <typename T>
static int set_attrs() {
  static bool done = false;
  if (done) return kOk;
  cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dec_attn_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_last_error("decode: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return kCudaError;
  }
  done = true;
  return kOk;
}

// The videos [b0, b0+nb) as an independent decode problem: every buffer pointer is advanced to the sub-batch.
static DecodeParams sub_batch(const DecodeParams& p, int b0, int nb, int idx) {
  DecodeParams s = p;
  const size_t es = p.dtype == 0 ? 4 : 2;
  const int dh = p.E / p.H;
  auto adv = [](const void* q, size_t bytes) { return q ? static_cast<const void*>(static_cast<const char*>(q) + bytes) : q; };
  auto advm = [](void* q, size_t bytes) { return q ? static_cast<void*>(static_cast<char*>(q) + bytes) : q; };
  s.B = nb;
  s.key = p.key + b0;
  s.gen = p.gen + (size_t)b0 * p.cap;
  s.gen_root = p.gen_root + (size_t)b0 * p.cap;
  s.gen_attr = p.gen_attr + (size_t)b0 * p.cap;
  s.step = p.step + idx;
  s.h = p.h + (size_t)b0 * p.E;
  s.r = advm(p.r, (size_t)b0 * p.E * es);
  s.qbuf = p.qbuf + (size_t)b0 * 3 * p.E;
  s.ctx = advm(p.ctx, (size_t)b0 * p.E * es);
  s.ff = advm(p.ff, (size_t)b0 * p.FF * es);
  s.xn = advm(p.xn, (size_t)b0 * p.E * es);
  s.logits = p.logits + (size_t)b0 * p.vocab;
  if (p.logits_all) s.logits_all = p.logits_all + (size_t)b0 * p.cap * p.vocab;
  for (int l = 0; l < p.n_layers; ++l) {
    s.layer[l].self_k = advm(p.layer[l].self_k, (size_t)b0 * p.H * p.cap * dh * es);
    s.layer[l].self_v = advm(p.layer[l].self_v, (size_t)b0 * p.H * p.cap * dh * es);
    s.layer[l].cross_k = adv(p.layer[l].cross_k, (size_t)b0 * p.H * p.S * dh * es);
    s.layer[l].cross_v = adv(p.layer[l].cross_v, (size_t)b0 * p.H * p.S * dh * es);
  }
  return s;
}

constexpr int kMaxSplit = 8;

// use_graph: 0 = plain launches; 1 = one CUDA graph per position (51 kernel nodes) replayed n_steps times;
// k in [2, 8] = the same graph with the batch cut into k sub-batches whose 51-kernel chains are independent branches,
// so that the latency-bound linear kernels of one sub-batch overlap the bandwidth-bound attention of another.
// p.step must point to kMaxSplit ints holding the same position.
template <typename T>
static int decode_run_t(const DecodeParams& p, int n_steps, int use_graph, cudaStream_t stream) {
  int rc = set_attrs<T>();
  if (rc) return rc;
  int n_split = use_graph > 1 ? use_graph : 1;
  if (n_split > kMaxSplit) n_split = kMaxSplit;
  if (n_split > p.B) n_split = p.B;
  const int per = (p.B + n_split - 1) / n_split;
  n_split = (p.B + per - 1) / per;
  DecodeParams subs[kMaxSplit];
  for (int i = 0; i < n_split; ++i) subs[i] = sub_batch(p, i * per, (i + 1) * per <= p.B ? per : p.B - i * per, i);
  if (!use_graph) {
    for (int i = 0; i < n_steps; ++i) {
      rc = enqueue_step<T>(subs[0], stream);
      if (rc) return rc;
    }
    return kOk;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t cs = stream;
  cudaStream_t own = nullptr;
  cudaStream_t side[kMaxSplit] = {nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[kMaxSplit] = {nullptr};
  auto cleanup = [&]() {
    for (int i = 1; i < n_split; ++i) {
      if (side[i]) cudaStreamDestroy(side[i]);
      if (join_ev[i]) cudaEventDestroy(join_ev[i]);
    }
    if (fork_ev) cudaEventDestroy(fork_ev);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (own) { cudaStreamSynchronize(own); cudaStreamDestroy(own); }
  };
  if (cs == nullptr || cs == cudaStreamLegacy) {   // the legacy default stream cannot be captured
    if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) {
      set_last_error("decode: cudaStreamCreate failed");
      return kCudaError;
    }
    cudaStreamSynchronize(stream);
    cs = own;
  }
  cudaError_t e = cudaSuccess;
  for (int i = 1; i < n_split && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess && n_split > 1) e = cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming);
  if (e != cudaSuccess) { set_last_error("decode: stream/event setup: %s", cudaGetErrorString(e)); cleanup(); return kCudaError; }
  e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { set_last_error("decode: begin capture: %s", cudaGetErrorString(e)); cleanup(); return kCudaError; }
  if (n_split > 1) {
    cudaEventRecord(fork_ev, cs);
    for (int i = 1; i < n_split; ++i) cudaStreamWaitEvent(side[i], fork_ev, 0);
  }
  rc = enqueue_step<T>(subs[0], cs);
  for (int i = 1; i < n_split && !rc; ++i) {
    rc = enqueue_step<T>(subs[i], side[i]);
    cudaEventRecord(join_ev[i], side[i]);
    cudaStreamWaitEvent(cs, join_ev[i], 0);
  }
  e = cudaStreamEndCapture(cs, &graph);
  if (rc || e != cudaSuccess) {
    if (!rc) set_last_error("decode: end capture: %s", cudaGetErrorString(e));
    cleanup();
    return rc ? rc : kCudaError;
  }
  e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) {
    set_last_error("decode: graph instantiate: %s", cudaGetErrorString(e));
    cleanup();
    return kCudaError;
  }
  for (int i = 0; i < n_steps && e == cudaSuccess; ++i) e = cudaGraphLaunch(exec, cs);
  if (!own) cudaStreamSynchronize(cs);           // the graph and its side streams are destroyed below
  cleanup();
  if (e != cudaSuccess) { set_last_error("decode: graph launch: %s", cudaGetErrorString(e)); return kCudaError; }
  return kOk;
}

// Launches one decode kernel kind `reps` rounds over all layers (bench.py times it with CUDA events for the roofline
// line).  kind 0: self-attention, 1: cross-attention, 2: QKV skinny GEMM, 3: FFN1 skinny GEMM.
template <typename T>
static int decode_probe_t(const DecodeParams& p, int kind, int reps, cudaStream_t s) {
  int rc = set_attrs<T>();
  if (rc) return rc;
  for (int r = 0; r < reps; ++r)
    for (int l = 0; l < p.n_layers; ++l) {
      const DecLayer& L = p.layer[l];
      if (kind == 0 || kind == 1) rc = launch_attn<T>(p, l, kind, s);
      else if (kind == 2) {
        SkinnyArgs a{L.w_qkv, L.b_qkv, 3 * p.E, p.E, p.xn, E_Q, l};
        rc = launch_skinny<T>(p, a, s);
      } else {
        SkinnyArgs a{L.w_f1, L.b_f1, p.FF, p.E, p.xn, E_RELU, l};
        rc = launch_skinny<T>(p, a, s);
      }
      if (rc) return rc;
    }
  return kOk;
}

int decode_probe(const DecodeParams& p, int kind, int reps, cudaStream_t stream) {
  V2M_REQUIRE(kind >= 0 && kind <= 3 && reps >= 1, "decode_probe: bad kind/reps");
  return p.dtype == 0 ? decode_probe_t<float>(p, kind, reps, stream) : decode_probe_t<bf16>(p, kind, reps, stream);
}

int decode_run(const DecodeParams& p, int n_steps, int use_graph, cudaStream_t stream) {
  g_use_pdl = getenv("V2M_NO_PDL") == nullptr;
  V2M_REQUIRE(p.n_layers >= 1 && p.n_layers <= kMaxDecLayers, "decode: n_layers %d out of range", p.n_layers);
  V2M_REQUIRE(p.E % p.H == 0 && p.E / p.H == 64, "decode: head_dim must be 64 (E=%d H=%d)", p.E, p.H);
  V2M_REQUIRE(p.E == 512 && p.FF % 512 == 0 && p.FF <= 2048, "decode: d_model must be 512 and dim_feedforward a multiple of 512 (E=%d FF=%d)", p.E, p.FF);
  V2M_REQUIRE(p.cap <= p.er_len, "decode: cap %d exceeds er_len %d (rpr.py:426-450 fails for L > er_len)", p.cap, p.er_len);
  V2M_REQUIRE(p.B >= 1 && n_steps >= 0, "decode: bad B=%d n_steps=%d", p.B, n_steps);
  V2M_REQUIRE(!p.sample || (p.uniforms && p.max_conseq_chord >= 1), "decode: sampling needs uniforms and max_conseq_chord >= 1");
  if (p.dtype == 0) return decode_run_t<float>(p, n_steps, use_graph, stream);
  if (p.dtype == 1) return decode_run_t<bf16>(p, n_steps, use_graph, stream);
  set_last_error("decode: dtype %d unsupported", p.dtype);
  return kUnsupported;
}

}  // namespace v2m
