// Persistent, cluster-scoped KV-cached decode: the WHOLE greedy generation loop in ONE kernel launch.
//
// Same dataflow and arithmetic as decode.cu (see there for the reference lines), different schedule.
// The 51-kernels-per-position version is bound by launch latency and by every CTA re-reading the
// 64-row activation tile from L2.  Here the batch is split over thread-block clusters instead:
//   * a cluster of CS CTAs (16, non-portable size; 8 as fall-back) owns R <= 8 videos for the whole
//     run; clusters never talk to each other, so there is no grid-wide barrier and no relaunch;
//   * inside a cluster every linear layer is split by output features (tile of 8 features per work
//     item, K split so that all 8 warps have one or three items), activations of the R rows are
//     exchanged through tiny L2-resident scratch rows (p.qbuf / p.r / p.ctx / p.ff / p.h) and phases
//     are separated by the hardware cluster barrier (barrier.cluster, release/acquire) instead of
//     kernel boundaries;
//   * attention over the caches: the R*H (video, head) problems of a cluster are dealt round-robin
//     to its CTAs; K and V blocks stream through a 2-deep shared-memory ring with cp.async.bulk
//     (TMA engine) + mbarriers, and the block for the *next* attention phase is already in flight
//     while the linear phases in between execute, so HBM keeps streaming across phase boundaries.
// Weights are read once per cluster per step (8 clusters -> 8 x 32 MB per step out of L2); the
// K/V caches are read exactly once per step from HBM.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>
#include <stdio.h>

namespace v2m {

namespace dc {

constexpr int THREADS = 256;
constexpr int RMAX = 16;           // rows (videos) per cluster = rows of one MMA tile
constexpr int DH = 64;
constexpr int AS = 1024 + 32;      // activation row pitch in bf16 elements (K <= 1024)
constexpr int MAX_ITEMS = 32;
constexpr int HC = 64;             // residual columns kept per CTA (owned feature tiles of a 512-wide layer)
constexpr int KV_ROWS = 304;       // capacity of one ring slot (rows of 64 bf16)

enum AMode { A_PLAIN_T = 0, A_PLAIN_F32 = 1, A_LN = 2, A_LN2 = 3, A_EMBED = 4 };
enum EMode { E_QKV = 0, E_RESID = 1, E_Q = 2, E_RELU = 3, E_EMBED = 4, E_LOGITS = 5 };

struct GArgs {
  const bf16* W; const float* bias; int N, K;
  int amode; const void* a_src;
  const float* g1; const float* b1; const float* g2; const float* b2;
  int emode; int layer;
};

struct Smem {
  bf16 act[RMAX * AS];                 // A operand of the current linear phase
  float hres[RMAX * HC];               // residual stream (latest LayerNorm output): only the columns this CTA owns
  float red[MAX_ITEMS * RMAX * 8];     // K-split partial sums
  bf16 kv[2][2][KV_ROWS * DH];         // ring: [slot][K|V][row][64]
  float sc[KV_ROWS + 16];
  float qs[DH];
  float pvred[32 * DH];
  float stat[32];
  uint64_t full[2];
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_idx() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// Hardware cluster barrier; release/acquire at cluster scope orders the global-memory exchange rows.
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// exchange rows are written by other CTAs of the cluster: read them around L1
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint4 ldcg4u(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void ln16(float* v, const float* g, const float* b) {
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) sum += v[e];
  const float mean = warp_sum(sum) * (1.f / 512.f);
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) { const float d = v[e] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / 512.f) + 1e-5f);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (v[e] - mean) * rstd * __ldg(g + e) + __ldg(b + e);
}

__device__ __forceinline__ void load16_bf16(const bf16* src, float* v) {      // 16 consecutive bf16, L1 bypass
  const uint4 a = ldcg4u(src), b = ldcg4u(src + 8);
  float2 f;
  f = bf16x2_to_f2(a.x); v[0] = f.x; v[1] = f.y;   f = bf16x2_to_f2(a.y); v[2] = f.x; v[3] = f.y;
  f = bf16x2_to_f2(a.z); v[4] = f.x; v[5] = f.y;   f = bf16x2_to_f2(a.w); v[6] = f.x; v[7] = f.y;
  f = bf16x2_to_f2(b.x); v[8] = f.x; v[9] = f.y;   f = bf16x2_to_f2(b.y); v[10] = f.x; v[11] = f.y;
  f = bf16x2_to_f2(b.z); v[12] = f.x; v[13] = f.y; f = bf16x2_to_f2(b.w); v[14] = f.x; v[15] = f.y;
}
__device__ __forceinline__ void load16_f32(const float* src, float* v, bool bypass) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = bypass ? ldcg4(src + 4 * q) : __ldg(reinterpret_cast<const float4*>(src) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void store16_bf16(bf16* dst, const float* v) {
  uint4 a, b;
  a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
  b.x = f2_to_bf16x2(v[8], v[9]);   b.y = f2_to_bf16x2(v[10], v[11]); b.z = f2_to_bf16x2(v[12], v[13]); b.w = f2_to_bf16x2(v[14], v[15]);
  *reinterpret_cast<uint4*>(dst) = a;
  *reinterpret_cast<uint4*>(dst + 8) = b;
}

// Optional phase timestamps (globaltimer, ns) written by CTA 0 of cluster 0: set with v2m_debug_set_timestamps().
__device__ unsigned long long* g_ts = nullptr;
__device__ int g_ts_cap = 0;
__device__ __forceinline__ void stamp(int& n) {
  if (g_ts && blockIdx.x == 0 && threadIdx.x == 0 && n < g_ts_cap) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_ts[n] = t;
  }
  ++n;
}

struct Ctx {
  int crank, csize, row0, R, t;
};

// ---- stage the R activation rows of this cluster into shared memory (8 warps, rows warp and warp + 8) ----------
__device__ __forceinline__ void keep_residual(Smem& sm, const Ctx& c, int r, int lane, const float* v) {
  // a lane holds features [16*lane, 16*lane+16) = feature tiles 2*lane and 2*lane+1; keep the tiles this CTA owns
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int tile = 2 * lane + half;
    if (tile % c.csize == c.crank) {
      float* h = sm.hres + (size_t)r * HC + (tile / c.csize) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = v[half * 8 + e];
    }
  }
}

__device__ __noinline__ void stage_rows(const DecodeParams& p, const GArgs& a, Smem& sm, const Ctx& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_ln = (a.amode == A_LN || a.amode == A_LN2);
  for (int seg = 0; seg < a.K; seg += 512) {
    float v[2][16];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = warp + 8 * i, row = c.row0 + r;
      if (r >= c.R) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[i][e] = 0.f;
        continue;
      }
      if (a.amode == A_PLAIN_T || is_ln) {
        load16_bf16(static_cast<const bf16*>(a.a_src) + (size_t)row * a.K + seg + lane * 16, v[i]);
      } else if (a.amode == A_PLAIN_F32) {
        load16_f32(static_cast<const float*>(a.a_src) + (size_t)row * a.K + seg + lane * 16, v[i], true);
      } else {                                 // A_EMBED (video_music_transformer.py:984-989)
        if (p.chord_embed) {
          const long long tok = __ldcg(p.gen + (size_t)row * p.cap + c.t);
          load16_f32(p.emb_chord + (size_t)tok * p.E + lane * 16, v[i], false);
        } else {
          float w[16];
          load16_f32(p.emb_root + (size_t)p.gen_root[(size_t)row * p.cap + c.t] * p.E + lane * 16, v[i], false);
          load16_f32(p.emb_attr + (size_t)p.gen_attr[(size_t)row * p.cap + c.t] * p.E + lane * 16, w, false);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[i][e] += w[e];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = warp + 8 * i;
      if (r < c.R) {
        if (is_ln) {
          ln16(v[i], a.g1 + lane * 16, a.b1 + lane * 16);
          keep_residual(sm, c, r, lane, v[i]);
          if (a.amode == A_LN2) ln16(v[i], a.g2 + lane * 16, a.b2 + lane * 16);
        } else if (a.amode == A_PLAIN_F32) {   // layer-0 input x_t is the residual of the first block
          keep_residual(sm, c, r, lane, v[i]);
        }
      }
      store16_bf16(sm.act + (size_t)r * AS + seg + lane * 16, v[i]);
    }
  }
}

__device__ __forceinline__ void gemm_epilogue(const DecodeParams& p, const GArgs& a, const Smem& sm, const Ctx& c, int r, int n,
                                              int nloc, float v) {
  const int row = c.row0 + r;
  v += __ldg(a.bias + n);
  const float scaling = 0.125f;                // float(64) ** -0.5, rpr.py:251
  switch (a.emode) {
    case E_QKV: {
      const DecLayer& L = p.layer[a.layer];
      if (n < p.E) {
        p.qbuf[(size_t)row * 3 * p.E + n] = v * scaling;
      } else {
        const int m = (n - p.E) % p.E, hh = m / DH, d = m % DH;
        bf16* dst = static_cast<bf16*>(n < 2 * p.E ? L.self_k : L.self_v);
        dst[(((size_t)row * p.H + hh) * p.cap + c.t) * DH + d] = __float2bfloat16_rn(v);
      }
      break;
    }
    case E_Q:
      p.qbuf[(size_t)row * 3 * p.E + n] = v * scaling;
      break;
    case E_RESID:
      static_cast<bf16*>(p.r)[(size_t)row * p.E + n] = __float2bfloat16_rn(v + sm.hres[r * HC + nloc]);
      break;
    case E_RELU:
      static_cast<bf16*>(p.ff)[(size_t)row * p.FF + n] = __float2bfloat16_rn(fmaxf(v, 0.f));
      break;
    case E_EMBED:
      p.h[(size_t)row * p.E + n] = v + __ldg(p.key + row) * __ldg(p.wc_key + n) + __ldg(p.pe + (size_t)c.t * p.E + n);
      break;
    case E_LOGITS:
      p.logits[(size_t)row * p.vocab + n] = v;
      if (p.logits_all) p.logits_all[((size_t)row * p.cap + c.t) * p.vocab + n] = v;
      break;
  }
}

// ---- one linear layer for the R rows of the cluster; this CTA computes feature tiles crank, crank+csize, ... --
__device__ __noinline__ void phase_gemm(const DecodeParams& p, const GArgs& a, Smem& sm, const Ctx& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int n_tiles = (a.N + 7) >> 3;
  const int n_my = (n_tiles - c.crank + c.csize - 1) / c.csize;            // tiles crank + i*csize
  // K split: minimise rounds(items over 8 warps) x K per item; at most 512 k per item (16 loads in flight per lane)
  int KS = a.K > 512 ? a.K / 512 : 1;
  {
    int best = KS, best_cost = 1 << 30;
    for (int ks = KS; ks <= 8 && a.K / ks >= 64 && n_my * ks <= MAX_ITEMS; ks <<= 1) {
      const int cost = ((n_my * ks + 7) >> 3) * (a.K / ks);
      if (cost < best_cost) { best_cost = cost; best = ks; }
    }
    KS = best;
  }
  const int k_item = a.K / KS, nl = k_item >> 5;                           // 16-byte weight loads per lane and item
  const int n_items = n_my * KS;

  // weights of my first item first: they do not depend on the previous phase
  uint4 wv[16];
  auto load_w = [&](int item) {
    const int tile = c.crank + (item / KS) * c.csize, ks = item % KS;
    const bf16* wrow = a.W + (size_t)min(tile * 8 + g, a.N - 1) * a.K + ks * k_item + q * 8;
#pragma unroll
    for (int ch = 0; ch < 16; ++ch)
      if (ch < nl) wv[ch] = ld_nc_v4(wrow + ch * 32);
  };
  if (warp < n_items) load_w(warp);
  stage_rows(p, a, sm, c);
  __syncthreads();
  for (int item = warp; item < n_items; item += 8) {
    if (item != warp) load_w(item);
    const int ks = item % KS;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* ar = sm.act + (size_t)g * AS + ks * k_item + q * 8;
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) {
      if (ch < nl) {
        const uint4 lo = *reinterpret_cast<const uint4*>(ar + ch * 32);
        const uint4 hi = *reinterpret_cast<const uint4*>(ar + 8 * AS + ch * 32);
        mma_bf16_16816(acc, lo.x, hi.x, lo.y, hi.y, wv[ch].x, wv[ch].y);
        mma_bf16_16816(acc, lo.z, hi.z, lo.w, hi.w, wv[ch].z, wv[ch].w);
      }
    }
    float* rd = sm.red + (size_t)item * (RMAX * 8);
    rd[g * 8 + 2 * q] = acc[0];
    rd[g * 8 + 2 * q + 1] = acc[1];
    rd[(g + 8) * 8 + 2 * q] = acc[2];
    rd[(g + 8) * 8 + 2 * q + 1] = acc[3];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < n_my * RMAX * 8; o += THREADS) {
    const int tl = o >> 7, r = (o >> 3) & 15, col = o & 7;
    const int n = (c.crank + tl * c.csize) * 8 + col;
    if (r < c.R && n < a.N) {
      float s = 0.f;
      for (int ks = 0; ks < KS; ++ks) s += sm.red[(size_t)(tl * KS + ks) * (RMAX * 8) + r * 8 + col];
      gemm_epilogue(p, a, sm, c, r, n, tl * 8 + col, s);
    }
  }
}

// ---- attention of one query row per (video, head) over the cached keys / values ------------------------------
struct Ring { uint32_t seq; };     // running number of attention tasks consumed by this CTA (slot = seq & 1)

__device__ __forceinline__ void task_coords(const Ctx& c, int k, int& r, int& h) { r = k % c.R; h = k / c.R; }

// thread 0: start the bulk copies of the first `n` cached rows of task k of (layer, is_cross) into ring slot `slot`
__device__ __noinline__ void issue_kv(const DecodeParams& p, Smem& sm, const Ctx& c, int layer, int is_cross, int k, int n, int slot) {
  int r, h;
  task_coords(c, k, r, h);
  const DecLayer& L = p.layer[layer];
  const int kv_cap = is_cross ? p.S : p.cap;
  const size_t off = (((size_t)(c.row0 + r)) * p.H + h) * kv_cap * DH;
  const bf16* kg = static_cast<const bf16*>(is_cross ? L.cross_k : L.self_k) + off;
  const bf16* vg = static_cast<const bf16*>(is_cross ? L.cross_v : L.self_v) + off;
  const uint32_t bytes = (uint32_t)n * DH * 2;
  // generic-proxy accesses (the ring slot just read by all threads; cache rows written by other CTAs in earlier
  // phases) must be ordered before the async-proxy copy
  asm volatile("fence.proxy.async;" ::: "memory");
  mbar_arrive_expect_tx(&sm.full[slot], 2 * bytes);
  bulk_g2s(sm.kv[slot][0], kg, bytes, &sm.full[slot]);
  bulk_g2s(sm.kv[slot][1], vg, bytes, &sm.full[slot]);
}

struct NextTask { int valid, layer, is_cross, k, n; };

__device__ __noinline__ void attn_task(const DecodeParams& p, Smem& sm, const Ctx& c, int layer, int is_cross, int k, uint32_t seq,
                          int n_prefetched, const NextTask& nx, uint32_t* ph) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int r, h;
  task_coords(c, k, r, h);
  const int b = c.row0 + r;
  const int slot = seq & 1;
  const DecLayer& L = p.layer[layer];
  const int n = is_cross ? p.S : c.t + 1;
  bf16* Ks = sm.kv[slot][0];
  bf16* Vs = sm.kv[slot][1];
  // the slot of the following task is free (its previous user finished before the last __syncthreads): prefetch now
  if (tid == 0 && nx.valid && nx.n > 0) issue_kv(p, sm, c, nx.layer, nx.is_cross, nx.k, nx.n, slot ^ 1);
  if (tid < DH) sm.qs[tid] = __ldcg(p.qbuf + (size_t)b * 3 * p.E + h * DH + tid);
  if (n_prefetched > 0) {                      // a slot's mbarrier completes once per copy actually issued into it
    mbar_wait(&sm.full[slot], ph[slot]);
    ph[slot] ^= 1u;
  }
  if (!is_cross) {
    // row t was produced in this step by the QKV phase (other CTAs of the cluster): plain loads around L1
    const size_t off = ((((size_t)b) * p.H + h) * p.cap + c.t) * DH;
    if (tid < 8) *reinterpret_cast<uint4*>(Ks + (size_t)c.t * DH + tid * 8) = ldcg4u(static_cast<const bf16*>(L.self_k) + off + tid * 8);
    else if (tid < 16) *reinterpret_cast<uint4*>(Vs + (size_t)c.t * DH + (tid - 8) * 8) = ldcg4u(static_cast<const bf16*>(L.self_v) + off + (tid - 8) * 8);
  }
  __syncthreads();
  const bf16* eg = static_cast<const bf16*>(L.er) + (size_t)(p.er_len - 1 - c.t) * DH;
  float lmax = -INFINITY;
  for (int j = tid; j < n; j += THREADS) {
    float dot = 0.f;
#pragma unroll
    for (int cc0 = 0; cc0 < 8; ++cc0) {
      const int cc = (cc0 + j) & 7;
      const uint4 ku = *reinterpret_cast<const uint4*>(Ks + (size_t)j * DH + cc * 8);
      float kv[8];
      float2 f;
      f = bf16x2_to_f2(ku.x); kv[0] = f.x; kv[1] = f.y; f = bf16x2_to_f2(ku.y); kv[2] = f.x; kv[3] = f.y;
      f = bf16x2_to_f2(ku.z); kv[4] = f.x; kv[5] = f.y; f = bf16x2_to_f2(ku.w); kv[6] = f.x; kv[7] = f.y;
      if (!is_cross) {                         // + Er[er_len-1-(t-j)]  (rpr.py:391-395,439-455)
        const uint4 eu = __ldg(reinterpret_cast<const uint4*>(eg + (size_t)j * DH + cc * 8));
        f = bf16x2_to_f2(eu.x); kv[0] += f.x; kv[1] += f.y; f = bf16x2_to_f2(eu.y); kv[2] += f.x; kv[3] += f.y;
        f = bf16x2_to_f2(eu.z); kv[4] += f.x; kv[5] += f.y; f = bf16x2_to_f2(eu.w); kv[6] += f.x; kv[7] += f.y;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) dot = fmaf(sm.qs[cc * 8 + e], kv[e], dot);
    }
    sm.sc[j] = dot;
    lmax = fmaxf(lmax, dot);
  }
  lmax = warp_max(lmax);
  if (lane == 0) sm.stat[warp] = lmax;
  __syncthreads();
  float mx = sm.stat[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sm.stat[w]);
  float lsum = 0.f;
  for (int j = tid; j < n; j += THREADS) {
    const float e = expf(sm.sc[j] - mx);
    sm.sc[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) sm.stat[8 + warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += sm.stat[8 + w];
  const int jg = tid >> 3, dcn = tid & 7;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int j = jg; j < n; j += 32) {
    const uint4 vu = *reinterpret_cast<const uint4*>(Vs + (size_t)j * DH + dcn * 8);
    const float pj = sm.sc[j];
    float2 f;
    f = bf16x2_to_f2(vu.x); acc[0] = fmaf(pj, f.x, acc[0]); acc[1] = fmaf(pj, f.y, acc[1]);
    f = bf16x2_to_f2(vu.y); acc[2] = fmaf(pj, f.x, acc[2]); acc[3] = fmaf(pj, f.y, acc[3]);
    f = bf16x2_to_f2(vu.z); acc[4] = fmaf(pj, f.x, acc[4]); acc[5] = fmaf(pj, f.y, acc[5]);
    f = bf16x2_to_f2(vu.w); acc[6] = fmaf(pj, f.x, acc[6]); acc[7] = fmaf(pj, f.y, acc[7]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm.pvred[jg * DH + dcn * 8 + e] = acc[e];
  __syncthreads();
  if (tid < DH) {
    float o = 0.f;
#pragma unroll 8
    for (int g2 = 0; g2 < 32; ++g2) o += sm.pvred[g2 * DH + tid];
    static_cast<bf16*>(p.ctx)[(size_t)b * p.E + h * DH + tid] = __float2bfloat16_rn(o / tot);
  }
  __syncthreads();                             // slot, sc, qs, pvred are free again
}

__global__ void __launch_bounds__(THREADS, 1) decode_cluster_kernel(const __grid_constant__ DecodeParams p, int t0, int n_steps, int rows_per_cluster) {
  extern __shared__ __align__(128) unsigned char dc_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(dc_smem_raw);
  Ctx c;
  c.crank = (int)cluster_ctarank();
  c.csize = (int)cluster_nctarank();
  c.row0 = (int)cluster_idx() * rows_per_cluster;
  c.R = min(rows_per_cluster, p.B - c.row0);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (c.R <= 0) return;                        // whole cluster idle (uniform across the cluster)
  const int n_tasks = c.R * p.H;
  const int n_mine = (n_tasks - c.crank + c.csize - 1) / c.csize;          // tasks crank + i*csize
  uint32_t seq = 0;
  uint32_t ph[2] = {0u, 0u};
  int ts_n = 0;
  stamp(ts_n);
  // very first block of the run
  if (tid == 0 && n_mine > 0 && t0 > 0) issue_kv(p, sm, c, 0, 0, c.crank, t0, 0);

  for (int step = 0; step < n_steps; ++step) {
    c.t = t0 + step;
    GArgs a;
    // x_t = Linear_chord([emb | key]) + pe[t]  -> p.h rows
    a = GArgs{static_cast<const bf16*>(p.w_chord), p.b_chord, p.E, p.E, A_EMBED, nullptr, nullptr, nullptr, nullptr, nullptr, E_EMBED, 0};
    phase_gemm(p, a, sm, c);
    cluster_sync();
    stamp(ts_n);
    for (int l = 0; l < p.n_layers; ++l) {
      const DecLayer& L = p.layer[l];
      const DecLayer* P = l > 0 ? &p.layer[l - 1] : nullptr;
      a = GArgs{static_cast<const bf16*>(L.w_qkv), L.b_qkv, 3 * p.E, p.E, l == 0 ? A_PLAIN_F32 : A_LN,
                l == 0 ? (const void*)p.h : (const void*)p.r, P ? P->ln3_g : nullptr, P ? P->ln3_b : nullptr, nullptr, nullptr, E_QKV, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
      for (int i = 0; i < n_mine; ++i, ++seq) {       // self-attention: rows [0,t) were prefetched, row t is fresh
        NextTask nx;
        if (i + 1 < n_mine) nx = NextTask{1, l, 0, c.crank + (i + 1) * c.csize, c.t};
        else nx = NextTask{1, l, 1, c.crank, p.S};
        attn_task(p, sm, c, l, 0, c.crank + i * c.csize, seq, c.t, nx, ph);
      }
      cluster_sync();
      stamp(ts_n);
      a = GArgs{static_cast<const bf16*>(L.w_so), L.b_so, p.E, p.E, A_PLAIN_T, p.ctx, nullptr, nullptr, nullptr, nullptr, E_RESID, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
      a = GArgs{static_cast<const bf16*>(L.w_cq), L.b_cq, p.E, p.E, A_LN, p.r, L.ln1_g, L.ln1_b, nullptr, nullptr, E_Q, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
      for (int i = 0; i < n_mine; ++i, ++seq) {       // cross-attention over the video memory
        NextTask nx;
        if (i + 1 < n_mine) nx = NextTask{1, l, 1, c.crank + (i + 1) * c.csize, p.S};
        else if (l + 1 < p.n_layers) nx = NextTask{1, l + 1, 0, c.crank, c.t};
        else if (step + 1 < n_steps) nx = NextTask{1, 0, 0, c.crank, c.t + 1};
        else nx = NextTask{0, 0, 0, 0, 0};
        attn_task(p, sm, c, l, 1, c.crank + i * c.csize, seq, p.S, nx, ph);
      }
      cluster_sync();
      stamp(ts_n);
      a = GArgs{static_cast<const bf16*>(L.w_co), L.b_co, p.E, p.E, A_PLAIN_T, p.ctx, nullptr, nullptr, nullptr, nullptr, E_RESID, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
      a = GArgs{static_cast<const bf16*>(L.w_f1), L.b_f1, p.FF, p.E, A_LN, p.r, L.ln2_g, L.ln2_b, nullptr, nullptr, E_RELU, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
      a = GArgs{static_cast<const bf16*>(L.w_f2), L.b_f2, p.E, p.FF, A_PLAIN_T, p.ff, nullptr, nullptr, nullptr, nullptr, E_RESID, l};
      phase_gemm(p, a, sm, c);
      cluster_sync();
      stamp(ts_n);
    }
    const DecLayer& LL = p.layer[p.n_layers - 1];
    a = GArgs{static_cast<const bf16*>(p.w_out), p.b_out, p.vocab, p.E, A_LN2, p.r, LL.ln3_g, LL.ln3_b, p.lnf_g, p.lnf_b, E_LOGITS, 0};
    phase_gemm(p, a, sm, c);
    cluster_sync();
    stamp(ts_n);
    // greedy arg-max over [:vocab_limit] (first index wins on ties), CTA 0 of the cluster, one warp per row
    if (c.crank == 0) {
      const int lane = tid & 31, warp = tid >> 5;
      for (int rr = warp; rr < c.R; rr += 8) {
        const int b = c.row0 + rr;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int n = lane; n < p.vocab_limit; n += 32) {
          const float v = __ldcg(p.logits + (size_t)b * p.vocab + n);
          if (v > best) { best = v; bi = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0 && c.t + 1 >= p.primer_len && c.t + 1 < p.cap) p.gen[(size_t)b * p.cap + c.t + 1] = bi;
      }
    }
    cluster_sync();
    stamp(ts_n);
  }
  if (blockIdx.x == 0 && tid == 0) *p.step = t0 + n_steps;
}

}  // namespace dc

// Runs n_steps positions starting at t0 (host value of the step counter) with the cluster kernel.
// Returns kUnsupported when the configuration is outside what the kernel covers (caller falls back to the
// per-kernel graph path of decode.cu, which is also the fp32 path).
int decode_debug_set_timestamps(unsigned long long* buf, int cap) {
  cudaError_t e = cudaMemcpyToSymbol(dc::g_ts, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(dc::g_ts_cap, &cap, sizeof(cap));
  if (e != cudaSuccess) { set_last_error("decode_debug_set_timestamps: %s", cudaGetErrorString(e)); return kCudaError; }
  return kOk;
}

int decode_run_cluster(const DecodeParams& p, int t0, int n_steps, cudaStream_t stream) {
  if (p.dtype != 1 || p.E != 512 || p.H != 8 || p.FF % 512 != 0 || p.FF > 1024 || p.S > dc::KV_ROWS || p.cap > dc::KV_ROWS)
    return kUnsupported;
  const size_t smem = sizeof(dc::Smem);
  // How many clusters of each size can be co-resident on this part (GPC floor-plan dependent, probed once).
  static int max_active[17];
  static bool probed = false;
  if (!probed) {
    cudaFuncSetAttribute(dc::decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dc::decode_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs = 1; cs <= 16; ++cs) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs, 1, 1);
      cfg.blockDim = dim3(dc::THREADS, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      max_active[cs] = (cudaOccupancyMaxActiveClusters(&ncl, dc::decode_cluster_kernel, &cfg) == cudaSuccess) ? ncl : 0;
    }
    cudaGetLastError();
    probed = true;
  }
  // Pick (cluster size, clusters): all clusters co-resident, <= 16 videos per cluster, as many SMs as possible;
  // among equals the fewest clusters (every cluster streams all weights from L2 once per position).
  int cs_sel = 0, ncl_sel = 0;
  if (const char* ov = getenv("V2M_CLUSTER")) {
    if (sscanf(ov, "%d,%d", &cs_sel, &ncl_sel) != 2) cs_sel = ncl_sel = 0;
  }
  if (cs_sel == 0) {
    long best = -1;
    for (int cs = 16; cs >= 8; --cs) {                    // cs >= 8: a CTA keeps at most HC = 64 residual columns
      if (max_active[cs] <= 0) continue;
      int ncl = max_active[cs];
      if (ncl > p.B) ncl = p.B;
      const int rows = (p.B + ncl - 1) / ncl;
      if (rows > dc::RMAX) continue;                       // would need more than one wave of clusters
      ncl = (p.B + rows - 1) / rows;                        // drop clusters that would be empty
      const long score = (long)cs * ncl * 100 - ncl;
      if (score > best) { best = score; cs_sel = cs; ncl_sel = ncl; }
    }
  }
  if (cs_sel < 8 || cs_sel > 16 || ncl_sel <= 0) return kUnsupported;
  const int rows = (p.B + ncl_sel - 1) / ncl_sel;
  if (rows > dc::RMAX) return kUnsupported;
  if (getenv("V2M_VERBOSE")) {
    fprintf(stderr, "v2m: decode cluster kernel: %d clusters x %d CTAs, %d videos per cluster, %zu B smem; max active:", ncl_sel,
            cs_sel, rows, smem);
    for (int cs = 4; cs <= 16; ++cs) fprintf(stderr, " %d:%d", cs, max_active[cs]);
    fprintf(stderr, "\n");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl_sel * cs_sel, 1, 1);
  cfg.blockDim = dim3(dc::THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs_sel; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dc::decode_cluster_kernel, p, t0, n_steps, rows);
  if (e != cudaSuccess) {
    set_last_error("decode_run_cluster: launch failed: %s", cudaGetErrorString(e));
    return kCudaError;
  }
  return check_launch("decode_cluster_kernel");
}

}  // namespace v2m
