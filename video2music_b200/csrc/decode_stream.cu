// Streamed cluster decode: the whole greedy generation loop in ONE launch, fed by bulk-copy rings instead of
// per-phase kernel launches.
//
// Same arithmetic as decode.cu (see there for the reference lines: video_music_transformer.py:984-1084,
// rpr.py:24-70,387-455); different decomposition:
//   * a cluster of 8 CTAs owns R <= 8 videos for the whole run; CTA c of the cluster owns HEAD c (q/k/v
//     projection rows, the self- and cross-attention of that head over the caches), the output features
//     [64c, 64c+64) of every 512-wide linear layer and the hidden features [128c, 128c+128) of the FFN.
//     QKV projection -> attention never leaves the CTA; everything else is an all-gather of a small slice
//     through distributed shared memory (st.shared::cluster) followed by a cluster barrier.
//   * weights are pre-packed (engine.pack_fragments) in mma.m16n8k16 A-fragment order: one tile of 16 output
//     features x 512 k is 16 KB of contiguous memory read with conflict-free LDS.128 (weights are the M=16
//     operand, the videos the N=8 operand).
//   * two producer warps walk the static schedule (weight tiles, Er rows, cache rows, in consumption order) and keep
//     a ring of 32 KB slots full with cp.async.bulk, so the ring always holds the next ~160 KB the consumers need,
//     whatever the phase.  Measured on B200 (profiles/r01_tma_ingest_microbench.txt): one thread sustains one bulk
//     copy per ~300 ns whatever its size, hence 32 KB copies (two weight tiles / 128 cache positions) and two
//     producer threads (slot s belongs to producer s % 2, so the uses of a slot stay ordered).  Producers never
//     synchronise with other CTAs: they run ahead as far as the ring allows.
//   * the caches of this path hold, per (video, head, position), one 256-byte row [K(64) | V(64)] whose eight
//     16-byte chunks are XOR-swizzled with (position & 7): a bulk copy lands rows that ldmatrix reads without bank
//     conflicts, so q.K^T and P.V run on the tensor cores (mma.sync m16n8k16, keys / dims as the M operand).
//     A chunk (128 positions of one video) is processed by the warp whose index equals its ring slot: it produces a partial softmax
//     (max, sum, unnormalised output) for it; the partials of a video are combined once per phase (split softmax).
//   * an all-gather IS its barrier: every CTA stages its slice in shared memory and sends it to each of the 8 CTAs with
//     one cp.async.bulk (shared::cta -> shared::cluster) that completes on the RECEIVER's mbarrier; a CTA waits on its own
//     barrier for the 8 slices.  No fence, no separate arrive, producers not involved.  (The one remaining plain cluster
//     barrier per position, for the logits in global memory, is an mbarrier with remote arrives.)
#include "common.cuh"
#include "kernels.h"
#include "decode_pick.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace v2m {

namespace ds {

constexpr int CS = 8;                 // CTAs per cluster == heads
constexpr int MAX_SLOTS = 6;          // ring slots (as many as fit, at least 4)
constexpr int SLOT_BYTES = 32768;     // two weight tiles (16 x 512 bf16 each), 128 cache positions or 256 Er rows
constexpr int TILE_BYTES = 16384;
constexpr int NCONS = 256;            // consumer threads (8 warps)
constexpr int NPROD = 2;              // producer warps (one issuing lane each): slot s is filled by producer s % NPROD
constexpr int THREADS = NCONS + 32 * NPROD;
constexpr int DH = 64, E = 512, FF = 1024;
constexpr int XP = E + 8;             // pitch (bf16) of 512-wide operand rows: conflict-free B-fragment loads
// gathered operands are slice-major: [8 source CTAs][R rows][slice pitch]; one source CTA's block is contiguous
constexpr int CP = 72;                // pitch (bf16) of a 64-wide attention-context slice row
constexpr int HSP = 136;              // pitch (bf16) of a 128-wide FFN-hidden slice row
constexpr int YP = 64;                // pitch (fp32) of a 64-wide pre-LayerNorm slice row
constexpr int KCH = 128;              // cache positions per ring slot (256 B each)
constexpr int ECH = 256;              // Er rows per ring slot (128 B each)
constexpr int SMEM_MAX = 227 * 1024;
constexpr int CW = 68;                // floats per softmax partial: m, l, pad, pad, o[64]

struct Layout {
  uint32_t ring, bars, bias, xn, ctxg, hg, yg, resid, stage, qs, knew, vnew, srel, comb, snew, tok, total;
  int srel_pitch, nslot, nch_max;
  uint32_t stage_bytes;
};
__host__ __device__ inline Layout make_layout(int R, int cap, int S) {
  Layout L;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
  L.bars = take(128);                 // full[MAX_SLOTS] empty[MAX_SLOTS] cbar gbar[3]
  L.bias = take(kMaxDecLayers * 576 * 4);   // this CTA's bias slices of every layer: q k v so cq co (64 each) f1 (128) f2 (64)
  L.xn = take(R * XP * 2);
  L.ctxg = take(CS * R * CP * 2);
  L.hg = take(CS * R * HSP * 2);
  L.yg = take(CS * R * YP * 4);
  L.resid = take(R * DH * 4);
  L.stage_bytes = (uint32_t)((R * HSP * 2 + 127) & ~127);      // largest slice block (FFN hidden)
  L.stage = take(2 * L.stage_bytes);                            // two staging buffers, used alternately
  L.qs = take(R * DH * 4);
  L.knew = take(R * DH * 2);
  L.vnew = take(R * DH * 2);
  L.srel_pitch = (cap + 3) & ~3;
  L.srel = take(R * L.srel_pitch * 4);
  L.nch_max = ((cap > S ? cap : S) + KCH - 1) / KCH;   // chunks per video in the longer of the two attention kinds
  L.comb = take(R * L.nch_max * CW * 4);
  L.snew = take(R * 4);
  L.tok = take(R * 64);               // per video: token, root, attr (int64) + the last 8 tokens (int32)
  o = (o + 1023u) & ~1023u;
  L.ring = o;
  int ns = ((int)SMEM_MAX - (int)o) / SLOT_BYTES;
  L.nslot = ns > MAX_SLOTS ? MAX_SLOTS : ns;
  L.total = o + (uint32_t)(L.nslot > 0 ? L.nslot : 0) * SLOT_BYTES;
  return L;
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_hw() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// arrive on the mbarrier at shared::cluster address `addr` (another CTA's).  Relaxed: the caller issues ONE
// fence.acq_rel.cluster before its 8 arrives (8 release-arrives cost 8 cluster-scope fences, ~2 us measured).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// non-blocking test: a thread suspended in try_wait is not woken by REMOTE arrivals before its time slice ends (~2 us measured)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// Debug markers (compile with -DV2M_DEBUG_MARKS): when the timestamp buffer has >= 8192 entries, warp w of CTA b stores
// its latest progress code at ts[4096 + b*16 + w] (pass a pinned host buffer: it survives a device fault).
#ifdef V2M_DEBUG_MARKS
__device__ unsigned long long* volatile g_mark = nullptr;
__device__ __forceinline__ void mark(int code) {
  if (g_mark && (threadIdx.x & 31) == 0) {
    g_mark[4096 + blockIdx.x * 16 + (threadIdx.x >> 5)] = (unsigned long long)code;
    __threadfence_system();
  }
}
#else
__device__ __forceinline__ void mark(int) {}
#endif
#ifdef V2M_FINE_STAMPS
#define FSTAMP(cx) stamp(cx)
#else
#define FSTAMP(cx)
#endif
#ifdef V2M_CHUNK_STAMPS
#define CSTAMP(cx) stamp(cx)
#else
#define CSTAMP(cx)
#endif

// One ring of 32 KB slots: every use of a slot flips its parity bit; slots are taken round-robin.  Producer and
// consumers walk the same static schedule, so both sides derive (slot, parity) of every copy without communication.
struct Ring {
  uint64_t* full;
  uint64_t* empty;
  unsigned char* base;
  uint32_t pm;      // bit s: parity of the next use of slot s
  int cur, nslot;
  __device__ __forceinline__ int next(uint32_t& par) {
    const int s = cur;
    cur = cur + 1 == nslot ? 0 : cur + 1;
    par = (pm >> s) & 1u;
    pm ^= 1u << s;
    return s;
  }
  __device__ __forceinline__ unsigned char* slot(int s) const { return base + (size_t)s * SLOT_BYTES; }
  // Consumer side: the previous use of the slot must have been CONSUMED before the parity test on `full` is
  // unambiguous (try_wait cannot tell use k from use k+2); this warp is among those that complete the current use.
  __device__ __forceinline__ void wait_full(int s, uint32_t par) const {
    mbar_wait(empty + s, par ^ 1u);
    mbar_wait(full + s, par);
  }
  // Producer side: both producers walk the whole schedule, producer `id` issues the copies of the slots it owns
  __device__ __forceinline__ void put(int id, const void* src, uint32_t bytes) {
    uint32_t par;
    const int s = next(par);
    if ((s % NPROD) != id) return;
    // the whole producer warp walks the schedule in convergent code and one elected lane issues: the operands of the bulk copy
    // stay in uniform registers (under `if (lane == 0)` every copy went through an ELECT / R2UR waterfall, see common.cuh)
    mbar_wait_u(empty + s, par ^ 1u);
    if (elect_one()) {
      mbar_arrive_expect_tx(full + s, bytes);
      bulk_g2s(slot(s), src, bytes, full + s);
    }
    __syncwarp();
  }
};

struct Ctx {
  unsigned char* sm;
  Layout L;
  Ring rg;
  uint64_t* cbar;
  uint32_t cphase;
  uint64_t* gbar;       // [3] gather barriers: 0 attention context, 1 pre-LayerNorm rows, 2 FFN hidden
  uint32_t gphase;      // bit i: parity of the next use of gbar[i]
  uint32_t n_gather;    // gathers issued so far (selects the staging buffer)
  __device__ __forceinline__ unsigned char* stage_buf() const { return sm + L.stage + (n_gather & 1u) * L.stage_bytes; }
  int c;          // CTA rank in the cluster == head
  int row0, R;    // videos [row0, row0 + R)
  int t;
  int ph;         // debug: phase counter
  unsigned long long* ts; int ts_cap; int ts_n;   // optional phase timestamps (thread 0 of CTA 0)
  template <typename T> __device__ __forceinline__ T* at(uint32_t off) const { return reinterpret_cast<T*>(sm + off); }
};

__device__ __forceinline__ void stamp(Ctx& cx);

// Cluster barrier for the consumer warps: every CTA's thread 0 arrives on the barrier of all 8 CTAs (its own included)
// after a CTA-level barrier, with release semantics at cluster scope (cumulative over the CTA barrier, so the
// st.shared::cluster writes of all its threads are covered); everybody then waits on the local barrier (acquire.cluster).
__device__ __forceinline__ void cluster_sync_cons(Ctx& cx) {
  mark(cx.ph * 1000 + 900);
  cons_sync();
  if (threadIdx.x == 0) {                                // one poller: an acquire at cluster scope is not free
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    const uint32_t a = smem_u32(cx.cbar);
#pragma unroll
    for (int peer = 0; peer < CS; ++peer) mbar_arrive_remote(mapa(a, peer));
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(cx.cbar, cx.cphase)) {
      if (++spins > (1u << 28)) { printf("v2m: cluster barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    }
  }
  cx.cphase ^= 1u;
  cons_sync();                                           // thread 0's acquire is cumulative over this CTA barrier
  mark(cx.ph * 1000 + 901);
  FSTAMP(cx);
}

// ---- one linear layer: n_tiles (<= 8) tiles of 16 output features, tile i -> warp i ----------------------------------
// K = 512: two tiles per 32 KB copy (single: one tile per copy); K = 1024: one tile per copy.
// epi(tile, c): c (bias already added): c[0] = (feature g, video 2q), c[1] = (g, 2q+1), c[2] = (g+8, 2q), c[3] = (g+8, 2q+1)
// bias_of(tile): pointer to the tile's 16 biases, loaded BEFORE the wait on the ring (global latency off the critical path)
template <class BiasOf, class Epi>
// Operand rows: element k of row r lives at x_off + ((k >> slice_log2) * R + r) * pitch + (k & (slice - 1))  (slice-major
// gathered buffers; slice_log2 >= log2(K) for a plain row-major operand).
__device__ __forceinline__ void gemm_phase(Ctx& cx, int n_tiles, int K, bool single, uint32_t x_off, int pitch, int slice_log2,
                                           BiasOf bias_of, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  cons_sync();                                   // operand rows written by other warps
  ++cx.ph; mark(cx.ph * 1000 + 100 + n_tiles);
  const int tpc = (K == 512 && !single) ? 2 : 1;                // tiles per copy
  const int n_copies = (n_tiles + tpc - 1) / tpc;
  int my_slot = 0;
  uint32_t my_par = 0;
  for (int j = 0; j < n_copies; ++j) {                          // identical in every warp: keeps the ring state in step
    uint32_t par;
    const int s = cx.rg.next(par);
    if (j == warp / tpc) { my_slot = s; my_par = par; }
  }
  if (warp < n_tiles) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bf16* xr = cx.at<bf16>(x_off) + (size_t)min(g, cx.R - 1) * pitch + 2 * q;
    const int slice_stride = cx.R * pitch, slice_mask = (1 << slice_log2) - 1;
    const float* bp = bias_of(warp);
    const float blo = bp[g], bhi = bp[g + 8];
    cx.rg.wait_full(my_slot, my_par);
    mark(cx.ph * 1000 + 120 + n_tiles);
    const uint4* ap = reinterpret_cast<const uint4*>(cx.rg.slot(my_slot) + (tpc == 2 ? (warp & 1) * TILE_BYTES : 0)) + lane;
    const int nks = K >> 4;
#pragma unroll 8
    for (int ks = 0; ks < nks; ++ks) {
      const uint4 a = ap[ks * 32];
      const bf16* xk = xr + ((ks * 16) >> slice_log2) * slice_stride + ((ks * 16) & slice_mask);
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(xk);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(xk + 8);
      mma_bf16_16816(acc[ks & 3], a.x, a.y, a.z, a.w, b0, b1);
    }
    __syncwarp();
    if (lane == 0) {
      const bool alone = tpc == 1 || ((warp & 1) == 0 && warp + 1 >= n_tiles);
      mbar_arrive_cnt(cx.rg.empty + my_slot, alone ? 8 : 4);
    }
    float c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
    c[0] += blo; c[1] += blo; c[2] += bhi; c[3] += bhi;
    epi(warp, c);
  }
  FSTAMP(cx);
}

// ---- all-gather with the barrier built in.  The caller has written this CTA's slice block (R rows, `block_bytes`
// contiguous bytes, a multiple of 16) to cx.stage_buf().  Warp w sends it to CTA w of the cluster with one bulk copy that
// completes (complete_tx) on the receiver's gather barrier `which`; everybody then waits for the 8 blocks addressed to
// this CTA.  dst_off: offset of the gathered buffer; block c of it belongs to source CTA c.
// Staging buffers alternate: buffer k is rewritten at gather k+2, after the wait of gather k+1 -- which every peer can
// only have sent after receiving block k from this CTA, i.e. after this CTA's copy k has completed.
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster_addr, const void* src, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_cluster_addr), "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void gather(Ctx& cx, int which, uint32_t dst_off, uint32_t block_bytes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows (generic proxy) -> bulk copy (async proxy)
  cons_sync();
  uint64_t* bar = cx.gbar + which;
  if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, CS * block_bytes);
  if (lane == 0) {
    const uint32_t base = smem_u32(cx.sm);
    bulk_s2c(mapa(base + dst_off + cx.c * block_bytes, warp), cx.stage_buf(), block_bytes, mapa(smem_u32(bar), warp));
  }
  ++cx.n_gather;
  mbar_wait(bar, (cx.gphase >> which) & 1u);
  cx.gphase ^= 1u << which;
  FSTAMP(cx);
}

// LayerNorm affine parameters of this lane's 16 columns, fetched BEFORE the cluster barrier (global latency hidden)
struct LnP { float g[16], b[16]; };
__device__ __forceinline__ void ln_load(LnP& lp, const float* g, const float* b) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + lane * 16) + q4), bb = __ldg(reinterpret_cast<const float4*>(b + lane * 16) + q4);
    lp.g[4 * q4] = gg.x; lp.g[4 * q4 + 1] = gg.y; lp.g[4 * q4 + 2] = gg.z; lp.g[4 * q4 + 3] = gg.w;
    lp.b[4 * q4] = bb.x; lp.b[4 * q4 + 1] = bb.y; lp.b[4 * q4 + 2] = bb.z; lp.b[4 * q4 + 3] = bb.w;
  }
}
__device__ __forceinline__ void ln_row16(float* v, const LnP& lp) {
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) sum += v[e];
  const float mean = warp_sum(sum) * (1.f / 512.f);
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) { const float d = v[e] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / 512.f) + 1e-5f);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (v[e] - mean) * rstd * lp.g[e] + lp.b[e];
}

// After the cluster barrier that completes a y all-gather: xn = LayerNorm(y) (optionally twice), resid = own slice (fp32).
// n_ln == 0: no normalisation (layer-0 input).  Warp r handles row r.
__device__ __forceinline__ void rows_norm(const Ctx& cx, int n_ln, const LnP& p1, const LnP& p2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < cx.R) {
    float v[16];
    const float* y = cx.at<float>(cx.L.yg) + ((size_t)(lane >> 2) * cx.R + warp) * YP + (lane & 3) * 16;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 t4 = reinterpret_cast<const float4*>(y)[q4];
      v[4 * q4] = t4.x; v[4 * q4 + 1] = t4.y; v[4 * q4 + 2] = t4.z; v[4 * q4 + 3] = t4.w;
    }
    if (n_ln >= 1) ln_row16(v, p1);
    if ((lane >> 2) == cx.c) {                       // lanes 4c..4c+3 hold this CTA's 64 residual columns
      float* rs = cx.at<float>(cx.L.resid) + warp * DH + (lane & 3) * 16;
#pragma unroll
      for (int e = 0; e < 16; ++e) rs[e] = v[e];
    }
    if (n_ln >= 2) ln_row16(v, p2);
    uint4 a, b;
    a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
    b.x = f2_to_bf16x2(v[8], v[9]);   b.y = f2_to_bf16x2(v[10], v[11]); b.z = f2_to_bf16x2(v[12], v[13]); b.w = f2_to_bf16x2(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(cx.at<bf16>(cx.L.xn) + (size_t)warp * XP + lane * 16);
    dst[0] = a;
    dst[1] = b;
  }
}

// ---- attention over the cache: head cx.c of the R videos -------------------------------------------------------------
// Cache-ring stream of the phase: (self only) ceil(t/256) copies of Er rows [er_len-1-t, er_len-1) from the copy of Er
// swizzled for that start row, consumed by all warps; then, video by video, ceil(n/128) copies of 128 positions;
// each copy is processed by the warp whose index equals its ring slot: it computes S = K q on the tensor cores (keys are the M operand, q replicated over the
// 8 columns), a chunk-local softmax and O_c = V^T p (dims are the M operand, V read through ldmatrix.trans).
__device__ __forceinline__ void attention_phase(Ctx& cx, const DecodeParams& p, int layer, bool is_self) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int n = is_self ? cx.t : p.S;                  // cached positions to visit
  const int nch = (n + KCH - 1) / KCH;
  const float* qs = cx.at<float>(cx.L.qs);
  float* srel = cx.at<float>(cx.L.srel);
  float* comb = cx.at<float>(cx.L.comb);
  float* snew = cx.at<float>(cx.L.snew);
  const int ri = lane & 7, mi = lane >> 3;             // ldmatrix: this lane addresses row ri of 8x8 matrix mi
  cons_sync();                                         // q_s / knew / vnew written
  ++cx.ph; mark(cx.ph * 1000 + (is_self ? 200 : 300));
#ifdef V2M_EXP_NOATTN
  if (false) {                                         // timing experiment: the linear chain + exchanges alone (wrong results)
#else
  if (is_self) {
#endif
    if (warp < cx.R) {                                 // score of position t itself (distance 0: Er[er_len-1])
      const DecLayer& Ld = p.layer[layer];
      const float2 kn = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(cx.at<bf16>(cx.L.knew) + warp * DH + 2 * lane));
      const float2 en = bf16x2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(Ld.er) + (size_t)(p.er_len - 1) * DH + 2 * lane)));
      const float2 qq = *reinterpret_cast<const float2*>(qs + warp * DH + 2 * lane);
      const float sn = warp_sum(qq.x * (kn.x + en.x) + qq.y * (kn.y + en.y));
      if (lane == 0) snew[warp] = sn;
    }
    // relative term for the cached keys of ALL videos at once: srel[video][j] = q_video . Er[er_len-1-t+j]  (rpr.py:426-455)
    // A = 16 Er rows, B = q of video (column n) -> one MMA serves 8 videos.  Warp w: rows [32w, 32w+32) of each copy.
    const int nec = (n + ECH - 1) / ECH;
    for (int ci = 0; ci < nec; ++ci) {
      uint32_t par;
      const int s = cx.rg.next(par);
      const int row_base = ci * ECH + warp * 32;
      cx.rg.wait_full(s, par);
      if (row_base < n) {                               // warp-uniform
        uint32_t qb[4][2];
        const float* qrow = qs + (size_t)min(g, cx.R - 1) * DH + 2 * q;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const float2 lo = *reinterpret_cast<const float2*>(qrow + ks * 16), hi = *reinterpret_cast<const float2*>(qrow + ks * 16 + 8);
          qb[ks][0] = f2_to_bf16x2(lo.x, lo.y);
          qb[ks][1] = f2_to_bf16x2(hi.x, hi.y);
        }
        const uint32_t base = smem_u32(cx.rg.slot(s));
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          float cacc[4] = {0.f, 0.f, 0.f, 0.f};
          const int row = warp * 32 + tl * 16 + (mi & 1) * 8 + ri;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(base + row * 128 + (((2 * ks + (mi >> 1)) ^ ri) << 4), a0, a1, a2, a3);
            mma_bf16_16816(cacc, a0, a1, a2, a3, qb[ks][0], qb[ks][1]);
          }
          const int j0 = row_base + tl * 16 + g;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = j0 + (i >> 1) * 8, vid = 2 * q + (i & 1);
            if (vid < cx.R && j < n) srel[vid * cx.L.srel_pitch + j] = cacc[i];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_cnt(cx.rg.empty + s, 1);
    }
    cons_sync();                                       // srel rows are read by other warps below
    FSTAMP(cx);
  }
  // a chunk is processed by the warp whose index equals its ring slot: that warp then meets every use of the slot during
  // the phase in order and can never be two uses ahead of it (the parity waits cannot tell use k from use k+2)
#ifdef V2M_EXP_NOATTN
  const int total = 0;
#else
  const int total = cx.R * nch;
#endif
  for (int i = 0; i < total; ++i) {
    uint32_t my_par;
    const int my_slot = cx.rg.next(my_par);            // ring state in step in every warp
    if (my_slot != warp) continue;
    const int r = i / nch, ci = i - r * nch;
    const int nvalid = min(KCH, n - ci * KCH);
    const int nkt = (nvalid + 15) >> 4;                // key tiles with at least one valid position
    uint32_t qb[4][2];
    {
      const float* qrow = qs + (size_t)r * DH + 2 * q;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 lo = *reinterpret_cast<const float2*>(qrow + ks * 16), hi = *reinterpret_cast<const float2*>(qrow + ks * 16 + 8);
        qb[ks][0] = f2_to_bf16x2(lo.x, lo.y);
        qb[ks][1] = f2_to_bf16x2(hi.x, hi.y);
      }
    }
    CSTAMP(cx);
    cx.rg.wait_full(my_slot, my_par);
    CSTAMP(cx);
    mark(cx.ph * 1000 + (is_self ? 210 : 310) + ci);
    const uint32_t base = smem_u32(cx.rg.slot(my_slot));
    // ---- scores of 16 keys per tile: lane (g, q) ends up with key g (c0 == c1) and key g + 8 (c2 == c3)
    float s[8][2];
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      s[kt][0] = -INFINITY; s[kt][1] = -INFINITY;
      if (kt < nkt) {
        float cacc[4] = {0.f, 0.f, 0.f, 0.f};
        const int row = kt * 16 + (mi & 1) * 8 + ri;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4(base + row * 256 + (((2 * ks + (mi >> 1)) ^ ri) << 4), a0, a1, a2, a3);
          mma_bf16_16816(cacc, a0, a1, a2, a3, qb[ks][0], qb[ks][1]);
        }
        const int k0 = kt * 16 + g, j0 = ci * KCH + k0;
        if (k0 < nvalid) s[kt][0] = cacc[0] + (is_self ? srel[r * cx.L.srel_pitch + j0] : 0.f);
        if (k0 + 8 < nvalid) s[kt][1] = cacc[2] + (is_self ? srel[r * cx.L.srel_pitch + j0 + 8] : 0.f);
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) m = fmaxf(m, fmaxf(s[kt][0], s[kt][1]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));    // finite: every chunk holds at least one valid key
    float l = 0.f;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      s[kt][0] = __expf(s[kt][0] - m);                 // masked keys: exp(-inf) = 0
      s[kt][1] = __expf(s[kt][1] - m);
      l += s[kt][0] + s[kt][1];
    }
    l += __shfl_xor_sync(0xffffffffu, l, 4);
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    // ---- O_c = V^T p: p of key k lives in the lanes with g == k % 8; the B fragment wants keys (2q, 2q+1, 2q+8, 2q+9)
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      if (kt < nkt) {
        const float p0 = __shfl_sync(0xffffffffu, s[kt][0], 8 * q), p1 = __shfl_sync(0xffffffffu, s[kt][0], 8 * q + 4);
        const float p2 = __shfl_sync(0xffffffffu, s[kt][1], 8 * q), p3 = __shfl_sync(0xffffffffu, s[kt][1], 8 * q + 4);
        const uint32_t b0 = f2_to_bf16x2(p0, p1), b1 = f2_to_bf16x2(p2, p3);
        const int row = kt * 16 + (mi >> 1) * 8 + ri;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4_t(base + row * 256 + 128 + (((2 * dt + (mi & 1)) ^ ri) << 4), a0, a1, a2, a3);
          mma_bf16_16816(o[dt], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_cnt(cx.rg.empty + my_slot, 8);
    CSTAMP(cx);
    mark(cx.ph * 1000 + (is_self ? 220 : 320) + ci);
    float* cb = comb + ((size_t)r * cx.L.nch_max + ci) * CW;
    if (lane == 0) { cb[0] = m; cb[1] = l; }
    if (q == 0) {
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) { cb[4 + dt * 16 + g] = o[dt][0]; cb[4 + dt * 16 + g + 8] = o[dt][2]; }
    }
  }
  cons_sync();
  FSTAMP(cx);
  mark(cx.ph * 1000 + (is_self ? 240 : 340));
  // ---- combine the chunk partials of every video (and, for self-attention, the new key / value): thread <-> (video, dim)
  bf16* st = reinterpret_cast<bf16*>(cx.stage_buf());
  for (int i = threadIdx.x; i < cx.R * DH; i += NCONS) {
    const int r = i >> 6, d = i & 63;
    const float* cb = comb + (size_t)r * cx.L.nch_max * CW;
    float M = is_self ? snew[r] : -INFINITY;
    for (int ci = 0; ci < nch; ++ci) M = fmaxf(M, cb[ci * CW]);
    float Lt = 0.f, ov = 0.f;
    if (is_self) {
      const float f = __expf(snew[r] - M);
      Lt = f;
      ov = f * __bfloat162float(cx.at<bf16>(cx.L.vnew)[r * DH + d]);
    }
    for (int ci = 0; ci < nch; ++ci) {
      const float f = __expf(cb[ci * CW] - M);
      Lt = fmaf(cb[ci * CW + 1], f, Lt);
      ov = fmaf(cb[ci * CW + 4 + d], f, ov);
    }
#ifdef V2M_EXP_NOATTN
    st[r * CP + d] = __float2bfloat16_rn(0.01f * (float)(d & 7));
    (void)ov; (void)Lt;
#else
    st[r * CP + d] = __float2bfloat16_rn(ov / Lt);
#endif
  }
  gather(cx, 0, cx.L.ctxg, (uint32_t)cx.R * CP * 2);
  mark(cx.ph * 1000 + (is_self ? 250 : 350));
}

// Optional phase timestamps (globaltimer ns) from thread 0 of CTA 0 of cluster 0.
__device__ __forceinline__ void stamp(Ctx& cx) {
  if (cx.ts && blockIdx.x == 0 && threadIdx.x == 0 && cx.ts_n < cx.ts_cap && cx.ts_n < 4096) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    cx.ts[cx.ts_n] = v;
  }
  ++cx.ts_n;
}

__global__ void __launch_bounds__(THREADS, 1) decode_stream_kernel(const __grid_constant__ DecodeParams p, int t0, int n_steps,
                                                                   int rows_per_cluster, unsigned long long* ts, int ts_cap) {
  extern __shared__ __align__(1024) unsigned char ds_smem[];
  Ctx cx;
  cx.sm = ds_smem;
  cx.c = (int)cluster_ctarank();
  cx.row0 = (int)cluster_id() * rows_per_cluster;
  cx.R = min(rows_per_cluster, p.B - cx.row0);
  cx.L = make_layout(rows_per_cluster, p.cap, p.S);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + cx.L.bars);
  cx.rg = Ring{bars, bars + MAX_SLOTS, ds_smem + cx.L.ring, 0u, 0, cx.L.nslot};
  cx.cbar = bars + 2 * MAX_SLOTS;
  cx.cphase = 0u;
  cx.gbar = cx.cbar + 1;
  cx.gphase = 0u;
  cx.n_gather = 0u;
  cx.ph = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  if (cx.R <= 0) return;                                 // uniform over the cluster
#ifdef V2M_DEBUG_MARKS
  if (ts_cap >= 8192 && tid == 0 && blockIdx.x == 0) g_mark = ts;
#endif
  if (tid == 0) {
    for (int i = 0; i < cx.L.nslot; ++i) { mbar_init(cx.rg.full + i, 1); mbar_init(cx.rg.empty + i, 8); }
    mbar_init(cx.cbar, CS);
    for (int i = 0; i < 3; ++i) mbar_init(cx.gbar + i, 1);
    fence_barrier_init();
  }
  // stale ring rows (beyond the valid positions of a chunk) reach the tensor cores multiplied by p = 0: they must be finite
  for (uint32_t i = tid; i < (uint32_t)cx.L.nslot * SLOT_BYTES / 16; i += THREADS)
    reinterpret_cast<uint4*>(ds_smem + cx.L.ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async;" ::: "memory");
  {
    float* bs = reinterpret_cast<float*>(ds_smem + cx.L.bias);
    for (int i = tid; i < p.n_layers * 576; i += THREADS) {
      const DecLayer& Ld = p.layer[i / 576];
      const int j = i % 576, hc = cx.c * DH;
      float v;
      if (j < 192) v = Ld.b_qkv[(j >> 6) * E + hc + (j & 63)];
      else if (j < 256) v = Ld.b_so[hc + j - 192];
      else if (j < 320) v = Ld.b_cq[hc + j - 256];
      else if (j < 384) v = Ld.b_co[hc + j - 320];
      else if (j < 512) v = Ld.b_f1[cx.c * 128 + j - 384];
      else v = Ld.b_f2[hc + j - 512];
      bs[i] = v;
    }
  }
  __syncthreads();
  cluster_sync_hw();                                     // every CTA's barriers are initialised before any remote arrive
  const int c = cx.c, NL = p.n_layers;
  const size_t hb = (size_t)p.H;
  const int n_vt = (p.vocab + 15) / 16;

  if (warp >= 8) {
    // ============================== producer warps: the whole schedule in consumption order ==============================
    {
      Ring& rg = cx.rg;
      const int id = warp - 8;
      auto tiles512 = [&](const void* base_w, int first, int count) {      // two K=512 tiles per copy
        const char* b = static_cast<const char*>(base_w) + (size_t)first * TILE_BYTES;
        for (int i = 0; i < count; i += 2) rg.put(id, b + (size_t)i * TILE_BYTES, (uint32_t)min(2, count - i) * TILE_BYTES);
      };
      for (int step = 0; step < n_steps; ++step) {
        const int t = t0 + step;
        tiles512(p.w_chord, c * 4, 4);
        for (int l = 0; l < NL; ++l) {
          const DecLayer& Ld = p.layer[l];
          tiles512(Ld.w_qkv, c * 4, 4);
          tiles512(Ld.w_qkv, 32 + c * 4, 4);
          tiles512(Ld.w_qkv, 64 + c * 4, 4);
          // Er rows [er_len-1-t, er_len-1) from the copy swizzled for that start row, then the cached self K|V rows
#ifndef V2M_EXP_NOATTN
          const int start = p.er_len - 1 - t;
          const char* er = static_cast<const char*>(Ld.er_sw) + ((size_t)(start & 7) * p.er_len + start) * (DH * 2);
          for (int r0 = 0; r0 < t; r0 += ECH) rg.put(id, er + (size_t)r0 * DH * 2, (uint32_t)min(ECH, t - r0) * DH * 2);
          for (int r = 0; r < cx.R; ++r)
            for (int r0 = 0; r0 < t; r0 += KCH) {
              const size_t off = ((((size_t)(cx.row0 + r)) * hb + c) * p.cap + r0) * (2 * DH);
              rg.put(id, static_cast<const bf16*>(Ld.self_k) + off, (uint32_t)min(KCH, t - r0) * 2 * DH * 2);
            }
#endif
          tiles512(Ld.w_so, c * 4, 4);
          tiles512(Ld.w_cq, c * 4, 4);
#ifndef V2M_EXP_NOATTN
          for (int r = 0; r < cx.R; ++r)
            for (int r0 = 0; r0 < p.S; r0 += KCH) {
              const size_t off = ((((size_t)(cx.row0 + r)) * hb + c) * p.S + r0) * (2 * DH);
              rg.put(id, static_cast<const bf16*>(Ld.cross_k) + off, (uint32_t)min(KCH, p.S - r0) * 2 * DH * 2);
            }
#endif
          tiles512(Ld.w_co, c * 4, 4);
          tiles512(Ld.w_f1, c * 8, 8);
          for (int i = 0; i < 4; ++i) rg.put(id, static_cast<const char*>(Ld.w_f2) + (size_t)(c * 4 + i) * SLOT_BYTES, SLOT_BYTES);
        }
        for (int tl = c; tl < n_vt; tl += CS) rg.put(id, static_cast<const char*>(p.w_out) + (size_t)tl * TILE_BYTES, TILE_BYTES);
      }
    }
    return;
  }

  // ===================================================== consumer warps =====================================================
  cx.ts = ts; cx.ts_cap = ts_cap; cx.ts_n = 0;
  // per-video token state, identical in every CTA of the cluster: [0] token, [1] root, [2] attr of the current position,
  // then (int32 view, index 6..13) the tokens of positions t, t-1, ..., t-7 for the repeat constraint
  long long* tok = cx.at<long long>(cx.L.tok);
  if (tid < cx.R) {
    const size_t row = (size_t)(cx.row0 + tid) * p.cap;
    long long* ts_ = tok + tid * 8;
    ts_[0] = p.gen[row + t0];
    ts_[1] = p.gen_root[row + t0];
    ts_[2] = p.gen_attr[row + t0];
    int* hist = reinterpret_cast<int*>(ts_ + 3);
    for (int k = 0; k < 8; ++k) hist[k] = t0 - k >= 0 ? (int)p.gen[row + t0 - k] : -1;
  }
  stamp(cx);

  for (int step = 0; step < n_steps; ++step) {
    const int t = t0 + step;
    cx.t = t;
    cons_sync();                                          // tok[] of this step visible
    // ---- x_t = Linear_chord([emb | key]) + pe[t]   (video_music_transformer.py:984-1001,1029)
    if (warp < cx.R) {
      const int b = cx.row0 + warp;
      float v[16];
      if (p.chord_embed) {
        const float* src = p.emb_chord + (size_t)tok[warp * 8] * E + lane * 16;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(src) + q4); v[4 * q4] = t4.x; v[4 * q4 + 1] = t4.y; v[4 * q4 + 2] = t4.z; v[4 * q4 + 3] = t4.w; }
      } else {
        // primer positions and (greedy mode) the never-updated PAD entries come from the token buffers; sampled positions
        // carry the root / attribute of the token chosen in the previous step (video_music_transformer.py:1105-1123)
        const bool own = p.sample && t >= p.primer_len;
        const long long ir = own ? tok[warp * 8 + 1] : p.gen_root[(size_t)b * p.cap + t];
        const long long ia = own ? tok[warp * 8 + 2] : p.gen_attr[(size_t)b * p.cap + t];
        const float* sa = p.emb_root + (size_t)ir * E + lane * 16;
        const float* sb = p.emb_attr + (size_t)ia * E + lane * 16;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(sa) + q4), y4 = __ldg(reinterpret_cast<const float4*>(sb) + q4);
          v[4 * q4] = x4.x + y4.x; v[4 * q4 + 1] = x4.y + y4.y; v[4 * q4 + 2] = x4.z + y4.z; v[4 * q4 + 3] = x4.w + y4.w;
        }
      }
      uint4 a, bb;
      a.x = f2_to_bf16x2(v[0], v[1]);   a.y = f2_to_bf16x2(v[2], v[3]);   a.z = f2_to_bf16x2(v[4], v[5]);   a.w = f2_to_bf16x2(v[6], v[7]);
      bb.x = f2_to_bf16x2(v[8], v[9]);  bb.y = f2_to_bf16x2(v[10], v[11]); bb.z = f2_to_bf16x2(v[12], v[13]); bb.w = f2_to_bf16x2(v[14], v[15]);
      uint4* dst = reinterpret_cast<uint4*>(cx.at<bf16>(cx.L.xn) + (size_t)warp * XP + lane * 16);
      dst[0] = a;
      dst[1] = bb;
    }
    const float* resid = cx.at<float>(cx.L.resid);
    gemm_phase(cx, 4, 512, false, cx.L.xn, XP, 10, [&](int tl) { return p.b_chord + c * DH + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1), n = c * DH + nl;
        if (vid < cx.R)
          reinterpret_cast<float*>(cx.stage_buf())[vid * YP + nl] = cc[i] + __ldg(p.key + cx.row0 + vid) * __ldg(p.wc_key + n) + __ldg(p.pe + (size_t)t * E + n);
      }
    });
    gather(cx, 1, cx.L.yg, (uint32_t)cx.R * YP * 4);
    LnP lp1, lp2;
    rows_norm(cx, 0, lp1, lp2);
    stamp(cx);

    for (int l = 0; l < NL; ++l) {
      const DecLayer& Ld = p.layer[l];
      const float* bl = cx.at<float>(cx.L.bias) + l * 576;
      // ---- q, k, v of head c for the R rows  (rpr.py:253,328)
      float* qsm = cx.at<float>(cx.L.qs);
      bf16* knew = cx.at<bf16>(cx.L.knew);
      bf16* vnew = cx.at<bf16>(cx.L.vnew);
      gemm_phase(cx, 8, 512, false, cx.L.xn, XP, 10, [&](int tl) { return bl + tl * 16; },
                 [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = (tl & 3) * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) {
            if (tl < 4) qsm[vid * DH + nl] = cc[i] * 0.125f;
            else knew[vid * DH + nl] = __float2bfloat16_rn(cc[i]);
          }
        }
      });
      gemm_phase(cx, 4, 512, false, cx.L.xn, XP, 10, [&](int tl) { return bl + 128 + tl * 16; },
                 [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) vnew[vid * DH + nl] = __float2bfloat16_rn(cc[i]);
        }
      });
      cons_sync();
      if (tid < cx.R * 16) {                              // row t of the cache (read by the TMA engine in later steps)
        const int r = tid >> 4, part = (tid >> 3) & 1, v8 = tid & 7;
        const uint4 val = *reinterpret_cast<const uint4*>((part ? vnew : knew) + r * DH + v8 * 8);
        bf16* dst = static_cast<bf16*>(Ld.self_k) + ((((size_t)(cx.row0 + r)) * hb + c) * p.cap + t) * (2 * DH) + part * DH + ((v8 ^ (t & 7)) << 3);
        *reinterpret_cast<uint4*>(dst) = val;
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      stamp(cx);
      attention_phase(cx, p, l, true);
      stamp(cx);
      // ---- out-proj + residual -> LayerNorm  (rpr.py:417,58-59)
      gemm_phase(cx, 4, 512, false, cx.L.ctxg, CP, 6, [&](int tl) { return bl + 192 + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) reinterpret_cast<float*>(cx.stage_buf())[vid * YP + nl] = cc[i] + resid[vid * DH + nl];
        }
      });
      if (warp < cx.R) ln_load(lp1, Ld.ln1_g, Ld.ln1_b);
      gather(cx, 1, cx.L.yg, (uint32_t)cx.R * YP * 4);
      rows_norm(cx, 1, lp1, lp2);
      stamp(cx);
      // ---- cross-attention over the video memory  (rpr.py:62-66)
      gemm_phase(cx, 4, 512, false, cx.L.xn, XP, 10, [&](int tl) { return bl + 256 + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) qsm[vid * DH + nl] = cc[i] * 0.125f;
        }
      });
      stamp(cx);
      attention_phase(cx, p, l, false);
      stamp(cx);
      gemm_phase(cx, 4, 512, false, cx.L.ctxg, CP, 6, [&](int tl) { return bl + 320 + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) reinterpret_cast<float*>(cx.stage_buf())[vid * YP + nl] = cc[i] + resid[vid * DH + nl];
        }
      });
      if (warp < cx.R) ln_load(lp1, Ld.ln2_g, Ld.ln2_b);
      gather(cx, 1, cx.L.yg, (uint32_t)cx.R * YP * 4);
      rows_norm(cx, 1, lp1, lp2);
      stamp(cx);
      // ---- feed-forward  (rpr.py:67-69): hidden features [128c, 128c+128), then output features [64c, 64c+64)
      gemm_phase(cx, 8, 512, false, cx.L.xn, XP, 10, [&](int tl) { return bl + 384 + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) reinterpret_cast<bf16*>(cx.stage_buf())[vid * HSP + nl] = __float2bfloat16_rn(fmaxf(cc[i], 0.f));
        }
      });
      gather(cx, 2, cx.L.hg, (uint32_t)cx.R * HSP * 2);
      stamp(cx);
      gemm_phase(cx, 4, 1024, false, cx.L.hg, HSP, 7, [&](int tl) { return bl + 512 + tl * 16; }, [&](int tl, const float* cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = tl * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
          if (vid < cx.R) reinterpret_cast<float*>(cx.stage_buf())[vid * YP + nl] = cc[i] + resid[vid * DH + nl];
        }
      });
      if (warp < cx.R) {
        ln_load(lp1, Ld.ln3_g, Ld.ln3_b);
        if (l + 1 == NL) ln_load(lp2, p.lnf_g, p.lnf_b);                     // + decoder final norm (rpr.py:32-33)
      }
      gather(cx, 1, cx.L.yg, (uint32_t)cx.R * YP * 4);
      rows_norm(cx, l + 1 < NL ? 1 : 2, lp1, lp2);
      stamp(cx);
    }
    // ---- logits: vocabulary tiles c, c + 8  (video_music_transformer.py:1042)
    const int my_vt = c < n_vt ? (n_vt - c + CS - 1) / CS : 0;
    gemm_phase(cx, my_vt, 512, true, cx.L.xn, XP, 10, [&](int tl) { return p.b_out + (c + tl * CS) * 16; }, [&](int tl, const float* cc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = (c + tl * CS) * 16 + g + (i >> 1) * 8, vid = 2 * q + (i & 1);
        if (vid < cx.R && n < p.vocab) {
          const float v = cc[i];
          const int b = cx.row0 + vid;
          p.logits[(size_t)b * p.vocab + n] = v;
          if (p.logits_all) p.logits_all[((size_t)b * p.cap + t) * p.vocab + n] = v;
        }
      }
    });
    __threadfence();                                      // logits of this CTA visible cluster-wide before the barrier
    cluster_sync_cons(cx);
    // ---- next token (decode_pick.cuh), computed redundantly and identically in every CTA; CTA 0 records it
    if (warp < cx.R) {
      const int b = cx.row0 + warp;
      long long* ts_ = tok + warp * 8;
      int* hist = reinterpret_cast<int*>(ts_ + 3);
      const float u = p.sample ? __ldg(p.uniforms + (size_t)b * p.cap + min(t + 1, p.cap - 1)) : 0.f;
      int nt = pick_token(p.logits + (size_t)b * p.vocab, p.vocab, p.vocab_limit, p.sample, p.max_conseq_N, p.max_conseq_chord, t, u,
                          [&](int k) { return hist[k]; }, lane);
      __syncwarp();
      if (lane == 0 && t + 1 < p.cap) {
        long long root, attr;
        if (t + 1 >= p.primer_len) {
          chord_root_attr(nt, root, attr);
          if (c == 0) {
            p.gen[(size_t)b * p.cap + t + 1] = nt;
            if (p.sample) { p.gen_root[(size_t)b * p.cap + t + 1] = root; p.gen_attr[(size_t)b * p.cap + t + 1] = attr; }
          }
        } else {
          nt = (int)p.gen[(size_t)b * p.cap + t + 1];
          root = p.gen_root[(size_t)b * p.cap + t + 1];
          attr = p.gen_attr[(size_t)b * p.cap + t + 1];
        }
        ts_[0] = nt; ts_[1] = root; ts_[2] = attr;
        for (int k = 7; k > 0; --k) hist[k] = hist[k - 1];
        hist[0] = nt;
      }
    }
    stamp(cx);
  }
  // the last remote write into this CTA's shared memory (and the last remote barrier arrive) precede the final logits
  // barrier of every CTA: exiting here is safe
  if (blockIdx.x == 0 && tid == 0) *p.step = t0 + n_steps;
}

}  // namespace ds

// Runs n_steps positions starting at t0 with the streamed cluster kernel.  `p` must be the stream-path parameter block
// (engine.build_decode): fragment-packed matrices, interleaved + swizzled K|V caches (self_k / cross_k point at them,
// self_v / cross_v are unused) and the 8 swizzled copies of Er in er_sw.  kUnsupported when the configuration is not covered.
int decode_run_stream(const DecodeParams& p, int t0, int n_steps, unsigned long long* ts, int ts_cap, cudaStream_t stream) {
  if (p.dtype != 1 || p.E != ds::E || p.H != ds::CS || p.FF != ds::FF || p.vocab > 256 || p.cap > p.er_len || p.B < 1)
    return kUnsupported;
  for (int l = 0; l < p.n_layers; ++l)
    if (p.layer[l].er_sw == nullptr) return kUnsupported;
  if (p.sample && (p.max_conseq_chord < 1 || p.max_conseq_chord > 8 || !p.uniforms)) return kUnsupported;
  if (n_steps <= 0) return kOk;
  static int max_clusters[9] = {0};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ds::decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ds::SMEM_MAX);
    if (e != cudaSuccess) { set_last_error("decode_run_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kCudaError; }
    attr_set = true;
  }
  // rows per cluster: the smallest R for which all clusters are co-resident (clusters are independent, so more clusters
  // than fit would still be correct, just run in waves)
  int R_sel = 0;
  for (int R = 1; R <= 8; ++R) {
    const ds::Layout L = ds::make_layout(R, p.cap, p.S);
    if (L.nslot < 4) break;
    if (max_clusters[R] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ds::CS, 1, 1);
      cfg.blockDim = dim3(ds::THREADS, 1, 1);
      cfg.dynamicSmemBytes = L.total;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = ds::CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, ds::decode_stream_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
      max_clusters[R] = ncl > 0 ? ncl : -1;
    }
    if (max_clusters[R] >= (p.B + R - 1) / R) { R_sel = R; break; }
  }
  if (const char* ov = getenv("V2M_STREAM_ROWS")) {       // measurement / test override
    const int R = atoi(ov);
    if (R >= 1 && R <= 8 && ds::make_layout(R, p.cap, p.S).nslot >= 4) R_sel = R;
  }
  if (R_sel == 0) return kUnsupported;
  const ds::Layout L = ds::make_layout(R_sel, p.cap, p.S);
  const int ncl = (p.B + R_sel - 1) / R_sel;
  if (getenv("V2M_VERBOSE"))
    fprintf(stderr, "v2m: decode stream kernel: %d clusters x 8 CTAs, %d videos per cluster, ring %d x 32 KB, %u B smem (max co-resident %d)\n",
            ncl, R_sel, L.nslot, L.total, max_clusters[R_sel]);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl * ds::CS, 1, 1);
  cfg.blockDim = dim3(ds::THREADS, 1, 1);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = ds::CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, ds::decode_stream_kernel, p, t0, n_steps, R_sel, ts, ts_cap);
  if (e != cudaSuccess) {
    set_last_error("decode_run_stream: launch failed: %s", cudaGetErrorString(e));
    return kCudaError;
  }
  return check_launch("decode_stream_kernel");
}

}  // namespace v2m
