// Fused RPR attention forward on tcgen05 / TMEM / TMA (bf16 operands, fp32 accumulate, sm_100a).
//
//   O = softmax(Q K^T + skew(Q Er^T) + causal) V      one CTA = one (video, head) x 128 query rows
//
// Replaces model/rpr.py:387-414 of the reference (bmm, einsum("hld,md->hlm", q, Er), _skew :439-455,
// + mask, softmax, bmm) -- three materialised (B*H, L, L) fp32 tensors and ~10 ATen kernels -- and,
// with Er == null, the stock attention cores (encoder self-attention, decoder cross-attention,
// grouped-query attention with Hkv < Hq).
//
// Because L <= 320 the whole score row of a query fits in tensor memory, so the softmax is exact
// (two passes over TMEM, no online rescaling):
//   TMEM columns [0,320)    S = Q K^T (fp32), rewritten in place as the masked scores; O lands in columns [0,64) at the end
//   TMEM columns [320,512)  QE = Q Er_band^T, 192 columns at a time (two overlapping halves), later P (packed bf16 pairs)
// The relative term needs, for query row i and key j, column c = j - i + imax of QE (imax = last row
// of the tile): a per-row shift.  tcgen05.ld addresses are warp-uniform, so each softmax warp pulls a
// 64-column window of QE into registers, parks it in its private shared-memory scratch rows and
// reads it back at the row-dependent offset -- the skew is a shifted read of shared memory
// (conflict-free: row pitch 68 words).
//
// Persistent: one CTA per SM walks the (video, head, query tile) items with stride gridDim.x.  The shared-memory
// operands of item n+1 are fetched as soon as the tensor core has consumed those of item n (Q, K, Er band right after
// the S / QE products, V after P V), so TMA latency, TMEM allocation and CTA launch are off the per-tile critical path;
// the next S product only waits for the epilogue warps to have pulled O out of tensor memory (bar_epi).
//
// Warp roles (32 + 128 * AT_NW threads): warp 0 = TMA producer + single-thread MMA issuer + TMEM allocator,
// warps 1.. = softmax/epilogue.  A warp can only touch the TMEM lane quadrant (warp id % 4), so AT_NW warps share each
// quadrant (32 query rows) and split its 32-key chunks round-robin; the row max / row sum are exchanged through shared
// memory behind a named barrier per quadrant.  The softmax warps are the critical path of a tile (one warp per SM
// sub-partition issues at ~0.24 IPC), hence several of them per sub-partition.  Chunks that lie entirely below the
// causal diagonal skip the mask / skew predicates, chunks entirely above it are zero-filled without being read.
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cstdlib>

namespace v2m {

constexpr int AT_M = 128;            // query rows per CTA
constexpr int AT_DH = 64;            // head dim (128-byte swizzled rows)
constexpr int AT_MAXK = 320;         // keys per row that fit the TMEM plan
constexpr int AT_QE_COLS = 192;      // width of one QE half
constexpr int AT_QE_OVERLAP = 128;   // half B starts at band column 128
// softmax warps per TMEM lane quadrant: 2 with RPR (each needs a 8.7 KB skew scratch), 4 without (encoder / cross attention)
constexpr int at_threads(int nw) { return 32 + 128 * nw; }
constexpr int AT_SCR_PITCH = 68;     // floats; 68 % 32 == 4 -> conflict-free STS.128 and shifted LDS.32

constexpr int AT_SMEM_Q = AT_M * 128;
constexpr int AT_SMEM_K = AT_MAXK * 128;
constexpr int at_smem_scr(int nw, bool er) { return er ? nw * 4 * 32 * AT_SCR_PITCH * 4 : 0; }
constexpr int at_smem_red(int nw) { return 2 * nw * AT_M * 4; }   // row max / row sum exchange
constexpr size_t at_smem(int nw, bool er) {
  return 1024 + AT_SMEM_Q + (er ? 3 : 2) * AT_SMEM_K + at_smem_scr(nw, er) + at_smem_red(nw) + 256;
}

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long batch,
                      long long row_pitch, long long batch_pitch, int box_rows, int swap);
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);

struct AttnTcArgs {
  void* o; long long o_sb, o_sl;
  float* lse;
  int B, Hq, Hkv, Lq, Lk, causal, has_er, er_len;
  int swap;   // tensor maps are (col, batch, row) instead of (col, row, batch): sequence-first layouts
  float drop_scale; unsigned int drop_thresh, drop_seed;   // probability dropout (training), 0 = off
  const unsigned int* drop_seed_dev;
  int n_q_tiles, n_items;   // work items = (video, head) x 128-row query tiles; CTAs walk them with stride gridDim.x
  int first_rows;           // query rows of tile 0 (the partial tile comes FIRST: see attn_fwd_bf16_tc)
};

template <int AT_NW, bool HAS_ER, bool DROP>
__global__ void __launch_bounds__(at_threads(AT_NW), 1)
attn_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                    const __grid_constant__ AttnTcArgs a) {
  extern __shared__ unsigned char at_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + AT_SMEM_Q;
  unsigned char* sV = sK + AT_SMEM_K;
  unsigned char* sE = sV + AT_SMEM_K;
  float* scr = reinterpret_cast<float*>(sE + (HAS_ER ? AT_SMEM_K : 0));   // no Er buffer / skew scratch without RPR
  float* red = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scr) + at_smem_scr(AT_NW, HAS_ER));   // [2][AT_NW][AT_M]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(red) + at_smem_red(AT_NW));
  uint64_t* bar_qk = bars + 0;     // Q and K tiles landed
  uint64_t* bar_e = bars + 1;      // Er band landed
  uint64_t* bar_v = bars + 2;      // V tile landed
  uint64_t* bar_s = bars + 3;      // S = Q K^T complete
  uint64_t* bar_qa = bars + 4;     // QE half A complete
  uint64_t* bar_qa_free = bars + 5;  // all softmax warps are done with half A (count 4 * AT_NW)
  uint64_t* bar_qb = bars + 6;     // QE half B complete
  uint64_t* bar_p = bars + 7;      // P written to TMEM by all softmax warps (count 4 * AT_NW)
  uint64_t* bar_o = bars + 8;      // O = P V complete
  uint64_t* bar_epi = bars + 9;    // O pulled out of TMEM by all softmax warps (count 4 * AT_NW): next item may overwrite S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  auto tma3 = [&](void* dst, const CUtensorMap* tm, int col, int row, int batch, uint64_t* bar) {
    if (a.swap) tma_load_3d(dst, tm, col, batch, row, bar);
    else tma_load_3d(dst, tm, col, row, batch, bar);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int coff = a.Lk - a.Lq;
  struct Item { int b, hq, hkv, bh, i0, imax, nk, nk16, nk64, e_base, qe_rows; bool need_b; };
  auto make_item = [&](int item) {
    Item t;
    t.bh = item / a.n_q_tiles;
    t.b = t.bh / a.Hq; t.hq = t.bh % a.Hq; t.hkv = t.hq / (a.Hq / a.Hkv);
    const int tq = item - t.bh * a.n_q_tiles;
    t.i0 = tq == 0 ? 0 : a.first_rows + (tq - 1) * AT_M;
    t.imax = tq == 0 ? a.first_rows - 1 : min(t.i0 + AT_M - 1, a.Lq - 1);
    t.nk = a.causal ? min(a.Lk, t.imax + coff + 1) : a.Lk;            // keys this tile needs
    t.nk16 = (t.nk + 15) & ~15;
    t.nk64 = (t.nk + 63) & ~63;
    t.e_base = a.er_len - 1 - t.imax;                                  // band column c <-> Er row e_base + c, needed c in [0, imax]
    t.need_b = HAS_ER && (t.imax >= AT_QE_COLS);                     // widest window start is clamped to 128 in half A
    t.qe_rows = t.need_b ? AT_QE_OVERLAP + AT_QE_COLS : AT_QE_COLS;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(bars + i, (i == 5 || i == 7 || i == 9) ? 4u * AT_NW : 1u);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    if (HAS_ER) tma_prefetch_desc(&tmE);
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  grid_dep_launch();
  grid_dep_wait();                                             // the prologue above overlapped the previous kernel's tail
  // P goes where QE was (free once the quadrant's warps have left pass A) and O over the first S columns (all read by the
  // time P is complete): with several warps per quadrant P chunk c would otherwise overwrite S chunk c/2 of another warp.
  const uint32_t T_S = tmem, T_QE = tmem + AT_MAXK, T_O = tmem, T_P = tmem + AT_MAXK;

  if (warp == 0) {
    // The whole warp walks the items in convergent code; one elected lane issues the TMA loads and the MMAs.  (Inside an
    // `if (lane == 0)` region with C++ spin loops the compiler cannot keep the tcgen05 operands in uniform registers and wraps
    // every MMA in an ELECT / R2UR x 7 / branch waterfall, ~25 instructions each: the ~35 MMAs of an item then cost the issuing
    // thread more time than the tensor pipe needs, on the critical path of the softmax warps -- see umma_kstep4 in common.cuh.)
    {
      auto load_qke = [&](const Item& t) {                             // Q + K, Er band (64-row boxes)
        mbar_arrive_expect_tx(bar_qk, AT_SMEM_Q + t.nk64 * 128);
        tma3(sQ, &tmQ, t.hq * AT_DH, t.i0, t.b, bar_qk);
        tma3(sQ + 64 * 128, &tmQ, t.hq * AT_DH, t.i0 + 64, t.b, bar_qk);
        for (int r = 0; r < t.nk64; r += 64) tma3(sK + r * 128, &tmK, t.hkv * AT_DH, r, t.b, bar_qk);
        if (HAS_ER) {
          mbar_arrive_expect_tx(bar_e, t.qe_rows * 128);
          for (int r = 0; r < t.qe_rows; r += 64) tma_load_2d(sE + r * 128, &tmE, 0, t.e_base + r, bar_e);
        }
      };
      auto load_v = [&](const Item& t) {
        mbar_arrive_expect_tx(bar_v, t.nk64 * 128);
        for (int r = 0; r < t.nk64; r += 64) tma3(sV + r * 128, &tmV, t.hkv * AT_DH, r, t.b, bar_v);
      };
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), e_addr = smem_u32(sE);
      uint32_t par = 0, parb = 0;                                      // phase of the once-per-item barriers / of the half-B pair
      if ((int)blockIdx.x < a.n_items) {
        const Item first = make_item(blockIdx.x);
        if (elect_one()) {
          load_qke(first);
          load_v(first);
        }
        __syncwarp();
      }
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const Item t = make_item(item);
        const bool has_next = item + (int)gridDim.x < a.n_items;
        // ---------------- MMA issue ----------------
        mbar_wait_u(bar_qk, par);
        tc_fence_after();
        if (elect_one()) {
          for (int n0 = 0; n0 < t.nk16; n0 += 256) {                   // S[:, n0:n0+n] = Q K[n0:n0+n]^T
            const int n = min(256, t.nk16 - n0);
            const uint32_t idesc = make_idesc_bf16(AT_M, n, 0, 0);
#pragma unroll
            for (int k = 0; k < AT_DH / 16; ++k)
              umma_bf16_ss(T_S + n0, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           make_smem_desc_sw128(k_addr + n0 * 128 + k * 32, 16, 1024), idesc, k != 0);
          }
          umma_commit(bar_s);
        }
        __syncwarp();
        if (HAS_ER) {
          const uint32_t idesc = make_idesc_bf16(AT_M, AT_QE_COLS, 0, 0);
          mbar_wait_u(bar_e, par);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_DH / 16; ++k)
              umma_bf16_ss(T_QE, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           make_smem_desc_sw128(e_addr + k * 32, 16, 1024), idesc, k != 0);
            umma_commit(bar_qa);
          }
          __syncwarp();
          if (t.need_b) {
            mbar_wait_u(bar_qa_free, parb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < AT_DH / 16; ++k)
                umma_bf16_ss(T_QE, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                             make_smem_desc_sw128(e_addr + AT_QE_OVERLAP * 128 + k * 32, 16, 1024), idesc, k != 0);
              umma_commit(bar_qb);
            }
            __syncwarp();
          }
        }
        // the products above are the last readers of sQ / sK / sE: once they have retired, fetch the next item's operands
        if (has_next) {
          if (t.need_b) mbar_wait_u(bar_qb, parb);
          else if (HAS_ER) mbar_wait_u(bar_qa, par);
          else mbar_wait_u(bar_s, par);
          if (elect_one()) load_qke(make_item(item + gridDim.x));
          __syncwarp();
        }
        // O = P V : A = P from TMEM (bf16 pairs, 8 columns per k16), B = V (d contiguous -> MN-major)
        mbar_wait_u(bar_v, par);
        mbar_wait_u(bar_p, par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc_bf16(AT_M, AT_DH, 0, 1);
          for (int k = 0; k < t.nk16 / 16; ++k)
            umma_bf16_ts(T_O, T_P + k * 8, make_smem_desc_sw128(v_addr + k * 2048, 1024, 1024), idesc, k != 0);
          umma_commit(bar_o);
        }
        __syncwarp();
        if (has_next) {
          mbar_wait_u(bar_o, par);                                     // P V retired: sV is free
          if (elect_one()) load_v(make_item(item + gridDim.x));
          __syncwarp();
          mbar_wait_u(bar_epi, par);                                   // O is out of TMEM: the next S may overwrite it
          tc_fence_after();
        }
        par ^= 1;
        if (t.need_b) parb ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ================= softmax / epilogue warps =================
    const int quad = warp & 3;                                   // TMEM lane quadrant of this warp
    const int sub = (warp - 1) >> 2;                             // which of the AT_NW warps of the quadrant
    const int u = lane;
    const int row = quad * 32 + u;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t my_scr = smem_u32(scr + (size_t)((sub * 4 + quad) * 32 + u) * AT_SCR_PITCH);   // this lane's private row
    const uint32_t red_a = smem_u32(red);
    const float LOG2E = 1.4426950408889634f;
    const uint32_t bar_id = 1 + quad;                            // named barrier of this quadrant's AT_NW warps
    uint32_t par = 0, parb = 0;
    const uint32_t dseed = DROP ? a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0u) : 0u;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const Item t = make_item(item);
    const int i0 = t.i0, imax = t.imax, nk16 = t.nk16, b = t.b, hq = t.hq, bh = t.bh;
    const bool need_b = t.need_b;
    const int i = i0 + row;                                      // my query row
    const bool row_ok = i <= imax;                                // (tile 0 may own fewer than 128 rows)
    const int jlim = row_ok ? (a.causal ? min(a.Lk, i + coff + 1) : a.Lk) : 0;   // keys [0, jlim) are visible
    const int wfirst = i0 + quad * 32;                           // first row of this quadrant
    const int wlast = min(imax, wfirst + 31);                    // last valid row of this quadrant
    const bool quad_ok = wfirst <= imax;
    const int wjlim = quad_ok ? (a.causal ? min(a.Lk, wlast + coff + 1) : a.Lk) : 0;     // widest row of the quadrant
    const int wjmin = (quad_ok && wfirst + 31 <= imax) ? (a.causal ? min(a.Lk, wfirst + coff + 1) : a.Lk) : 0;  // narrowest
    const int nchunks = (nk16 + 31) / 32;                        // chunks the P V product reads
    const int live = min(nchunks, (wjlim + 31) / 32);            // chunks with at least one visible key for this quadrant
    const int base_w = imax - i0 - quad * 32 - 31;               // window start (band column) for key chunk j0 is j0 + base_w

    mbar_wait(bar_s, par);
    tc_fence_after();
    float mx = -INFINITY;
    // ---- pass A: s = S + Srel, mask, running max, s written back to TMEM
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && HAS_ER) { mbar_wait(bar_qa, par); tc_fence_after(); }
      if (phase == 1) {
        if (!need_b) break;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_qa_free);
        mbar_wait(bar_qb, parb);
        tc_fence_after();
      }
      for (int c = sub; c < live; c += AT_NW) {
        const int j0 = c * 32;
        const int start = j0 + base_w;                           // un-clamped window start
        const bool rel_chunk = HAS_ER && (j0 <= wlast);        // chunk holds some (i, j<=i) of this quadrant
        const bool in_a = !need_b || !rel_chunk || start <= AT_QE_OVERLAP;
        if ((phase == 0) != in_a) continue;
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_off + j0, r);
        if (rel_chunk) {
          const int cs = in_a ? max(0, min(start, AT_QE_COLS - 64)) : max(AT_QE_OVERLAP, min(start, AT_QE_OVERLAP + AT_QE_COLS - 64));
          const int tcol = in_a ? cs : cs - AT_QE_OVERLAP;      // column inside the resident half
          uint32_t w[64];
          tmem_ld_32x32(T_QE + lane_off + tcol, w);
          tmem_ld_32x32(T_QE + lane_off + tcol + 32, w + 32);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 16; ++q4)
            sts_v4(my_scr + q4 * 16, w[q4 * 4], w[q4 * 4 + 1], w[q4 * 4 + 2], w[q4 * 4 + 3]);
          // band column of (i, j) is j - i + imax; offset inside the window = that - cs
          const int off = j0 - i + imax - cs;                    // in [0, 63 - k] for every (i, j <= i) that is needed
          if (j0 + 31 <= wfirst) {                               // whole chunk at or below the diagonal for every row
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) + lds_f32(my_scr + (off + k) * 4));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float rel = lds_f32(my_scr + min(max(off + k, 0), 63) * 4);  // clamped reads only hit masked / invalid entries
              if (j0 + k <= i) r[k] = __float_as_uint(__uint_as_float(r[k]) + rel);
            }
          }
        } else {
          tmem_ld_wait();
        }
        bool modified = rel_chunk;                               // scores that changed go back to TMEM for pass B
        float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};     // four independent chains instead of 32 dependent FMNMX
        if (j0 + 32 <= wjmin) {                                  // every key of the chunk is visible to every row
#pragma unroll
          for (int k = 0; k < 32; ++k) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
        } else {
          modified = true;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            float sv = __uint_as_float(r[k]);
            sv = (j0 + k < jlim) ? sv : -INFINITY;
            m4[k & 3] = fmaxf(m4[k & 3], sv);
            r[k] = __float_as_uint(sv);
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (modified) tmem_st_32x32(T_S + lane_off + j0, r);
      }
    }
    tmem_st_wait();
    if (AT_NW > 1) {                                             // row max over the warps of the quadrant; also orders
      sts_f32(red_a + (sub * AT_M + row) * 4, mx);               // every QE read of the quadrant before the P writes
      named_bar_sync(bar_id, 32 * AT_NW);
#pragma unroll
      for (int w2 = 0; w2 < AT_NW; ++w2) mx = fmaxf(mx, lds_f32(red_a + (w2 * AT_M + row) * 4));
    }
    // ---- pass B: p = exp(s - max), row sum, P (bf16 pairs) written over S
    const float mneg = (mx == -INFINITY) ? 0.f : mx * LOG2E;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};                          // independent partial row sums (short FADD chains)
    for (int c = sub; c < nchunks; c += AT_NW) {
      const int j0 = c * 32;
      uint32_t pk[16];
      if (c < live) {
        uint32_t r[32];
        tmem_ld_32x32(T_S + lane_off + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * k]), LOG2E, -mneg));
          float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * k + 1]), LOG2E, -mneg));
          s4[k & 3] += p0 + p1;
          if (DROP) {                                          // dropout acts on the normalised probabilities: O stays / sum
            const uint32_t rr = (uint32_t)bh * (uint32_t)a.Lq + (uint32_t)i;
            const uint32_t hh = drop_hash4(dseed, rr, (j0 >> 2) + (k >> 1));   // columns j0 + 2k, +1 (j0 % 32 == 0); CSE'd per two k
            p0 = drop_keep_byte(hh, 2 * k, a.drop_thresh) ? p0 * a.drop_scale : 0.f;
            p1 = drop_keep_byte(hh, 2 * k + 1, a.drop_thresh) ? p1 * a.drop_scale : 0.f;
          }
          pk[k] = f2_to_bf16x2(p0, p1);
        }
      } else {                                                   // above the diagonal for the whole quadrant: P = 0
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = 0u;
      }
      tmem_st_32x16(T_P + lane_off + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    if (AT_NW > 1) {                                             // row sum over the warps of the quadrant
      sts_f32(red_a + ((AT_NW + sub) * AT_M + row) * 4, sum);
      named_bar_sync(bar_id, 32 * AT_NW);
      sum = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < AT_NW; ++w2) sum += lds_f32(red_a + ((AT_NW + w2) * AT_M + row) * 4);
    }
    // ---- epilogue: O / sum -> bf16 -> global; the warps of a quadrant split the 64 columns
    constexpr int OC = AT_DH / AT_NW;
    mbar_wait(bar_o, par);
    tc_fence_after();
    uint32_t o[OC];
#pragma unroll
    for (int c32 = 0; c32 < OC; c32 += 32) {
      if (OC - c32 >= 32) tmem_ld_32x32(T_O + lane_off + sub * OC + c32, o + c32);
      else tmem_ld_32x16(T_O + lane_off + sub * OC + c32, o + c32);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_epi);                         // the stores below overlap the next item's products
    if (row_ok) {
      const float inv = 1.f / sum;
      bf16* dst = static_cast<bf16*>(a.o) + (size_t)b * a.o_sb + (size_t)i * a.o_sl + (size_t)hq * AT_DH + sub * OC;
#pragma unroll
      for (int g8 = 0; g8 < OC / 8; ++g8) {
        uint4 v;
        v.x = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 0]) * inv, __uint_as_float(o[g8 * 8 + 1]) * inv);
        v.y = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 2]) * inv, __uint_as_float(o[g8 * 8 + 3]) * inv);
        v.z = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 4]) * inv, __uint_as_float(o[g8 * 8 + 5]) * inv);
        v.w = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 6]) * inv, __uint_as_float(o[g8 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + g8 * 8) = v;
      }
      if (a.lse && sub == 0) a.lse[(size_t)bh * a.Lq + i] = mx + logf(sum);
    }
    par ^= 1;
    if (need_b) parb ^= 1;
    }  // items
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int attn_fwd_bf16_tc(const AttnParams& p, cudaStream_t stream) {
  V2M_REQUIRE(p.dh == AT_DH, "attn_fwd_bf16_tc: head_dim %d unsupported (64)", p.dh);
  V2M_REQUIRE(p.B > 0 && p.Hq > 0 && p.Hkv > 0 && p.Hq % p.Hkv == 0, "attn_fwd_bf16_tc: bad heads %d/%d", p.Hq, p.Hkv);
  V2M_REQUIRE(p.Lq > 0 && p.Lk > 0 && p.Lk <= AT_MAXK, "attn_fwd_bf16_tc: Lk=%d unsupported (1..%d)", p.Lk, AT_MAXK);
  V2M_REQUIRE(!p.Er || (p.Lq == p.Lk && p.Lq <= p.er_len), "attn_fwd_bf16_tc: RPR needs Lq == Lk <= er_len (%d, %d, %d)",
              p.Lq, p.Lk, p.er_len);
  V2M_REQUIRE(!p.causal || p.Lk >= p.Lq, "attn_fwd_bf16_tc: causal needs Lk >= Lq");
  V2M_REQUIRE(p.q_scale == 1.0f, "attn_fwd_bf16_tc: q must be pre-scaled (fold the scale into the projection epilogue)");
  V2M_REQUIRE(p.p_out == nullptr, "attn_fwd_bf16_tc: need_weights output is only available on the fp32 path");
  V2M_REQUIRE(p.lk_dev == nullptr, "attn_fwd_bf16_tc: a device-side key count is only available on the fp32 path");
  V2M_REQUIRE(p.o_sl % 8 == 0 && p.o_sb % 8 == 0 && reinterpret_cast<uintptr_t>(p.o) % 16 == 0,
              "attn_fwd_bf16_tc: output must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV, tmE;
  int rc;
  // TMA wants ascending strides: sequence-first tensors (row pitch > batch pitch) get (col, batch, row) maps
  const int swap = p.B > 1 && p.q_sl > p.q_sb;
  V2M_REQUIRE(p.B == 1 || ((p.k_sl > p.k_sb) == (swap != 0) && (p.v_sl > p.v_sb) == (swap != 0)),
              "attn_fwd_bf16_tc: q, k, v must share one layout family (batch-first or sequence-first)");
  if ((rc = make_tmap_3d_bf16(&tmQ, p.q, (long long)p.Hq * AT_DH, p.Lq, p.B, p.q_sl, p.q_sb, 64, swap))) return rc;
  if ((rc = make_tmap_3d_bf16(&tmK, p.k, (long long)p.Hkv * AT_DH, p.Lk, p.B, p.k_sl, p.k_sb, 64, swap))) return rc;
  if ((rc = make_tmap_3d_bf16(&tmV, p.v, (long long)p.Hkv * AT_DH, p.Lk, p.B, p.v_sl, p.v_sb, 64, swap))) return rc;
  if (p.Er) {
    if ((rc = make_tmap_2d_bf16(&tmE, p.Er, p.er_len, AT_DH, AT_DH, 64))) return rc;
  } else {
    tmE = tmQ;
  }
  AttnTcArgs a;
  a.o = p.o; a.o_sb = p.o_sb; a.o_sl = p.o_sl; a.lse = p.lse;
  a.B = p.B; a.Hq = p.Hq; a.Hkv = p.Hkv; a.Lq = p.Lq; a.Lk = p.Lk; a.causal = p.causal;
  a.has_er = p.Er != nullptr; a.er_len = p.er_len; a.swap = swap;
  a.drop_scale = p.drop_scale; a.drop_thresh = p.drop_thresh; a.drop_seed = p.drop_seed; a.drop_seed_dev = p.drop_seed_dev;
  static bool attr = false;
  static int nw_plain = 4;                 // softmax warps per quadrant without RPR (V2M_ATTN_NW=2 for A/B measurements)
  if (!attr) {
    if (const char* e = getenv("V2M_ATTN_NW")) nw_plain = atoi(e) == 2 ? 2 : 4;
    cudaError_t e = cudaSuccess;
#define V2M_ATTR(NW, ER, DR)                                                                                                      \
    if (e == cudaSuccess)                                                                                                          \
      e = cudaFuncSetAttribute(attn_bf16_tc_kernel<NW, ER, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem(NW, ER))
    V2M_ATTR(2, true, false); V2M_ATTR(2, true, true); V2M_ATTR(4, false, false); V2M_ATTR(4, false, true);
    V2M_ATTR(2, false, false); V2M_ATTR(2, false, true);
#undef V2M_ATTR
    if (e != cudaSuccess) { set_last_error("attn_fwd_bf16_tc: smem attribute: %s", cudaGetErrorString(e)); return kCudaError; }
    attr = true;
  }
  a.n_q_tiles = (p.Lq + AT_M - 1) / AT_M;
  // Lq % 128 leftover rows: they form tile 0, not the last tile.  Under the causal mask a tile costs what its LONGEST rows
  // cost (per TMEM lane quadrant: a warp cannot help another quadrant), so at L = 299 the leftover 43 rows are cheap as rows
  // 0..42 (2 key chunks, one QE half, 48 S columns) and expensive as rows 256..298 (10 chunks, both QE halves, 304 columns):
  // critical-path chunks per (video, head) 2 + 6 + 10 = 18 instead of 4 + 8 + 10 = 22, and 20 % fewer MMA columns.
  static int shift_tiles = -1;
  if (shift_tiles < 0) { const char* e = getenv("V2M_ATTN_SHIFT"); shift_tiles = e ? atoi(e) != 0 : 1; }
  a.first_rows = shift_tiles ? p.Lq - (a.n_q_tiles - 1) * AT_M : (p.Lq < AT_M ? p.Lq : AT_M);
  const long long items = (long long)a.n_q_tiles * p.B * p.Hq;
  V2M_REQUIRE(items < (1ll << 31), "attn_fwd_bf16_tc: too many tiles");
  a.n_items = (int)items;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = a.n_items < num_sms ? a.n_items : num_sms;
  const bool drop = a.drop_scale != 0.f;
#define V2M_GO(NW, ER, DR) launch_dep(attn_bf16_tc_kernel<NW, ER, DR>, dim3(grid), dim3(at_threads(NW)), at_smem(NW, ER), stream, tmQ, tmK, tmV, tmE, a)
  if (a.has_er) { if (drop) V2M_GO(2, true, true); else V2M_GO(2, true, false); }
  else if (nw_plain == 4) { if (drop) V2M_GO(4, false, true); else V2M_GO(4, false, false); }
  else { if (drop) V2M_GO(2, false, true); else V2M_GO(2, false, false); }
#undef V2M_GO
  return check_launch("attn_fwd_bf16_tc");
}

}  // namespace v2m
