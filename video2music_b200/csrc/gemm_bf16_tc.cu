// bf16 tensor-core GEMM  C[M,N] = epi(A[M,K] * W[N,K]^T)  on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces every nn.Linear of the AMT on the bf16 path: QKV / out projections
// (model/rpr.py:253,277,417 and the stock encoder layers), FFN (rpr.py:67), Linear_vis /
// Linear_chord / Wout (model/video_music_transformer.py:1001,1022,1042) with the bias, q-scaling,
// ReLU, residual(+positional encoding) and KV-cache scatter fused into the epilogue.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0     TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a kStages ring
//   warp 1     MMA issuer     : one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//                               accumulators in TMEM, 2 accumulator stages (2*BN columns)
//   warp 2     TMEM allocator
//   warps 4-7  epilogue       : tcgen05.ld 32x32b -> registers -> fused epilogue -> global
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// Ragged edges: TMA zero-fills out-of-bounds rows/columns (M, N, K tails), stores are guarded.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <cuda.h>

namespace v2m {

constexpr int GM = 128, GK = 64;
constexpr int kGemmThreads = 384;   // 4 control warps + 8 epilogue warps

template <int BN> struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = GM * GK * 2;
  static constexpr int kBBytes = BN * GK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;      // two accumulator stages
  static constexpr int kBarBytes = 1024;        // barriers + TMEM slot (keeps the staging buffers 1024-byte aligned)
  static constexpr int kStoreBytes = 8 * 2 * 2048;   // per epilogue warp: two 32 x 32 bf16 boxes on their way out through TMA
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + kBarBytes + kStoreBytes;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = box_rows x 64 columns, 128B swizzle.
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not found"); return kCudaError; }
  V2M_REQUIRE((ld * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0,
              "TMA needs 16-byte aligned base and row pitch (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return kCudaError; }
  return kOk;
}

// Store-side map of a bf16 row-major [rows, cols] matrix: box = 32 rows x 32 columns (64-byte rows, 64B swizzle).
static int make_tmap_store_bf16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int wide = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not found"); return kCudaError; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {wide ? 64u : 32u, 32u};            // wide: 32 rows x 64 columns (128-byte rows, 128B swizzle)
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(store) failed (%d)", (int)r); return kCudaError; }
  return kOk;
}

// 3-D bf16 view (cols, rows, batch) of a strided matrix stack; box = 64 columns x box_rows rows x 1 batch,
// 128B swizzle.  swap != 0 orders the map (cols, batch, rows) for sequence-first tensors (ascending strides).
int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long batch,
                      long long row_pitch, long long batch_pitch, int box_rows, int swap) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not found"); return kCudaError; }
  if (batch == 1) batch_pitch = row_pitch * rows;      // any valid pitch will do
  V2M_REQUIRE((row_pitch * 2) % 16 == 0 && (batch_pitch * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(base) % 16 == 0,
              "TMA needs 16-byte aligned base and pitches (row %lld, batch %lld elements)", row_pitch, batch_pitch);
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1u, 1u, 1u};
  dims[0] = (cuuint64_t)cols; box[0] = 64u;
  if (!swap) {
    dims[1] = (cuuint64_t)rows;  strides[0] = (cuuint64_t)row_pitch * 2;   box[1] = (cuuint32_t)box_rows;
    dims[2] = (cuuint64_t)batch; strides[1] = (cuuint64_t)batch_pitch * 2; box[2] = 1u;
  } else {
    dims[1] = (cuuint64_t)batch; strides[0] = (cuuint64_t)batch_pitch * 2; box[1] = 1u;
    dims[2] = (cuuint64_t)rows;  strides[1] = (cuuint64_t)row_pitch * 2;   box[2] = (cuuint32_t)box_rows;
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(3d) failed (%d)", (int)r); return kCudaError; }
  return kOk;
}

// A_MN / B_MN: the operand is stored transposed ([K, M] resp. [K, N] row-major, "MN-major" for the tensor core):
// TMA fetches 64(k) x 64(m|n) boxes, the smem descriptor walks 8-k-row groups (SBO 1024 B) and 64-wide chunks (LBO 8 KB).
// This is what lets the backward pass compute dX = dY W and dW = dY^T X without materialising a transpose.
// PAIR: the CTA-pair variant (cta_group::2, plain GEMMs with many rounds of tiles only: no K split, groups or tail sub-tiles).
// A cluster of two CTAs owns a 256 x BN output tile: CTA r holds rows [128 r, 128 r + 128) of A and of the accumulator (its own
// TMEM) and rows [BN/2 r, BN/2 (r + 1)) of the B tile; the leader (rank 0) issues ONE M = 256 MMA per k-step that reads both
// CTAs' shared memory.  Per k-step a CTA ingests 16 + 16 KB instead of 16 + 32 KB: the 1-CTA kernel moves 48.7 B/clk/SM from
// L2 at K = 512 (ncu: l1tex__m_xbar2l1tex_read_bytes 2.82 GB in 240 us) and is bound by that, not by the tensor pipe (54 %).
// Barriers: both producers' TMA bytes are counted on the LEADER's full barrier (the leader expects the sum); tcgen05.commit
// multicasts to both CTAs' empty / tmem-full barriers; both CTAs' epilogue warps arrive on the leader's tmem-empty barrier.
template <int BN, bool A_MN, bool B_MN, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, void* __restrict__ C,
                    int ldc, int out_bf16, int vec_ok, int M, int N, int K, const __grid_constant__ GemmEpilogue ep, int k_split,
                    const int* __restrict__ tile_group, const __grid_constant__ CUtensorMap tmC, int tma_out,
                    const int* __restrict__ k_off, long long c_gstride, const __grid_constant__ CUtensorMap tmBs, int tail_first,
                    const __grid_constant__ CUtensorMap tmC64) {
  // tail_first >= 0 (plain GEMMs only: k_split == 1, no groups): the big tiles [tail_first, tiles) -- the ones that would make up
  // a mostly idle last round of the persistent grid (M = 19136 gives 150 row tiles on 148 SMs: 300 tiles = 2.03 rounds at
  // N = 512) -- are cut into BN / 64 sub-tiles of 128 x 64 each, so that round costs a quarter of a tile per SM instead of a
  // whole one.  Sub-tiles load 64 rows of B (tmBs: the same matrix with a 64-row box), multiply with N = 64 into the first
  // columns of the accumulator stage and drain two 32-column chunks.
  // k_off != null ("K-grouped", the ragged dW of the MoE experts: dW_e = dY_e^T X_e over the rows of group e): k_split is the
  // number of groups, work item (tile, g) multiplies over the K range [k_off[g], k_off[g+1]) (multiples of 64, read on the
  // device: group sizes never visit the host) and adds its tile to C + g * c_gstride (zero-initialised by the launcher);
  // empty groups are skipped by all three roles.
  // tma_out: plain bf16 row-major C with 16-byte aligned rows.  Each epilogue warp packs its 32 x 32 chunk into a private
  // swizzled shared-memory box and one lane hands it to the TMA engine: the per-thread row stores (32 rows x 16 B per
  // instruction) were 40 % of the time of a K = 512 GEMM.
  // tile_group != null: grouped (ragged) GEMM for the MoE experts.  Rows of A are expert-contiguous with every group
  // starting on a 128-row boundary; tile_group[m_blk] names the group of a row tile (-1: unused tile, skipped by all three
  // roles), whose weights are rows [g*N, (g+1)*N) of the stacked W and whose bias is ep.bias + g*N.
  // k_split > 1: a work item is (output tile, K slice); partial sums are added to the zero-initialised fp32 C with
  // red.global (small outputs with a long K, e.g. dW = dY^T X over all tokens, would otherwise keep a handful of SMs busy)
  using Cfg = GemmCfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // pair: a stage holds A (16 KB) + half of B; the shared memory this frees deepens the ring of outgoing TMA-store boxes
  constexpr int kStride = PAIR ? Cfg::kABytes + Cfg::kBBytes / 2 : Cfg::kStageBytes;
  constexpr int NBUF = 2;                                   // staging boxes per epilogue warp
  // pair + TMA-store epilogue: 32 rows x 64 columns per box (whole 128-byte lines, half as many TMA operations); a warp then
  // drains PAIRS of adjacent 32-column chunks instead of every other chunk
  constexpr int BOXB = PAIR ? 4096 : 2048;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::kStages * kStride);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + GM - 1) / GM, n_tiles = (N + BN - 1) / BN;
  const int num_kb = (K + GK - 1) / GK;
  constexpr int SUB = BN / 64;                             // sub-tiles per tail tile
  const int big_tiles = m_tiles * n_tiles;
  const uint32_t rank = PAIR ? pair_rank() : 0u;           // CTA of the pair
  const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, cta_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_tiles = PAIR ? ((m_tiles + 1) / 2) * n_tiles
                             : (tail_first >= 0 ? tail_first + (big_tiles - tail_first) * SUB : big_tiles * k_split);       // work items
  // work item -> (row tile, first column, width); k_split == 1 whenever tail_first >= 0
  auto decode = [&](int tile, int& m_blk, int& n_base, int& bn) {
    if (PAIR) {
      m_blk = 2 * (tile / n_tiles) + (int)rank; n_base = (tile % n_tiles) * BN; bn = BN;
    } else if (tail_first >= 0 && tile >= tail_first) {
      const int j = tile - tail_first, t = tail_first + j / SUB;
      m_blk = t / n_tiles; n_base = (t % n_tiles) * BN + (j % SUB) * 64; bn = 64;
    } else {
      m_blk = tile / n_tiles; n_base = (tile % n_tiles) * BN; bn = BN;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (tma_out) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, PAIR ? 16 : 8); }
    fence_barrier_init();
  }
  if (warp == 2) { if (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot); else tmem_alloc<Cfg::kTmemCols>(tmem_slot); }
  tc_fence_before();
  if (PAIR) pair_sync(); else __syncthreads();             // (pair: the peer's barriers are initialised before anybody signals them)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_launch();                                       // the next kernel of the chain may start its own prologue
  grid_dep_wait();                                         // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ================= TMA producer =================
    // (whole warp in convergent code, one elected lane issues: the TMA operands stay in uniform registers -- under
    //  `if (lane == 0)` each of the up to 6 loads of a k-step went through an ELECT / R2UR waterfall, see umma_kstep4)
    {
      int stage = 0; uint32_t phase = 0;
      for (int item = cta_id; item < num_tiles; item += cta_stride) {
        const int tile = item / k_split, ks = item - tile * k_split;
        int m_blk, n_base, bn;
        decode(tile, m_blk, n_base, bn);
        if (PAIR) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_u(empty_bar + stage, phase ^ 1);
            unsigned char* sa = smem + (size_t)stage * kStride;
            const uint32_t fb = pair_mapa(smem_u32(full_bar + stage), 0);       // the leader's barrier counts both CTAs' bytes
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * (Cfg::kABytes + (BN / 2) * GK * 2));
              tma_load_2d_pair(sa, &tmA, kb * GK, m_blk * GM, fb);
              if (B_MN) {
#pragma unroll
                for (int c = 0; c < BN / 128; ++c)
                  tma_load_2d_pair(sa + Cfg::kABytes + c * 8192, &tmB, n_base + (int)rank * (BN / 2) + c * 64, kb * GK, fb);
              } else {
                tma_load_2d_pair(sa + Cfg::kABytes, &tmBs, kb * GK, n_base + (int)rank * (BN / 2), fb);   // tmBs: BN/2-row box
              }
            }
            __syncwarp();
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        int kb0 = (int)((long long)ks * num_kb / k_split), kb1 = (int)((long long)(ks + 1) * num_kb / k_split);
        if (k_off) { kb0 = k_off[ks] / GK; kb1 = k_off[ks + 1] / GK; if (kb0 >= kb1) continue; }
        int g = 0;
        if (tile_group) { g = tile_group[m_blk]; if (g < 0) continue; }
        const int b_row = g * N + n_base;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_u(empty_bar + stage, phase ^ 1);
          unsigned char* sa = smem + (size_t)stage * kStride;
          if (elect_one()) {
          mbar_arrive_expect_tx(full_bar + stage, Cfg::kABytes + bn * GK * 2);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < GM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m_blk * GM + c * 64, kb * GK, full_bar + stage);
          } else {
            tma_load_2d(sa, &tmA, kb * GK, m_blk * GM, full_bar + stage);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              if (c * 64 < bn) tma_load_2d(sa + Cfg::kABytes + c * 8192, &tmB, n_base + c * 64, kb * GK, full_bar + stage);
          } else {
            tma_load_2d(sa + Cfg::kABytes, bn == BN ? &tmB : &tmBs, kb * GK, b_row, full_bar + stage);
          }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && PAIR) {
    // ================= MMA issuer of the pair: the leader only =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GM, BN, 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = cta_id; item < num_tiles; item += cta_stride) {
        mbar_wait_u(tempty_bar + acc, acc_phase ^ 1);            // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_u(full_bar + stage, phase);
          tc_fence_after();
          if (elect_one()) {                                     // one asm statement per k-step (see umma_kstep4)
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStride);
            const uint32_t b_addr = a_addr + Cfg::kABytes;
            const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
            umma_kstep4_pair(tmem_base + acc * BN, da, db, idesc, kb != 0 ? 1u : 0u, 2u, B_MN ? 128u : 2u, empty_bar + stage,
                             tfull_bar + acc, kb == num_kb - 1 ? 1u : 0u);   // commits reach both CTAs
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_full = make_idesc_bf16(GM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t idesc_sub = make_idesc_bf16(GM, 64, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
      const int ks = item % k_split;
      const uint32_t idesc = (tail_first >= 0 && item >= tail_first) ? idesc_sub : idesc_full;
      int kb0 = (int)((long long)ks * num_kb / k_split), kb1 = (int)((long long)(ks + 1) * num_kb / k_split);
      if (k_off) { kb0 = k_off[ks] / GK; kb1 = k_off[ks + 1] / GK; if (kb0 >= kb1) continue; }
      if (tile_group && tile_group[(item / k_split) / n_tiles] < 0) continue;
      mbar_wait_u(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_u(full_bar + stage, phase);
        tc_fence_after();
        if (elect_one()) {                                       // one asm statement per k-step (see umma_kstep4)
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStride);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
          const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr, 8192, 1024) : make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 16, 1024);
          // frees the smem slot when these MMAs retire; on the last k-step the accumulator goes to the epilogue
          umma_kstep4(tmem_base + acc * BN, da, db, idesc, kb != kb0 ? 1u : 0u, A_MN ? 128u : 2u, B_MN ? 128u : 2u, empty_bar + stage,
                      tfull_bar + acc, kb == kb1 - 1 ? 1u : 0u);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue: 8 warps, two per TMEM lane quadrant, alternating 32-column chunks ============
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;          // which chunks of the tile this warp drains
    unsigned char* stage = smem + (size_t)Cfg::kStages * kStride + Cfg::kBarBytes + (warp - 4) * (NBUF * BOXB);
    uint32_t n_store = 0;                      // boxes this warp has sent (selects the staging buffer)
    const uint32_t dseed = ep.drop_seed + (ep.drop_seed_dev ? *ep.drop_seed_dev : 0u);
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = cta_id; item < num_tiles; item += cta_stride) {
      const int tile = item / k_split;
      int m_blk, n_base, bn;
      decode(tile, m_blk, n_base, bn);
      const int m = m_blk * GM + quad * 32 + lane;
      size_t c_goff = 0;
      if (k_off) {
        const int ks = item - tile * k_split;
        if (k_off[ks] / GK >= k_off[ks + 1] / GK) continue;
        c_goff = (size_t)ks * (size_t)c_gstride;
      }
      const float* bias = ep.bias;
      if (tile_group) {
        const int g = tile_group[m_blk];
        if (g < 0) continue;
        if (bias) bias += (size_t)g * N;
      }
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const bool row_ok = m < M;
      const float rs = (row_ok && ep.row_scale) ? ep.row_scale[m] : 0.f;
      const size_t res_row = ep.residual ? (size_t)(ep.res_mod > 0 ? m % ep.res_mod : m) * ep.ldr : 0;
      // The global operands of the epilogue (bias row, bf16 residual rows) do not depend on the accumulators: their loads
      // are issued BEFORE waiting for the TMEM load, and the residual of the NEXT chunk is fetched while this one is
      // processed (an L2 round trip per chunk otherwise sits on the critical path of an epilogue-bound tile).
      const bool pre_res = vec_ok && row_ok && ep.residual && ep.residual_bf16 && k_split == 1;
      uint4 res_nx[4];
      auto fetch_res = [&](int cc, uint4* dst) {
        const int nn = n_base + cc * 32;
        if (pre_res && cc < bn / 32 && nn + 32 <= N) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.residual) + res_row + nn);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = __ldg(rp + q);
        }
      };
      const bool w64 = PAIR && tma_out;                       // (pair tiles: N % 256 == 0, every chunk takes the same path)
      auto chunk_of = [&](int t) { return w64 ? 2 * (half + 2 * (t >> 1)) + (t & 1) : half + 2 * t; };
      fetch_res(chunk_of(0), res_nx);
#pragma unroll 1
      for (int t = 0; chunk_of(t) < bn / 32; ++t) {
        const int c = chunk_of(t);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(quad * 32) << 16), r);
        const int n0 = n_base + c * 32;
        const bool fast = vec_ok && n0 + 32 <= N && k_split == 1 && !ep.accumulate && !k_off;
        float4 bv[8];
        if (fast && row_ok && bias) {
          const float4* bp = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q) bv[q] = __ldg(bp + q);
        }
        uint4 res_cur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) res_cur[q] = res_nx[q];
        fetch_res(chunk_of(t + 1), res_nx);
        tmem_ld_wait();
        const bool tma_chunk = tma_out && fast;                // warp-uniform
        const bool red_chunk = (k_split > 1 || ep.accumulate || k_off) && vec_ok && n0 + 32 <= N;   // warp-uniform: all lanes take part in the transpose
        if ((!row_ok && !tma_chunk && !red_chunk) || n0 >= N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (k_split > 1 || ep.accumulate || k_off) {          // partial sum of one K slice / accumulation into C (plain fp32, no epilogue ops)
          if (vec_ok && n0 + 32 <= N) {
            // Coalesced vector reductions: the 32 x 32 chunk (lane = row) is transposed through this warp's staging buffer so
            // that one red.global.add.v4.f32 instruction covers four whole 128-byte row segments (8 lanes x 16 B each) instead
            // of 32 scalar atomics scattered over 32 rows: 8 instructions per chunk instead of 32, every one fully coalesced.
            const uint32_t sbuf = smem_u32(stage);               // explicit shared-space accesses (see common.cuh: sts_f32)
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; ++i) sts_f32(sbuf + (lane * 32 + (i ^ lane)) * 4, v[i]);      // XOR swizzle: conflict-free both ways
            __syncwarp();
            const int c0 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
            for (int r4 = 0; r4 < 8; ++r4) {
              const int row = r4 * 4 + rsub;
              const int mm = m_blk * GM + quad * 32 + row;
              const float a0 = lds_f32(sbuf + (row * 32 + ((c0 + 0) ^ row)) * 4), a1 = lds_f32(sbuf + (row * 32 + ((c0 + 1) ^ row)) * 4);
              const float a2 = lds_f32(sbuf + (row * 32 + ((c0 + 2) ^ row)) * 4), a3 = lds_f32(sbuf + (row * 32 + ((c0 + 3) ^ row)) * 4);
              if (mm < M) {
                float* dst = static_cast<float*>(C) + c_goff + (size_t)mm * ldc + n0 + c0;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
              }
            }
            __syncwarp();
          } else if (row_ok) {
            float* dst = static_cast<float*>(C) + c_goff + (size_t)m * ldc + n0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < N) atomicAdd(dst + i, v[i]);
          }
          continue;
        }
        if (vec_ok && n0 + 32 <= N) {
          // ---- fast path: whole chunk in range, every pointer 16-byte aligned -> 128-bit loads / stores only
          if (row_ok) {                                          // rows past M only exist on the TMA-store path (clipped there)
          if (bias) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = bv[q];
              v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
            }
          }
          if (n0 < ep.alpha_cols) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (n0 + i < ep.alpha_cols) ? v[i] * ep.alpha : v[i];
          }
          if (ep.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (ep.row_scale) {
            const float4* cp = reinterpret_cast<const float4*>(ep.col_vec + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = __ldg(cp + q);
              v[4 * q] = fmaf(rs, t.x, v[4 * q]); v[4 * q + 1] = fmaf(rs, t.y, v[4 * q + 1]);
              v[4 * q + 2] = fmaf(rs, t.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(rs, t.w, v[4 * q + 3]);
            }
          }
          if (ep.drop_scale != 0.f && !ep.drop_after_res) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {                     // one hash per four columns (n0 is a multiple of 32)
              const uint32_t h = drop_hash4(dseed, m, (n0 >> 2) + i4);
#pragma unroll
              for (int e = 0; e < 4; ++e) v[4 * i4 + e] = drop_keep_byte(h, e, ep.drop_thresh) ? v[4 * i4 + e] * ep.drop_scale : 0.f;
            }
          }
          if (ep.residual) {
            if (ep.residual_bf16) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 t = res_cur[q];
                const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = bf16x2_to_f2(tw[e]);
                  if (ep.residual_gate) {                         // backward of ReLU (+ dropout) of the layer below, see kernels.h
                    v[8 * q + 2 * e] = f.x > 0.f ? v[8 * q + 2 * e] * ep.gate_scale : 0.f;
                    v[8 * q + 2 * e + 1] = f.y > 0.f ? v[8 * q + 2 * e + 1] * ep.gate_scale : 0.f;
                  } else {
                    v[8 * q + 2 * e] += f.x; v[8 * q + 2 * e + 1] += f.y;
                  }
                }
              }
            } else {
              const float4* rp = reinterpret_cast<const float4*>(ep.residual + res_row + n0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(rp + q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
              }
            }
          }
          if (ep.drop_scale != 0.f && ep.drop_after_res) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {                     // one hash per four columns (n0 is a multiple of 32)
              const uint32_t h = drop_hash4(dseed, m, (n0 >> 2) + i4);
#pragma unroll
              for (int e = 0; e < 4; ++e) v[4 * i4 + e] = drop_keep_byte(h, e, ep.drop_thresh) ? v[4 * i4 + e] * ep.drop_scale : 0.f;
            }
          }
          }
          const long long o = epi_out_index(ep, m, n0, ldc);     // 32 columns never straddle a head (dh % 32 == 0)
          if (tma_chunk && PAIR) {
            unsigned char* buf = stage + ((n_store >> 1) % NBUF) * BOXB;
            if ((t & 1) == 0) {
              if (lane == 0) tma_store_wait_read<NBUF - 1>();    // the box sent NBUF pairs ago has left this buffer
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {                        // row = lane (128 B), 16-byte chunk j at its 128B-swizzled place
              uint4 pk;
              pk.x = f2_to_bf16x2(v[8 * q + 0], v[8 * q + 1]);
              pk.y = f2_to_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              pk.z = f2_to_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              pk.w = f2_to_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              const int j = (t & 1) * 4 + q;
              sts_v4(smem_u32(buf) + lane * 128 + ((j ^ (lane & 7)) << 4), pk.x, pk.y, pk.z, pk.w);
            }
            if (t & 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmC64, buf, n0 - 32, m_blk * GM + quad * 32);   // rows >= M are clipped by the tensor map
                tma_store_commit();
              }
            }
            ++n_store;
          } else if (tma_chunk) {
            unsigned char* buf = stage + (n_store % NBUF) * 2048;
            if (lane == 0) tma_store_wait_read<NBUF - 1>();      // the box sent NBUF chunks ago has left this buffer
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {                        // row = lane, 16-byte chunk q at its 64B-swizzled place
              uint4 pk;
              pk.x = f2_to_bf16x2(v[8 * q + 0], v[8 * q + 1]);
              pk.y = f2_to_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              pk.z = f2_to_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              pk.w = f2_to_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              sts_v4(smem_u32(buf) + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), pk.x, pk.y, pk.z, pk.w);
            }
            fence_proxy_async();
            __syncwarp();
#ifndef V2M_EXP_NOSTORE
            if (lane == 0) {
              tma_store_2d(&tmC, buf, n0, m_blk * GM + quad * 32);   // rows >= M are clipped by the tensor map
              tma_store_commit();
            }
#endif
            ++n_store;
          } else if (out_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<bf16*>(C) + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 pk;
              pk.x = f2_to_bf16x2(v[8 * q + 0], v[8 * q + 1]);
              pk.y = f2_to_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              pk.z = f2_to_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              pk.w = f2_to_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              dst[q] = pk;
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(C) + o);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else {
          // ---- edge path (N tail or unaligned operands): scalar, guarded.  Fully unrolled: a dynamically indexed v[]
          // would live in local memory for the fast path as well.
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int n = n0 + i;
            if (n >= N) continue;
            float x = v[i];
            if (bias) x += __ldg(bias + n);
            if (n < ep.alpha_cols) x *= ep.alpha;
            if (ep.relu) x = fmaxf(x, 0.f);
            if (ep.row_scale) x = fmaf(rs, __ldg(ep.col_vec + n), x);
            if (ep.drop_scale != 0.f && !ep.drop_after_res) x = drop_keep(dseed, m, n, ep.drop_thresh) ? x * ep.drop_scale : 0.f;
            if (ep.residual) {
              const float rv = ep.residual_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(ep.residual)[res_row + n])
                                                : __ldg(ep.residual + res_row + n);
              x = ep.residual_gate ? (rv > 0.f ? x * ep.gate_scale : 0.f) : x + rv;
            }
            if (ep.drop_scale != 0.f && ep.drop_after_res) x = drop_keep(dseed, m, n, ep.drop_thresh) ? x * ep.drop_scale : 0.f;
            const long long o = epi_out_index(ep, m, n, ldc);
            if (out_bf16) static_cast<bf16*>(C)[o] = __float2bfloat16_rn(x);
            else static_cast<float*>(C)[o] = x;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) pair_mbar_arrive(pair_mapa(smem_u32(tempty_bar + acc), 0)); else mbar_arrive(tempty_bar + acc); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (tma_out && lane == 0) tma_store_wait_read<0>();      // shared memory must outlive the reads of the last boxes
  }
  tc_fence_before();
  if (PAIR) pair_sync(); else __syncthreads();             // (pair: nobody leaves while the peer may still read its shared memory)
  if (warp == 2) { if (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base); }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, void* C, int ldc, int out_bf16, int vec_ok, int M, int N, int K,
                       const GemmEpilogue& ep, cudaStream_t stream, bool allow_split, const int* tile_group = nullptr,
                       const int* k_off = nullptr, int n_kgroups = 0, long long c_gstride = 0, const CUtensorMap* tmB_sub = nullptr,
                       const CUtensorMap* tmB_half = nullptr) {
  // TMA-store epilogue for plain bf16 outputs (the store map only exists in that case; otherwise it aliases tmA, unused)
  CUtensorMap tmC = tmA;
  int tma_out = 0;
  if (out_bf16 && vec_ok && !ep.head_scatter && make_tmap_store_bf16(&tmC, C, M, N, ldc) == kOk) tma_out = 1;
  using Cfg = GemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, A_MN, B_MN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if constexpr (BN == 256 && !A_MN)
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, A_MN, B_MN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) { set_last_error("gemm_bf16_tc: smem attribute: %s", cudaGetErrorString(e)); return kCudaError; }
    attr = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = ((M + GM - 1) / GM) * ((N + BN - 1) / BN);
  int k_split = 1;
  const int num_kb = (K + GK - 1) / GK;
  if (allow_split && tiles * 2 <= num_sms && num_kb >= 16) {
    k_split = num_sms / tiles;
    if (k_split > num_kb / 4) k_split = num_kb / 4;           // at least 4 k-blocks per slice
    if (k_split < 1) k_split = 1;
  }
  if (k_off) k_split = n_kgroups;                            // K-grouped: one K range per group, outputs C + g * c_gstride (cleared by the caller)
  if (k_split > 1 && !ep.accumulate && !k_off) {
    cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, stream);
    if (e != cudaSuccess) { set_last_error("gemm_bf16_tc: clearing the split-K output failed: %s", cudaGetErrorString(e)); return kCudaError; }
  }
  // CTA pairs (see the kernel): plain GEMMs with at least pair_min_rounds rounds of 256 x BN tiles on the 74 SM pairs -- the
  // small-batch shapes keep the 1-CTA kernel, whose tail sub-tiles matter more there.  V2M_GEMM_PAIR=0 turns the pair kernel off,
  // =2 forces it for every eligible shape (tests).
  if constexpr (BN == 256 && !A_MN) {
    static int pair_mode = -1;
    if (pair_mode < 0) { const char* e = getenv("V2M_GEMM_PAIR"); pair_mode = e ? atoi(e) : 1; }
    const int pair_tiles = (((M + GM - 1) / GM + 1) / 2) * ((N + BN - 1) / BN);
    const int n_pairs = num_sms / 2;
    const bool eligible = pair_mode > 0 && k_split == 1 && !tile_group && !k_off && (B_MN || tmB_half) && num_sms % 2 == 0;
    if (eligible && (pair_mode >= 2 || pair_tiles >= 8 * n_pairs)) {
      const int grid2 = 2 * (pair_tiles < n_pairs ? pair_tiles : n_pairs);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid2);
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = Cfg::kSmem;
      cfg.stream = stream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = dep_launch_enabled() ? 2 : 1;
      CUtensorMap tmC64 = tmC;
      if (tma_out && make_tmap_store_bf16(&tmC64, C, M, N, ldc, 1) != kOk) return kCudaError;
      cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<BN, A_MN, B_MN, true>, tmA, tmB, C, ldc, out_bf16, vec_ok, M, N, K, ep, 1,
                                          (const int*)nullptr, tmC, tma_out, (const int*)nullptr, 0ll, B_MN ? tmB : *tmB_half, -1, tmC64);
      if (le == cudaSuccess) return check_launch("gemm_bf16_tc (pair)");
      // a device / driver configuration that refuses the 2-CTA cluster launch: the 1-CTA kernel below computes the same result
      (void)cudaGetLastError();
      pair_mode = 0;
    }
  }
  // Tail sub-tiles (see the kernel): when the last round of the persistent grid would hold only a few 128 x BN tiles, those are
  // cut into 128 x 64 pieces that spread over the SMs.  V2M_GEMM_TAIL=0 turns it off (A/B runs).
  static int tail_on = -1;
  if (tail_on < 0) { const char* e = getenv("V2M_GEMM_TAIL"); tail_on = (e && e[0] == '0') ? 0 : 1; }
  int tail_first = -1;
  int items = tiles * k_split;
  if (tail_on && BN == 256 && k_split == 1 && !tile_group && !k_off && tmB_sub && tiles > num_sms) {
    const int r = tiles % num_sms;
    if (r > 0 && r * (BN / 64) <= num_sms) {
      tail_first = tiles - r;
      items = tail_first + r * (BN / 64);
    }
  }
  const int grid = items < num_sms ? items : num_sms;
  cudaError_t le = launch_dep(gemm_bf16_tc_kernel<BN, A_MN, B_MN, false>, dim3(grid), dim3(kGemmThreads), Cfg::kSmem, stream, tmA, tmB, C, ldc, out_bf16,
                              vec_ok, M, N, K, ep, k_split, tile_group, tmC, tma_out, k_off, c_gstride, tmB_sub ? *tmB_sub : tmB, tail_first, tmC);
  if (le != cudaSuccess) { set_last_error("gemm_bf16_tc: %s", cudaGetErrorString(le)); return kCudaError; }
  return check_launch("gemm_bf16_tc");
}

int gemm_bf16_tc_general(const void* A, int lda, int a_mn, const void* W, int ldw, int b_mn, void* C, int ldc, int out_bf16,
                         int M, int N, int K, const GemmEpilogue& ep, cudaStream_t stream) {
  V2M_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_bf16_tc: bad dims M=%d N=%d K=%d", M, N, K);
  if (M == 0) return kOk;
  int bn = (N % 256 == 0) ? 256 : 128;
  if (bn == 256) {
    // Wave quantisation: the persistent grid hands every SM ceil(tiles / SMs) tiles.  When 128 x 256 tiles leave most SMs idle
    // in the last round (e.g. M = 19136 = 150 row tiles on 148 SMs with N = 512: 300 tiles, 3 rounds for 2.03 rounds of
    // work) the 128 x 128 tiling wastes less, although a narrow tile costs a little more than half a wide one.
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long long mt = (M + GM - 1) / GM;
    const long long t256 = mt * (N / 256), t128 = mt * (N / 128);
    const double c256 = (double)((t256 + sms - 1) / sms), c128 = 0.56 * (double)((t128 + sms - 1) / sms);
    if (t256 > sms && c128 < 0.93 * c256) bn = 128;           // (small outputs keep the wide tile: they are split along K)
  }
  CUtensorMap tmA, tmB;
  // K-major operand: [rows = M|N, cols = K], box (rows x 64 k);  MN-major operand: [rows = K, cols = M|N], box (64 k x 64)
  int rc = a_mn ? make_tmap_2d_bf16(&tmA, A, K, M, lda, 64) : make_tmap_2d_bf16(&tmA, A, M, K, lda, GM);
  if (rc) return rc;
  rc = b_mn ? make_tmap_2d_bf16(&tmB, W, K, N, ldw, 64) : make_tmap_2d_bf16(&tmB, W, N, K, ldw, bn);
  if (rc) return rc;
  // 128-bit epilogue accesses need 16-byte aligned vectors, residual rows and output rows
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const int osz = out_bf16 ? 2 : 4;
  bool vec_ok = al16(ep.bias) && al16(ep.col_vec) && al16(ep.residual) && al16(C);
  vec_ok = vec_ok && (!ep.residual || ((long long)ep.ldr * (ep.residual_bf16 ? 2 : 4)) % 16 == 0);
  if (ep.head_scatter) vec_ok = vec_ok && ep.dh % 32 == 0 && ((long long)ep.part_stride * osz) % 16 == 0;
  else vec_ok = vec_ok && ((long long)ldc * osz) % 16 == 0;
  // split-K only for plain fp32 outputs (no epilogue operation commutes with partial sums except the identity)
  const bool allow_split = !out_bf16 && !ep.bias && !ep.residual && !ep.row_scale && !ep.relu && ep.alpha_cols == 0 && !ep.head_scatter &&
                           ep.drop_scale == 0.f;
  V2M_REQUIRE(!ep.accumulate || allow_split, "gemm_bf16_tc: accumulate needs a plain fp32 output (no bias / residual / activation / dropout)");
  V2M_REQUIRE(!ep.residual_gate || (ep.residual && ep.residual_bf16), "gemm_bf16_tc: the gate operand is a bf16 matrix");
  // 64-row box over the same B operand for the tail sub-tiles (the MN-major operand is fetched in 64-column boxes anyway)
  CUtensorMap tmBs = tmB;
  if (bn == 256 && !b_mn && (rc = make_tmap_2d_bf16(&tmBs, W, N, K, ldw, 64))) return rc;
  // 128-row box for the CTA-pair kernel (each CTA of a pair fetches half of the 256 rows of a B tile)
  CUtensorMap tmBh = tmB;
  const bool has_half = bn == 256 && !b_mn;
  if (has_half && (rc = make_tmap_2d_bf16(&tmBh, W, N, K, ldw, 128))) return rc;
#define V2M_GO(BN_, AM_, BM_) launch_gemm<BN_, AM_, BM_>(tmA, tmB, C, ldc, out_bf16, vec_ok, M, N, K, ep, stream, allow_split, nullptr, nullptr, 0, 0, &tmBs, has_half ? &tmBh : nullptr)
  if (bn == 256) {
    if (!a_mn && !b_mn) return V2M_GO(256, false, false);
    if (!a_mn && b_mn) return V2M_GO(256, false, true);
    if (a_mn && b_mn) return V2M_GO(256, true, true);
    return V2M_GO(256, true, false);
  }
  if (!a_mn && !b_mn) return V2M_GO(128, false, false);
  if (!a_mn && b_mn) return V2M_GO(128, false, true);
  if (a_mn && b_mn) return V2M_GO(128, true, true);
  return V2M_GO(128, true, false);
#undef V2M_GO
}

// Grouped GEMM over expert-contiguous rows (see the kernel comment): A [M_cap, K], W [n_groups * N, K] stacked, bias [n_groups * N],
// tile_group [ceil(M_cap / 128)] on the device.  Bias and ReLU are the only epilogue operations.
int gemm_bf16_tc_grouped(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int out_bf16, int M_cap, int N, int K,
                         int n_groups, const int* tile_group, const float* bias, int relu, cudaStream_t stream) {
  V2M_REQUIRE(M_cap >= 0 && N > 0 && K > 0 && n_groups > 0 && tile_group, "gemm_bf16_tc_grouped: bad arguments");
  V2M_REQUIRE(N % 128 == 0, "gemm_bf16_tc_grouped: N=%d must be a multiple of 128 (a weight tile must not straddle two experts)", N);
  if (M_cap == 0) return kOk;
  const int bn = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, A, M_cap, K, lda, GM);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, W, (long long)n_groups * N, K, ldw, bn);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.bias = bias;
  ep.relu = relu;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const int vec_ok = al16(bias) && al16(C) && ((long long)ldc * (out_bf16 ? 2 : 4)) % 16 == 0;
  if (bn == 256) return launch_gemm<256, false, false>(tmA, tmB, C, ldc, out_bf16, vec_ok, M_cap, N, K, ep, stream, false, tile_group);
  return launch_gemm<128, false, false>(tmA, tmB, C, ldc, out_bf16, vec_ok, M_cap, N, K, ep, stream, false, tile_group);
}

// K-grouped dW: C[g] (M x N fp32, group stride c_gstride elements) = A[rows of group g]^T B[rows of group g], both operands
// stored row-major over the rows ([R, M] and [R, N] bf16: "MN-major"), group g = rows [k_off[g], k_off[g+1]) with every bound a
// multiple of 64 (the 128-row aligned expert groups of moe_permute; padding rows are zero in at least one operand).
int gemm_bf16_tc_kgrouped(const void* A, int lda, const void* B, int ldb, float* C, int ldc, long long c_gstride, int M, int N,
                          int R, int n_groups, const int* k_off, cudaStream_t stream) {
  V2M_REQUIRE(M > 0 && N > 0 && R >= 0 && n_groups > 0 && k_off, "gemm_bf16_tc_kgrouped: bad arguments");
  cudaError_t e = cudaMemsetAsync(C, 0, (size_t)n_groups * (size_t)c_gstride * 4, stream);
  if (e != cudaSuccess) { set_last_error("gemm_bf16_tc_kgrouped: clearing the output failed: %s", cudaGetErrorString(e)); return kCudaError; }
  if (R == 0) return kOk;
  const int bn = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, A, R, M, lda, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, B, R, N, ldb, 64);
  if (rc) return rc;
  GemmEpilogue ep;
  const int vec_ok = reinterpret_cast<uintptr_t>(C) % 16 == 0 && ((long long)ldc * 4) % 16 == 0 && (c_gstride * 4) % 16 == 0;
  if (bn == 256) return launch_gemm<256, true, true>(tmA, tmB, C, ldc, 0, vec_ok, M, N, R, ep, stream, false, nullptr, k_off, n_groups, c_gstride);
  return launch_gemm<128, true, true>(tmA, tmB, C, ldc, 0, vec_ok, M, N, R, ep, stream, false, nullptr, k_off, n_groups, c_gstride);
}

int gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int out_bf16, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream) {
  return gemm_bf16_tc_general(A, lda, 0, W, ldw, 0, C, ldc, out_bf16, M, N, K, ep, stream);
}

}  // namespace v2m
