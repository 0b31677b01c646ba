// Attention backward on tcgen05 / TMEM / TMA (bf16 operands, fp32 accumulate, sm_100a) -- the stock attentions of the AMT
// (encoder self-attention and decoder cross-attention: no relative term, no mask), i.e. the gradient torch autograd derives
// for F.multi_head_attention_forward (model/rpr.py:42,62-63; nn.TransformerEncoderLayer):
//
//   P = exp(Q K^T - lse),  Pd = dropout(P),  D_i = dO_i . O_i,  dS = P o (dropout'(dO V^T) - D)
//   dV = Pd^T dO,   dK = dS^T Q,   dQ = dS K
//
// One CTA owns one (video, head): Q, K, V, dO of the whole problem (L <= 320) are resident in shared memory, every product
// runs on the tensor core and NOTHING goes through HBM between them (the mma.sync kernels of attn_bwd_tc.cu write the P and dS
// tiles -- 1.8 GB per launch at B = 512 -- to a workspace for a second kernel).
//
// Orientation: keys on the TMEM lanes.  For key tile j (128 keys) and query block i (64 queries):
//   S^T  = K_j Q_i^T          [128 x 64]   TMEM cols [  0, 64)      A = K_j (K-major), B = Q_i (K-major)
//   dP^T = V_j dO_i^T         [128 x 64]   TMEM cols [ 64,128)
//   the 8 softmax warps (two per lane quadrant, 32 query columns each) turn them into P^T (dropped) and dS^T, bf16, written
//   to shared memory as [128 keys x 64 queries] tiles in the 128-byte-swizzled operand layout;
//   dV_j += P^T dO_i          TMEM cols [128,192)   A = P^T tile (K-major),  B = dO_i (MN-major: [queries x dims])
//   dK_j += dS^T Q_i          TMEM cols [192,256)   A = dS^T tile (K-major), B = Q_i  (MN-major)
//   dQ_I += dS_I K_j          TMEM cols [256,448)   for the pair I = (i-1, i) of query blocks: A = the two dS^T tiles read as
//                                                  an MN-major operand (M = 128 queries, K = 128 keys), B = K_j (MN-major)
// dQ accumulates over the key tiles in TMEM (3 x 64 columns), dK / dV leave after each key tile.
// An odd last query block issues its dQ product with a stale second half: only accumulator rows >= Lq see it, and those
// are never stored.  Keys >= Lk (zero rows of K / V) are masked in P^T and dS^T.
//
// Pipeline per step n = (j, i):   issuer: S^T / dP^T of step n+1 are issued as soon as the softmax warps have pulled step n
// out of TMEM (bar_sfree); the three accumulating products of step n follow when the tiles are in shared memory (bar_pds)
// and free them again when they retire (bar_tfree).
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cstdlib>

namespace v2m {

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, long long cols, long long rows, long long batch,
                      long long row_pitch, long long batch_pitch, int box_rows, int swap);

namespace ab5 {

constexpr int DH = 64;
constexpr int MAXL = 320;                 // queries / keys per (video, head)
constexpr int QROWS = 320, KROWS = 384;   // rows held in shared memory (keys padded to whole 128-row tiles, zero rows)
constexpr int SM_Q = QROWS * 128, SM_K = KROWS * 128, SM_TILE = 128 * 128;
constexpr int threads(int nw) { return 32 + 128 * nw; }   // nw = softmax warps per TMEM lane quadrant (64 / nw query columns each)
constexpr size_t SMEM = 1024 + 2 * SM_Q + 2 * SM_K + 3 * SM_TILE + 256;

struct Args {
  const float* lse; const float* delta;
  void* dq; void* dk; void* dv;
  long long dq_sb, dq_sl, dkv_sb, dkv_sl;
  int B, H, Lq, Lk, swap;
  float drop_scale; unsigned int drop_thresh, drop_seed; const unsigned int* drop_seed_dev;
  float dq_scale;
};

// delta[bh][i] = sum_d dO[b,i,h,d] * O[b,i,h,d]: eight lanes per (video, position, head) row of 64 dims (one 16-byte load of
// each operand per lane); consecutive lane groups take consecutive heads of the same position, so a warp reads 512 contiguous
// bytes of O and of dO.
__global__ void __launch_bounds__(256) delta_kernel(const bf16* __restrict__ o, long long o_sb, long long o_sl, const bf16* __restrict__ dO,
                                                    long long do_sb, long long do_sl, float* __restrict__ delta, int B, int H, int Lq) {
  grid_dep_launch();
  grid_dep_wait();
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;        // (b, i, h) with h fastest
  const int sub = threadIdx.x & 7;
  const bool ok = grp < (long long)B * H * Lq;
  float s = 0.f;
  long long b = 0; int i = 0, h = 0;
  if (ok) {
    h = (int)(grp % H);
    const long long bi = grp / H;
    i = (int)(bi % Lq);
    b = bi / Lq;
    const uint4 a = *reinterpret_cast<const uint4*>(o + b * o_sb + (long long)i * o_sl + h * DH + 8 * sub);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + b * do_sb + (long long)i * do_sl + h * DH + 8 * sub);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = bf16x2_to_f2(aw[e]), y = bf16x2_to_f2(gw[e]);
      s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) delta[((size_t)b * H + h) * Lq + i] = s;
}

template <bool DROP, int NW>
__global__ void __launch_bounds__(threads(NW), 1)
attn_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ Args a) {
  extern __shared__ unsigned char ab5_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ab5_smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sDO = sQ + SM_Q;
  unsigned char* sK = sDO + SM_Q;
  unsigned char* sV = sK + SM_K;
  unsigned char* sP = sV + SM_K;            // P^T tile
  unsigned char* sDS = sP + SM_TILE;        // two dS^T tiles (even / odd query block of a pair)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * SM_TILE);
  uint64_t* bar_ld = bars + 0;       // operands of the item landed
  uint64_t* bar_s = bars + 1;        // S^T and dP^T of a step complete
  uint64_t* bar_sfree = bars + 2;    // all softmax warps have pulled them out of TMEM (count 4 * NW)
  uint64_t* bar_pds = bars + 3;      // P^T / dS^T tiles written (count 4 * NW)
  uint64_t* bar_tfree = bars + 4;    // the accumulating products of a step have retired: tiles (and, after a key tile, dV / dK) final
  uint64_t* bar_kvfree = bars + 5;   // dV / dK pulled out of TMEM (count 4 * NW)
  uint64_t* bar_qfree = bars + 6;    // dQ pulled out of TMEM (count 4 * NW): next item may start
  uint64_t* bar_done = bars + 7;     // all products of the item retired (issuer only): the operand buffers may be reloaded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = (a.Lq + 63) / 64;          // query blocks of 64
  const int nkt = (a.Lk + 127) / 128;        // key tiles of 128
  const int nq64 = nqb, nk64 = (a.Lk + 63) / 64;
  const int n_items = a.B * a.H;
  const int steps = nkt * nqb;

  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_s, 1); mbar_init(bar_sfree, 4 * NW); mbar_init(bar_pds, 4 * NW);
    mbar_init(bar_tfree, 1); mbar_init(bar_kvfree, 4 * NW); mbar_init(bar_qfree, 4 * NW); mbar_init(bar_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
  }
  // rows of K / V beyond the loaded boxes are read by the MN-major dQ product (times dS = 0): they must be finite
  constexpr int THREADS = threads(NW);
  constexpr int CW = 64 / NW;               // query columns per softmax warp
  for (int i = threadIdx.x; i < 2 * SM_K / 16; i += THREADS) reinterpret_cast<uint4*>(sK)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  grid_dep_launch();
  grid_dep_wait();                                             // the prologue above overlapped the previous kernel's tail
  const uint32_t T_S = tmem, T_DP = tmem + 64, T_DV = tmem + 128, T_DK = tmem + 192, T_DQ = tmem + 256;

  auto tma3 = [&](void* dst, const CUtensorMap* tm, int col, int row, int batch, uint64_t* bar) {
    if (a.swap) tma_load_3d(dst, tm, col, batch, row, bar);
    else tma_load_3d(dst, tm, col, row, batch, bar);
  };

  if (warp == 0) {
    // The whole warp walks the steps in convergent code (asm-loop barrier waits); one elected lane issues the TMA loads and the
    // ~20 MMAs of a 64-query step back to back from uniform registers.  Under `if (lane == 0)` with C++ spin loops every MMA
    // was wrapped in an ELECT / R2UR x 7 / branch waterfall (~25 instructions): the issuing thread, not the tensor pipe or the
    // softmax warps, bounded the kernel (see umma_kstep4 in common.cuh).
    {
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);      // S^T, dP^T: both operands K-major
      const uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);     // dV, dK: A K-major tile, B MN-major
      const uint32_t idesc_q = make_idesc_bf16(128, 64, 1, 1);      // dQ: A = dS^T tiles read MN-major, B = K_j MN-major
      uint32_t ph_ld = 0, ph_sfree = 0, ph_pds = 0, ph_kvfree = 0, ph_qfree = 0, ph_done = 0;
      auto issue_s = [&](int n) {                                    // S^T and dP^T of step n
        const int j = n / nqb, i = n - j * nqb;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss(T_S, make_smem_desc_sw128(k_addr + j * SM_TILE + k * 32, 16, 1024),
                       make_smem_desc_sw128(q_addr + i * 8192 + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss(T_DP, make_smem_desc_sw128(v_addr + j * SM_TILE + k * 32, 16, 1024),
                       make_smem_desc_sw128(do_addr + i * 8192 + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(bar_s);
      };
      bool first_item = true;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / a.H, h = item - b * a.H;
        // every product of the previous item has retired (its last commit): the operand buffers may be overwritten while the
        // softmax warps are still storing that item's dQ
        if (!first_item) { mbar_wait_u(bar_done, ph_done); ph_done ^= 1; }
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_ld, (uint32_t)(2 * nq64 + 2 * nk64) * 8192u);
          for (int r = 0; r < nq64; ++r) {
            tma3(sQ + r * 8192, &tmQ, h * DH, r * 64, b, bar_ld);
            tma3(sDO + r * 8192, &tmDO, h * DH, r * 64, b, bar_ld);
          }
          for (int r = 0; r < nk64; ++r) {
            tma3(sK + r * 8192, &tmK, h * DH, r * 64, b, bar_ld);
            tma3(sV + r * 8192, &tmV, h * DH, r * 64, b, bar_ld);
          }
        }
        __syncwarp();
        if (!first_item) {                                           // dQ of the previous item is out of TMEM (the stores of the
          mbar_wait_u(bar_qfree, ph_qfree); ph_qfree ^= 1;             // epilogue overlapped the loads above)
          tc_fence_after();
        }
        first_item = false;
        mbar_wait_u(bar_ld, ph_ld); ph_ld ^= 1;
        tc_fence_after();
        if (elect_one()) issue_s(0);
        __syncwarp();
        for (int n = 0; n < steps; ++n) {
          const int j = n / nqb, i = n - j * nqb;
          mbar_wait_u(bar_sfree, ph_sfree); ph_sfree ^= 1;             // step n is in the warps' registers
          tc_fence_after();
          if (n + 1 < steps) {
            if (elect_one()) issue_s(n + 1);
            __syncwarp();
          }
          if (i == 0 && j > 0) {                                     // dV / dK of the previous key tile are out of TMEM
            mbar_wait_u(bar_kvfree, ph_kvfree); ph_kvfree ^= 1;
            tc_fence_after();
          }
          mbar_wait_u(bar_pds, ph_pds); ph_pds ^= 1;                   // P^T and dS^T[i & 1] are in shared memory
          tc_fence_after();
          const uint32_t ds_i = ds_addr + (i & 1) * SM_TILE;
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {                              // K = 64 queries
            umma_bf16_ss(T_DV, make_smem_desc_sw128(p_addr + k * 32, 16, 1024),
                         make_smem_desc_sw128(do_addr + i * 8192 + k * 2048, 1024, 1024), idesc_kv, (i | k) != 0);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss(T_DK, make_smem_desc_sw128(ds_i + k * 32, 16, 1024),
                         make_smem_desc_sw128(q_addr + i * 8192 + k * 2048, 1024, 1024), idesc_kv, (i | k) != 0);
          }
          if ((i & 1) || i == nqb - 1) {                             // pair complete (or odd last block: stale second half)
            const int I = i >> 1;
            const uint32_t a_base = ds_addr;                         // chunk 0 = even block, chunk 1 = odd block
#pragma unroll
            for (int k = 0; k < 8; ++k) {                            // K = 128 keys
              umma_bf16_ss(T_DQ + I * 64, make_smem_desc_sw128(a_base + k * 2048, SM_TILE, 1024),
                           make_smem_desc_sw128(k_addr + j * SM_TILE + k * 2048, 1024, 1024), idesc_q, (j | k) != 0);
            }
          }
          umma_commit(bar_tfree);
          if (n == steps - 1) umma_commit(bar_done);                 // every product of the item has retired: operands may be replaced
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    // ================= softmax warps =================
    const int quad = warp & 3;
    const int sub = (warp - 1) >> 2;                                 // which CW query columns of the block
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int r = quad * 32 + lane;                                  // key row inside the tile == TMEM lane
    const float LOG2E = 1.4426950408889634f;
    uint32_t ph_s = 0, ph_tfree = 0;
    const uint32_t dseed = DROP ? a.drop_seed + (a.drop_seed_dev ? *a.drop_seed_dev : 0u) : 0u;
    unsigned char* my_p = sP + r * 128;
    auto ld_cols = [&](uint32_t taddr, uint32_t* dst) {
      if (CW == 32) tmem_ld_32x32(taddr, dst); else tmem_ld_32x16(taddr, dst);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / a.H, h = item - b * a.H;
      const float* lse = a.lse + (size_t)item * a.Lq;
      const float* delta = a.delta + (size_t)item * a.Lq;
      // per-column constants: lane c (< CW) holds those of column q0 + c and the loop below broadcasts them with shuffles; the
      // values of the NEXT step are fetched while this one is being processed (global latency off the critical path).
      // Queries >= Lq: lse = +inf -> P = 0, dS = 0 (their Q / dO rows are zero).  Keys >= Lk need no mask: their K / V rows are
      // zero, so P^T and dS^T stay finite and only touch dV / dK rows that are never stored, while dQ sees them times K = 0.
      auto fetch = [&](int ii, float& l2, float& dl) {
        const int qc = ii * 64 + sub * CW + (lane & (CW - 1));
        l2 = qc < a.Lq ? __ldg(lse + qc) * LOG2E : INFINITY;
        dl = qc < a.Lq ? __ldg(delta + qc) : 0.f;
      };
      float my_l2, my_dl;
      fetch(0, my_l2, my_dl);
      int j = 0, i = 0;
      for (int n = 0; n < steps; ++n, ++i) {
        if (i == nqb) { i = 0; ++j; }
        const int key = j * 128 + r;
        const bool key_ok = key < a.Lk;
        const int q0 = i * 64 + sub * CW;                            // first query column of this warp
        mbar_wait(bar_s, ph_s); ph_s ^= 1;
        tc_fence_after();
        uint32_t s[CW], dp[CW];
        ld_cols(T_S + lane_off + sub * CW, s);
        ld_cols(T_DP + lane_off + sub * CW, dp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfree);
        float nx_l2, nx_dl;
        fetch(i + 1 == nqb ? 0 : i + 1, nx_l2, nx_dl);
        uint32_t pk[CW / 2], dk_[CW / 2];
        // Dropout keep mask = byte (key & 3) of hash(query row, key >> 2): the four lanes of a key quad need the SAME hash for a
        // given query, so each computes it for a quarter of the queries and the quad exchanges them by shuffle (one hash per
        // four elements, as in the forward kernel, instead of one per element).
        uint32_t hq[DROP ? CW / 4 : 1];
        if (DROP) {
#pragma unroll
          for (int c4 = 0; c4 < CW / 4; ++c4)
            hq[c4] = drop_hash4(dseed, (uint32_t)item * (uint32_t)a.Lq + (uint32_t)(q0 + c4 * 4 + (lane & 3)), (uint32_t)key >> 2);
        }
#pragma unroll
        for (int c = 0; c < CW; c += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int qi = q0 + c + e;
            const float l2 = __shfl_sync(0xffffffffu, my_l2, c + e);
            const float dl = __shfl_sync(0xffffffffu, my_dl, c + e);
            float p = ex2_approx(fmaf(__uint_as_float(s[c + e]), LOG2E, -l2));
            float g = __uint_as_float(dp[c + e]);
            float pd = p;
            if (DROP) {
              const uint32_t hh = __shfl_sync(0xffffffffu, hq[(c + e) >> 2], (lane & ~3) | ((c + e) & 3));
              const bool keep = drop_keep_byte(hh, (uint32_t)key, a.drop_thresh);
              pd = keep ? p * a.drop_scale : 0.f;
              g = keep ? g * a.drop_scale : 0.f;
            }
            pv[e] = pd;
            dv[e] = p * (g - dl);
          }
          pk[c >> 1] = f2_to_bf16x2(pv[0], pv[1]);
          dk_[c >> 1] = f2_to_bf16x2(dv[0], dv[1]);
        }
        my_l2 = nx_l2; my_dl = nx_dl;
        if (i > 0) {                                                 // products of step n-1 have retired: the tiles are free
          mbar_wait(bar_tfree, ph_tfree); ph_tfree ^= 1;             // (i == 0: the end-of-key-tile wait below already saw them)
        }
        unsigned char* my_ds = sDS + (i & 1) * SM_TILE + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < CW / 8; ++q4) {                        // this warp's CW * 2 bytes of row r: 16-byte chunks sub*CW/8 ..
          const int ch = ((sub * (CW / 8) + q4) ^ (r & 7)) << 4;
          sts_v4(smem_u32(my_p) + ch, pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);          // explicit shared-space
          sts_v4(smem_u32(my_ds) + ch, dk_[4 * q4], dk_[4 * q4 + 1], dk_[4 * q4 + 2], dk_[4 * q4 + 3]);   // stores (common.cuh)
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
        if (i == nqb - 1) {
          // ---- dV_j / dK_j complete once this step's products retire: pull them out, store bf16 rows
          mbar_wait(bar_tfree, ph_tfree); ph_tfree ^= 1;
          tc_fence_after();
          uint32_t o[CW];
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            ld_cols((which ? T_DK : T_DV) + lane_off + sub * CW, o);
            tmem_ld_wait();
            if (key_ok) {
              bf16* dst = static_cast<bf16*>(which ? a.dk : a.dv) + (size_t)b * a.dkv_sb + (size_t)key * a.dkv_sl + (size_t)h * DH + sub * CW;
#pragma unroll
              for (int g8 = 0; g8 < CW / 8; ++g8) {
                uint4 v;
                v.x = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 0]), __uint_as_float(o[g8 * 8 + 1]));
                v.y = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 2]), __uint_as_float(o[g8 * 8 + 3]));
                v.z = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 4]), __uint_as_float(o[g8 * 8 + 5]));
                v.w = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 6]), __uint_as_float(o[g8 * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + g8 * 8) = v;
              }
            }
          }
          if (j == nkt - 1) {
            // ---- dQ complete: accumulator I holds queries [128 I, 128 I + 128) on the lanes
            for (int I = 0; I < (nqb + 1) / 2; ++I) {
              ld_cols(T_DQ + I * 64 + lane_off + sub * CW, o);
              tmem_ld_wait();
              const int qi = I * 128 + r;
              if (qi < a.Lq) {
                bf16* dst = static_cast<bf16*>(a.dq) + (size_t)b * a.dq_sb + (size_t)qi * a.dq_sl + (size_t)h * DH + sub * CW;
                const float qs = a.dq_scale;
#pragma unroll
                for (int g8 = 0; g8 < CW / 8; ++g8) {
                  uint4 v;
                  v.x = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 0]) * qs, __uint_as_float(o[g8 * 8 + 1]) * qs);
                  v.y = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 2]) * qs, __uint_as_float(o[g8 * 8 + 3]) * qs);
                  v.z = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 4]) * qs, __uint_as_float(o[g8 * 8 + 5]) * qs);
                  v.w = f2_to_bf16x2(__uint_as_float(o[g8 * 8 + 6]) * qs, __uint_as_float(o[g8 * 8 + 7]) * qs);
                  *reinterpret_cast<uint4*>(dst + g8 * 8) = v;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (j == nkt - 1) mbar_arrive(bar_qfree); else mbar_arrive(bar_kvfree); }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace ab5

// Workspace: delta (fp32, B * H * Lq).
long long attn_bwd_tc5_workspace(int B, int H, int Lq) { return (long long)B * H * Lq * 4; }

bool attn_bwd_tc5_supported(const AttnBwdParams& p) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("V2M_ATTN_BWD_TC5");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled && p.dtype == 1 && p.dh == 64 && !p.Er && !p.causal && p.Hq == p.Hkv && p.Lq <= ab5::MAXL && p.Lk <= ab5::MAXL &&
         p.q_scale == 1.0f && p.dq_sl % 8 == 0 && p.dq_sb % 8 == 0 && p.dkv_sl % 8 == 0 && p.dkv_sb % 8 == 0 &&
         reinterpret_cast<uintptr_t>(p.dq) % 16 == 0 && reinterpret_cast<uintptr_t>(p.dk) % 16 == 0 && reinterpret_cast<uintptr_t>(p.dv) % 16 == 0;
}

int attn_bwd_tc5(const AttnBwdParams& p, void* ws, long long ws_bytes, cudaStream_t stream) {
  V2M_REQUIRE(attn_bwd_tc5_supported(p), "attn_bwd_tc5: configuration not covered");
  V2M_REQUIRE(ws && ws_bytes >= attn_bwd_tc5_workspace(p.B, p.Hq, p.Lq), "attn_bwd_tc5: workspace too small");
  float* delta = static_cast<float*>(ws);
  const long long rows = (long long)p.B * p.Hq * p.Lq;
  launch_dep(ab5::delta_kernel, dim3((unsigned)((rows * 8 + 255) / 256)), dim3(256), 0, stream, static_cast<const bf16*>(p.o), p.o_sb, p.o_sl,
             static_cast<const bf16*>(p.dO), p.do_sb, p.do_sl, delta, p.B, p.Hq, p.Lq);
  int rc = check_launch("attn_bwd_tc5 delta");
  if (rc) return rc;
  CUtensorMap tmQ, tmK, tmV, tmDO;
  const int swap = p.B > 1 && p.q_sl > p.q_sb;
  V2M_REQUIRE(p.B == 1 || ((p.k_sl > p.k_sb) == (swap != 0) && (p.v_sl > p.v_sb) == (swap != 0) && (p.do_sl > p.do_sb) == (swap != 0)),
              "attn_bwd_tc5: q, k, v, dO must share one layout family (batch-first or sequence-first)");
  if ((rc = make_tmap_3d_bf16(&tmQ, p.q, (long long)p.Hq * 64, p.Lq, p.B, p.q_sl, p.q_sb, 64, swap))) return rc;
  if ((rc = make_tmap_3d_bf16(&tmK, p.k, (long long)p.Hkv * 64, p.Lk, p.B, p.k_sl, p.k_sb, 64, swap))) return rc;
  if ((rc = make_tmap_3d_bf16(&tmV, p.v, (long long)p.Hkv * 64, p.Lk, p.B, p.v_sl, p.v_sb, 64, swap))) return rc;
  if ((rc = make_tmap_3d_bf16(&tmDO, p.dO, (long long)p.Hq * 64, p.Lq, p.B, p.do_sl, p.do_sb, 64, swap))) return rc;
  ab5::Args a;
  a.lse = p.lse; a.delta = delta;
  a.dq = p.dq; a.dk = p.dk; a.dv = p.dv;
  a.dq_sb = p.dq_sb; a.dq_sl = p.dq_sl; a.dkv_sb = p.dkv_sb; a.dkv_sl = p.dkv_sl;
  a.B = p.B; a.H = p.Hq; a.Lq = p.Lq; a.Lk = p.Lk; a.swap = swap;
  a.drop_scale = p.drop_scale; a.drop_thresh = p.drop_thresh; a.drop_seed = p.drop_seed; a.drop_seed_dev = p.drop_seed_dev;
  a.dq_scale = p.dq_scale == 0.f ? 1.f : p.dq_scale;
  static bool attr = false;
  static int nw = 2;
  if (!attr) {
    if (const char* ev = getenv("V2M_TC5_NW")) nw = atoi(ev) == 4 ? 4 : 2;       // A/B switch: softmax warps per lane quadrant
    cudaError_t e = cudaFuncSetAttribute(ab5::attn_bwd_tc5_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab5::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ab5::attn_bwd_tc5_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab5::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ab5::attn_bwd_tc5_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab5::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ab5::attn_bwd_tc5_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab5::SMEM);
    if (e != cudaSuccess) { set_last_error("attn_bwd_tc5: smem attribute: %s", cudaGetErrorString(e)); return kCudaError; }
    attr = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int items = p.B * p.Hq;
  const int grid = items < num_sms ? items : num_sms;
  const bool drop = a.drop_scale != 0.f;
#define V2M_GO(DR, NW_) launch_dep(ab5::attn_bwd_tc5_kernel<DR, NW_>, dim3(grid), dim3(ab5::threads(NW_)), ab5::SMEM, stream, tmQ, tmK, tmV, tmDO, a)
  if (nw == 4) { if (drop) V2M_GO(true, 4); else V2M_GO(false, 4); }
  else { if (drop) V2M_GO(true, 2); else V2M_GO(false, 2); }
#undef V2M_GO
  return check_launch("attn_bwd_tc5");
}

}  // namespace v2m
