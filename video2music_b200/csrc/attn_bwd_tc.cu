// Attention backward on the tensor cores (bf16 operands, fp32 accumulation): dQ, dK, dV and dEr of
//   O = softmax(Q K^T + skew(Q Er^T) + causal) V        (model/rpr.py:387-414, _skew :439-455)
// which the reference obtains from autograd over the materialised (B*H, L, L) tensors.  Same math as the exact SIMT
// kernel in attn_bwd.cu (see the formulas there); three kernels, no atomics except the final dEr reduction:
//   rows kernel : one CTA = (batch, head) x 64 query rows, one warp = 16 rows (P and dS never leave the registers of the
//                 warp that owns the rows).  Recomputes S = Q K^T (+ Srel from a per-key-tile Q Er_band^T product), P,
//                 dP = dO V^T, dS = P o (dP - D); accumulates dQ = dS K (+ skewed dS . Er_band) in registers; writes the
//                 P and dS tiles (bf16) to a workspace.
//   cols kernel : one CTA = (batch, kv head) x 64 keys: dV = P^T dO, dK = dS^T Q from the workspace tiles (operands
//                 transposed on the fly by ldmatrix.trans), summed over the query heads of the group (GQA).
//   dEr kernel  : dEr[er_len-1-d] = sum_{b,h,i} dS[i][i-d] q_i read skewed from the dS workspace, split over (batch, head)
//                 groups, one fp32 atomic per output.
// Shared-memory tiles are 64 x 64 bf16 with a pitch of 72 elements: ldmatrix and 32-bit fragment loads are conflict-free.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

namespace abt {

constexpr int TP = 72;                 // tile pitch (bf16 elements)
constexpr int TILE = 64 * TP;          // elements per tile
constexpr int THREADS = 128;

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_t(const bf16* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4(const bf16* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;                         // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// asynchronous version of load_tile (completion: cp_async_wait + __syncthreads)
__device__ __forceinline__ void load_tile_async(bf16* dst, const bf16* src, long long ld, int n_valid) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += THREADS) {
    const int r = idx >> 3, c = idx & 7;
    const bool ok = r < n_valid;
    cp_async16(dst + r * TP + c * 8, src + (ok ? (long long)r * ld + c * 8 : 0), ok);
  }
}

// acc[nt][4] (16 rows x 64 columns = 8 n-tiles) = A (16 x 64, four k-step fragments) * T^T where T holds the n index in
// its rows and the k index in its columns (K / V tiles: rows = keys, columns = head dims): one ldmatrix.x4 feeds two MMAs.
__device__ __forceinline__ void mma_a_times_colmajor(float acc[8][4], const uint32_t a[4][4], const bf16* t, int lane) {
  const int ri = lane & 7, mi = lane >> 3;
#pragma unroll
  for (int ntp = 0; ntp < 4; ++ntp) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t r0, r1, r2, r3;
      ldsm_x4(t + ((2 * ntp + (mi >> 1)) * 8 + ri) * TP + ks * 16 + (mi & 1) * 8, r0, r1, r2, r3);
      mma16816(acc[2 * ntp], a[ks], r0, r1);
      mma16816(acc[2 * ntp + 1], a[ks], r2, r3);
    }
  }
}

// 64 rows x 64 bf16 from global (row stride `ld` elements, rows >= n_valid zero-filled) into a pitch-72 tile
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long ld, int n_valid) {
  for (int idx = threadIdx.x; idx < 64 * 8; idx += THREADS) {
    const int r = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < n_valid) v = *reinterpret_cast<const uint4*>(src + (long long)r * ld + c * 8);
    *reinterpret_cast<uint4*>(dst + r * TP + c * 8) = v;
  }
}

// A fragment (rows r0+g, r0+g+8; k-step ks) of a row-major tile
__device__ __forceinline__ void frag_a(const bf16* t, int r0, int ks, int g, int q, uint32_t* a) {
  const bf16* p = t + (r0 + g) * TP + ks * 16 + 2 * q;
  a[0] = lds32(p); a[1] = lds32(p + 8 * TP); a[2] = lds32(p + 8); a[3] = lds32(p + 8 * TP + 8);
}

// acc[8][4] (16 x 64, n over the 64 columns of `t`) += A(16 x 64, from four k-steps of a_frag) * T  where T is row-major
// [k = 64 rows][n = 64 cols] in shared memory (operand transposed by ldmatrix.trans)
__device__ __forceinline__ void mma_a_times_rowmajor(float acc[8][4], const uint32_t a[4][4], const bf16* t, int lane) {
  const int ri = lane & 7, mi = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int ntp = 0; ntp < 4; ++ntp) {
      uint32_t r0, r1, r2, r3;
      ldsm_x4_t(t + (kk * 16 + (mi & 1) * 8 + ri) * TP + (2 * ntp + (mi >> 1)) * 8, r0, r1, r2, r3);
      mma16816(acc[2 * ntp], a[kk], r0, r1);
      mma16816(acc[2 * ntp + 1], a[kk], r2, r3);
    }
  }
}

struct Ws {          // workspace views (bf16), padded to multiples of 64 in both dims
  bf16* P; bf16* dS; bf16* dQE;
  int Lqp, Lkp;
};

// ------------------------------------------------------------------------------------------------ rows kernel
// RPR (HAS_ER): the relative term needs, for this warp's 16 rows and the 64 keys of tile J0, the distances
// d = i - j in [dhi - 78, dhi], dhi = (I0 + r0 + 15) - J0: an 80-wide band of Er rows.  Band column e <-> d = dhi - e <-> row
// k0 + e of the shared-memory Er slice (ascending Er rows), and element (row r, key column c) of the tile sits at e = 15 - r + c.
//   * QE band  = q (16 x 64) . Er_band^T (80 rows): 40 MMAs per tile, parked as fp32 in the warp's scratch and read back
//     skewed when the scores are formed;
//   * dQ      += dS_skewed (16 x 80) . Er_band: the A fragments are read straight out of the staged dS rows (the same
//     bf16 tile that goes to the workspace) at column c = e - 15 + r, zero outside the tile / above the diagonal.
// No per-CTA (64 x L) tables: 105 KB of shared memory for L = 320, two CTAs per SM (the table version had one).  dEr is
// computed by the dEr kernel from the dS workspace itself.
constexpr int QB = 88;                 // fp32 pitch of the QE band scratch (16 rows x 80 columns)
constexpr int WBUF = 16 * QB * 4;      // per-warp scratch bytes: QE band (fp32), later the staged P / dS rows (2 x 16 x TP bf16)

template <bool HAS_ER, bool DROP>
__global__ void __launch_bounds__(THREADS, HAS_ER ? 2 : 3) attn_bwd_rows_kernel(AttnBwdParams p, Ws ws) {
  v2m::grid_dep_launch();
  v2m::grid_dep_wait();
  extern __shared__ __align__(16) unsigned char abt_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(abt_smem);
  bf16* sdO = sQ + TILE;
  bf16* sK = sdO + TILE;                // first holds the O tile (RPR variant)
  bf16* sV = sK + TILE;
  // second K / V buffer (cp.async double buffering); the RPR variant spends the shared memory on occupancy instead
  constexpr bool DBUF = !HAS_ER;
  bf16* sK2 = DBUF ? sV + TILE : sK;
  bf16* sV2 = DBUF ? sK2 + TILE : sV;
  unsigned char* sWarp = reinterpret_cast<unsigned char*>(sV + (DBUF ? 3 : 1) * TILE);   // 4 x WBUF
  bf16* sEr = reinterpret_cast<bf16*>(sWarp + 4 * WBUF);     // RPR only: [I0 + 64][TP], row k <-> distance d = I0 + 63 - k

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  const int bh = blockIdx.y, b = bh / p.Hq, hq = bh % p.Hq, hkv = hq / (p.Hq / p.Hkv);
  const int I0 = blockIdx.x * 64;
  const int coff = p.Lk - p.Lq;
  const bf16* Qg = static_cast<const bf16*>(p.q) + (long long)b * p.q_sb + (long long)hq * 64;
  const bf16* Kg = static_cast<const bf16*>(p.k) + (long long)b * p.k_sb + (long long)hkv * 64;
  const bf16* Vg = static_cast<const bf16*>(p.v) + (long long)b * p.v_sb + (long long)hkv * 64;
  const bf16* Og = static_cast<const bf16*>(p.o) + (long long)b * p.o_sb + (long long)hq * 64;
  const bf16* dOg = static_cast<const bf16*>(p.dO) + (long long)b * p.do_sb + (long long)hq * 64;
  const int nq = max(0, min(64, p.Lq - I0));
  const int n_jt = p.causal ? min(ws.Lkp / 64, (I0 + 63 + coff) / 64 + 1) : ws.Lkp / 64;
  // K / V tiles are double-buffered with cp.async: tile jt+1 is in flight while tile jt is processed
  auto issue_kv = [&](int jt) {
    const int J0 = jt * 64;
    const int nk = max(0, min(64, p.Lk - J0));
    bf16* kb = (jt & 1) ? sK2 : sK;
    bf16* vb = (jt & 1) ? sV2 : sV;
    load_tile_async(kb, Kg + (long long)J0 * p.k_sl, p.k_sl, nk);
    load_tile_async(vb, Vg + (long long)J0 * p.v_sl, p.v_sl, nk);
    cp_async_commit();
  };
  // Prologue: Q, dO, O (and the Er slice) are fetched asynchronously and together -- a chain of synchronous tile loads
  // was a third of the kernel's stall samples.  The O tile borrows a K buffer that is not needed yet.
  bf16* sO = DBUF ? sK2 : sK;
  load_tile_async(sQ, Qg + (long long)I0 * p.q_sl, p.q_sl, nq);
  load_tile_async(sdO, dOg + (long long)I0 * p.do_sl, p.do_sl, nq);
  load_tile_async(sO, Og + (long long)I0 * p.o_sl, p.o_sl, nq);
  const int er_rows = min(ws.Lkp, I0 + 64);                 // distances d = i - j <= I0 + 63 can occur in this row tile
  if (HAS_ER) {
    const bf16* Er = static_cast<const bf16*>(p.Er);
    for (int idx = tid; idx < er_rows * 8; idx += THREADS) {
      const int k = idx >> 3, c = idx & 7, d = I0 + 63 - k;
      const bool ok = d < p.er_len && d < p.Lq;              // other distances never meet a valid (i, j <= i): zero rows
      cp_async16(sEr + k * TP + c * 8, Er + (ok ? (long long)(p.er_len - 1 - d) * 64 + c * 8 : 0), ok);
    }
  }
  cp_async_commit();
  if (DBUF) issue_kv(0);
  if (DBUF) cp_async_wait<1>(); else cp_async_wait<0>();
  __syncthreads();

  const int r0 = w * 16;
  const uint32_t dseed = DROP ? p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0u) : 0u;
  uint32_t qa[4][4], doa[4][4];
  float Dlo = 0.f, Dhi = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    frag_a(sQ, r0, ks, g, q, qa[ks]);
    frag_a(sdO, r0, ks, g, q, doa[ks]);
    uint32_t oa[4];
    frag_a(sO, r0, ks, g, q, oa);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = bf16x2_to_f2(doa[ks][e]), y = bf16x2_to_f2(oa[e]);
      const float s = x.x * y.x + x.y * y.y;
      if (e & 1) Dhi += s; else Dlo += s;
    }
  }
  Dlo += __shfl_xor_sync(0xffffffffu, Dlo, 1); Dlo += __shfl_xor_sync(0xffffffffu, Dlo, 2);
  Dhi += __shfl_xor_sync(0xffffffffu, Dhi, 1); Dhi += __shfl_xor_sync(0xffffffffu, Dhi, 2);
  const int ilo = I0 + r0 + g, ihi = ilo + 8;
  const float lse_lo = ilo < p.Lq ? p.lse[(long long)bh * p.Lq + ilo] : 0.f;
  const float lse_hi = ihi < p.Lq ? p.lse[(long long)bh * p.Lq + ihi] : 0.f;
  __syncthreads();                                         // the O tile is no longer needed: K / V tiles go there

  float dq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
  bf16* Pw = ws.P + ((long long)bh * ws.Lqp + I0 + r0) * ws.Lkp;
  bf16* dSw = ws.dS + ((long long)bh * ws.Lqp + I0 + r0) * ws.Lkp;
  float* wQE = reinterpret_cast<float*>(sWarp + w * WBUF);            // [16][QB] fp32 QE band ...
  bf16* wSt = reinterpret_cast<bf16*>(sWarp + w * WBUF);              // ... then 16 rows of P + 16 rows of dS (pitch TP)
  const int ri = lane & 7, mi = lane >> 3;
  for (int jt = 0; jt < n_jt; ++jt) {
    const int J0 = jt * 64;
    const bf16* kb = (jt & 1) ? sK2 : sK;
    const bf16* vb = (jt & 1) ? sV2 : sV;
    if (!DBUF) { issue_kv(jt); cp_async_wait<0>(); }
    else if (jt + 1 < n_jt) { issue_kv(jt + 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int dhi = I0 + r0 + 15 - J0;                       // largest distance of this warp's rows in this key tile
    const bool rel_tile = HAS_ER && dhi >= 0;                // some (i, j <= i) in the tile
    const int k0 = 48 - r0 + J0;                             // Er slice row of band column 0 (>= 0)
    if (rel_tile) {
      // QE band: 10 n-tiles of 8 band columns
#pragma unroll
      for (int nt = 0; nt < 10; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        const bf16* bp = sEr + min(k0 + nt * 8 + g, er_rows - 1) * TP + 2 * q;   // clamped rows are above the diagonal (unused)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma16816(c, qa[ks], lds32(bp + ks * 16), lds32(bp + ks * 16 + 8));
        float* dst = wQE + g * QB + nt * 8 + 2 * q;
        *reinterpret_cast<float2*>(dst) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(dst + 8 * QB) = make_float2(c[2], c[3]);
      }
      __syncwarp();
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
    mma_a_times_colmajor(s, qa, kb, lane);
    mma_a_times_colmajor(dp, doa, vb, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int hi = e >> 1, i = hi ? ihi : ilo, c = nt * 8 + 2 * q + (e & 1), j = J0 + c;
        const bool valid = i < p.Lq && j < p.Lk && (!p.causal || j <= i + coff);
        float sc = s[nt][e];
        const bool rel = HAS_ER && valid && j <= i;            // _skew contributes only for j <= i (rpr.py:439-455)
        if (rel) sc += wQE[(g + hi * 8) * QB + 15 - (g + hi * 8) + c];
        const float pv = valid ? __expf(sc - (hi ? lse_hi : lse_lo)) : 0.f;
        // forward dropout of the probabilities: O = (P o M) V, so dV needs P o M and dP = (dO V^T) o M; D = rowsum(dO o O) as is
        float mk = 1.f;
        if (DROP)                                              // j and j ^ 1 share one hash (the compiler merges the two calls)
          mk = drop_keep_byte(drop_hash4(dseed, (uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i, (uint32_t)j >> 2), (uint32_t)j,
                              p.drop_thresh) ? p.drop_scale : 0.f;
        s[nt][e] = pv * mk;
        dp[nt][e] = pv * (dp[nt][e] * mk - (hi ? Dhi : Dlo));
      }
    }
    if (rel_tile) __syncwarp();                                // every lane has read its QE values: the scratch becomes the stage
    // P and dS fragments -> this warp's staging rows (bf16, pitch 72): written out below as whole 128-byte rows
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bf16* sp = wSt + g * TP + nt * 8 + 2 * q;
      *reinterpret_cast<uint32_t*>(sp) = f2_to_bf16x2(s[nt][0], s[nt][1]);
      *reinterpret_cast<uint32_t*>(sp + 8 * TP) = f2_to_bf16x2(s[nt][2], s[nt][3]);
      *reinterpret_cast<uint32_t*>(sp + 16 * TP) = f2_to_bf16x2(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(sp + 24 * TP) = f2_to_bf16x2(dp[nt][2], dp[nt][3]);
    }
    __syncwarp();
    // P and dS tiles for the cols kernel: 16 rows x 128 B each, 16-byte stores, 8 lanes per row
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane, r = idx >> 3, c8 = idx & 7;
      const uint4 pv4 = *reinterpret_cast<const uint4*>(wSt + r * TP + c8 * 8);
      const uint4 dv4 = *reinterpret_cast<const uint4*>(wSt + (16 + r) * TP + c8 * 8);
      *reinterpret_cast<uint4*>(Pw + (long long)r * ws.Lkp + J0 + c8 * 8) = pv4;
      *reinterpret_cast<uint4*>(dSw + (long long)r * ws.Lkp + J0 + c8 * 8) = dv4;
    }
    // dQ += dS K_J : the C fragments of two adjacent key tiles form the A fragment of one k-step
    uint32_t a[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      a[kk][0] = f2_to_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
      a[kk][1] = f2_to_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      a[kk][2] = f2_to_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      a[kk][3] = f2_to_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
    }
    mma_a_times_rowmajor(dq, a, kb, lane);
    if (rel_tile) {
      // dQ += dS_skewed . Er_band : A[r][e] = dS[r][c = e - 15 + r] for 0 <= c < 64 and e <= dhi (j <= i), else 0
      const unsigned short* dsr = reinterpret_cast<const unsigned short*>(wSt + 16 * TP);
      // chk_c: band columns 16..63 always fall inside the tile (c = e - 15 + r in [1, 63]); chk_d: only the tile on the
      // diagonal has band columns beyond dhi.  Both are resolved at compile time / per warp.
      auto skew2 = [&](int r, int e, bool chk_c, bool chk_d) -> uint32_t {   // band columns e, e + 1 of row r, packed
        const int c = e - 15 + r;
        const bool ok_lo = (!chk_c || (c >= 0 && c < 64)) && (!chk_d || e <= dhi);
        const bool ok_hi = (!chk_c || (c + 1 >= 0 && c + 1 < 64)) && (!chk_d || e + 1 <= dhi);
        const uint32_t lo = ok_lo ? dsr[r * TP + c] : 0u;
        const uint32_t hi = ok_hi ? dsr[r * TP + c + 1] : 0u;
        return lo | (hi << 16);
      };
      auto band = [&](bool chk_d) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          uint32_t af[4];
          const int e0 = ks * 16 + 2 * q;
          const bool chk_c = ks == 0 || ks == 4;
          af[0] = skew2(g, e0, chk_c, chk_d); af[1] = skew2(g + 8, e0, chk_c, chk_d);
          af[2] = skew2(g, e0 + 8, chk_c, chk_d); af[3] = skew2(g + 8, e0 + 8, chk_c, chk_d);
          const int kr = min(k0 + ks * 16 + (mi & 1) * 8 + ri, er_rows - 1);
#pragma unroll
          for (int ntp = 0; ntp < 4; ++ntp) {
            uint32_t x0, x1, x2, x3;
            ldsm_x4_t(sEr + kr * TP + (2 * ntp + (mi >> 1)) * 8, x0, x1, x2, x3);
            mma16816(dq[2 * ntp], af, x0, x1);
            mma16816(dq[2 * ntp + 1], af, x2, x3);
          }
        }
      };
      if (dhi >= 79) band(false); else band(true);
    }
    __syncthreads();                                       // everyone is done with this buffer before the next tile lands in it
  }
  bf16* dQg = static_cast<bf16*>(p.dq) + (long long)b * p.dq_sb + (long long)hq * 64;
  const float dqs = p.q_scale * (p.dq_scale == 0.f ? 1.f : p.dq_scale);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (ilo < p.Lq) *reinterpret_cast<uint32_t*>(dQg + (long long)ilo * p.dq_sl + nt * 8 + 2 * q) = f2_to_bf16x2(dq[nt][0] * dqs, dq[nt][1] * dqs);
    if (ihi < p.Lq) *reinterpret_cast<uint32_t*>(dQg + (long long)ihi * p.dq_sl + nt * 8 + 2 * q) = f2_to_bf16x2(dq[nt][2] * dqs, dq[nt][3] * dqs);
  }
}

// A fragments (m = column c0.. of the tile, k = its rows): acc += T^T-style products.  a[kk] = A(16 m x 16 k) with
// A[m][k] = t[kk*16 + k][c0 + m]
__device__ __forceinline__ void frag_a_transposed(const bf16* t, int c0, int lane, uint32_t a[4][4]) {
  const int ri = lane & 7, mi = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4_t(t + (kk * 16 + (mi >> 1) * 8 + ri) * TP + c0 + (mi & 1) * 8, a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}

// ------------------------------------------------------------------------------------------------ cols kernel
__global__ void __launch_bounds__(THREADS) attn_bwd_cols_kernel(AttnBwdParams p, Ws ws) {
  v2m::grid_dep_launch();
  v2m::grid_dep_wait();
  extern __shared__ __align__(16) unsigned char abt_smem[];
  bf16* sP = reinterpret_cast<bf16*>(abt_smem);         // two stages of [P | dS | dO | Q] tiles
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  const int b = blockIdx.y / p.Hkv, hkv = blockIdx.y % p.Hkv, grp = p.Hq / p.Hkv;
  const int J0 = blockIdx.x * 64;
  const int coff = p.Lk - p.Lq;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[nt][e] = 0.f; dv[nt][e] = 0.f; }
  // query tiles that the rows kernel wrote for this key tile (causal: the others are fully masked) form a contiguous range;
  // the (query head of the group, query tile) pairs are walked with cp.async double buffering (4 tiles per stage)
  const int it_begin = p.causal ? max(0, (J0 - 63 - coff + 63) / 64) : 0;
  const int n_it = ws.Lqp / 64 - it_begin, total = grp * max(n_it, 0);
  auto issue = [&](int k) {
    const int hh = k / n_it, it = it_begin + k % n_it;
    const int hq = hkv * grp + hh, bh = b * p.Hq + hq, I0 = it * 64;
    const int nq = max(0, min(64, p.Lq - I0));
    bf16* base = sP + (k & 1) * 4 * TILE;
    load_tile_async(base, ws.P + ((long long)bh * ws.Lqp + I0) * ws.Lkp + J0, ws.Lkp, 64);
    load_tile_async(base + TILE, ws.dS + ((long long)bh * ws.Lqp + I0) * ws.Lkp + J0, ws.Lkp, 64);
    load_tile_async(base + 2 * TILE, static_cast<const bf16*>(p.dO) + (long long)b * p.do_sb + (long long)hq * 64 + (long long)I0 * p.do_sl, p.do_sl, nq);
    load_tile_async(base + 3 * TILE, static_cast<const bf16*>(p.q) + (long long)b * p.q_sb + (long long)hq * 64 + (long long)I0 * p.q_sl, p.q_sl, nq);
    cp_async_commit();
  };
  if (total > 0) issue(0);
  for (int k = 0; k < total; ++k) {
    if (k + 1 < total) { issue(k + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const bf16* base = sP + (k & 1) * 4 * TILE;
    uint32_t a[4][4];
    frag_a_transposed(base, w * 16, lane, a);                         // A[m = key][k = query] = P[query][key]
    mma_a_times_rowmajor(dv, a, base + 2 * TILE, lane);
    frag_a_transposed(base + TILE, w * 16, lane, a);
    mma_a_times_rowmajor(dk, a, base + 3 * TILE, lane);
    __syncthreads();
  }
  const int jlo = J0 + w * 16 + g, jhi = jlo + 8;
  bf16* dKg = reinterpret_cast<bf16*>(p.dk) + (long long)b * p.dkv_sb + (long long)hkv * 64;
  bf16* dVg = reinterpret_cast<bf16*>(p.dv) + (long long)b * p.dkv_sb + (long long)hkv * 64;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * q;
    if (jlo < p.Lk) {
      *reinterpret_cast<uint32_t*>(dKg + (long long)jlo * p.dkv_sl + c) = f2_to_bf16x2(dk[nt][0], dk[nt][1]);
      *reinterpret_cast<uint32_t*>(dVg + (long long)jlo * p.dkv_sl + c) = f2_to_bf16x2(dv[nt][0], dv[nt][1]);
    }
    if (jhi < p.Lk) {
      *reinterpret_cast<uint32_t*>(dKg + (long long)jhi * p.dkv_sl + c) = f2_to_bf16x2(dk[nt][2], dk[nt][3]);
      *reinterpret_cast<uint32_t*>(dVg + (long long)jhi * p.dkv_sl + c) = f2_to_bf16x2(dv[nt][2], dv[nt][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ dEr kernel
// dEr[er_len-1-d] = sum_{b,h,i} dS[i][i-d] q_i.  For the distance block [D0, D0+64) and the query tile I0 the needed keys
// j = i - d lie in the two dS tiles starting at I0 - D0 - 64 and I0 - D0: both go to shared memory side by side
// ([64 rows][128 columns], pitch TP2) and the A operand A[m = d][k = i] = T[i][i - d + 64] is gathered with 16-bit loads.
constexpr int TP2 = 136;
__global__ void __launch_bounds__(THREADS) attn_bwd_der_kernel(AttnBwdParams p, Ws ws, int n_split) {
  v2m::grid_dep_launch();
  v2m::grid_dep_wait();
  extern __shared__ __align__(16) unsigned char abt_smem[];
  constexpr int STAGE = 64 * TP2 + TILE;                  // elements: [dS pair | Q tile]
  bf16* sE = reinterpret_cast<bf16*>(abt_smem);           // two stages
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
  const int D0 = blockIdx.x * 64;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  // (batch*head, query tile) pairs of this split, rows i >= d only, cp.async double buffering
  const int it_begin = D0 / 64, n_it = ws.Lqp / 64 - it_begin;
  const int n_bh = (p.B * p.Hq - (int)blockIdx.y + n_split - 1) / n_split, total = n_bh * n_it;
  auto issue = [&](int k) {
    const int bh = blockIdx.y + (k / n_it) * n_split, I0 = (it_begin + k % n_it) * 64;
    const int b = bh / p.Hq, hq = bh % p.Hq;
    const int nq = max(0, min(64, p.Lq - I0));
    bf16* base = sE + (k & 1) * STAGE;
    const int Ja = I0 - D0 - 64;                           // multiple of 64; negative only when I0 == D0 (keys j < 0: zeros)
    const bf16* src = ws.dS + ((long long)bh * ws.Lqp + I0) * ws.Lkp;
    for (int idx = threadIdx.x; idx < 64 * 16; idx += THREADS) {
      const int r = idx >> 4, c = idx & 15;                // 16 chunks of 8 columns per row
      const bool ok = Ja >= 0 || c >= 8;
      cp_async16(base + r * TP2 + c * 8, src + (ok ? (long long)r * ws.Lkp + Ja + c * 8 : 0), ok);
    }
    load_tile_async(base + 64 * TP2, static_cast<const bf16*>(p.q) + (long long)b * p.q_sb + (long long)hq * 64 + (long long)I0 * p.q_sl, p.q_sl, nq);
    cp_async_commit();
  };
  if (total > 0) issue(0);
  for (int k = 0; k < total; ++k) {
    if (k + 1 < total) { issue(k + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const unsigned short* T = reinterpret_cast<const unsigned short*>(sE + (k & 1) * STAGE);
    uint32_t a[4][4];
    const int m_lo = w * 16 + g, m_hi = m_lo + 8;           // distance rows of this thread's fragments
    auto pair = [&](int m, int kk) -> uint32_t {            // A[m][kk], A[m][kk+1] packed
      const uint32_t lo = T[kk * TP2 + kk - m + 64], hi = T[(kk + 1) * TP2 + kk + 1 - m + 64];
      return lo | (hi << 16);
    };
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int kk = ks * 16 + 2 * q;
      a[ks][0] = pair(m_lo, kk); a[ks][1] = pair(m_hi, kk); a[ks][2] = pair(m_lo, kk + 8); a[ks][3] = pair(m_hi, kk + 8);
    }
    mma_a_times_rowmajor(acc, a, sE + (k & 1) * STAGE + 64 * TP2, lane);
    __syncthreads();
  }
  const int dlo = D0 + w * 16 + g, dhi = dlo + 8;
  const int dlim = min(p.Lq, p.er_len);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * q;
    if (dlo < dlim) {
      atomicAdd(p.dEr + (long long)(p.er_len - 1 - dlo) * 64 + c, acc[nt][0] * p.q_scale);
      atomicAdd(p.dEr + (long long)(p.er_len - 1 - dlo) * 64 + c + 1, acc[nt][1] * p.q_scale);
    }
    if (dhi < dlim) {
      atomicAdd(p.dEr + (long long)(p.er_len - 1 - dhi) * 64 + c, acc[nt][2] * p.q_scale);
      atomicAdd(p.dEr + (long long)(p.er_len - 1 - dhi) * 64 + c + 1, acc[nt][3] * p.q_scale);
    }
  }
}

}  // namespace abt

// Workspace bytes attn_bwd_tc needs for these dimensions.
long long attn_bwd_tc_workspace(int B, int Hq, int Lq, int Lk, int has_er) {
  const long long Lqp = (Lq + 63) / 64 * 64, Lkp = (Lk + 63) / 64 * 64;
  (void)has_er;
  return (long long)B * Hq * Lqp * Lkp * 2 * 2;            // P and dS planes (bf16)
}

// bf16 only, head_dim 64.  p.dk / p.dv are BF16 outputs here (written once, no accumulation; strides dkv_sb / dkv_sl in
// elements), p.dEr fp32 [er_len, 64] must be zeroed by the caller.  ws: attn_bwd_tc_workspace() bytes of device memory.
int attn_bwd_tc(const AttnBwdParams& p, void* ws_ptr, long long ws_bytes, cudaStream_t stream) {
  V2M_REQUIRE(p.dtype == 1 && p.dh == 64, "attn_bwd_tc: bf16 with head_dim 64 only (dtype %d, dh %d)", p.dtype, p.dh);
  V2M_REQUIRE(p.B > 0 && p.Hq > 0 && p.Hkv > 0 && p.Hq % p.Hkv == 0, "attn_bwd_tc: bad heads %d/%d", p.Hq, p.Hkv);
  V2M_REQUIRE(p.Lq > 0 && p.Lk > 0, "attn_bwd_tc: empty sequence");
  V2M_REQUIRE(p.q_scale == 1.0f, "attn_bwd_tc: q must be pre-scaled (q_scale 1)");
  const bool has_er = p.Er != nullptr;
  V2M_REQUIRE(!has_er || (p.Lq == p.Lk && p.Lq <= p.er_len && p.dEr), "attn_bwd_tc: RPR needs Lq == Lk <= er_len and a dEr buffer");
  V2M_REQUIRE(p.q_sl % 8 == 0 && p.k_sl % 8 == 0 && p.v_sl % 8 == 0 && p.o_sl % 8 == 0 && p.do_sl % 8 == 0 && p.q_sb % 8 == 0 &&
              p.k_sb % 8 == 0 && p.v_sb % 8 == 0 && p.o_sb % 8 == 0 && p.do_sb % 8 == 0,
              "attn_bwd_tc: operand strides must be multiples of 8 elements (16-byte rows)");
  V2M_REQUIRE(p.dq_sl % 2 == 0 && p.dkv_sl % 2 == 0 && p.dq_sb % 2 == 0 && p.dkv_sb % 2 == 0, "attn_bwd_tc: odd gradient strides");
  const long long need = attn_bwd_tc_workspace(p.B, p.Hq, p.Lq, p.Lk, has_er);
  V2M_REQUIRE(ws_ptr && ws_bytes >= need, "attn_bwd_tc: workspace of %lld B needed, %lld given", need, ws_bytes);
  // stock attentions (encoder self-attention, decoder cross-attention): one tcgen05 kernel, nothing through HBM
  if (attn_bwd_tc5_supported(p)) return attn_bwd_tc5(p, ws_ptr, ws_bytes, stream);
  abt::Ws ws;
  ws.Lqp = (p.Lq + 63) / 64 * 64;
  ws.Lkp = (p.Lk + 63) / 64 * 64;
  const long long plane = (long long)p.B * p.Hq * ws.Lqp * ws.Lkp;
  ws.P = static_cast<bf16*>(ws_ptr);
  ws.dS = ws.P + plane;
  ws.dQE = nullptr;
  const size_t tiles4 = 4 * abt::TILE * sizeof(bf16);
  const size_t smem_plain = 6 * abt::TILE * sizeof(bf16) + 4 * abt::WBUF;
  const size_t smem_er = 4 * abt::TILE * sizeof(bf16) + 4 * abt::WBUF + (size_t)ws.Lkp * abt::TP * 2;
  const size_t smem_der = 2 * (64 * abt::TP2 + abt::TILE) * sizeof(bf16);
  V2M_REQUIRE(!has_er || smem_er <= 227 * 1024, "attn_bwd_tc: L=%d needs %zu B of shared memory with RPR (> 227 KB)", p.Lk, smem_er);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(abt::attn_bwd_rows_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(abt::attn_bwd_rows_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(abt::attn_bwd_rows_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plain);
    cudaFuncSetAttribute(abt::attn_bwd_rows_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plain);
    cudaFuncSetAttribute(abt::attn_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * tiles4));
    cudaFuncSetAttribute(abt::attn_bwd_der_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_der);
    attr_set = true;
  }
  dim3 grid_r(ws.Lqp / 64, p.B * p.Hq);
  const bool drop = p.drop_scale != 0.f;
  if (has_er) {
    if (drop) launch_dep(abt::attn_bwd_rows_kernel<true, true>, grid_r, dim3(abt::THREADS), smem_er, stream, p, ws);
    else launch_dep(abt::attn_bwd_rows_kernel<true, false>, grid_r, dim3(abt::THREADS), smem_er, stream, p, ws);
  } else {
    if (drop) launch_dep(abt::attn_bwd_rows_kernel<false, true>, grid_r, dim3(abt::THREADS), smem_plain, stream, p, ws);
    else launch_dep(abt::attn_bwd_rows_kernel<false, false>, grid_r, dim3(abt::THREADS), smem_plain, stream, p, ws);
  }
  int rc = check_launch("attn_bwd_rows");
  if (rc) return rc;
  dim3 grid_c(ws.Lkp / 64, p.B * p.Hkv);
  launch_dep(abt::attn_bwd_cols_kernel, grid_c, dim3(abt::THREADS), 2 * tiles4, stream, p, ws);
  rc = check_launch("attn_bwd_cols");
  if (rc || !has_er) return rc;
  const int n_split = p.B * p.Hq < 148 ? p.B * p.Hq : 148;
  dim3 grid_e(ws.Lkp / 64, n_split);
  launch_dep(abt::attn_bwd_der_kernel, grid_e, dim3(abt::THREADS), smem_der, stream, p, ws, n_split);
  return check_launch("attn_bwd_der");
}

}  // namespace v2m
