// Selective-scan recurrence  H[t] = A[t] * H[t-1] + X[t]  along L of (B, L, D, N) fp32 tensors,
// forward and backward -- replaces PScan.forward / PScan.backward (model/pscan.py:154-226), i.e. the
// Blelloch up/down sweep the reference runs as ~2*log2(L) strided in-place PyTorch kernels over
// tensors padded to the next power of two (pscan.py:15-35,170-176).  No padding is needed here.
//
// HBM-bound (12 B per element forward, 20 B backward), so the layout is all about coalescing:
// the (D, N) plane is contiguous, so a lane owns VEC consecutive channels (a warp touches
// 128*VEC contiguous bytes per time step) and the L axis is split into one chunk per warp:
//   1. every warp scans its chunk with a zero carry-in and records the chunk aggregate
//      (prod A, local H_end);
//   2. the per-chunk aggregates are combined by an inclusive warp-shuffle scan with the operator
//      (P2,h2) o (P1,h1) = (P1*P2, P2*h1 + h2), lane <-> chunk, giving every chunk's carry-in;
//   3. the chunk is replayed with its carry-in and H is written.
// Short sequences (chunk <= 32 steps, e.g. L = 300 with 10 warps) keep the chunk's A and X in
// registers between 1 and 3, so each input byte is read from HBM exactly once; long sequences
// (L = 4096) re-read the chunk, which is served by L2 (a CTA's working set is L * 256 * VEC bytes).
//
// Backward (pscan.py:191-226): gX[t] = gH[t] + A[t+1] * gX[t+1] is the same recurrence run from the
// end with A shifted by one step (:218), and gA[t] = H[t-1] * gX[t], gA[0] = 0 (:223-224).
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cstdlib>

namespace v2m {

constexpr int PS_MAXREG = 32;   // chunk length kept in registers

template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float type; };
template <> struct VecT<4> { typedef float4 type; };

template <int VEC> __device__ __forceinline__ void ldv(const float* p, float* o);
template <> __device__ __forceinline__ void ldv<1>(const float* p, float* o) { o[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<4>(const float* p, float* o) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <int VEC> __device__ __forceinline__ void stv(float* p, const float* o);
template <> __device__ __forceinline__ void stv<1>(float* p, const float* o) { *p = o[0]; }
template <> __device__ __forceinline__ void stv<4>(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

// BWD == false: a(t) = A[t],   x(t) = X[t],  out H[t]
// BWD == true : runs over u = L-1-t; a(u) = A[t+1] (0 at t = L-1), x(u) = gH[t]; out gX[t], gA[t] = H[t-1]*gX[t]
template <int VEC, bool REG, bool BWD>
__global__ void __launch_bounds__(REG ? 640 : 1024) pscan_kernel(const float* __restrict__ A, const float* __restrict__ X,
                                                     const float* __restrict__ Hprev, float* __restrict__ out,
                                                     float* __restrict__ out_gA, int L, int DN, int chunk) {
  extern __shared__ float sm[];                 // [2][nw][32*VEC]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cpb = 32 * VEC;                     // channels per block
  const int blocks_per_batch = DN / cpb;
  const int b = blockIdx.x / blocks_per_batch;
  const int c0 = (blockIdx.x % blocks_per_batch) * cpb + lane * VEC;
  const size_t base = (size_t)b * L * DN + c0;
  const int u0 = warp * chunk, u1 = min(L, u0 + chunk);

  float ra[REG ? PS_MAXREG : 1][VEC], rx[REG ? PS_MAXREG : 1][VEC];
  float P[VEC], h[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { P[e] = 1.f; h[e] = 0.f; }

  auto load_ax = [&](int u, float* a, float* x) {
    if (!BWD) {
      ldv<VEC>(A + base + (size_t)u * DN, a);
      ldv<VEC>(X + base + (size_t)u * DN, x);
    } else {
      const int t = L - 1 - u;
      if (t + 1 < L) ldv<VEC>(A + base + (size_t)(t + 1) * DN, a);
      else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] = 0.f;
      }
      ldv<VEC>(X + base + (size_t)t * DN, x);
    }
  };

  // ---- 1. local scan of the chunk (zero carry-in)
  if (REG) {
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) load_ax(u0 + i, ra[i], rx[i]);
    }
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { h[e] = fmaf(ra[i][e], h[e], rx[i][e]); P[e] *= ra[i][e]; }
      }
    }
  } else {
#pragma unroll 4
    for (int u = u0; u < u1; ++u) {
      float a[VEC], x[VEC];
      load_ax(u, a, x);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { h[e] = fmaf(a[e], h[e], x[e]); P[e] *= a[e]; }
    }
  }
  float* sP = sm;
  float* sH = sm + nw * cpb;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    sP[warp * cpb + lane * VEC + e] = P[e];
    sH[warp * cpb + lane * VEC + e] = h[e];
  }
  __syncthreads();
  // ---- 2. scan of the chunk aggregates: lane <-> chunk, one channel at a time per warp
  for (int ch = warp; ch < cpb; ch += nw) {
    float p = lane < nw ? sP[lane * cpb + ch] : 1.f;
    float g = lane < nw ? sH[lane * cpb + ch] : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float pp = __shfl_up_sync(0xffffffffu, p, o);
      const float gp = __shfl_up_sync(0xffffffffu, g, o);
      if (lane >= o) { g = fmaf(p, gp, g); p *= pp; }
    }
    // exclusive carry-in of chunk `lane` = inclusive result of chunk lane-1
    const float carry = __shfl_up_sync(0xffffffffu, g, 1);
    if (lane < nw) sH[lane * cpb + ch] = lane == 0 ? 0.f : carry;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < VEC; ++e) h[e] = sH[warp * cpb + lane * VEC + e];
  // ---- 3. replay with the carry-in and write
  auto emit = [&](int u, const float* hv) {
    if (!BWD) {
      stv<VEC>(out + base + (size_t)u * DN, hv);
    } else {
      const int t = L - 1 - u;
      stv<VEC>(out + base + (size_t)t * DN, hv);
      float ga[VEC];
      if (t >= 1) {
        float hp[VEC];
        ldv<VEC>(Hprev + base + (size_t)(t - 1) * DN, hp);
#pragma unroll
        for (int e = 0; e < VEC; ++e) ga[e] = hp[e] * hv[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) ga[e] = 0.f;
      }
      stv<VEC>(out_gA + base + (size_t)t * DN, ga);
    }
  };
  if (REG) {
#pragma unroll
    for (int i = 0; i < PS_MAXREG; ++i) {
      if (u0 + i < u1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) h[e] = fmaf(ra[i][e], h[e], rx[i][e]);
        emit(u0 + i, h);
      }
    }
  } else {
#pragma unroll 4
    for (int u = u0; u < u1; ++u) {
      float a[VEC], x[VEC];
      load_ax(u, a, x);
#pragma unroll
      for (int e = 0; e < VEC; ++e) h[e] = fmaf(a[e], h[e], x[e]);
      emit(u, h);
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Streaming version (the one that runs whenever the pointers are 16-byte aligned): every global access goes through the
// TMA engine, the compute warps only touch shared memory.
//
//   work item   = (batch b, block of CPI = 32 | 64 channels); persistent CTAs (one per SM) walk the items with stride gridDim.x
//   chunk       = Lc consecutive time steps of the item: one 3-D TMA box (CPI channels x Lc steps x 1 batch) per operand,
//                 landing in a ring of S stages; steps outside [0, L) are zero-filled by the TMA unit and clipped on the
//                 way out, which is exactly the boundary of the recurrence (a = 0, x = 0) and of the backward's one-step
//                 shifts (A[t+1] at t = L-1, H[t-1] at t = 0): no edge code
//   load threads (one per operand) -> full[s] -> 8 compute warps (warp = Lc/8 steps, lane = channel(s): local scan, aggregates exchanged through
//                 shared memory behind ONE named barrier per chunk, carry kept in registers across the chunks of an item,
//                 results written IN PLACE over X / gH / H) -> done[s] -> store threads (TMA store, bulk group) -> empty[s]
//   so loads of chunk q+S-1, compute of chunk q and the store of chunk q-1 overlap, and each byte crosses HBM once whatever L.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(BWD ? 416 : 352, 1)
pscan_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmHp, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmOutA, int L, int Lc, int n_chunks, int S, int n_items,
                 int blocks_per_batch) {
  constexpr int CPI = 32 * VEC;
  constexpr int NIN = BWD ? 3 : 2;
  constexpr int CW = 8;                          // compute warps
  constexpr int MAXSPW = 16 / VEC;               // steps per warp and chunk (Lc <= 128 / VEC)
  extern __shared__ unsigned char ps_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ps_raw) + 127) & ~uintptr_t(127));
  const uint32_t arr_bytes = (uint32_t)Lc * CPI * 4u;
  const uint32_t stage_bytes = NIN * arr_bytes;
  float* aggr = reinterpret_cast<float*>(sm + (size_t)S * stage_bytes);          // [2 parities][CW][2][CPI]
  uint64_t* bars = reinterpret_cast<uint64_t*>(aggr + 2 * CW * 2 * CPI);
  uint64_t* full = bars;
  uint64_t* done = bars + S;
  uint64_t* empty = bars + 2 * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, NIN); mbar_init(done + s, CW); mbar_init(empty + s, BWD ? 2 : 1); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmOut);
    if (BWD) { tma_prefetch_desc(&tmHp); tma_prefetch_desc(&tmOutA); }
  }
  __syncthreads();
  const int my_items = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int Q = my_items * n_chunks;             // chunks this CTA streams, in processing order
  // chunk q -> item, first time step t0 (the backward walks the chunks of an item from the end)
  auto where = [&](int q, int& b, int& c0, int& t0) {
    const int it = (int)blockIdx.x + (q / n_chunks) * (int)gridDim.x;
    const int k = q % n_chunks;
    b = it / blocks_per_batch;
    c0 = (it % blocks_per_batch) * CPI;
    t0 = (BWD ? n_chunks - 1 - k : k) * Lc;
  };

  if (warp >= CW && warp < CW + NIN) {           // ---- load threads: one per operand (a thread issues one TMA op per ~0.5 us)
    if (lane == 0) {
      const int which = warp - CW;               // 0: A (shifted by +1 in the backward), 1: X / gH, 2: H (shifted by -1)
      const CUtensorMap* tm = which == 0 ? &tmA : which == 1 ? &tmX : &tmHp;
      const int shift = BWD ? (which == 0 ? 1 : which == 2 ? -1 : 0) : 0;
      for (int q = 0; q < Q; ++q) {
        const int s = q % S, use = q / S;
        if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
        int b, c0, t0;
        where(q, b, c0, t0);
        mbar_arrive_expect_tx(full + s, arr_bytes);
        tma_load_3d(sm + (size_t)s * stage_bytes + (size_t)which * arr_bytes, tm, c0, t0 + shift, b, full + s);
      }
    }
  } else if (warp >= CW + NIN) {                 // ---- store threads: one per output
    if (lane == 0) {
      const int which = warp - CW - NIN;         // 0: H / gX (over the X buffer), 1: gA (over the H buffer)
      const CUtensorMap* tm = which == 0 ? &tmOut : &tmOutA;
      for (int q = 0; q < Q; ++q) {
        const int s = q % S;
        mbar_wait(done + s, (q / S) & 1);
        int b, c0, t0;
        where(q, b, c0, t0);
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(tm), "r"(smem_u32(sm + (size_t)s * stage_bytes + (size_t)(1 + which) * arr_bytes)), "r"(c0), "r"(t0), "r"(b)
                     : "memory");
        tma_store_commit();
        if (q > 0) {                             // the store of chunk q-1 has read its stage: hand it back to the loaders
          tma_store_wait_read<1>();              // (waiting for the newest store instead, or freeing the A buffer at `done`, measured slower)
          mbar_arrive(empty + (q - 1) % S);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {                                       // ---- compute warps
    const int spw = Lc / CW;                     // Lc is a multiple of 8
    float carry[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) carry[e] = 0.f;
    for (int q = 0; q < Q; ++q) {
      const int s = q % S;
      if (q % n_chunks == 0) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) carry[e] = 0.f;
      }
      float* sa = reinterpret_cast<float*>(sm + (size_t)s * stage_bytes);
      float* sx = sa + (size_t)Lc * CPI;
      float* shp = sx + (size_t)Lc * CPI;
      mbar_wait(full + s, (q / S) & 1);
      // scan step i of this warp <-> buffer row (the backward runs the rows of the chunk from the last to the first)
      const int i0 = warp * spw;
      float ra[MAXSPW][VEC], rx[MAXSPW][VEC];
      float P[VEC], h[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { P[e] = 1.f; h[e] = 0.f; }
#pragma unroll
      for (int i = 0; i < MAXSPW; ++i) {
        if (i < spw) {
          const int r = BWD ? Lc - 1 - (i0 + i) : i0 + i;
          if (VEC == 1) {
            ra[i][0] = sa[r * CPI + lane];
            rx[i][0] = sx[r * CPI + lane];
          } else {
            const float2 va = *reinterpret_cast<const float2*>(sa + r * CPI + lane * 2);
            const float2 vx = *reinterpret_cast<const float2*>(sx + r * CPI + lane * 2);
            ra[i][0] = va.x; ra[i][VEC - 1] = va.y;
            rx[i][0] = vx.x; rx[i][VEC - 1] = vx.y;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < MAXSPW; ++i) {
        if (i < spw) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) { h[e] = fmaf(ra[i][e], h[e], rx[i][e]); P[e] *= ra[i][e]; }
        }
      }
      float* ag = aggr + (size_t)(q & 1) * CW * 2 * CPI;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        ag[(warp * 2 + 0) * CPI + lane * VEC + e] = P[e];
        ag[(warp * 2 + 1) * CPI + lane * VEC + e] = h[e];
      }
      named_bar_sync(1, 32 * CW);
      // every warp folds the aggregates in the same order: its own carry-in on the way, the chunk's end state at the end
#pragma unroll
      for (int e = 0; e < VEC; ++e) h[e] = carry[e];
#pragma unroll
      for (int w2 = 0; w2 < CW; ++w2) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (w2 == warp) h[e] = carry[e];
          carry[e] = fmaf(ag[(w2 * 2 + 0) * CPI + lane * VEC + e], carry[e], ag[(w2 * 2 + 1) * CPI + lane * VEC + e]);
        }
      }
#pragma unroll
      for (int i = 0; i < MAXSPW; ++i) {
        if (i < spw) {
          const int r = BWD ? Lc - 1 - (i0 + i) : i0 + i;
#pragma unroll
          for (int e = 0; e < VEC; ++e) h[e] = fmaf(ra[i][e], h[e], rx[i][e]);
          if (VEC == 1) {
            sx[r * CPI + lane] = h[0];
            if (BWD) shp[r * CPI + lane] *= h[0];
          } else {
            *reinterpret_cast<float2*>(sx + r * CPI + lane * 2) = make_float2(h[0], h[VEC - 1]);
            if (BWD) {
              float2 hp = *reinterpret_cast<const float2*>(shp + r * CPI + lane * 2);
              hp.x *= h[0]; hp.y *= h[VEC - 1];
              *reinterpret_cast<float2*>(shp + r * CPI + lane * 2) = hp;
            }
          }
        }
      }
      fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) mbar_arrive(done + s);
    }
  }
}

typedef CUresult (*PsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (channel, step, batch) fp32 view of a contiguous (B, L, DN) tensor; box = cpi channels x lc steps x 1 batch, no swizzle.
static int make_tmap_scan(CUtensorMap* tm, const float* base, int B, int L, int DN, int cpi, int lc) {
  static PsEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PsEncodeFn>(p);
  }
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not found"); return kCudaError; }
  cuuint64_t dims[3] = {(cuuint64_t)DN, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)DN * 4, (cuuint64_t)L * DN * 4};
  cuuint32_t box[3] = {(cuuint32_t)cpi, (cuuint32_t)lc, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(scan) failed (%d)", (int)r); return kCudaError; }
  return kOk;
}

template <bool BWD>
static int pscan_launch_tma(const float* A, const float* X, const float* Hprev, float* out, float* out_gA, int B, int L,
                            int DN, cudaStream_t stream) {
  static int num_sms = 0, cpi_env = 0, lc_env = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (const char* e = getenv("V2M_PSCAN_CPI")) cpi_env = atoi(e);     // A/B measurements only
    if (const char* e = getenv("V2M_PSCAN_LC")) lc_env = atoi(e);
  }
  // 64 channels per item (256-byte rows) when that still leaves >= 8 items per SM, else 32
  int cpi = (DN % 64 == 0 && (long long)B * (DN / 64) >= 8ll * num_sms) ? 64 : 32;
  if (cpi_env == 32 || (cpi_env == 64 && DN % 64 == 0)) cpi = cpi_env;
  const int lc_cap = cpi == 64 ? 64 : 128;                              // 16 / VEC steps per compute warp in registers
  const int lc_max = lc_env >= 8 && lc_env <= lc_cap ? (lc_env & ~7) : lc_cap;
  const int n_chunks = (L + lc_max - 1) / lc_max;
  const int Lc = (((L + n_chunks - 1) / n_chunks) + 7) & ~7;            // equal chunks, multiple of the 8 compute warps
  const int nin = BWD ? 3 : 2;
  const int stage = nin * Lc * cpi * 4;
  int S = (200 * 1024) / stage;
  if (S > 8) S = 8;
  if (S > n_chunks * 64) S = n_chunks * 64;
  V2M_REQUIRE(S >= 2, "pscan: stage of %d bytes does not fit twice", stage);
  const size_t smem = 128 + (size_t)S * stage + 2 * 8 * 2 * cpi * 4 + 3 * S * 8 + 64;
  const int bpb = DN / cpi;
  const long long items = (long long)B * bpb;
  V2M_REQUIRE(items < (1ll << 30), "pscan: too many items");
  CUtensorMap tmA, tmX, tmHp, tmOut, tmOutA;
  int rc;
  if ((rc = make_tmap_scan(&tmA, A, B, L, DN, cpi, Lc))) return rc;
  if ((rc = make_tmap_scan(&tmX, X, B, L, DN, cpi, Lc))) return rc;
  if ((rc = make_tmap_scan(&tmOut, out, B, L, DN, cpi, Lc))) return rc;
  if (BWD) {
    if ((rc = make_tmap_scan(&tmHp, Hprev, B, L, DN, cpi, Lc))) return rc;
    if ((rc = make_tmap_scan(&tmOutA, out_gA, B, L, DN, cpi, Lc))) return rc;
  } else {
    tmHp = tmA; tmOutA = tmOut;
  }
  const int grid = items < num_sms ? (int)items : num_sms;
#define PS_TMA(V)                                                                                                     \
  do {                                                                                                                \
    static bool attr = false;                                                                                         \
    if (!attr) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(pscan_tma_kernel<V, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) { set_last_error("pscan: smem attribute: %s", cudaGetErrorString(e)); return kCudaError; } \
      attr = true;                                                                                                    \
    }                                                                                                                 \
    pscan_tma_kernel<V, BWD><<<grid, BWD ? 416 : 352, smem, stream>>>(tmA, tmX, tmHp, tmOut, tmOutA, L, Lc, n_chunks, S, (int)items, bpb); \
  } while (0)
  if (cpi == 64) PS_TMA(2); else PS_TMA(1);
#undef PS_TMA
  return check_launch("pscan(tma)");
}

template <bool BWD>
static int pscan_launch(const float* A, const float* X, const float* Hprev, float* out, float* out_gA, int B, int L, int D,
                        int N, cudaStream_t stream) {
  V2M_REQUIRE(B >= 0 && L >= 1 && D >= 1 && N >= 1, "pscan: bad shape (%d,%d,%d,%d)", B, L, D, N);
  if (B == 0) return kOk;
  const int DN = D * N;
  V2M_REQUIRE(DN % 32 == 0, "pscan: D*N=%d must be a multiple of 32", DN);
  {
    static int use_tma = -1;
    if (use_tma < 0) { const char* e = getenv("V2M_PSCAN_TMA"); use_tma = (e && atoi(e) == 0) ? 0 : 1; }
    const bool aligned = (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                         (!BWD || (reinterpret_cast<uintptr_t>(Hprev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_gA) % 16 == 0));
    if (use_tma && aligned) return pscan_launch_tma<BWD>(A, X, Hprev, out, out_gA, B, L, DN, stream);
  }
  // Short sequences (L <= 640): one channel per lane, chunk (<= 32 steps) resident in registers, <= 20 warps.
  // Long sequences: two-pass chunks; float4 channels per lane when that still fills 148 SMs twice over.
  int nw = (L + PS_MAXREG - 1) / PS_MAXREG;
  const bool reg = nw <= 20;
  const bool vec4 = !reg && (DN % 128 == 0) && ((long long)B * (DN / 128) >= 2 * 148) &&
                    (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                    (!BWD || (reinterpret_cast<uintptr_t>(Hprev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_gA) % 16 == 0));
  const int vec = vec4 ? 4 : 1;
  if (!reg) nw = 32;
  if (nw < 1) nw = 1;
  const int chunk = (L + nw - 1) / nw;
  const int grid = B * (DN / (32 * vec));
  const size_t smem = sizeof(float) * 2 * nw * 32 * vec;
#define PS_GO(V, R) pscan_kernel<V, R, BWD><<<grid, nw * 32, smem, stream>>>(A, X, Hprev, out, out_gA, L, DN, chunk)
  if (vec4) PS_GO(4, false);
  else if (reg) PS_GO(1, true);
  else PS_GO(1, false);
#undef PS_GO
  return check_launch("pscan");
}

int pscan_fwd(const float* A, const float* X, float* H, int B, int L, int D, int N, cudaStream_t stream) {
  return pscan_launch<false>(A, X, nullptr, H, nullptr, B, L, D, N, stream);
}

int pscan_bwd(const float* A, const float* H, const float* gH, float* gA, float* gX, int B, int L, int D, int N,
              cudaStream_t stream) {
  return pscan_launch<true>(A, gH, H, gX, gA, B, L, D, N, stream);
}

}  // namespace v2m
