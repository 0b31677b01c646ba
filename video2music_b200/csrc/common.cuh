// Common device helpers for the v2m_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
// host pass / other passes: nothing
#endif

namespace v2m {

typedef __nv_bfloat16 bf16;

// ---- status codes of the C ABI (include/v2m_b200.h) -------------------------------------------
enum Status : int {
  kOk = 0,
  kBadArg = 1,
  kCudaError = 2,
  kUnsupported = 3,
  kNoDevice = 4,
};

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

#define V2M_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      v2m::set_last_error(__VA_ARGS__);                \
      return v2m::kBadArg;                             \
    }                                                  \
  } while (0)

// ---- programmatic dependent launch (training / forward chain) ---------------------------------
// A kernel launched through launch_dep() may be scheduled while its predecessor in the stream is still draining: its CTAs start
// on SMs the predecessor has left, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and block in
// grid_dep_wait() until the predecessor has completed and flushed -- launch latency and the predecessor's tail are hidden.
// Rules: grid_dep_wait() precedes EVERY global-memory access except kernel parameters; grid_dep_launch() (early trigger) may be
// called at the very top because dependents guard themselves.  Without the launch attribute both are no-ops.
// Measured on the training step (profiles/r02_train_pdl_ab.txt): eager launches 13.58 -> 13.15 ms at 64 videos (launch gaps
// hidden), but NO gain when the step is replayed from a CUDA graph (12.5 ms either way; 77.0 vs 79.6 ms at 512 videos), which
// is how the trainer runs -- so the attribute is opt-in (V2M_PDL=1) and the default launch is a plain one.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool dep_launch_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = dep_launch_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- small device utilities ------------------------------------------------------------------
// Dropout masks are a stateless function of (seed, row, column): the forward kernel and the backward kernel that needs the
// same mask recompute it instead of storing it.  One 32-bit hash (lowbias32-style integer mix) of (seed, row, column / 4)
// serves four neighbouring columns, one byte each: keep <=> byte >= thresh8, thresh8 = round(p * 256) -- the drop
// probability is quantised to 1/256 and the kept values are scaled by 256 / (256 - thresh8).
__device__ __forceinline__ uint32_t drop_hash4(uint32_t seed, uint32_t row, uint32_t col4) {
  uint32_t h = seed ^ (row * 0x9E3779B1u);
  h ^= col4 * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool drop_keep_byte(uint32_t h, uint32_t col, uint32_t thresh8) {
  return ((h >> ((col & 3u) * 8u)) & 0xFFu) >= thresh8;
}
__device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t row, uint32_t col, uint32_t thresh8) {
  return drop_keep_byte(drop_hash4(seed, row, col >> 2), col, thresh8);
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// v[0..31]: one partial per lane of 32 DIFFERENT sums -> lane l returns sum #l over the 32 lanes (transpose-reduce butterfly:
// 31 shuffles for 32 sums instead of 5 per sum; fixed summation order)
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte streaming load that does not pollute L1 (KV-cache / weight streams).
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- mbarrier / bulk-copy (TMA engine) wrappers -----------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a lost arrival turns into a trap (reported as a CUDA error by the host wrapper)
// instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("v2m: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// The same bounded wait as ONE asm statement (the spin loop is invisible to the compiler's divergence analysis): used by
// warp-uniform issue loops, where a C++ `while (!try_wait)` makes everything downstream "possibly divergent" and forces
// tcgen05 operands through ELECT / R2UR waterfalls (see umma_kstep4).
__device__ __forceinline__ void mbar_wait_u(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1, P2;\n\t.reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n"
      "V2M_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra V2M_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.gt.u32 P2, cnt, 0x4000000;\n\t"
      "@P2 trap;\n\t"
      "bra V2M_WAIT;\n"
      "V2M_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one elected lane of a converged warp (the predicate ptxas recognises as uniform-safe)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// 1-D bulk copy global -> shared, completion on an mbarrier (bytes multiple of 16, 16B aligned).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 2-D / 3-D tiled TMA loads (tensor map in kernel param / global memory).
// TMA store of one box from shared memory (bulk-group completion): issue, commit, and wait until at most N groups
// of this thread are still READING their shared-memory source.
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tcgen05 (5th-gen tensor core) wrappers ---------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand taken from tensor memory (used for P*V where P was written back with tcgen05.st).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// One k-step of a GEMM main loop: four K = 16 MMAs (descriptors advanced by a_step / b_step in the 16-byte address field), the
// commit that frees the shared-memory slot and, on the last k-step of a tile, the commit that hands the accumulator to the
// epilogue.  LEAN ISSUE (session 5 of round 2): tcgen05 / TMA operands must sit in uniform registers.  In code the compiler
// cannot prove convergent -- `if (lane == 0) { ... }` around C++ `while (!try_wait)` spin loops -- ptxas wraps EVERY such
// instruction in an ELECT / R2UR x 7 / branch "waterfall" (~25 instructions each, whether or not several share one asm
// statement): the issuing thread then spent ~900 clocks per k-step where the tensor pipe needs 512 (ncu source page: the MMA
// warp never waits on a barrier, it executes its ~130-instruction loop).  The cure is control flow, not syntax: the whole warp
// runs the loop, waits with mbar_wait_u (spin loop inside one asm statement) and issues under elect_one(); the operands are
// then computed by the uniform datapath and the MMAs leave back to back (~45 instructions per k-step).
__device__ __forceinline__ void umma_kstep4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc_first,
                                            uint32_t a_step, uint32_t b_step, uint64_t* bar_empty, uint64_t* bar_tfull,
                                            uint32_t last) {
  asm volatile(
      "{\n\t.reg .pred p, t, r;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, sa, sb;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "setp.ne.b32 r, %9, 0;\n\t"
      "cvt.u64.u32 sa, %5;\n\tcvt.u64.u32 sb, %6;\n\t"
      "add.u64 a1, %1, sa;\n\tadd.u64 a2, a1, sa;\n\tadd.u64 a3, a2, sa;\n\t"
      "add.u64 b1, %2, sb;\n\tadd.u64 b2, b1, sb;\n\tadd.u64 b3, b2, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
      "@r tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc_first), "r"(a_step), "r"(b_step), "r"(smem_u32(bar_empty)),
        "r"(smem_u32(bar_tfull)), "r"(last) : "memory");
}
// the same for a CTA pair (issued by the leader; commits multicast to both CTAs)
__device__ __forceinline__ void umma_kstep4_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc_first,
                                                 uint32_t a_step, uint32_t b_step, uint64_t* bar_empty, uint64_t* bar_tfull,
                                                 uint32_t last) {
  asm volatile(
      "{\n\t.reg .pred p, t, r;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, sa, sb;\n\t.reg .b16 m;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "setp.ne.b32 r, %9, 0;\n\t"
      "mov.b16 m, 3;\n\t"
      "cvt.u64.u32 sa, %5;\n\tcvt.u64.u32 sb, %6;\n\t"
      "add.u64 a1, %1, sa;\n\tadd.u64 a2, a1, sa;\n\tadd.u64 a3, a2, sa;\n\t"
      "add.u64 b1, %2, sb;\n\tadd.u64 b2, b1, sb;\n\tadd.u64 b3, b2, sb;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], m;\n\t"
      "@r tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], m;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc_first), "r"(a_step), "r"(b_step), "r"(smem_u32(bar_empty)),
        "r"(smem_u32(bar_tfull)), "r"(last) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster (one TPC) run ONE M = 256 MMA; each holds its 128 rows of A and D and
// half of the N rows of B, so a CTA ingests (128 + N/2) x K operand bytes per tile instead of (128 + N) x K.
__device__ __forceinline__ uint32_t pair_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pair_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pair_mapa(uint32_t addr, uint32_t rank) {     // shared::cluster address of CTA `rank`'s copy
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Relaxed: the arrival only tells the leader's MMA issuer that this warp's tcgen05.ld of the accumulator stage have completed
// (tcgen05.wait::ld + tcgen05.fence::before_thread_sync precede it); no memory is published.  A .release.cluster arrive
// compiles to a cluster-scope memory barrier (ERRBAR) that waited for the warp's outstanding stores: 7.7 % of the pair
// kernel's stall samples (ncu source page, profiles/r02_final_ncu_key_metrics_s5.txt).
__device__ __forceinline__ void pair_mbar_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {         // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
// TMA tile into THIS CTA's shared memory, bytes counted on the barrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}

// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread t <-> TMEM lane base+t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// bar.sync on a named barrier (ids 1..15; 0 is __syncthreads) for `count` threads (multiple of 32)
// The dynamic shared-memory base is re-aligned through an integer cast, after which the compiler no longer knows the address
// space and emits GENERIC loads / stores (LD.E / ST.E through the local/global queue: 'stall_lg' on every scratch store in
// the ncu source page).  The skew scratch and the row-statistics exchange therefore use explicit shared-space accesses.
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major : rows of 128 B (64 bf16 along K); 8-row groups are `sbo_bytes` apart (1024 when dense).
//   MN-major: rows of 128 B (64 bf16 along M/N), one row per k; 8-k groups are `sbo_bytes` apart,
//             64-element M/N chunks are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // D format fp32
         | (1u << 7)                     // A = bf16
         | (1u << 10)                    // B = bf16
         | (uint32_t(a_mn_major) << 15)  // A major
         | (uint32_t(b_mn_major) << 16)  // B major
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

}  // namespace v2m
