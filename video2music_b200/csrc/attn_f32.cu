// fp32 SIMT attention forward (exact-precision path): softmax(q k^T + skew(q Er^T) + causal) v.
//
// Restates model/rpr.py:387-414 (bmm, einsum with Er, _skew, mask, softmax, bmm) for one
// (batch, head) and a block of RB query rows per CTA without materialising the (BH, L, L)
// tensors.  The same kernel serves the stock attentions (Er = null: encoder self-attention and
// decoder cross-attention) and grouped-query attention (Hkv < Hq,
// model/grouped_query_attention.py:122-156).
//
// The skew of rpr.py:439-455 is the closed form Srel[i,j] = q_i . Er[er_len-1-(i-j)] (j <= i):
// the CTA stages the rows Erev[r] = Er[er_len-1-r] in shared memory, computes qe[i][r] for its
// rows exactly like a second "key" pass, and adds qe[i][i-j] into the score row held in shared
// memory -- the skew is an index shift inside shared memory.
#include "common.cuh"
#include "kernels.h"

namespace v2m {

constexpr int RB = 32;       // query rows per CTA
constexpr int NW = 8;        // warps per CTA, 4 rows each
constexpr int RPW = RB / NW;

template <int DH>
__global__ void __launch_bounds__(NW * 32) attn_f32_kernel(AttnParams p, int lk_pad) {
  extern __shared__ __align__(16) float smem[];
  const int KS = DH + 1;                        // padded row stride of the K / Er tiles
  float* Qs = smem;                             // [RB][DH]
  float* Ps = Qs + RB * DH;                     // [RB][lk_pad]
  float* Ks = Ps + RB * lk_pad;                 // [lk_pad][KS]   (re-used for V with stride DH)
  float* Es = Ks + (size_t)lk_pad * KS;         // [lk_pad][KS]   (only when Er)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bh = blockIdx.y, b = bh / p.Hq, hq = bh % p.Hq;
  const int hkv = hq / (p.Hq / p.Hkv);
  const int i0 = blockIdx.x * RB;
  const int nrows = min(RB, p.Lq - i0);
  // causal: query i sees keys j <= i + (Lk - Lq)   (Lq == Lk for self-attention)
  const int Lk = p.lk_dev ? min(p.Lk, __ldg(p.lk_dev)) : p.Lk;     // keys that exist (device-side count: graph replays)
  const int coff = Lk - p.Lq;
  const int nk = p.causal ? min(Lk, i0 + nrows + coff) : Lk;       // keys needed by this CTA

  const float* q = static_cast<const float*>(p.q) + (size_t)b * p.q_sb + (size_t)hq * DH;
  const float* k = static_cast<const float*>(p.k) + (size_t)b * p.k_sb + (size_t)hkv * DH;
  const float* v = static_cast<const float*>(p.v) + (size_t)b * p.v_sb + (size_t)hkv * DH;
  const float* Er = static_cast<const float*>(p.Er);

  for (int idx = tid; idx < RB * DH; idx += NW * 32) {
    const int r = idx / DH, d = idx % DH;
    Qs[idx] = (r < nrows) ? q[(size_t)(i0 + r) * p.q_sl + d] * p.q_scale : 0.f;
  }
  for (int idx = tid; idx < nk * DH; idx += NW * 32) {
    const int j = idx / DH, d = idx % DH;
    Ks[j * KS + d] = k[(size_t)j * p.k_sl + d];
  }
  if (Er) {
    // relative distance r = i - j in [0, i0 + nrows - 1]
    const int nr = min(i0 + nrows, p.er_len);
    for (int idx = tid; idx < nr * DH; idx += NW * 32) {
      const int r = idx / DH, d = idx % DH;
      Es[r * KS + d] = Er[(size_t)(p.er_len - 1 - r) * DH + d];
    }
  }
  __syncthreads();

  const int r0 = warp * RPW;                     // this warp's first row inside the block
  // ---- pass 1: raw scores into Ps -----------------------------------------------------------
  for (int j = lane; j < nk; j += 32) {
    float acc[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) acc[r] = 0.f;
    const float* kr = Ks + j * KS;
#pragma unroll 4
    for (int d = 0; d < DH; d += 4) {
      const float k0 = kr[d], k1 = kr[d + 1], k2 = kr[d + 2], k3 = kr[d + 3];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + (r0 + r) * DH + d);
        acc[r] = fmaf(qv.x, k0, acc[r]);
        acc[r] = fmaf(qv.y, k1, acc[r]);
        acc[r] = fmaf(qv.z, k2, acc[r]);
        acc[r] = fmaf(qv.w, k3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) Ps[(r0 + r) * lk_pad + j] = acc[r];
  }
  __syncwarp();
  // ---- pass 2: relative-position term, skewed into place -----------------------------------
  if (Er) {
    // largest distance this warp needs; only distances [0, i0 + nrows) were staged in Es
    const int rel_end = min(min(i0 + r0 + RPW, i0 + nrows), p.er_len);
    for (int rel = lane; rel < rel_end; rel += 32) {
      float acc[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) acc[r] = 0.f;
      const float* er = Es + rel * KS;
#pragma unroll 4
      for (int d = 0; d < DH; d += 4) {
        const float e0 = er[d], e1 = er[d + 1], e2 = er[d + 2], e3 = er[d + 3];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + (r0 + r) * DH + d);
          acc[r] = fmaf(qv.x, e0, acc[r]);
          acc[r] = fmaf(qv.y, e1, acc[r]);
          acc[r] = fmaf(qv.z, e2, acc[r]);
          acc[r] = fmaf(qv.w, e3, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int i = i0 + r0 + r;
        const int j = i - rel;                   // Srel[i, j] = qe[i][i - j]
        if (j >= 0 && j < nk) Ps[(r0 + r) * lk_pad + j] += acc[r];
      }
    }
    __syncwarp();
  }
  // ---- softmax (row by row, one warp) -------------------------------------------------------
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int i = i0 + r0 + r;
    float* pr = Ps + (r0 + r) * lk_pad;
    const int lim = (i < p.Lq) ? (p.causal ? min(Lk, i + coff + 1) : Lk) : 0;   // keys [0, lim)
    float mx = -INFINITY;
    for (int j = lane; j < lim; j += 32) mx = fmaxf(mx, pr[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < lim; j += 32) {
      const float e = expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = lim > 0 ? 1.f / sum : 0.f;
    if (p.drop_scale != 0.f) {
      // training: inverted dropout of the probabilities (rpr.py:407 / F.multi_head_attention_forward): the same stateless
      // mask (seed, (b*Hq+h)*Lq + i, j) as the tensor-core kernels; the returned weights are the dropped ones, as in torch
      const uint32_t dseed = p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0u);
      const uint32_t drow = (uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i;
      for (int j = lane; j < nk; j += 32)
        pr[j] = (j < lim && drop_keep(dseed, drow, (uint32_t)j, p.drop_thresh)) ? pr[j] * inv * p.drop_scale : 0.f;
    } else {
      for (int j = lane; j < nk; j += 32) pr[j] = (j < lim) ? pr[j] * inv : 0.f;
    }
    if (p.lse && lane == 0 && i < p.Lq) p.lse[(size_t)bh * p.Lq + i] = mx + logf(sum);
    if (p.p_out && i < p.Lq) {
      float* po = p.p_out + ((size_t)bh * p.Lq + i) * p.Lk;
      for (int j = lane; j < p.Lk; j += 32) po[j] = (j < nk) ? pr[j] : 0.f;
    }
  }
  __syncthreads();   // everybody is done with Ks -> overwrite with V
  float* Vs = Ks;    // [nk][DH]
  for (int idx = tid; idx < nk * DH; idx += NW * 32) {
    const int j = idx / DH, d = idx % DH;
    Vs[idx] = v[(size_t)j * p.v_sl + d];
  }
  __syncthreads();
  // ---- O = P V -----------------------------------------------------------------------------
  constexpr int DPL = DH / 32;                   // output dims per lane
  float o[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int c = 0; c < DPL; ++c) o[r][c] = 0.f;
  const int jend = p.causal ? min(nk, i0 + r0 + RPW + coff) : nk;
  for (int j = 0; j < jend; ++j) {
    float vv[DPL];
#pragma unroll
    for (int c = 0; c < DPL; ++c) vv[c] = Vs[j * DH + lane + 32 * c];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float pj = Ps[(r0 + r) * lk_pad + j];
#pragma unroll
      for (int c = 0; c < DPL; ++c) o[r][c] = fmaf(pj, vv[c], o[r][c]);
    }
  }
  float* out = static_cast<float*>(p.o) + (size_t)b * p.o_sb + (size_t)hq * DH;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int i = i0 + r0 + r;
    if (i < p.Lq) {
#pragma unroll
      for (int c = 0; c < DPL; ++c) out[(size_t)i * p.o_sl + lane + 32 * c] = o[r][c];
    }
  }
}

template <int DH>
static int launch(const AttnParams& p, cudaStream_t stream) {
  const int lk_pad = (p.Lk + 3) & ~3;
  size_t smem = sizeof(float) * ((size_t)RB * DH + (size_t)RB * lk_pad + (size_t)lk_pad * (DH + 1) * (p.Er ? 2 : 1));
  V2M_REQUIRE(smem <= 227 * 1024, "attn_fwd_f32: Lk=%d needs %zu B of shared memory (> 227 KB)", p.Lk, smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_f32_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  dim3 grid((p.Lq + RB - 1) / RB, p.B * p.Hq);
  attn_f32_kernel<DH><<<grid, NW * 32, smem, stream>>>(p, lk_pad);
  return check_launch("attn_fwd_f32");
}

int attn_fwd_f32(const AttnParams& p, cudaStream_t stream) {
  V2M_REQUIRE(p.B > 0 && p.Hq > 0 && p.Hkv > 0 && p.Hq % p.Hkv == 0, "attn_fwd_f32: bad heads %d/%d", p.Hq, p.Hkv);
  V2M_REQUIRE(p.Lq > 0 && p.Lk > 0, "attn_fwd_f32: empty sequence Lq=%d Lk=%d", p.Lq, p.Lk);
  V2M_REQUIRE(!p.Er || (p.Lq == p.Lk && p.Lq <= p.er_len), "attn_fwd_f32: RPR needs Lq == Lk <= er_len (%d, %d, %d)",
              p.Lq, p.Lk, p.er_len);
  V2M_REQUIRE(!p.lk_dev || (!p.Er && !p.p_out), "attn_fwd_f32: a device-side key count excludes Er and need_weights");
  switch (p.dh) {
    case 32: return launch<32>(p, stream);
    case 64: return launch<64>(p, stream);
    case 128: return launch<128>(p, stream);
    default:
      set_last_error("attn_fwd_f32: unsupported head_dim %d (32, 64, 128)", p.dh);
      return kUnsupported;
  }
}

}  // namespace v2m
