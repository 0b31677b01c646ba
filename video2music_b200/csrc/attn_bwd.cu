// Attention backward (SIMT, fp32 math; fp32 or bf16 I/O): dQ, dK, dV and dEr of
//   O = softmax(Q K^T + skew(Q Er^T) + causal) V
// i.e. the gradient of model/rpr.py:387-414 (+ _skew :439-455) that the reference gets from autograd over the
// materialised (B*H, L, L) tensors.  With Srel[i,j] = q_i . Er[er_len-1-(i-j)] (j <= i):
//   P  = exp(S - lse),  D_i = sum_d dO_id O_id,  dS = P o (dO V^T - D)
//   dV = P^T dO,  dK = dS^T Q,  dQ_i = sum_j dS_ij K_j + sum_{j<=i} dS_ij Er[er_len-1-(i-j)]
//   dEr[er_len-1-r] += sum_i dS_{i,i-r} q_i          (reduction over batch, heads and rows: fp32 atomics)
// One CTA = one (batch, head) x 32 query rows.  K (then reused), the reversed Er band / V, the 32-row P/dS block and
// the Q / dO rows live in shared memory (213 KB at L = 300).  dK / dV / dEr are accumulated with fp32 atomics into
// caller-zeroed buffers (dK, dV in the layout [B, Lk, Hkv, dh] = strides of the fp32 gradient buffers).
#include "common.cuh"
#include "kernels.h"

namespace v2m {

constexpr int BRB = 32;
constexpr int BNW = 8;
constexpr int BRPW = BRB / BNW;

__device__ __forceinline__ float ldT(const void* p, int dtype, size_t i) {
  return dtype == 0 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void stT(void* p, int dtype, size_t i, float v) {
  if (dtype == 0) static_cast<float*>(p)[i] = v;
  else static_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

template <int DH>
__global__ void __launch_bounds__(BNW * 32) attn_bwd_kernel(AttnBwdParams p, int lk_pad) {
  extern __shared__ __align__(16) float bsm[];
  const int KS = DH + 1;
  float* Qs = bsm;                               // [BRB][DH]
  float* dOs = Qs + BRB * DH;                    // [BRB][DH]
  float* Ps = dOs + BRB * DH;                    // [BRB][lk_pad]   P, then dS
  float* Ks = Ps + BRB * lk_pad;                 // [lk_pad][KS]
  float* Xs = Ks + (size_t)lk_pad * KS;          // [lk_pad][KS]    reversed Er band, then V, then Er band again
  __shared__ float Dsh[BRB], lsesh[BRB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bh = blockIdx.y, b = bh / p.Hq, hq = bh % p.Hq, hkv = hq / (p.Hq / p.Hkv);
  const int i0 = blockIdx.x * BRB;
  const int nrows = min(BRB, p.Lq - i0);
  const int coff = p.Lk - p.Lq;
  const int nk = p.causal ? min(p.Lk, i0 + nrows + coff) : p.Lk;
  const int dt = p.dtype;
  const size_t qo = (size_t)b * p.q_sb + (size_t)hq * DH, ko = (size_t)b * p.k_sb + (size_t)hkv * DH;
  const size_t vo = (size_t)b * p.v_sb + (size_t)hkv * DH, oo = (size_t)b * p.o_sb + (size_t)hq * DH;
  const size_t doo = (size_t)b * p.do_sb + (size_t)hq * DH, dqo = (size_t)b * p.dq_sb + (size_t)hq * DH;
  const bool has_er = p.Er != nullptr;
  const int nr = has_er ? min(i0 + nrows, p.er_len) : 0;          // distances [0, nr) are needed
  // the forward's probability dropout: same stateless mask (seed, (b*Hq+h)*Lq + i, j).  With Pd = P o M * s:
  //   dV = Pd^T dO,  dS = P o (M * s * (dO V^T) - D),  D_i = dO_i . O_i  (unchanged, O = Pd V)
  const bool drop_on = p.drop_scale != 0.f;
  const uint32_t dseed = p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0u);
  const uint32_t drow0 = (uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i0;

  auto load_er = [&]() {
    for (int idx = tid; idx < nr * DH; idx += BNW * 32) {
      const int r = idx / DH, d = idx % DH;
      Xs[r * KS + d] = ldT(p.Er, dt, (size_t)(p.er_len - 1 - r) * DH + d);
    }
  };
  for (int idx = tid; idx < BRB * DH; idx += BNW * 32) {
    const int r = idx / DH, d = idx % DH;
    const bool ok = r < nrows;
    Qs[idx] = ok ? ldT(p.q, dt, qo + (size_t)(i0 + r) * p.q_sl + d) * p.q_scale : 0.f;
    dOs[idx] = ok ? ldT(p.dO, dt, doo + (size_t)(i0 + r) * p.do_sl + d) : 0.f;
  }
  for (int idx = tid; idx < nk * DH; idx += BNW * 32) {
    const int j = idx / DH, d = idx % DH;
    Ks[j * KS + d] = ldT(p.k, dt, ko + (size_t)j * p.k_sl + d);
  }
  if (has_er) load_er();
  if (tid < BRB) lsesh[tid] = (tid < nrows) ? p.lse[(size_t)bh * p.Lq + i0 + tid] : 0.f;
  __syncthreads();
  // D_i = sum_d dO_id * O_id  (one warp per 4 rows, lanes over d)
#pragma unroll
  for (int r = 0; r < BRPW; ++r) {
    const int rr = warp * BRPW + r;
    float s = 0.f;
    if (rr < nrows)
      for (int d = lane; d < DH; d += 32) s += dOs[rr * DH + d] * ldT(p.o, dt, oo + (size_t)(i0 + rr) * p.o_sl + d);
    s = warp_sum(s);
    if (lane == 0) Dsh[rr] = s;
  }
  const int r0 = warp * BRPW;
  // ---- phase 1: P = exp(q k^T + Srel - lse), masked
  for (int j = lane; j < nk; j += 32) {
    float acc[BRPW];
#pragma unroll
    for (int r = 0; r < BRPW; ++r) acc[r] = 0.f;
    const float* kr = Ks + j * KS;
#pragma unroll 4
    for (int d = 0; d < DH; d += 4) {
      const float k0 = kr[d], k1 = kr[d + 1], k2 = kr[d + 2], k3 = kr[d + 3];
#pragma unroll
      for (int r = 0; r < BRPW; ++r) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + (r0 + r) * DH + d);
        acc[r] = fmaf(qv.x, k0, acc[r]); acc[r] = fmaf(qv.y, k1, acc[r]);
        acc[r] = fmaf(qv.z, k2, acc[r]); acc[r] = fmaf(qv.w, k3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < BRPW; ++r) Ps[(r0 + r) * lk_pad + j] = acc[r];
  }
  __syncwarp();
  if (has_er) {
    const int rel_end = min(min(i0 + r0 + BRPW, i0 + nrows), p.er_len);
    for (int rel = lane; rel < rel_end; rel += 32) {
      float acc[BRPW];
#pragma unroll
      for (int r = 0; r < BRPW; ++r) acc[r] = 0.f;
      const float* er = Xs + rel * KS;
#pragma unroll 4
      for (int d = 0; d < DH; d += 4) {
        const float e0 = er[d], e1 = er[d + 1], e2 = er[d + 2], e3 = er[d + 3];
#pragma unroll
        for (int r = 0; r < BRPW; ++r) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + (r0 + r) * DH + d);
          acc[r] = fmaf(qv.x, e0, acc[r]); acc[r] = fmaf(qv.y, e1, acc[r]);
          acc[r] = fmaf(qv.z, e2, acc[r]); acc[r] = fmaf(qv.w, e3, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < BRPW; ++r) {
        const int j = i0 + r0 + r - rel;
        if (j >= 0 && j < nk) Ps[(r0 + r) * lk_pad + j] += acc[r];
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < BRPW; ++r) {
    const int i = i0 + r0 + r;
    const int lim = (i < p.Lq) ? (p.causal ? min(p.Lk, i + coff + 1) : p.Lk) : 0;
    float* pr = Ps + (r0 + r) * lk_pad;
    const float l = lsesh[r0 + r];
    for (int j = lane; j < nk; j += 32) pr[j] = (j < lim) ? expf(pr[j] - l) : 0.f;
  }
  __syncthreads();                               // P complete, Er band no longer needed -> load V over it
  for (int idx = tid; idx < nk * DH; idx += BNW * 32) {
    const int j = idx / DH, d = idx % DH;
    Xs[j * KS + d] = ldT(p.v, dt, vo + (size_t)j * p.v_sl + d);
  }
  // ---- phase 2a: dV[j] += sum_i P_ij dO_i   (thread <-> key j = lane + 32*jj, 8 dims d0 = 8*warp)
  {
    const int d0 = warp * (DH / 8);
    for (int j = lane; j < nk; j += 32) {
      float acc[DH / 8];
#pragma unroll
      for (int e = 0; e < DH / 8; ++e) acc[e] = 0.f;
      for (int i = 0; i < nrows; ++i) {
        float pij = Ps[i * lk_pad + j];
        if (drop_on) pij = drop_keep(dseed, drow0 + (uint32_t)i, (uint32_t)j, p.drop_thresh) ? pij * p.drop_scale : 0.f;   // dV uses the dropped P
#pragma unroll
        for (int e = 0; e < DH / 8; ++e) acc[e] = fmaf(pij, dOs[i * DH + d0 + e], acc[e]);
      }
      float* dst = p.dv + (size_t)b * p.dkv_sb + (size_t)j * p.dkv_sl + (size_t)hkv * DH + d0;
#pragma unroll
      for (int e = 0; e < DH / 8; ++e) atomicAdd(dst + e, acc[e]);
    }
  }
  __syncthreads();                               // V staged
  // ---- phase 2b: dS = P o (dO V^T - D)
  for (int j = lane; j < nk; j += 32) {
    float acc[BRPW];
#pragma unroll
    for (int r = 0; r < BRPW; ++r) acc[r] = 0.f;
    const float* vr = Xs + j * KS;
#pragma unroll 4
    for (int d = 0; d < DH; d += 4) {
      const float v0 = vr[d], v1 = vr[d + 1], v2 = vr[d + 2], v3 = vr[d + 3];
#pragma unroll
      for (int r = 0; r < BRPW; ++r) {
        const float4 gv = *reinterpret_cast<const float4*>(dOs + (r0 + r) * DH + d);
        acc[r] = fmaf(gv.x, v0, acc[r]); acc[r] = fmaf(gv.y, v1, acc[r]);
        acc[r] = fmaf(gv.z, v2, acc[r]); acc[r] = fmaf(gv.w, v3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < BRPW; ++r) {
      float* ps = Ps + (r0 + r) * lk_pad + j;
      float dp = acc[r];                            // d(dropped P); d(P) = mask * scale * it
      if (drop_on) dp = drop_keep(dseed, drow0 + (uint32_t)(r0 + r), (uint32_t)j, p.drop_thresh) ? dp * p.drop_scale : 0.f;
      *ps = *ps * (dp - Dsh[r0 + r]);
    }
  }
  __syncthreads();                               // dS complete, V no longer needed -> Er band again
  if (has_er) load_er();
  __syncthreads();
  // ---- phase 3a: dQ_i = sum_j dS_ij K_j + sum_rel dS_{i,i-rel} Erev_rel   (warp rows, lanes over d)
  {
    constexpr int DPL = DH / 32;
    float dq[BRPW][DPL];
#pragma unroll
    for (int r = 0; r < BRPW; ++r)
#pragma unroll
      for (int c = 0; c < DPL; ++c) dq[r][c] = 0.f;
    const int jend = p.causal ? min(nk, i0 + r0 + BRPW + coff) : nk;
    for (int j = 0; j < jend; ++j) {
      float kk[DPL];
#pragma unroll
      for (int c = 0; c < DPL; ++c) kk[c] = Ks[j * KS + lane + 32 * c];
#pragma unroll
      for (int r = 0; r < BRPW; ++r) {
        const float s = Ps[(r0 + r) * lk_pad + j];
#pragma unroll
        for (int c = 0; c < DPL; ++c) dq[r][c] = fmaf(s, kk[c], dq[r][c]);
      }
    }
    if (has_er) {
      const int rel_end = min(min(i0 + r0 + BRPW, i0 + nrows), p.er_len);
      for (int rel = 0; rel < rel_end; ++rel) {
        float ee[DPL];
#pragma unroll
        for (int c = 0; c < DPL; ++c) ee[c] = Xs[rel * KS + lane + 32 * c];
#pragma unroll
        for (int r = 0; r < BRPW; ++r) {
          const int j = i0 + r0 + r - rel;
          const float s = (j >= 0 && j < nk) ? Ps[(r0 + r) * lk_pad + j] : 0.f;
#pragma unroll
          for (int c = 0; c < DPL; ++c) dq[r][c] = fmaf(s, ee[c], dq[r][c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < BRPW; ++r) {
      const int i = i0 + r0 + r;
      if (i < p.Lq) {
#pragma unroll
        for (int c = 0; c < DPL; ++c) stT(p.dq, dt, dqo + (size_t)i * p.dq_sl + lane + 32 * c, dq[r][c] * p.q_scale * (p.dq_scale == 0.f ? 1.f : p.dq_scale));
      }
    }
  }
  // ---- phase 3b: dK[j] += sum_i dS_ij q_i ;  3c: dEr[er_len-1-rel] += sum_i dS_{i,i-rel} q_i
  {
    const int d0 = warp * (DH / 8);
    for (int j = lane; j < nk; j += 32) {
      float acc[DH / 8];
#pragma unroll
      for (int e = 0; e < DH / 8; ++e) acc[e] = 0.f;
      for (int i = 0; i < nrows; ++i) {
        const float s = Ps[i * lk_pad + j];
#pragma unroll
        for (int e = 0; e < DH / 8; ++e) acc[e] = fmaf(s, Qs[i * DH + d0 + e], acc[e]);
      }
      float* dst = p.dk + (size_t)b * p.dkv_sb + (size_t)j * p.dkv_sl + (size_t)hkv * DH + d0;
#pragma unroll
      for (int e = 0; e < DH / 8; ++e) atomicAdd(dst + e, acc[e]);
    }
    if (has_er) {
      for (int rel = lane; rel < nr; rel += 32) {
        float acc[DH / 8];
#pragma unroll
        for (int e = 0; e < DH / 8; ++e) acc[e] = 0.f;
        for (int i = 0; i < nrows; ++i) {
          const int j = i0 + i - rel;
          if (j >= 0 && j < nk) {
            const float s = Ps[i * lk_pad + j];
#pragma unroll
            for (int e = 0; e < DH / 8; ++e) acc[e] = fmaf(s, Qs[i * DH + d0 + e], acc[e]);
          }
        }
        float* dst = p.dEr + (size_t)(p.er_len - 1 - rel) * DH + d0;
#pragma unroll
        for (int e = 0; e < DH / 8; ++e) atomicAdd(dst + e, acc[e]);
      }
    }
  }
}

template <int DH>
static int launch_bwd(const AttnBwdParams& p, cudaStream_t stream) {
  const int lk_pad = (p.Lk + 3) & ~3;
  const size_t smem = sizeof(float) * (2 * (size_t)BRB * DH + (size_t)BRB * lk_pad + 2 * (size_t)lk_pad * (DH + 1));
  V2M_REQUIRE(smem <= 226 * 1024, "attn_bwd: Lk=%d needs %zu B of shared memory (> 226 KB)", p.Lk, smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);  // + 256 B static
    attr_set = true;
  }
  dim3 grid((p.Lq + BRB - 1) / BRB, p.B * p.Hq);
  attn_bwd_kernel<DH><<<grid, BNW * 32, smem, stream>>>(p, lk_pad);
  return check_launch("attn_bwd");
}

int attn_bwd(const AttnBwdParams& p, cudaStream_t stream) {
  V2M_REQUIRE(p.B > 0 && p.Hq > 0 && p.Hkv > 0 && p.Hq % p.Hkv == 0, "attn_bwd: bad heads %d/%d", p.Hq, p.Hkv);
  V2M_REQUIRE(p.Lq > 0 && p.Lk > 0, "attn_bwd: empty sequence");
  V2M_REQUIRE(!p.Er || (p.Lq == p.Lk && p.Lq <= p.er_len && p.dEr), "attn_bwd: RPR needs Lq == Lk <= er_len and a dEr buffer");
  V2M_REQUIRE(p.dtype == 0 || p.dtype == 1, "attn_bwd: dtype %d", p.dtype);
  switch (p.dh) {
    case 32: return launch_bwd<32>(p, stream);
    case 64: return launch_bwd<64>(p, stream);
    default:
      set_last_error("attn_bwd: unsupported head_dim %d (32, 64)", p.dh);
      return kUnsupported;
  }
}

}  // namespace v2m
