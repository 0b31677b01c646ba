"""CPU restatement (torch fp32, functional) of the reference's AMT hot path.

TEST INFRASTRUCTURE ONLY -- this file is the *checker*, never the product.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.  video2music_b200/ must not.

Parity pin: the reference (khangklj/Video2Music) ships no tests, golden
vectors or checkpoints, so this restatement is pinned by
(1) tests/test_oracle_vs_reference.py, which runs the unmodified reference
live through oracle/ref_shim.py whenever /root/reference is mounted, and
(2) the fixtures under tests/golden/ that oracle/make_golden.py produced from
the unmodified reference in that same container.

Every function cites the reference lines it restates (paths relative to the
reference root).  Weights are passed as a plain dict with the reference's
state_dict keys.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

CHORD_END, CHORD_PAD, CHORD_SIZE = 157, 158, 159      # utilities/constants.py:50-52
CHORD_ROOT_PAD, CHORD_ATTR_PAD = 14, 15               # utilities/constants.py:55-62

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# RPR attention (model/rpr.py)
# --------------------------------------------------------------------------
def get_valid_embedding(Er: torch.Tensor, len_q: int) -> torch.Tensor:
    """model/rpr.py:426-437."""
    return Er[max(0, Er.shape[0] - len_q):, :]


def skew_literal(qe: torch.Tensor) -> torch.Tensor:
    """model/rpr.py:439-455, step by step (mask, pad one column, reshape, drop row)."""
    sz = qe.shape[1]
    mask = (torch.triu(torch.ones(sz, sz)) == 1).float().flip(0)
    qe = mask * qe
    qe = F.pad(qe, (1, 0, 0, 0, 0, 0))
    qe = torch.reshape(qe, (qe.shape[0], qe.shape[2], qe.shape[1]))
    return qe[:, 1:, :]


def skew_closed_form(q: torch.Tensor, Er_valid: torch.Tensor) -> torch.Tensor:
    """Closed form of rpr.py:393-395: Srel[h,i,j] = q[h,i] . Er_valid[Lv-1-(i-j)] for
    j <= i and 0 for j > i (what the CUDA kernels implement)."""
    H, L, _ = q.shape
    Lv = Er_valid.shape[0]
    i = torch.arange(L).view(L, 1)
    j = torch.arange(L).view(1, L)
    idx = (Lv - 1 - (i - j)).clamp(0, Lv - 1)
    e = Er_valid[idx]                                    # (L, L, dh)
    s = torch.einsum("hid,ijd->hij", q, e)
    return s * (j <= i).float()


def mha_forward(query: torch.Tensor, key: torch.Tensor, in_w: torch.Tensor, in_b: torch.Tensor,
                out_w: torch.Tensor, out_b: torch.Tensor, num_heads: int,
                Er: Optional[torch.Tensor] = None, attn_mask: Optional[torch.Tensor] = None,
                need_weights: bool = False):
    """multi_head_attention_forward_rpr, model/rpr.py:201-424 (eval mode, no
    bias_kv / zero_attn / padding mask), which for Er=None is also the arithmetic of
    the stock nn.MultiheadAttention used for cross attention (rpr.py:42,62) and of
    the stock encoder layers.  query (L,B,E), key=value (S,B,E)."""
    L, B, E = query.shape
    S = key.shape[0]
    dh = E // num_heads
    scaling = float(dh) ** -0.5
    q = F.linear(query, in_w[:E], in_b[:E])                       # rpr.py:253 / :263
    k, v = F.linear(key, in_w[E:], in_b[E:]).chunk(2, dim=-1)     # rpr.py:277
    q = q * scaling                                               # rpr.py:328
    q = q.contiguous().view(L, B * num_heads, dh).transpose(0, 1)  # rpr.py:349-353
    k = k.contiguous().view(S, B * num_heads, dh).transpose(0, 1)
    v = v.contiguous().view(S, B * num_heads, dh).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))                           # rpr.py:387
    if Er is not None:                                            # rpr.py:391-395
        Erv = get_valid_embedding(Er, L)
        qe = torch.einsum("hld,md->hlm", q, Erv)
        w = w + skew_literal(qe)
    if attn_mask is not None:                                     # rpr.py:397-399
        w = w + attn_mask.unsqueeze(0)
    w = torch.softmax(w, dim=-1)                                  # rpr.py:409
    o = torch.bmm(w, v)                                           # rpr.py:414
    o = o.transpose(0, 1).contiguous().view(L, B, E)
    o = F.linear(o, out_w, out_b)                                 # rpr.py:417
    if need_weights:
        return o, w.view(B, num_heads, L, S).sum(dim=1) / num_heads   # rpr.py:419-422
    return o, None


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def encoder_layer(x: torch.Tensor, sd: SD, p: str, H: int) -> torch.Tensor:
    """Stock nn.TransformerEncoderLayer as instantiated by nn.Transformer at
    video_music_transformer.py:967-971 (post-norm, ReLU, eps 1e-5)."""
    a, _ = mha_forward(x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                       sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], H)
    x = _ln(x + a, sd, p + "norm1")
    f = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return _ln(x + f, sd, p + "norm2")


def decoder_layer_rpr(tgt: torch.Tensor, memory: torch.Tensor, sd: SD, p: str, H: int,
                      tgt_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """TransformerDecoderLayerRPR.forward, model/rpr.py:55-70 (eval: dropouts are identity)."""
    a, _ = mha_forward(tgt, tgt, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                       sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], H,
                       Er=sd.get(p + "self_attn.Er"), attn_mask=tgt_mask)
    tgt = _ln(tgt + a, sd, p + "norm1")
    c, _ = mha_forward(tgt, memory, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                       sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], H)
    tgt = _ln(tgt + c, sd, p + "norm2")
    f = F.linear(F.relu(F.linear(tgt, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return _ln(tgt + f, sd, p + "norm3")


def sinusoid_pe(max_len: int, d_model: int) -> torch.Tensor:
    """model/positional_encoding.py:13-19 -> (max_len, d_model)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def count_layers(sd: SD, prefix: str) -> int:
    n = 0
    while (prefix + "%d.linear1.weight" % n) in sd:
        n += 1
    return n


def video_features(sem, scene, motion, emotion) -> torch.Tensor:
    """video_music_transformer.py:1003-1018: concat semantic | scene offset | motion | emotion."""
    vf = torch.cat([sem.float(), scene.unsqueeze(-1).float()], dim=-1)
    if motion.dim() == 2:
        vf = torch.cat([vf, motion.unsqueeze(-1).float()], dim=-1)
    else:
        vf = torch.cat([vf, motion], dim=-1)
    return torch.cat([vf, emotion.float()], dim=-1)


def chord_inputs(sd: SD, x, x_root, x_attr, feature_key, chord_embed: bool) -> torch.Tensor:
    """video_music_transformer.py:984-1001 -> Linear_chord([embed | key])  (B,T,d)."""
    if not chord_embed:
        e = F.embedding(x_root, sd["embedding_root.weight"]) + F.embedding(x_attr, sd["embedding_attr.weight"])
    else:
        e = F.embedding(x, sd["chord_embedding_model.weight"])
    B, T = e.shape[0], e.shape[1]
    key = feature_key.reshape(B, 1, 1).float().expand(B, T, 1)
    return F.linear(torch.cat([e, key], dim=-1), sd["Linear_chord.weight"], sd["Linear_chord.bias"])


def encode_memory(sd: SD, sem, scene, motion, emotion, num_heads: int = 8) -> torch.Tensor:
    """Encoder half of VideoMusicTransformer.forward (video_music_transformer.py:1003-1033):
    Linear_vis, + PE, 6 stock encoder layers, final LayerNorm -> memory (S,B,d)."""
    vf = F.linear(video_features(sem, scene, motion, emotion), sd["Linear_vis.weight"], sd["Linear_vis.bias"])
    vf = vf.permute(1, 0, 2)
    d = vf.shape[-1]
    vf = vf + sinusoid_pe(max(300, vf.shape[0]), d)[: vf.shape[0]].unsqueeze(1)
    nl = count_layers(sd, "transformer.encoder.layers.")
    for l in range(nl):
        vf = encoder_layer(vf, sd, "transformer.encoder.layers.%d." % l, num_heads)
    return _ln(vf, sd, "transformer.encoder.norm")


def amt_forward(sd: SD, x, x_root, x_attr, sem, key, scene, motion, emotion,
                num_heads: int = 8, chord_embed: bool = False, mask: bool = True) -> torch.Tensor:
    """VideoMusicTransformer.forward, video_music_transformer.py:978-1044 (IS_SEPERATED=False)."""
    T = x.shape[1]
    tgt_mask = torch.triu(torch.full((T, T), float("-inf")), diagonal=1) if mask else None  # :980
    xf = chord_inputs(sd, x, x_root, x_attr, key, chord_embed).permute(1, 0, 2)
    d = xf.shape[-1]
    xf = xf + sinusoid_pe(max(300, T), d)[:T].unsqueeze(1)                                # :1029
    memory = encode_memory(sd, sem, scene, motion, emotion, num_heads)
    nl = count_layers(sd, "transformer.decoder.layers.")
    out = xf
    for l in range(nl):                                                                    # rpr.py:24-35
        out = decoder_layer_rpr(out, memory, sd, "transformer.decoder.layers.%d." % l, num_heads, tgt_mask)
    out = _ln(out, sd, "transformer.decoder.norm")
    return F.linear(out.permute(1, 0, 2), sd["Wout.weight"], sd["Wout.bias"])           # :1042


def generate_greedy_literal(sd: SD, sem, key, scene, motion, emotion, primer, primer_root, primer_attr,
                            target_seq_length: int = 300, num_heads: int = 8, chord_embed: bool = False,
                            return_margins: bool = False):
    """VideoMusicTransformer.generate with beam=1, beam_chance=1.0
    (video_music_transformer.py:1046-1084,1129-1132): batch 1, the whole model is
    re-run on the growing prefix every step, the next token is the arg-max of
    softmax(logits)[..., :CHORD_END] at the last position; gen_seq_root / gen_seq_attr
    are NOT updated in this branch (they stay PAD for generated positions)."""
    gen = torch.full((1, target_seq_length), CHORD_PAD, dtype=torch.int64)
    gen_root = torch.full((1, target_seq_length), CHORD_ROOT_PAD, dtype=torch.int64)
    gen_attr = torch.full((1, target_seq_length), CHORD_ATTR_PAD, dtype=torch.int64)
    n = len(primer)
    gen[..., :n] = primer
    gen_root[..., :n] = primer_root
    gen_attr[..., :n] = primer_attr
    margins = []
    cur = n
    while cur < target_seq_length:
        y = torch.softmax(amt_forward(sd, gen[..., :cur], gen_root[..., :cur], gen_attr[..., :cur],
                                      sem, key, scene, motion, emotion, num_heads, chord_embed), dim=-1)[..., :CHORD_END]
        probs = y[:, cur - 1, :].flatten()
        top, idx = torch.topk(probs, 2)
        gen[..., cur] = idx[0] % CHORD_SIZE
        margins.append(float(top[0] - top[1]))
        cur += 1
    if return_margins:
        return gen[:, :cur], torch.tensor(margins)
    return gen[:, :cur]


def generate_greedy_cached(sd: SD, sem, key, scene, motion, emotion, primer, primer_root, primer_attr,
                           target_seq_length: int = 300, num_heads: int = 8, chord_embed: bool = False,
                           return_logits: bool = False, uniforms: Optional[torch.Tensor] = None, max_conseq_N: int = 0,
                           max_conseq_chord: int = 2, return_root_attr: bool = False):
    """KV-cached, batched restatement of the same greedy loop (what the decode
    kernels implement).  With `uniforms` (B, target_seq_length) the next token comes from the SAMPLING branch instead
    (video_music_transformer.py:1085-1123): softmax over the vocabulary restricted to [:CHORD_END], P(N) = 0 when
    max_conseq_N == 0, P(previous chord) = 0 after max_conseq_chord equal tokens, one Categorical draw -- stated as the
    inverse CDF at uniforms[b, t+1] -- and root / attr ids of the drawn chord (closed form of the JSON maps, :1105-1123).  Mathematically identical to generate_greedy_literal because
    the decoder is causal: row t of the prefix forward depends only on rows <= t, and
    Srel[t,j] = q_t . Er[er_len-1-(t-j)] does not depend on the prefix length
    (rpr.py:426-455).  primer* are (B,P) or (P,).  Returns (B, target_seq_length)."""
    B = sem.shape[0]
    E = sd["Linear_chord.weight"].shape[0]
    H = num_heads
    dh = E // H
    if primer.dim() == 1:
        primer, primer_root, primer_attr = (t.unsqueeze(0).expand(B, -1) for t in (primer, primer_root, primer_attr))
    P = primer.shape[1]
    gen = torch.full((B, target_seq_length), CHORD_PAD, dtype=torch.int64)
    gen_root = torch.full((B, target_seq_length), CHORD_ROOT_PAD, dtype=torch.int64)
    gen_attr = torch.full((B, target_seq_length), CHORD_ATTR_PAD, dtype=torch.int64)
    gen[:, :P], gen_root[:, :P], gen_attr[:, :P] = primer, primer_root, primer_attr
    memory = encode_memory(sd, sem, scene, motion, emotion, H)            # (S,B,E)
    S = memory.shape[0]
    nl = count_layers(sd, "transformer.decoder.layers.")
    pe = sinusoid_pe(max(300, target_seq_length), E)
    scaling = float(dh) ** -0.5
    ck, cv, sk, sv = [], [], [], []
    for l in range(nl):
        p = "transformer.decoder.layers.%d.multihead_attn." % l
        k, v = F.linear(memory, sd[p + "in_proj_weight"][E:], sd[p + "in_proj_bias"][E:]).chunk(2, dim=-1)
        ck.append(k.view(S, B, H, dh).permute(1, 2, 0, 3))                # (B,H,S,dh)
        cv.append(v.view(S, B, H, dh).permute(1, 2, 0, 3))
        sk.append(torch.zeros(B, H, target_seq_length, dh))
        sv.append(torch.zeros(B, H, target_seq_length, dh))
    all_logits = []
    for t in range(target_seq_length - 1):
        xt = chord_inputs(sd, gen[:, t:t + 1], gen_root[:, t:t + 1], gen_attr[:, t:t + 1], key, chord_embed)[:, 0]
        h = xt + pe[t]
        for l in range(nl):
            p = "transformer.decoder.layers.%d." % l
            qkv = F.linear(h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
            q, k, v = qkv.chunk(3, dim=-1)
            q = (q * scaling).view(B, H, dh)
            sk[l][:, :, t] = k.view(B, H, dh)
            sv[l][:, :, t] = v.view(B, H, dh)
            s = torch.einsum("bhd,bhjd->bhj", q, sk[l][:, :, :t + 1])
            Er = sd[p + "self_attn.Er"]
            er_rows = Er[Er.shape[0] - 1 - (t - torch.arange(t + 1))]     # Er[er_len-1-(t-j)]
            s = s + torch.einsum("bhd,jd->bhj", q, er_rows)
            a = torch.einsum("bhj,bhjd->bhd", torch.softmax(s, dim=-1), sv[l][:, :, :t + 1]).reshape(B, E)
            a = F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
            h = _ln(h + a, sd, p + "norm1")
            q = (F.linear(h, sd[p + "multihead_attn.in_proj_weight"][:E], sd[p + "multihead_attn.in_proj_bias"][:E])
                 * scaling).view(B, H, dh)
            s = torch.einsum("bhd,bhjd->bhj", q, ck[l])
            c = torch.einsum("bhj,bhjd->bhd", torch.softmax(s, dim=-1), cv[l]).reshape(B, E)
            c = F.linear(c, sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"])
            h = _ln(h + c, sd, p + "norm2")
            f = F.linear(F.relu(F.linear(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                         sd[p + "linear2.weight"], sd[p + "linear2.bias"])
            h = _ln(h + f, sd, p + "norm3")
        h = _ln(h, sd, "transformer.decoder.norm")
        logits = F.linear(h, sd["Wout.weight"], sd["Wout.bias"])
        if return_logits:
            all_logits.append(logits)
        if t + 1 >= P:
            if uniforms is None:
                gen[:, t + 1] = torch.argmax(logits[:, :CHORD_END], dim=-1)
            else:
                probs = torch.softmax(logits, dim=-1)[:, :CHORD_END].clone()                    # :1069
                if max_conseq_N == 0:
                    probs[:, 0] = 0.0                                                           # :1090-1091
                if t + 1 >= max_conseq_chord:                                                   # :1093-1104
                    prev = gen[:, t]
                    same = torch.ones(B, dtype=torch.bool)
                    for k in range(1, max_conseq_chord):
                        same &= gen[:, t - k] == prev
                    probs[same, prev[same]] = 0.0
                cdf = torch.cumsum(probs, dim=-1)
                target = uniforms[:, t + 1:t + 2] * cdf[:, -1:]
                nxt = ((cdf > target) & (probs > 0)).float().argmax(dim=-1)
                gen[:, t + 1] = nxt
                gen_root[:, t + 1] = torch.where(nxt > 0, (nxt - 1) // 13 + 1, torch.zeros_like(nxt))   # chord_inv/root/attr.json
                gen_attr[:, t + 1] = torch.where(nxt > 0, (nxt - 1) % 13 + 1, torch.ones_like(nxt))
    out = (gen,)
    if return_logits:
        out += (torch.stack(all_logits, dim=1),)
    if return_root_attr:
        out += (gen_root, gen_attr)
    return out if len(out) > 1 else gen


# --------------------------------------------------------------------------
# MoE (model/moe.py)
# --------------------------------------------------------------------------
def glu_expert(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """GLUExpert.forward, model/moe.py:44-49 (eval)."""
    x_ff = F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])
    x_g = F.linear(x, sd[p + "gate.weight"], sd[p + "gate.bias"])
    return F.linear(x_ff * F.silu(x_g), sd[p + "linear2.weight"], sd[p + "linear2.bias"])


def moe_route(x: torch.Tensor, gate_w: torch.Tensor, gate_b: torch.Tensor, k: int, t: float = 1.0,
              divide_before_topk: bool = True):
    """MoELayer: gate_logits = gate(x)/t; topk; softmax(fp32) (moe.py:180-190);
    SharedMoELayer (non-balancing): topk(gate(x)); softmax(weights/t) (moe.py:244-247,288)."""
    logits = F.linear(x, gate_w, gate_b)
    if divide_before_topk:
        logits = logits / t
    w, idx = torch.topk(logits, k)
    if not divide_before_topk:
        w = w / t
    return torch.softmax(w, dim=-1, dtype=torch.float), idx, logits


def moe_layer(x: torch.Tensor, sd: SD, p: str, n_experts: int, k: int, shared: bool = False,
              t: float = 1.0):
    """MoELayer.forward (moe.py:167-200) / SharedMoELayer.forward non-balancing
    (moe.py:231-302), eval mode.  x (L,B,d).  Returns (out, selected_experts, weights)."""
    w, idx, _ = moe_route(x, sd[p + "gate.weight"], sd[p + "gate.bias"], k, t, divide_before_topk=not shared)
    out = torch.zeros_like(x)
    for i in range(n_experts):
        ti, bi, ki = torch.where(idx == i)
        if ti.shape[0] == 0:
            continue
        out[ti, bi] += w[ti, bi, ki].unsqueeze(1) * glu_expert(x[ti, bi], sd, p + "experts.%d." % i)
    if shared:
        out = out + (1.0 / k) * glu_expert(x, sd, p + "shared_expert.")
    return out, idx, w


# --------------------------------------------------------------------------
# Grouped-query attention (model/grouped_query_attention.py)
# --------------------------------------------------------------------------
def sdp_gqa(query, key, value, is_causal: bool = False, scale: Optional[float] = None):
    """scaled_dot_product_gqa, grouped_query_attention.py:19-170 without dropout/masks.
    query (b,n,hq,d), key/value (b,s,hk,d) -> out (n,b,hq,d)  [sic: the reference
    returns sequence-first, :159].  Query head index = h*g + gi where h is the kv head."""
    b, n, hq, d = query.shape
    s, hk = key.shape[1], key.shape[2]
    g = hq // hk
    if scale is None:
        scale = d ** 0.5
    q = (query / scale).permute(0, 2, 1, 3).reshape(b, hk, g, n, d)         # (b,h,g,n,d)
    k = key.permute(0, 2, 1, 3)
    v = value.permute(0, 2, 1, 3)
    sim = torch.einsum("bhgnd,bhsd->bhgns", q, k)
    if is_causal:
        m = torch.ones(n, s, dtype=torch.bool).tril_()
        sim = sim.masked_fill(~m, torch.finfo(sim.dtype).min)                  # :149
    att = torch.softmax(sim, dim=-1)
    out = torch.einsum("bhgns,bhsd->bhgnd", att, v)                           # (b,h,g,n,d)
    return out.reshape(b, hq, n, d).permute(2, 0, 1, 3).contiguous()           # (n,b,(h g),d)


def mhgqa_forward(query, key, value, sd: SD, p: str, query_heads: int, kv_heads: int,
                  is_causal: bool = False, layer_norm: bool = True):
    """MultiheadGQA.forward, grouped_query_attention.py:285-358 (RoPE=None), including its
    literal `.view` reinterpretation of the (L,B,E) projections as (B,L,E) (:316-326)."""
    q = F.linear(query, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(key, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(value, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    tgt_len, bsz = q.shape[0], q.shape[1]
    src_len = k.shape[0]
    dh = q.shape[-1] // query_heads
    q = q.contiguous().view(bsz, tgt_len, query_heads, dh)
    k = k.contiguous().view(bsz, src_len, kv_heads, dh)
    v = v.contiguous().view(bsz, src_len, kv_heads, dh)
    x = sdp_gqa(q, k, v, is_causal=is_causal)              # (n,b,h,d) labelled "b n h d" by the caller (:343)
    x = x.reshape(x.shape[0], x.shape[1], query_heads * dh)
    if layer_norm:
        x = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


# --------------------------------------------------------------------------
# Generic layer wrappers (model/custom_transformer.py) -- BASELINE config 4: GQA attention + MoE FFN
# --------------------------------------------------------------------------
def _norm_generic(x, sd: SD, p: str, rms: bool):
    """LayerNorm(eps 1e-5) or the RMSNorm of custom_transformer.py:27-47 (eps 1e-6)."""
    if rms:
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * sd[p + ".weight"]
    return _ln(x, sd, p)


def variant_encoder_layer(x, sd: SD, p: str, hq: int, hk: int, n_experts: int, k: int, shared: bool, pre_norm: bool, rms: bool):
    """TransformerEncoderLayer(att=MultiheadGQA, ff=MoELayer|SharedMoELayer).forward, custom_transformer.py:1230-1248.
    attn_mask is passed by the wrapper and ignored by MultiheadGQA (grouped_query_attention.py:339): non-causal."""
    ff = lambda t: moe_layer(t, sd, p + "ff.", n_experts, k, shared=shared)[0]
    att = lambda t: mhgqa_forward(t, t, t, sd, p + "self_attn.", hq, hk)
    if not pre_norm:
        x = _norm_generic(x + att(x), sd, p + "norm1", rms)
        return _norm_generic(x + ff(x), sd, p + "norm2", rms)
    x = x + att(_norm_generic(x, sd, p + "norm1", rms))
    return x + ff(_norm_generic(x, sd, p + "norm2", rms))


def variant_decoder_layer(x, mem, sd: SD, p: str, hq: int, hk: int, n_experts: int, k: int, shared: bool, pre_norm: bool, rms: bool):
    """TransformerDecoderLayer.forward, custom_transformer.py:1261-1292."""
    ff = lambda t: moe_layer(t, sd, p + "ff.", n_experts, k, shared=shared)[0]
    satt = lambda t: mhgqa_forward(t, t, t, sd, p + "self_attn.", hq, hk)
    catt = lambda t: mhgqa_forward(t, mem, mem, sd, p + "cross_attn.", hq, hk)
    if not pre_norm:
        x = _norm_generic(x + satt(x), sd, p + "norm1", rms)
        x = _norm_generic(x + catt(x), sd, p + "norm2", rms)
        return _norm_generic(x + ff(x), sd, p + "norm3", rms)
    x = x + satt(_norm_generic(x, sd, p + "norm1", rms))
    x = x + catt(_norm_generic(x, sd, p + "norm2", rms))
    return x + ff(_norm_generic(x, sd, p + "norm3", rms))


def variant_stack_forward(sd: SD, src, tgt, n_layers: int, hq: int, hk: int, n_experts: int, k: int, shared: bool,
                          pre_norm: bool, rms: bool):
    """TransformerEncoder + TransformerDecoder (custom_transformer.py:1371-1401) over the layers above, each with a
    final norm; state_dict prefixes `enc.` / `dec.`.  Returns (memory, decoder output)."""
    m = src
    for i in range(n_layers):
        m = variant_encoder_layer(m, sd, "enc.layers.%d." % i, hq, hk, n_experts, k, shared, pre_norm, rms)
    m = _norm_generic(m, sd, "enc.norm", rms)
    y = tgt
    for i in range(n_layers):
        y = variant_decoder_layer(y, m, sd, "dec.layers.%d." % i, hq, hk, n_experts, k, shared, pre_norm, rms)
    return m, _norm_generic(y, sd, "dec.norm", rms)


# --------------------------------------------------------------------------
# Selective scan (model/pscan.py)
# --------------------------------------------------------------------------
def pscan_forward(A: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """What PScan.forward computes (pscan.py:154-188): H[t] = A[t]*H[t-1] + X[t], H[-1]=0,
    along dim 1 of (B,L,D,N).  Sequential restatement (cf. selective_scan_seq, mamba.py:353-383)."""
    H = torch.empty_like(X)
    h = torch.zeros_like(X[:, 0])
    for t in range(X.shape[1]):
        h = A[:, t] * h + X[:, t]
        H[:, t] = h
    return H


def pscan_backward(A: torch.Tensor, H: torch.Tensor, gH: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """PScan.backward (pscan.py:191-226): gX[t] = gH[t] + A[t+1]*gX[t+1] (reverse scan with A
    shifted by one, :218-221), gA[t] = H[t-1]*gX[t] with gA[0]=0 (:223-224)."""
    L = A.shape[1]
    gX = torch.empty_like(gH)
    g = torch.zeros_like(gH[:, 0])
    for t in range(L - 1, -1, -1):
        g = gH[:, t] + (A[:, t + 1] * g if t + 1 < L else 0)
        gX[:, t] = g
    gA = torch.zeros_like(A)
    gA[:, 1:] = H[:, :-1] * gX[:, 1:]
    return gA, gX


# --------------------------------------------------------------------------
# Mamba block (model/mamba.py, model/bimamba.py)
# --------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight, eps: float = 1e-5) -> torch.Tensor:
    """RMSNorm.forward (mamba.py:483-489)."""
    out = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return out * weight if weight is not None else out


def mamba_block_forward(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, dt_rank: int, d_state: int = 16,
                        use_version: int = 0) -> torch.Tensor:
    """MambaBlock.forward + ssm + selective_scan (mamba.py:259-351) with the scan stated sequentially
    (selective_scan_seq, :353-383, equals pscan mode).  x (B, L, D); sd[p + name] are the block's parameters."""
    B, L, _ = x.shape
    g = lambda n: sd.get(p + n)
    xz = F.linear(x, g("in_proj.weight"), g("in_proj.bias"))                       # :266
    xs, z = xz.chunk(2, dim=-1)                                                    # :267
    ED = xs.shape[-1]
    xs = F.conv1d(xs.transpose(1, 2), g("conv1d.weight"), g("conv1d.bias"), padding=g("conv1d.weight").shape[-1] - 1,
                  groups=ED)[:, :, :L].transpose(1, 2)                             # :270-272
    xs = F.silu(xs)                                                                # :274
    A = -torch.exp(g("A_log").float())                                             # :298
    D = g("D").float()
    dbc = F.linear(xs, g("x_proj.weight"))                                         # :301
    delta, Bm, Cm = torch.split(dbc, [dt_rank, d_state, d_state], dim=-1)          # :302
    delta = (g("dt_proj.weight") @ delta.transpose(1, 2)).transpose(1, 2)          # :304, :322
    delta = F.softplus(delta + g("dt_proj.bias"))                                  # :323
    deltaA = torch.exp(delta.unsqueeze(-1) * A)                                    # :343
    BX = delta.unsqueeze(-1) * Bm.unsqueeze(2) * xs.unsqueeze(-1)                  # :344-346
    hs = pscan_forward(deltaA, BX)                                                 # :348
    y = (hs @ Cm.unsqueeze(-1)).squeeze(3) + D * xs                                # :350-352
    zs = F.silu(z)                                                                 # :282
    out = y * zs + xs * (1 - torch.sigmoid(zs)) if use_version == 1 else y * zs    # :283-287
    return F.linear(out, g("out_proj.weight"), g("out_proj.bias"))                 # :288


def mamba_block_step(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, cache, dt_rank: int, d_state: int = 16):
    """MambaBlock.step + ssm_step (mamba.py:407-470): one token x (B, D) against cache = (h (B, ED, N) or None, inputs (B, ED,
    d_conv - 1)); returns (output (B, D), new cache).  The gate is y * silu(z) whatever use_version (:430)."""
    g = lambda n: sd.get(p + n)
    h, inputs = cache
    xz = F.linear(x, g("in_proj.weight"), g("in_proj.bias"))                       # :417
    xs, z = xz.chunk(2, dim=1)                                                     # :418
    ED = xs.shape[1]
    w = g("conv1d.weight")                                                         # (ED, 1, d_conv)
    window = torch.cat([inputs, xs.unsqueeze(2)], dim=2)                           # :421-422
    xc = F.conv1d(window, w, g("conv1d.bias"), padding=w.shape[-1] - 1, groups=ED)[:, :, w.shape[-1] - 1]
    xc = F.silu(xc)                                                                # :424
    A = -torch.exp(g("A_log").float())                                             # :443
    dbc = F.linear(xc, g("x_proj.weight"))                                         # :446
    delta, Bm, Cm = torch.split(dbc, [dt_rank, d_state, d_state], dim=-1)          # :448
    delta = F.softplus(F.linear(delta, g("dt_proj.weight"), g("dt_proj.bias")))    # :450
    dA = torch.exp(delta.unsqueeze(-1) * A)                                        # :452
    BX = delta.unsqueeze(-1) * Bm.unsqueeze(1) * xc.unsqueeze(-1)                  # :453-455
    if h is None:
        h = torch.zeros(x.shape[0], ED, d_state)                                   # :457-458
    h = dA * h + BX                                                                # :460
    y = (h @ Cm.unsqueeze(-1)).squeeze(2) + g("D").float() * xc                    # :462-464
    out = F.linear(y * F.silu(z), g("out_proj.weight"), g("out_proj.bias"))        # :428-431
    return out, (h, torch.cat([inputs[:, :, 1:], xs.unsqueeze(2)], dim=2))         # :434-435


def mamba_step(sd: Dict[str, torch.Tensor], x: torch.Tensor, caches, n_layers: int, dt_rank: int, d_state: int = 16, eps: float = 1e-5):
    """Mamba.step over ResidualBlock.step (mamba.py:100-108,151-159): x = mixer.step(norm(x)) + x per layer."""
    caches = list(caches)
    for l in range(n_layers):
        p = "layers.%d." % l
        y, caches[l] = mamba_block_step(sd, p + "mixer.", rmsnorm(x, sd.get(p + "norm.weight"), eps), caches[l], dt_rank, d_state)
        x = y + x
    return x, caches


def mamba_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, n_layers: int, dt_rank: int, d_state: int = 16,
                  eps: float = 1e-5, use_version: int = 0) -> torch.Tensor:
    """Mamba.forward over ResidualBlocks (mamba.py:91-98, 144-149): x = mixer(norm(x)) + x."""
    for l in range(n_layers):
        p = "layers.%d." % l
        x = mamba_block_forward(sd, p + "mixer.", rmsnorm(x, sd.get(p + "norm.weight"), eps), dt_rank, d_state, use_version=use_version) + x
    return x


def bimamba_layer_forward(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, dt_rank: int, d_state: int = 16) -> torch.Tensor:
    """BiMambaEncoderLayer.forward in eval mode (bimamba.py:64-99), literal: ffn2 reads the forward branch (:91)."""
    D = x.shape[-1]
    ln = lambda n, t: F.layer_norm(t, (D,), sd[p + n + ".weight"], sd[p + n + ".bias"], 1e-5)
    ffn = lambda n, t: F.linear(F.relu(F.linear(t, sd[p + n + ".0.weight"], sd[p + n + ".0.bias"])), sd[p + n + ".3.weight"],
                                sd[p + n + ".3.bias"])
    x_f = ln("norm1", mamba_block_forward(sd, p + "mamba_forward.", x, dt_rank, d_state) + x)
    x_f = ln("norm2", ffn("ffn1", x_f) + x_f)
    x_b = torch.flip(mamba_block_forward(sd, p + "mamba_backward.", torch.flip(x, dims=[1]), dt_rank, d_state), dims=[1])
    x_b = ln("norm3", x_b + x)
    x_b = ln("norm4", ffn("ffn2", x_f) + x_b)
    return x_f + x_b


def bimamba_v1_layer_forward(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, dt_rank: int, norm_first: bool,
                             moe: Optional[dict] = None, d_state: int = 16) -> torch.Tensor:
    """BiMambaEncoderLayer_V1.forward (bimamba.py:136-191), eval mode.  moe = dict(n_experts, k, shared) when the
    feed-forward is a (Shared)MoELayer, None for Linear-ReLU-Linear (`ffn.0`, `ffn.3`)."""
    mb = lambda name, t: mamba_block_forward(sd, p + name + ".", t, dt_rank, d_state, use_version=1)
    ln = lambda name, t: F.layer_norm(t, (t.shape[-1],), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5)

    def ffn(t):
        if moe is None:
            h = F.relu(F.linear(t, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"]))
            return F.linear(h, sd[p + "ffn.3.weight"], sd[p + "ffn.3.bias"])
        return moe_layer(t, sd, p + "ffn.", moe["n_experts"], moe["k"], shared=moe["shared"])[0]
    x_flip = torch.flip(x, dims=[1])
    if norm_first:
        x_f = x + mb("mamba_forward", ln("norm1", x))
        x_b = x + torch.flip(mb("mamba_backward", ln("norm2", x_flip)), dims=[1])
        x = x_f + x_b
        return x + ffn(ln("norm3", x))
    x_f = ln("norm1", mb("mamba_forward", x) + x)
    x_b = ln("norm2", torch.flip(mb("mamba_backward", x_flip), dims=[1]) + x)
    x = x_f + x_b
    return ln("norm3", ffn(x) + x)


# --------------------------------------------------------------------------
# Evaluation metrics (dataset/vevo_dataset.py)
# --------------------------------------------------------------------------
def vevo_accuracy(out: torch.Tensor, tgt: torch.Tensor, pad: int = 158) -> float:
    """compute_vevo_accuracy, vevo_dataset.py:653-673."""
    pred = torch.argmax(torch.softmax(out, dim=-1), dim=-1).flatten()
    t = tgt.flatten()
    m = t != pad
    if int(m.sum()) == 0:
        return 1.0
    return float((pred[m] == t[m]).sum()) / int(m.sum())


def hits_k(out: torch.Tensor, tgt: torch.Tensor, k: int, pad: int = 158) -> float:
    """compute_hits_k, vevo_dataset.py:675-701 (vectorised: target among the top-k classes of each non-pad row)."""
    top = torch.topk(torch.softmax(out, dim=-1), k, dim=-1).indices.reshape(-1, k)
    t = tgt.flatten()
    m = t != pad
    hit = (top == t.unsqueeze(1)).any(dim=1) & m
    return float(hit.sum()) / max(int(m.sum()), 1)


def vevo_correspondence(out: torch.Tensor, tgt_emotion: torch.Tensor, tgt_emotion_prob: torch.Tensor, threshold: float):
    """compute_vevo_correspondence, vevo_dataset.py:747-810, position by position as the reference loops (the chord_inv /
    chord_attr JSON lookups :775-790 reduce to quality = 1 for chord 0 and (c - 1) % 13 + 1 otherwise; checked against the
    reference's tables by oracle/make_golden.py).  Returns (value, pt, right)."""
    emo = tgt_emotion.reshape(-1, tgt_emotion.shape[-1])
    prob = tgt_emotion_prob.reshape(-1)
    pred = torch.argmax(torch.softmax(out, dim=-1), dim=-1).flatten()
    if emo.shape[0] == 0:
        return 1.0, 0, 0
    pt = right = 0
    for i in range(pred.numel()):
        if emo[i, -1] == 1 or bool(torch.all(emo[i, :14] == 0)) or prob[i] < threshold:      # :781-783
            continue
        pt += 1
        c = int(pred[i])
        if c != 157 and c != 158:                                                                 # CHORD_END / CHORD_PAD, :786
            quality = 1 if c == 0 else (c - 1) % 13 + 1
            if emo[i, quality] == 1:
                right += 1
    return (-1 if pt == 0 else right / pt), pt, right


# --------------------------------------------------------------------------
# V2 / V3 attention: nn.MultiheadAttention + RoPE (model/custom_transformer.py:51-321, 864-1218; rotate_operation.py)
# --------------------------------------------------------------------------
def rope_cache(dim: int, max_seq_len: int, base: int = 10000) -> torch.Tensor:
    """RotaryPositionalEmbeddings._rope_init / build_rope_cache (rotate_operation.py:89-110): [max_seq_len, dim/2, 2]."""
    theta = 1.0 / (base ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    idx_theta = torch.einsum("i, j -> ij", torch.arange(max_seq_len, dtype=theta.dtype), theta).float()
    return torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1)


def rope_literal(x: torch.Tensor, cache: torch.Tensor) -> torch.Tensor:
    """RotaryPositionalEmbeddings.forward (rotate_operation.py:112-165) on the [n_heads, len, bsz, head_dim] VIEW that
    custom_multi_head_attention_forward hands it (custom_transformer.py:1044-1050)."""
    seq_len = x.size(1)
    rc = cache[:seq_len]
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    rc = rc.view(-1, xs.size(1), 1, xs.size(3), 2)[:xs.size(0)]
    out = torch.stack([xs[..., 0] * rc[..., 0] - xs[..., 1] * rc[..., 1], xs[..., 1] * rc[..., 0] + xs[..., 0] * rc[..., 1]], -1)
    return out.flatten(3).type_as(x)


def custom_mha_forward(query, key, value, sd: SD, p: str, num_heads: int, cache: Optional[torch.Tensor], causal: bool):
    """custom_multi_head_attention_forward (custom_transformer.py:864-1218), eval, no padding masks: returns
    (output (L,B,E), head-averaged weights (B,L,S))."""
    L, B, E = query.shape
    S = key.shape[0]
    dh = E // num_heads
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    if cache is not None:
        q = rope_literal(q.contiguous().view(num_heads, L, B, dh), cache).view(L, B, E)
        k = rope_literal(k.contiguous().view(num_heads, S, B, dh), cache).view(S, B, E)
    q = q.contiguous().view(L, B * num_heads, dh).transpose(0, 1) * (dh ** -0.5)
    k = k.contiguous().view(S, B * num_heads, dh).transpose(0, 1)
    v = v.contiguous().view(S, B * num_heads, dh).transpose(0, 1)
    a = torch.bmm(q, k.transpose(1, 2))
    if causal:
        a = a + torch.triu(torch.full((L, S), float("-inf")), diagonal=1)
    a = torch.softmax(a, dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).contiguous().view(L * B, E)
    o = F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"]).view(L, B, E)
    return o, a.view(B, num_heads, L, S).mean(dim=1)


def diff_mha_forward(query, key, value, sd: SD, p: str, num_heads: int, cache: Optional[torch.Tensor], causal: bool, depth: int):
    """DifferentialMultiheadAttention.forward (custom_transformer.py:763-830), eval: literal views included."""
    import math
    L, B, E = query.shape
    S = key.shape[0]
    dh = E // num_heads
    k = F.linear(key, sd[p + "k_proj.weight"])
    q = F.linear(query, sd[p + "q_proj.weight"])
    v = F.linear(value, sd[p + "v_proj.weight"])
    q = q.contiguous().view(2 * num_heads, L, B, dh)
    k = k.contiguous().view(2 * num_heads, S, B, dh)
    if cache is not None:
        q, k = rope_literal(q, cache), rope_literal(k, cache)
    q = q.view(B, L, 2 * num_heads, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.view(B, S, 2 * num_heads, dh).transpose(1, 2)
    v = v.contiguous().view(B, S, num_heads, dh).transpose(1, 2)
    a = torch.matmul(q, k.transpose(-1, -2))
    if causal:
        a = a + torch.triu(torch.full((L, S), float("-inf")), diagonal=1 + S - L)
    a = torch.softmax(a, dim=-1)
    lam_init = 0.8 - 0.6 * math.exp(-0.3 * depth)
    lam = torch.exp(torch.sum(sd[p + "lambda_q1"] * sd[p + "lambda_k1"])) - torch.exp(torch.sum(sd[p + "lambda_q2"] * sd[p + "lambda_k2"])) \
        + lam_init
    a = a.view(B, num_heads, 2, L, S)
    a = a[:, :, 0] - lam * a[:, :, 1]
    o = torch.matmul(a, v)                                                            # (B, H, L, dh)
    o = o * torch.rsqrt(o.pow(2).mean(-1, keepdim=True) + 1e-5) * sd[p + "subln.weight"]
    o = (o * (1 - lam_init)).contiguous().view(L, B, E)
    return F.linear(o, sd[p + "out_proj.weight"])


def zoo_forward(sd: SD, x_root, x_attr, sem, key, scene, motion, emotion, n_layers: int, num_heads: int, ff_kind,
                rope: bool, pos_tables: bool, rms: bool = False, max_seq_video: int = 300, mask: bool = True,
                rope_dim: Optional[int] = None, diff_enc: bool = False, diff_dec: bool = False, pre_norm: bool = False,
                gqa_kv_heads: int = 0, moe_k: int = 2) -> torch.Tensor:
    """Shared body of VideoMusicTransformer_V1.forward / _V2.forward (video_music_transformer.py:141-225, 437-520), eval:
    embeddings + key column -> Linear_chord, video features -> Linear_vis, learned position tables or RoPE inside the
    attention, post-norm wrappers (custom_transformer.py:1220-1292) with feed-forward ff_kind(layer) in {"glu", "moe",
    "shared"}, final norms, Wout."""
    B, T = x_root.shape
    E = sd["Wout.weight"].shape[1]
    emb = sd["embedding_root.weight"][x_root] + sd["embedding_attr.weight"][x_attr]
    keycol = key.reshape(B, -1)[:, :1].float().expand(B, T).unsqueeze(-1)
    xf = F.linear(torch.cat([emb, keycol], dim=-1), sd["Linear_chord.weight"], sd["Linear_chord.bias"]).permute(1, 0, 2)
    vf = F.linear(video_features(sem, scene, motion, emotion), sd["Linear_vis.weight"], sd["Linear_vis.bias"]).permute(1, 0, 2)
    S = vf.shape[0]
    if pos_tables:
        xf = xf + sd["positional_embedding.weight"][:T].unsqueeze(1)
        vf = vf + sd["positional_embedding_video.weight"][:S].unsqueeze(1)
    cache = rope_cache(rope_dim or E, max_seq_video) if rope else None
    ln = lambda t, p: _norm_generic(t, sd, p, rms)

    def att(qx, kx, p, causal, diff, depth):
        if gqa_kv_heads:            # BASELINE config 4: MultiheadGQA in every layer; decoder self-attention causal (is_causal=True)
            return mhgqa_forward(qx, kx, kx, sd, p, num_heads, gqa_kv_heads, is_causal=causal)
        if diff:
            return diff_mha_forward(qx, kx, kx, sd, p, num_heads, cache, causal, depth)
        return custom_mha_forward(qx, kx, kx, sd, p, num_heads, cache, causal)[0]

    def ff(t, p, l):
        kind = ff_kind(l)
        if kind == "glu":
            return glu_expert(t, sd, p + "ff.")
        return moe_layer(t, sd, p + "ff.", 6, moe_k, shared=(kind == "shared"))[0]   # moe_k: the top-k scheduler's k in train() mode
    m = vf
    for l in range(n_layers):
        p = "transformer.encoder.layers.%d." % l
        if not pre_norm:
            m = ln(m + att(m, m, p + "self_attn.", False, diff_enc, l), p + "norm1")
            m = ln(m + ff(m, p, l), p + "norm2")
        else:                                                               # custom_transformer.py:1239-1247
            t = ln(m, p + "norm1")
            m = m + att(t, t, p + "self_attn.", False, diff_enc, l)
            m = m + ff(ln(m, p + "norm2"), p, l)
    m = ln(m, "transformer.encoder.norm")
    y = xf
    for l in range(n_layers):
        p = "transformer.decoder.layers.%d." % l
        if not pre_norm:
            y = ln(y + att(y, y, p + "self_attn.", bool(mask), diff_dec, l), p + "norm1")
            y = ln(y + att(y, m, p + "cross_attn.", False, diff_dec, l), p + "norm2")
            y = ln(y + ff(y, p, l), p + "norm3")
        else:                                                               # :1278-1291
            t = ln(y, p + "norm1")
            y = y + att(t, t, p + "self_attn.", bool(mask), diff_dec, l)
            y = y + att(ln(y, p + "norm2"), m, p + "cross_attn.", False, diff_dec, l)
            y = y + ff(ln(y, p + "norm3"), p, l)
    y = ln(y, "transformer.decoder.norm")
    return F.linear(y.permute(1, 0, 2), sd["Wout.weight"], sd["Wout.bias"])


def gqa_moe_forward(sd: SD, x_root, x_attr, sem, key, scene, motion, emotion, n_layers: int = 6, num_heads: int = 8, kv_heads: int = 2,
                    shared: bool = False, rms: bool = False, pre_norm: bool = False, mask: bool = True) -> torch.Tensor:
    """BASELINE config 4: the V1 shell (video_music_transformer.py:77-118) over TransformerEncoderLayer /
    TransformerDecoderLayer(att=MultiheadGQA(d, heads, kv_heads), ff=MoELayer | SharedMoELayer) (custom_transformer.py:1220-1292),
    decoder self-attention with is_causal=True (grouped_query_attention.py:106-121), everything else non-causal."""
    return zoo_forward(sd, x_root, x_attr, sem, key, scene, motion, emotion, n_layers, num_heads,
                       lambda l: "shared" if shared else "moe", rope=False, pos_tables=True, rms=rms, mask=mask, pre_norm=pre_norm,
                       gqa_kv_heads=kv_heads)


def zoo_generate_greedy_literal(forward, sem, key, scene, motion, emotion, primer, primer_root, primer_attr, target_seq_length: int):
    """generate(beam=1, beam_chance=1.0) of the V1 / V2 shells (video_music_transformer.py:227-315, 522-610): batch of one, one
    full forward per token, arg-max over the first 157 classes of softmax(logits[-1]), root / attribute of generated positions
    stay PAD.  `forward(x_root, x_attr)` -> logits (1, T, 159)."""
    gen = torch.full((1, target_seq_length), 158, dtype=torch.long)
    gr = torch.full((1, target_seq_length), 14, dtype=torch.long)
    ga = torch.full((1, target_seq_length), 15, dtype=torch.long)
    n0 = len(primer)
    gen[0, :n0], gr[0, :n0], ga[0, :n0] = primer, primer_root, primer_attr
    cur = n0
    while cur < target_seq_length:
        y = forward(gr[:, :cur], ga[:, :cur])
        gen[0, cur] = int(torch.argmax(torch.softmax(y[0, cur - 1], -1)[:157]))
        cur += 1
    return gen


def v2_forward(sd: SD, x_root, x_attr, sem, key, scene, motion, emotion, n_layers: int = 6, num_heads: int = 8,
               version: str = "2.2", max_seq_video: int = 300, mask: bool = True, moe_k: int = 2) -> torch.Tensor:
    """VideoMusicTransformer_V2 (versions 2.0 / 2.1 / 2.2): three shallow layers (GLUExpert) then SharedMoELayer layers (:399-416).
    moe_k: experts per token -- 2 in eval mode; in train() mode versions 2.0 / 2.1 take it from their TopKScheduler (moe.py:66-82,
    232-236: 6 for the first 31 forward calls)."""
    return zoo_forward(sd, x_root, x_attr, sem, key, scene, motion, emotion, n_layers, num_heads,
                       lambda l: "glu" if l < 3 else "shared", rope=version != "2.0", pos_tables=version == "2.0",
                       max_seq_video=max_seq_video, mask=mask, moe_k=moe_k)


def v1_forward(sd: SD, x_root, x_attr, sem, key, scene, motion, emotion, n_layers: int = 6, num_heads: int = 8,
               version: str = "1.1", rms: bool = False, mask: bool = True) -> torch.Tensor:
    """VideoMusicTransformer_V1 versions 1.1 (MoELayer) / 1.3 (SharedMoELayer): video_music_transformer.py:77-118."""
    return zoo_forward(sd, x_root, x_attr, sem, key, scene, motion, emotion, n_layers, num_heads,
                       lambda l: "moe" if version == "1.1" else "shared", rope=False, pos_tables=True, rms=rms, mask=mask)


def v3_forward(sd: SD, x_root, x_attr, sem, key, scene, motion, emotion, n_layers: int = 6, num_heads: int = 8,
               version: str = "3.0", mask: bool = True) -> torch.Tensor:
    """VideoMusicTransformer_V3 (video_music_transformer.py:611-760): RMSNorm, RoPE cache of dimension 2 * d_model, differential
    attention in the decoder (all versions) and in the encoder (3.1, 3.2), three GLU layers then SharedMoE layers, pre-norm for
    3.2.  The cross-attention of a differential layer carries the decoder mask?  No: memory_mask is None (:1271)."""
    return zoo_forward(sd, x_root, x_attr, sem, key, scene, motion, emotion, n_layers, num_heads, lambda l: "glu" if l < 3 else "shared",
                       rope=True, pos_tables=False, rms=True, mask=mask, rope_dim=2 * sd["Wout.weight"].shape[1],
                       diff_enc=version != "3.0", diff_dec=True, pre_norm=version == "3.2")


def video_regression_forward(sd: SD, sem, emotion, reg_model: str, n_layers: int, dt_rank: int):
    """VideoRegression.forward (video_regression.py:208-245) for regModel "mamba" / "mamba+" (Mamba stack) and "bimamba+"
    (BiMambaEncoder of Bi-Mamba+ layers with the FFN feed-forward) and "moe_bimamba+" / "sharedmoe_bimamba+" (MoE feed-forward): returns (loudness / note density (B,L,2), instruments)."""
    vf = F.linear(torch.cat([sem.float(), emotion.float()], dim=-1), sd["in_proj.0.weight"], sd["in_proj.0.bias"])
    if reg_model in ("mamba", "mamba+"):
        msd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        out = mamba_forward(msd, vf, n_layers, dt_rank, use_version=1 if reg_model == "mamba+" else 0)
    elif reg_model in ("bimamba+", "moe_bimamba+", "sharedmoe_bimamba+"):              # video_regression.py:158-186
        moe = None if reg_model == "bimamba+" else dict(n_experts=6, k=2, shared=reg_model.startswith("shared"))
        out = vf
        for i in range(n_layers):
            out = bimamba_v1_layer_forward(sd, "model.layers.%d." % i, out, dt_rank, norm_first=False, moe=moe)
    else:
        raise NotImplementedError(reg_model)
    return F.linear(out, sd["regressor.weight"], sd["regressor.bias"]), torch.sigmoid(F.linear(out, sd["classifier.0.weight"], sd["classifier.0.bias"]))
