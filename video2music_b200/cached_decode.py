"""KV-cached, batched autoregressive generation for the decoder stacks built from the reference's generic wrappers
(model/custom_transformer.py:1250-1292,1401-1433) -- BASELINE config 4 "generation".

The reference generates with a batch of one and one FULL forward per token over the growing prefix
(video_music_transformer.py:227-315 / 522-610): O(n^2) work, nothing reused.  For the stacks whose self-attention is causal the
keys / values of earlier positions never change, so one position per step is enough:

  * per decoder layer, a self-attention cache K | V of shape (videos, cap, kv_heads * head_dim) -- grouped-query attention
    (grouped_query_attention.py:172-358) keeps only `kv_heads` heads, i.e. 4x less cache than the 8 query heads at kv_heads = 2;
  * the cross-attention K | V of the encoder memory, projected once per generation instead of once per token and layer;
  * one step = one token per video through every layer: q / k / v projection of the new rows, attention of ONE query row per
    (video, head) over the cached rows (the fp32 attention kernel with Lq = 1), out-projection, the layer's feed-forward
    (GLUExpert, MoELayer or SharedMoELayer: fused router + permute + grouped expert GEMMs over the B new tokens).

The step runs in fp32 on kernels made for a handful of rows (csrc/step_f32.cu: the linear layers as weight streams over the whole
chip, one query row per (video, head) over the cache; fixed summation orders, results independent of the batch size) plus the
MoE kernels of the full forward; the greedy tokens equal those of the literal re-forward loop (tests/test_gpu_cached_decode.py),
which in turn equals the unmodified reference's generate() for the V1 models (tests/golden/v2.pt).

Cacheable attention modules: `CustomMultiheadAttention` without RoPE (V1 '1.1' / '1.3', V2 '2.0'; the RoPE variant rotates a
(position, head) element by an angle that depends on the CURRENT prefix length, custom_transformer.py:1044-1053, so its keys
change every step) and `MultiheadGQA` with causal self-attention.
"""
import weakref
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .custom_transformer import CustomMultiheadAttention, _norm
from .grouped_query_attention import MultiheadGQA
from .video_music_transformer import CHORD_END, CHORD_PAD

CHORD_ROOT_PAD, CHORD_ATTR_PAD = 14, 15            # utilities/constants.py


def _kind(att: nn.Module) -> Optional[str]:
    if isinstance(att, MultiheadGQA):
        return "gqa"
    if isinstance(att, CustomMultiheadAttention) and att.RoPE is None:
        return "mha"
    return None


def cacheable(model: nn.Module) -> bool:
    """True when `model` (a V1 / V2 / GQA shell: .transformer.decoder.layers of generic TransformerDecoderLayer) can be decoded
    with a KV cache: learned position tables, cacheable attention modules, causal decoder self-attention."""
    dec = getattr(getattr(model, "transformer", None), "decoder", None)
    if dec is None or not getattr(model, "_pos_tables", False):
        return False
    for layer in dec.layers:
        ks, kc = _kind(layer.self_attn), _kind(layer.cross_attn)
        if ks is None or kc is None:
            return False
        if ks == "gqa" and not getattr(layer.self_attn, "force_causal", False):
            return False                       # the literal GQA module ignores the mask (grouped_query_attention.py:339): not causal
    return True


class _Att:
    """One attention module of one decoder layer in step form."""

    def __init__(self, att: nn.Module):
        self.att, self.kind = att, _kind(att)
        det = lambda p: None if p is None else p.detach()
        if self.kind == "gqa":
            self.Hq, self.Hk = att.query_heads, att.kv_heads
            self.E = att.q_proj.in_features
            self.dh = self.E // self.Hq
            self.wq, self.bq = det(att.q_proj.weight), det(att.q_proj.bias)
            self.wk, self.bk = det(att.k_proj.weight), det(att.k_proj.bias)
            self.wv, self.bv = det(att.v_proj.weight), det(att.v_proj.bias)
            if self.bq is not None and self.bk is not None and self.bv is not None:
                # one launch for the three projections of a step (rows of the stacked weight keep their own k order)
                self.w_qkv = torch.cat([self.wq, self.wk, self.wv], 0).contiguous()
                self.b_qkv = torch.cat([self.bq, self.bk, self.bv], 0).contiguous()
        else:
            self.Hq = self.Hk = att.num_heads
            self.E, self.dh = att.embed_dim, att.head_dim
            w, b = det(att.in_proj_weight), det(att.in_proj_bias)
            E = self.E
            self.wq, self.bq, self.wk, self.bk, self.wv, self.bv = w[:E], b[:E], w[E:2 * E], b[E:2 * E], w[2 * E:], b[2 * E:]
            self.w_qkv, self.b_qkv, self.w_kv, self.b_kv = w, b, w[E:], b[E:]
        self.Wk = self.Hk * self.dh                      # cache row width
        self.q_scale = float(self.dh) ** -0.5            # q / sqrt(d) (grouped_query_attention.py:93-96; F.multi_head_attention_forward)

    def qkv(self, x: torch.Tensor):
        """q (B, E), k, v (B, Wk) of the new rows x (B, E)."""
        if self.kind == "mha":                           # packed projection, as the module's self-attention path
            y = ops.step_linear(x, self.w_qkv, self.b_qkv)
            E = self.E
            return y[:, :E], y[:, E:2 * E], y[:, 2 * E:]
        if getattr(self, "w_qkv", None) is not None:
            y = ops.step_linear(x, self.w_qkv, self.b_qkv)
            E, W = self.E, self.Hk * self.dh
            return y[:, :E], y[:, E:E + W], y[:, E + W:]
        return ops.step_linear(x, self.wq, self.bq), ops.step_linear(x, self.wk, self.bk), ops.step_linear(x, self.wv, self.bv)

    def q_only(self, x: torch.Tensor) -> torch.Tensor:
        return ops.step_linear(x, self.wq, self.bq)

    def memory_kv(self, mem_rows: torch.Tensor):
        """K, V (S*B, Wk) of the encoder memory rows (s, b) -- once per generation."""
        if self.kind == "mha":
            kv = ops.linear(mem_rows, self.w_kv, self.b_kv)
            return kv[:, :self.E], kv[:, self.E:]
        return ops.linear(mem_rows, self.wk, self.bk), ops.linear(mem_rows, self.wv, self.bv)

    def attend(self, q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, n: int, k_strides, n_dev=None) -> torch.Tensor:
        """One query row per (video, head) over the first n cached rows (n_dev: the count lives in an int32 device word and n is
        the capacity -- the launch is then the same for every position), then the module's output path."""
        B, E = q.shape[0], self.E
        assert V.stride() == K.stride() and q.stride(1) == 1
        if self.dh == 64:                                # one query row per (video, head): csrc/step_f32.cu
            ctx = ops.step_attention(q, K, V, Hq=self.Hq, Hkv=self.Hk, dh=64, n_max=n, kv_strides=k_strides, n_dev=n_dev,
                                     q_scale=self.q_scale)
        else:
            ctx = torch.empty((B, E), device=q.device, dtype=torch.float32)
            ops.attention(q, K, V, ctx, B=B, Hq=self.Hq, Hkv=self.Hk, Lq=1, Lk=n, dh=self.dh, q_strides=(q.stride(0), q.stride(0)),
                          k_strides=k_strides, v_strides=k_strides, o_strides=(E, E), causal=False, q_scale=self.q_scale, lk_dev=n_dev)
        a = self.att
        if self.kind == "gqa":
            if a.layer_norm:                             # grouped_query_attention.py:347-349
                ctx = ops.layernorm(ctx, a.norm.weight.detach(), a.norm.bias.detach(), eps=a.norm.eps)
            return ops.step_linear(ctx, a.out_proj.weight.detach(), None if a.out_proj.bias is None else a.out_proj.bias.detach(), k=E)
        return ops.step_linear(ctx, a.out_proj.weight.detach(), a.out_proj.bias.detach())


class CachedDecoder:
    """Self-attention caches + projected memory of every decoder layer, and the per-position step."""

    def __init__(self, model: nn.Module, memory: torch.Tensor, cap: int):
        S, B, E = memory.shape
        self.model, self.B, self.S, self.E, self.cap = model, B, S, E, cap
        dec = model.transformer.decoder
        self.layers = list(dec.layers)
        self.final_norm = dec.norm
        mem_rows = memory.reshape(S * B, E).float().contiguous()
        dev = memory.device
        self.sa: List[_Att] = [_Att(l.self_attn) for l in self.layers]
        self.ca: List[_Att] = [_Att(l.cross_attn) for l in self.layers]
        self.mem_kv = [a.memory_kv(mem_rows) for a in self.ca]
        self.K = [torch.zeros((B, cap, a.Wk), device=dev, dtype=torch.float32) for a in self.sa]
        self.V = [torch.zeros((B, cap, a.Wk), device=dev, dtype=torch.float32) for a in self.sa]
        self.cache_bytes = sum(k.numel() * 8 for k in self.K) + sum(kv[0].numel() * 8 for kv in self.mem_kv)

    def load_memory(self, memory: torch.Tensor) -> None:
        """New encoder memory into the SAME cross-attention K | V buffers (a captured graph keeps reading them); the self-attention
        caches need no reset: only their first n_keys rows are ever read."""
        S, B, E = memory.shape
        assert (S, B, E) == (self.S, self.B, self.E)
        mem_rows = memory.reshape(S * B, E).float().contiguous()
        for a, (k, v) in zip(self.ca, self.mem_kv):
            nk, nv = a.memory_kv(mem_rows)
            k.copy_(nk)
            v.copy_(nv)

    def step(self, x: torch.Tensor, t: torch.Tensor, n_keys: torch.Tensor) -> torch.Tensor:
        """x (B, E): embedded token of position t of every video -> decoder output rows (B, E) after the final norm.
        t: int64 device tensor (1,), n_keys = t + 1 as an int32 device tensor (1,): the position never reaches the host, so the
        launches of a step are the same for every position (one CUDA graph, replayed)."""
        B = self.B
        for i, layer in enumerate(self.layers):
            sa, ca = self.sa[i], self.ca[i]
            K, V = self.K[i], self.V[i]
            mk, mv = self.mem_kv[i]

            def self_att(z):
                q, k, v = sa.qkv(z)
                K.index_copy_(1, t, k.unsqueeze(1))
                V.index_copy_(1, t, v.unsqueeze(1))
                return sa.attend(q, K, V, self.cap, (K.stride(0), K.stride(1)), n_dev=n_keys)

            def cross_att(z):
                # memory rows are ordered (s, b): video b's row s sits at (s * B + b)
                return ca.attend(ca.q_only(z), mk, mv, self.S, (mk.stride(0), B * mk.stride(0)))

            ff = lambda z: layer.ff(z.view(1, B, self.E)).reshape(B, -1)
            if not layer.pre_norm:                                           # custom_transformer.py:1263-1277
                x = _norm(layer.norm1, x, self_att(x))
                x = _norm(layer.norm2, x, cross_att(x))
                x = _norm(layer.norm3, x, ff(x))
            else:                                                            # :1278-1291
                x = ops.axpy(x, self_att(_norm(layer.norm1, x)), 1.0)
                x = ops.axpy(x, cross_att(_norm(layer.norm2, x)), 1.0)
                x = ops.axpy(x, ff(_norm(layer.norm3, x)), 1.0)
        if self.final_norm is not None:
            x = _norm(self.final_norm, x)
        return x


@torch.no_grad()
def generate_cached(model: nn.Module, feature_semantic_list, feature_key, feature_scene_offset, feature_motion, feature_emotion,
                    primer, primer_root, primer_attr, target_seq_length: int = 300, beam: int = 0, beam_chance: float = 1.0,
                    max_conseq_N: int = 0, max_conseq_chord: int = 2, temperature: float = 1.0,
                    uniforms: Optional[torch.Tensor] = None, use_graph: Optional[bool] = None) -> torch.Tensor:
    """Batched `generate` of the V1 / V2 '2.0' / GQA shells (video_music_transformer.py:227-315, 522-610) with a KV cache.
    Features carry a batch dimension of B videos (the reference: 1); the primer (P,) is shared by all videos or (B, P).
    beam=1 (beam_chance >= 1): arg-max over the first 157 classes, root / attribute inputs of generated positions stay PAD
    (literal); beam=0: the sampling branch (no-"N" / no-repeat constraints, inverse-CDF draw from `uniforms` (B, T) or
    torch.rand), root / attribute updated.  Returns (B, target_seq_length) int64.
    The position is a device word (caches are written with index_copy_, the attention kernel reads its key count from it), so
    one position is ONE CUDA graph captured after the first step and replayed (use_graph; default: on unless a module carries
    a Python-side temperature scheduler, which advances per call, moe.py:238-242).  The buffers and the graph of a configuration
    (batch, lengths, options, parameter versions) are kept with the model (weakly) and reused by later calls."""
    assert not model.training, "Cannot generate while in training mode"
    if not (beam == 0 or (beam == 1 and beam_chance >= 1.0)):
        raise NotImplementedError("beam > 1 / 0 < beam_chance < 1 are not reproduced")
    if not cacheable(model):
        raise NotImplementedError("this model's decoder cannot be decoded with a KV cache (RoPE / differential attention / "
                                  "non-causal self-attention); use generate()")
    dev = model.Wout.weight.device
    sem, key, scene, motion, emotion = (t.to(dev) for t in (feature_semantic_list, feature_key, feature_scene_offset,
                                                             feature_motion, feature_emotion))
    B, S, E = sem.shape[0], sem.shape[1], model.d_model
    T = target_seq_length
    prim = lambda p: (p.long().to(dev).reshape(1, -1).expand(B, -1) if p.dim() == 1 else p.long().to(dev))
    primer, primer_root, primer_attr = prim(primer), prim(primer_root), prim(primer_attr)
    n0 = primer.shape[1]
    if beam == 0 and uniforms is None:
        uniforms = torch.rand((B, T), device=dev)
    # ---- encoder memory, once (the literal loop recomputes it for every token)
    tr = lambda t: t.transpose(0, 1).contiguous()
    vin = ops.concat_features(tr(sem), tr(scene), tr(motion), tr(emotion), torch.float32, model.total_vf_dim)
    vf = ops.linear(vin, model.Linear_vis.weight.detach(), model.Linear_vis.bias.detach())
    pos_v = model.positional_embedding_video.weight.detach()[:S].unsqueeze(1).expand(S, B, E).contiguous()
    vf = ops.axpy(vf.view(S, B, E), pos_v, 1.0)
    # the videos of the batch are independent generations: MultiheadGQA's literal (len, batch) -> (batch, len) `.view` would mix
    # them for B > 1, so the encoder runs in its batch-independent form (identical to the reference for every single video)
    gqa = [mod for mod in model.transformer.encoder.modules() if isinstance(mod, MultiheadGQA)]
    prev = [getattr(mod, "batch_independent", False) for mod in gqa]
    try:
        for mod in gqa:
            mod.batch_independent = True
        memory = model.transformer.encoder(vf)
    finally:
        for mod, pv in zip(gqa, prev):
            mod.batch_independent = pv
    if use_graph is None:                                # Python-side schedulers that advance in eval mode cannot be replayed
        use_graph = not any(hasattr(mod, "temperature_scheduler") for mod in model.modules())
    key_rows = key.reshape(B, -1)[:, 0].float().contiguous()
    cfg = (B, S, T, n0, beam, max_conseq_N, max_conseq_chord, float(temperature), bool(use_graph), str(dev),
           tuple((p.data_ptr(), p._version) for p in model.parameters()))
    sessions = _SESSIONS.setdefault(model, {})
    sess = sessions.get(cfg)
    if sess is None:
        sessions.clear()                                 # one live session per model: its buffers are the KV caches
        sess = sessions[cfg] = _GenSession(model, memory, T, n0, beam, max_conseq_N, max_conseq_chord, float(temperature))
    else:
        sess.dec.load_memory(memory)
    return sess.run(primer, primer_root, primer_attr, key_rows, uniforms, use_graph)


_SESSIONS = weakref.WeakKeyDictionary()      # model -> {configuration: _GenSession}; not part of the module's state (deepcopy-safe)


class _GenSession:
    """Persistent buffers of one generation configuration (token arrays, position words, KV caches, projected memory) and the
    CUDA graph of one position captured over them: the first generation captures, later ones only replay."""

    def __init__(self, model, memory, T, n0, beam, max_conseq_N, max_conseq_chord, temperature):
        S, B, E = memory.shape
        dev = memory.device
        self.model, self.B, self.T, self.E, self.n0 = model, B, T, E, n0
        self.beam, self.max_conseq_N, self.max_conseq_chord, self.temperature = beam, max_conseq_N, max_conseq_chord, temperature
        self.dec = CachedDecoder(model, memory, T)
        self.gen = torch.empty((B, T), dtype=torch.long, device=dev)
        self.gen_root, self.gen_attr = torch.empty_like(self.gen), torch.empty_like(self.gen)
        self.uniforms = torch.zeros((B, T), device=dev)
        self.key_rows = torch.zeros((B,), device=dev)
        # the position lives on the device: t (int64, index of the token that goes in), n_keys = t + 1 (int32, rows of the caches)
        self.t_dev = torch.zeros(1, dtype=torch.long, device=dev)
        self.n_keys = torch.ones(1, dtype=torch.int32, device=dev)
        self.ar_back = torch.arange(1, max(1, max_conseq_chord) + 1, device=dev).view(1, -1)     # 1 .. max_conseq_chord
        wc = model.Linear_chord.weight.detach()
        self.wkey = wc[:, E].contiguous()
        self.wc_main = wc[:, :E].contiguous()            # 16-byte aligned rows for the step kernel (the weight has 513 columns)
        self.graph, self.per_step = None, 0

    def step(self):
        """Token of position t in, token of position t + 1 out; then t += 1.  No host-visible value depends on t."""
        m, B, E = self.model, self.B, self.E
        gen, gen_root, gen_attr, t_dev = self.gen, self.gen_root, self.gen_attr, self.t_dev
        col = t_dev.view(1, 1).expand(B, 1)
        xin = ops.embed_sum(gen_root.gather(1, col).view(B), m.embedding_root.weight.detach(), gen_attr.gather(1, col).view(B),
                            m.embedding_attr.weight.detach(), torch.float32)
        x = ops.step_linear(xin, self.wc_main, m.Linear_chord.bias.detach(), k=E, row_scale=self.key_rows, col_vec=self.wkey)
        x = ops.axpy(x, m.positional_embedding.weight.detach().index_select(0, t_dev).expand(B, E).contiguous(), 1.0)
        h = self.dec.step(x, t_dev, self.n_keys)
        cur = col + 1                                    # (B, 1)
        logits = ops.step_linear(h, m.Wout.weight.detach(), m.Wout.bias.detach())
        probs = torch.softmax(logits / self.temperature, dim=-1)[:, :CHORD_END]
        old = gen.gather(1, cur)
        in_primer = cur < self.n0                        # primer positions only fill the caches
        if self.beam == 1:
            tok = torch.argmax(probs, dim=-1, keepdim=True)
        else:
            probs = probs.clone()
            if self.max_conseq_N == 0:
                probs[:, 0] = 0.0
            # the last max_conseq_chord tokens all equal the previous one -> it may not be drawn again
            back = (cur - self.ar_back).clamp_min(0)     # (B, max_conseq_chord)
            prev = gen.gather(1, cur - 1)
            rep = (gen.gather(1, back) == prev).all(dim=1, keepdim=True) & (cur >= self.max_conseq_chord)
            pidx = prev.clamp(0, CHORD_END - 1)
            probs.scatter_(1, pidx, torch.where(rep & (prev < CHORD_END), torch.zeros_like(probs[:, :1]), probs.gather(1, pidx)))
            cdf = torch.cumsum(probs / probs.sum(dim=1, keepdim=True), dim=1)
            tok = (cdf <= self.uniforms.gather(1, cur)).sum(dim=1, keepdim=True).clamp_max(CHORD_END - 1)
            root = torch.where(tok <= 0, torch.zeros_like(tok), (tok - 1) // 13 + 1)
            attr = torch.where(tok <= 0, torch.ones_like(tok), (tok - 1) % 13 + 1)
            gen_root.scatter_(1, cur, torch.where(in_primer, gen_root.gather(1, cur), root))
            gen_attr.scatter_(1, cur, torch.where(in_primer, gen_attr.gather(1, cur), attr))
        gen.scatter_(1, cur, torch.where(in_primer, old, tok))
        t_dev.add_(1)
        self.n_keys.add_(1)

    def run(self, primer, primer_root, primer_attr, key_rows, uniforms, use_graph) -> torch.Tensor:
        from . import _lib
        n0, T = self.n0, self.T
        dev = self.gen.device
        self.gen.fill_(CHORD_PAD); self.gen_root.fill_(CHORD_ROOT_PAD); self.gen_attr.fill_(CHORD_ATTR_PAD)
        self.gen[:, :n0], self.gen_root[:, :n0], self.gen_attr[:, :n0] = primer, primer_root, primer_attr
        self.key_rows.copy_(key_rows)
        if uniforms is not None:
            self.uniforms.copy_(uniforms)
        self.t_dev.zero_()
        self.n_keys.fill_(1)
        n_steps = T - 1
        if not use_graph or n_steps < 3:
            for _ in range(n_steps):
                self.step()
            return self.gen.clone()
        done = 0
        if self.graph is None:
            # one eager step (allocator / weight stacks warm), then the step is captured once; every later position -- and every
            # later generation of this configuration -- replays it
            self.step()
            done = 1
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(device=dev)
            graph = torch.cuda.CUDAGraph()
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                n_before = _lib.launches()
                with torch.cuda.graph(graph, stream=side):
                    self.step()
                self.per_step = _lib.launches() - n_before
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = graph
            _lib.count_launches(-self.per_step)          # the capture enqueued nothing
        for _ in range(n_steps - done):
            self.graph.replay()
        _lib.count_launches(self.per_step * (n_steps - done))
        return self.gen.clone()
