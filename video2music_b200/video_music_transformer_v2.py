"""Drop-in for `VideoMusicTransformer_V2` (model/video_music_transformer.py:317-610), the reference's shipped inference
default (version '2.2', argument_generate_funcs.py:82): root + attribute embeddings and the key column through
`Linear_chord`, the concatenated per-second video features through `Linear_vis`, then an encoder / decoder built from the
generic wrappers of custom_transformer.py -- three "shallow" layers (CustomMultiheadAttention with RoPE + GLUExpert) followed
by `n_layers - 3` "deep" layers whose feed-forward is a SharedMoELayer (6 experts, top-2) -- and `Wout`.

Same constructor signature, attribute and `state_dict` names (`transformer.encoder.layers.N.self_attn.in_proj_weight`,
`...ff.experts.M.linear1.weight`, ...), so a reference checkpoint loads.  Inference only: eval-mode forward and `generate`
(the literal loop of the reference: one full forward per generated token, batch of one).  Versions '2.0' (learned positional
embeddings), '2.1' and '2.2' are built; '2.3' needs efficient_kan (out of scope), `scene_embed`, `dropTokenRate` > 0 in
training and the logging side channel `get_highest_emotion_indices` are not reproduced.
"""
import copy
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .custom_transformer import (CustomMultiheadAttention, RotaryPositionalEmbeddings, TransformerDecoderLayer,
                                 TransformerDecoderShorter, TransformerEncoderLayer, TransformerEncoderShorter)
from .moe import GLUExpert, SharedMoELayer, TopKScheduler
from .video_music_transformer import CHORD_ATTR_SIZE, CHORD_END, CHORD_PAD, CHORD_ROOT_SIZE, CHORD_SIZE

CHORD_ROOT_PAD, CHORD_ATTR_PAD = 14, 15            # utilities/constants.py


def chord_root_attr(c: int):
    """chord id -> (root id, attribute id): closed form of dataset/vevo_meta/chord_inv.json + chord_root.json + chord_attr.json
    (video_music_transformer.py:585-600; checked against those JSON tables in the CPU tests)."""
    return (0, 1) if c <= 0 else ((c - 1) // 13 + 1, (c - 1) % 13 + 1)


class _Transformer(nn.Module):
    """Parameter layout of nn.Transformer(custom_encoder=..., custom_decoder=...): `encoder.*`, `decoder.*`."""

    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder

    def forward(self, src, tgt, tgt_mask=None):
        return self.decoder(tgt, self.encoder(src), tgt_mask=tgt_mask)


class _ZooModel(nn.Module):
    """forward / generate shared by the V1 and V2 classes (their bodies are the same in the reference up to the positional
    scheme: learned position tables in V1 and V2 '2.0', RoPE inside the attention otherwise)."""

    def _embeddings(self, d_model, total_vf_dim, max_sequence_chord, max_sequence_video, pos_tables: bool):
        self.embedding = nn.Embedding(CHORD_SIZE, d_model)
        self.embedding_root = nn.Embedding(CHORD_ROOT_SIZE, d_model)
        self.embedding_attr = nn.Embedding(CHORD_ATTR_SIZE, d_model)
        self.total_vf_dim = total_vf_dim
        self.Linear_vis = nn.Linear(total_vf_dim, d_model)
        self.Linear_chord = nn.Linear(d_model + 1, d_model)
        if pos_tables:
            self.positional_embedding = nn.Embedding(max_sequence_chord, d_model)
            self.positional_embedding_video = nn.Embedding(max_sequence_video, d_model)
        self.condition_linear = nn.Linear(1, d_model)
        self._pos_tables = pos_tables


class VideoMusicTransformer_V2(_ZooModel):
    def __init__(self, version_name='2.0', n_layers=6, num_heads=8, d_model=512, dim_feedforward=1024, dropout=0.1,
                 max_sequence_midi=2048, max_sequence_video=300, max_sequence_chord=300, total_vf_dim=0, rms_norm=False,
                 scene_embed=False, chord_embed=False, dropTokenRate=0.0, balancing=False):
        super().__init__()
        if version_name not in ('2.0', '2.1', '2.2'):
            raise NotImplementedError("version %r (2.3 needs efficient_kan: out of scope)" % (version_name,))
        if scene_embed or chord_embed:
            raise NotImplementedError("scene_embed / chord_embed are not built for the V2 model")
        self.nlayers, self.nhead, self.d_model, self.d_ff, self.dropout = n_layers, num_heads, d_model, dim_feedforward, dropout
        self.max_seq_midi, self.max_seq_video, self.max_seq_chord = max_sequence_midi, max_sequence_video, max_sequence_chord
        self.scene_embed, self.dropTokenRate, self.chord_embed, self.version_name = scene_embed, dropTokenRate, chord_embed, version_name
        self._embeddings(d_model, total_vf_dim, max_sequence_chord, max_sequence_video, pos_tables=version_name == '2.0')
        norm = nn.LayerNorm(d_model)
        RoPE = None if version_name == '2.0' else RotaryPositionalEmbeddings(d_model, max_sequence_video)   # :379
        self.n_experts, self.n_experts_per_token = 6, 2
        expert = GLUExpert(d_model, dim_feedforward, dropout)
        att = CustomMultiheadAttention(d_model, num_heads, dropout, RoPE=RoPE)
        topk = None if version_name == '2.2' else TopKScheduler(n_experts=6, min_n_experts_per_token=2, update_step=32)
        moelayer = SharedMoELayer(expert=expert, d_model=d_model, n_experts=6, n_experts_per_token=2, dropout=dropout,
                                  balancing=balancing, topk_scheduler=topk, temperature_scheduler=None, use_KAN=False)
        swiglu = GLUExpert(d_model, dim_feedforward, dropout)
        mk_e = lambda ff: TransformerEncoderLayer(att, ff, pre_norm=False, norm=norm, dropout=dropout)
        mk_d = lambda ff: TransformerDecoderLayer(att, att, ff, pre_norm=False, norm=norm, dropout=dropout)
        rate = 3
        enc = nn.ModuleList([copy.deepcopy(mk_e(swiglu)) for _ in range(rate)] + [copy.deepcopy(mk_e(moelayer)) for _ in range(n_layers - rate)])
        dec = nn.ModuleList([copy.deepcopy(mk_d(swiglu)) for _ in range(rate)] + [copy.deepcopy(mk_d(moelayer)) for _ in range(n_layers - rate)])
        self.transformer = _Transformer(TransformerEncoderShorter(enc, norm), TransformerDecoderShorter(dec, norm))
        self.Wout = nn.Linear(d_model, CHORD_SIZE)
        self.softmax = nn.Softmax(dim=-1)

    # ------------------------------------------------------------------ forward (video_music_transformer.py:437-520)
    def forward(self, x, x_root, x_attr, feature_semantic_list, feature_key, feature_scene_offset, feature_motion, feature_emotion,
                mask=True):
        dev = self.Wout.weight.device
        x_root, x_attr, sem, key, scene, motion, emotion = (t.to(dev) for t in (x_root, x_attr, feature_semantic_list, feature_key,
                                                                                feature_scene_offset, feature_motion, feature_emotion))
        B, T = x_root.shape
        S, E = sem.shape[1], self.d_model
        from . import autograd as ag
        if ag.tracking(self):
            # training (fp32): the same kernels inside autograd Functions -- embeddings + key column, Linear_chord / Linear_vis / Wout
            # through LinearFn, learned position tables, and the attention / feed-forward modules' own training paths
            if getattr(self, "dropTokenRate", 0.0) != 0.0:
                raise NotImplementedError("dropTokenRate > 0 (video_music_transformer.py:486-491) is not built")
            F32 = torch.float32
            key_rows = key.reshape(B, -1)[:, 0].float().unsqueeze(0).expand(T, B).reshape(-1).contiguous()
            xin = ag.EmbedKeyFn.apply(x_root.t().reshape(-1).contiguous(), self.embedding_root.weight, x_attr.t().reshape(-1).contiguous(),
                                      self.embedding_attr.weight, key_rows, F32)                    # (T*B, pad8(E + 1)): [emb | key | 0]
            wc = self.Linear_chord.weight
            xf = ag.LinearFn.apply(xin, wc, self.Linear_chord.bias, wc, E + 1, False, 1.0, 0, None, 0, F32, None)
            tr = lambda t: t.transpose(0, 1).contiguous()
            vin = ops.concat_features(tr(sem), tr(scene), tr(motion), tr(emotion), F32, self.total_vf_dim)
            vf = ag.linear_fn(vin, self.Linear_vis)
            if self._pos_tables:                                                # learned positions (:496-505, :196-201)
                xf = xf.view(T, B, E) + self.positional_embedding.weight[:T].unsqueeze(1)
                vf = vf.view(S, B, E) + self.positional_embedding_video.weight[:S].unsqueeze(1)
            tgt_mask = torch.triu(torch.full((T, T), float("-inf"), device=dev), diagonal=1) if mask is True else None
            out = self.transformer(src=vf.reshape(S, B, E), tgt=xf.reshape(T, B, E), tgt_mask=tgt_mask)
            y = ag.linear_fn(out.reshape(T * B, E).contiguous(), self.Wout)
            return y.view(T, B, CHORD_SIZE).permute(1, 0, 2).contiguous()
        # chords, sequence-first rows (t, b): (emb_root + emb_attr | key) -> Linear_chord; the key column is a rank-1 epilogue term
        xin = ops.embed_sum(x_root.t().reshape(-1), self.embedding_root.weight.detach(), x_attr.t().reshape(-1),
                            self.embedding_attr.weight.detach(), torch.float32)
        key_rows = key.reshape(B, -1)[:, 0].float().unsqueeze(0).expand(T, B).reshape(-1).contiguous()
        wc = self.Linear_chord.weight.detach()
        xf = ops.linear(xin, wc, self.Linear_chord.bias.detach(), k=E, row_scale=key_rows, col_vec=wc[:, E].contiguous())
        # video features, rows (s, b)
        tr = lambda t: t.transpose(0, 1).contiguous()
        vin = ops.concat_features(tr(sem), tr(scene), tr(motion), tr(emotion), torch.float32, self.total_vf_dim)
        vf = ops.linear(vin, self.Linear_vis.weight.detach(), self.Linear_vis.bias.detach())
        if self._pos_tables:                                                    # learned positions (:496-505, :196-201)
            xf = ops.axpy(xf.view(T, B, E), self.positional_embedding.weight.detach()[:T].unsqueeze(1).expand(T, B, E).contiguous(), 1.0)
            vf = ops.axpy(vf.view(S, B, E), self.positional_embedding_video.weight.detach()[:S].unsqueeze(1).expand(S, B, E).contiguous(), 1.0)
        tgt_mask = torch.triu(torch.full((T, T), float("-inf"), device=dev), diagonal=1) if mask is True else None
        out = self.transformer(src=vf.view(S, B, E), tgt=xf.view(T, B, E), tgt_mask=tgt_mask)      # (T, B, E)
        y = ops.linear(out.reshape(T * B, E).contiguous(), self.Wout.weight.detach(), self.Wout.bias.detach())
        return y.view(T, B, CHORD_SIZE).permute(1, 0, 2).contiguous()

    # ------------------------------------------------------------------ generate (video_music_transformer.py:522-610)
    @torch.no_grad()
    def generate(self, feature_semantic_list=[], feature_key=None, feature_scene_offset=None, feature_motion=None,
                 feature_emotion=None, primer=None, primer_root=None, primer_attr=None, target_seq_length=300, beam=0,
                 beam_chance=1.0, max_conseq_N=0, max_conseq_chord=2, temperature=1.0, uniforms: Optional[torch.Tensor] = None):
        """The reference's loop: batch of one, one full forward per token.  beam=1 (beam_chance >= 1): arg-max over the first
        157 classes, root / attribute inputs of generated positions stay PAD (literal); beam=0: the sampling branch with its
        no-"N" / no-repeat constraints, drawn by inverse CDF from `uniforms` (or torch.rand), root / attribute updated."""
        assert (not self.training), "Cannot generate while in training mode"
        if not (beam == 0 or (beam == 1 and beam_chance >= 1.0)):
            raise NotImplementedError("beam > 1 / 0 < beam_chance < 1 are not reproduced")
        dev = self.Wout.weight.device
        gen = torch.full((1, target_seq_length), CHORD_PAD, dtype=torch.long, device=dev)
        gen_root = torch.full((1, target_seq_length), CHORD_ROOT_PAD, dtype=torch.long, device=dev)
        gen_attr = torch.full((1, target_seq_length), CHORD_ATTR_PAD, dtype=torch.long, device=dev)
        n0 = len(primer)
        gen[..., :n0], gen_root[..., :n0], gen_attr[..., :n0] = primer.long().to(dev), primer_root.long().to(dev), primer_attr.long().to(dev)
        if beam == 0 and uniforms is None:
            uniforms = torch.rand(target_seq_length, device=dev)
        cur = n0
        while cur < target_seq_length:
            logits = self.forward(gen[..., :cur], gen_root[..., :cur], gen_attr[..., :cur], feature_semantic_list, feature_key,
                                  feature_scene_offset, feature_motion, feature_emotion)
            probs = torch.softmax(logits[0, cur - 1] / temperature, dim=-1)[:CHORD_END]
            if beam == 1:
                gen[0, cur] = int(torch.argmax(probs))
            else:
                probs = probs.clone()
                if max_conseq_N == 0:
                    probs[0] = 0.0
                if cur >= max_conseq_chord and bool((gen[0, cur - max_conseq_chord:cur] == gen[0, cur - 1]).all()):
                    probs[int(gen[0, cur - 1])] = 0.0
                cdf = torch.cumsum(probs / probs.sum(), dim=0)
                tok = min(int((cdf <= float(uniforms[cur])).sum()), CHORD_END - 1)
                gen[0, cur] = tok
                r, a = chord_root_attr(tok)
                gen_root[0, cur], gen_attr[0, cur] = r, a
            cur += 1
        return gen[:, :cur]


class VideoMusicTransformer_V1(_ZooModel):
    """Drop-in for `VideoMusicTransformer_V1` (model/video_music_transformer.py:22-315), versions '1.1' (MoELayer) and '1.3'
    (SharedMoELayer): stock multi-head attention, GLU experts (6, top-2) in every layer of both stacks, learned position tables,
    LayerNorm or RMSNorm (`rms_norm`).  The versions with nn.Sequential SiLU experts or RoPE ('1.0', '1.2.x', '1.3.3', '1.3.4')
    are not built.  Inference and (fp32, autograd over the same kernels) training."""

    def __init__(self, version_name='1.1', n_layers=6, num_heads=8, d_model=512, dim_feedforward=1024, dropout=0.1,
                 max_sequence_midi=2048, max_sequence_video=300, max_sequence_chord=300, total_vf_dim=0, rms_norm=False,
                 scene_embed=False, chord_embed=False, dropTokenRate=0.0):
        super().__init__()
        if version_name not in ('1.1', '1.3'):
            raise NotImplementedError("version %r (built: '1.1', '1.3')" % (version_name,))
        if scene_embed or chord_embed:
            raise NotImplementedError("scene_embed / chord_embed are not built for the V1 model")
        from .custom_transformer import RMSNorm, TransformerDecoder, TransformerEncoder
        from .moe import MoELayer
        self.nlayers, self.nhead, self.d_model, self.d_ff, self.dropout = n_layers, num_heads, d_model, dim_feedforward, dropout
        self.max_seq_midi, self.max_seq_video, self.max_seq_chord = max_sequence_midi, max_sequence_video, max_sequence_chord
        self.scene_embed, self.dropTokenRate, self.chord_embed, self.version_name = scene_embed, dropTokenRate, chord_embed, version_name
        self._embeddings(d_model, total_vf_dim, max_sequence_chord, max_sequence_video, pos_tables=True)
        norm = RMSNorm(d_model) if rms_norm else nn.LayerNorm(d_model)
        self.n_experts, self.n_experts_per_token = 6, 2
        expert = GLUExpert(d_model, dim_feedforward, dropout)
        att = CustomMultiheadAttention(d_model, num_heads, dropout)             # == nn.MultiheadAttention (same parameters)
        if version_name == '1.1':
            moelayer = MoELayer(expert, d_model, 6, 2, dropout)
        else:
            moelayer = SharedMoELayer(expert, d_model, n_experts=6, n_experts_per_token=2, balancing=False, dropout=dropout)
        encoder = TransformerEncoder(TransformerEncoderLayer(att, moelayer, pre_norm=False, norm=norm, dropout=dropout), n_layers, norm)
        decoder = TransformerDecoder(TransformerDecoderLayer(att, att, moelayer, pre_norm=False, norm=norm, dropout=dropout), n_layers, norm)
        self.transformer = _Transformer(encoder, decoder)
        self.Wout = nn.Linear(d_model, CHORD_SIZE)
        self.softmax = nn.Softmax(dim=-1)


class VideoMusicTransformer_GQA(_ZooModel):
    """BASELINE config 4: the AMT built from the reference's own blocks with grouped-query attention and MoE feed-forwards --
    `TransformerEncoderLayer / TransformerDecoderLayer(att=MultiheadGQA(d_model, num_heads, kv_heads), ff=MoELayer | SharedMoELayer
    (GLUExpert, 6 experts, top-2))` (custom_transformer.py:1220-1292, grouped_query_attention.py:172-358, moe.py:150-302) inside
    the V1 shell (learned position tables, root + attribute embeddings, video_music_transformer.py:77-118).  The reference ships
    no such class; it is the composition SURVEY.md 8d defines.  One deliberate difference from a literal composition: decoder
    self-attention is causal (`force_causal`), because MultiheadGQA drops the mask its wrapper passes (see there) -- without
    it next-chord training would see the future and generation would have no KV cache."""

    def __init__(self, n_layers=6, num_heads=8, kv_heads=2, d_model=512, dim_feedforward=1024, dropout=0.1, max_sequence_video=300,
                 max_sequence_chord=300, total_vf_dim=0, shared_moe=False, rms_norm=False, pre_norm=False):
        super().__init__()
        from .custom_transformer import RMSNorm, TransformerDecoder, TransformerEncoder
        from .grouped_query_attention import MultiheadGQA
        from .moe import MoELayer
        self.nlayers, self.nhead, self.d_model, self.d_ff, self.dropout = n_layers, num_heads, d_model, dim_feedforward, dropout
        self.max_seq_video, self.max_seq_chord = max_sequence_video, max_sequence_chord
        self._embeddings(d_model, total_vf_dim, max_sequence_chord, max_sequence_video, pos_tables=True)
        norm = RMSNorm(d_model) if rms_norm else nn.LayerNorm(d_model)
        self.n_experts, self.n_experts_per_token = 6, 2
        expert = GLUExpert(d_model, dim_feedforward, dropout)
        att = MultiheadGQA(d_model, num_heads, kv_heads, dropout=dropout)
        if shared_moe:
            moelayer = SharedMoELayer(expert, d_model, n_experts=6, n_experts_per_token=2, balancing=False, dropout=dropout)
        else:
            moelayer = MoELayer(expert, d_model, 6, 2, dropout)
        encoder = TransformerEncoder(TransformerEncoderLayer(att, moelayer, pre_norm=pre_norm, norm=norm, dropout=dropout), n_layers, norm)
        decoder = TransformerDecoder(TransformerDecoderLayer(att, att, moelayer, pre_norm=pre_norm, norm=norm, dropout=dropout), n_layers, norm)
        for layer in decoder.layers:
            layer.self_attn.force_causal = True
        self.transformer = _Transformer(encoder, decoder)
        self.Wout = nn.Linear(d_model, CHORD_SIZE)
        self.softmax = nn.Softmax(dim=-1)


class VideoMusicTransformer_V3(_ZooModel):
    """Drop-in for `VideoMusicTransformer_V3` (model/video_music_transformer.py:611-905), versions '3.0', '3.1', '3.2': RMSNorm,
    a RoPE cache of dimension 2 * d_model, DifferentialMultiheadAttention (depth = layer index) in every decoder layer and --
    for 3.1 / 3.2 -- in the encoder ('3.0' keeps CustomMultiheadAttention there), three GLUExpert layers then SharedMoELayer
    layers (balancing buffers registered as in the reference), pre-norm wrappers for '3.2'.  Inference and fp32 training (dropout 0 in the differential attention)."""

    def __init__(self, version_name='3.0', n_layers=6, num_heads=8, d_model=512, dim_feedforward=1024, dropout=0.1,
                 max_sequence_midi=2048, max_sequence_video=300, max_sequence_chord=300, total_vf_dim=0, rms_norm=False,
                 scene_embed=False, chord_embed=False, dropTokenRate=0.0):
        super().__init__()
        if version_name not in ('3.0', '3.1', '3.2'):
            raise NotImplementedError("version %r (built: '3.0', '3.1', '3.2')" % (version_name,))
        if scene_embed or chord_embed:
            raise NotImplementedError("scene_embed / chord_embed are not built for the V3 model")
        from .custom_transformer import DifferentialMultiheadAttention, RMSNorm
        self.nlayers, self.nhead, self.d_model, self.d_ff, self.dropout = n_layers, num_heads, d_model, dim_feedforward, dropout
        self.max_seq_midi, self.max_seq_video, self.max_seq_chord = max_sequence_midi, max_sequence_video, max_sequence_chord
        self.scene_embed, self.dropTokenRate, self.chord_embed, self.version_name = scene_embed, dropTokenRate, chord_embed, version_name
        self._embeddings(d_model, total_vf_dim, max_sequence_chord, max_sequence_video, pos_tables=False)
        RoPE = RotaryPositionalEmbeddings(d_model * 2, max_sequence_video)      # video_music_transformer.py:662
        norm = RMSNorm(d_model, elementwise_affine=True)
        self.n_experts, self.n_experts_per_token = 6, 2
        expert = GLUExpert(d_model, dim_feedforward, dropout)
        moelayer = SharedMoELayer(expert=expert, d_model=d_model, n_experts=6, n_experts_per_token=2, dropout=dropout, balancing=True,
                                  topk_scheduler=None, temperature_scheduler=None, use_KAN=False)
        swiglu = GLUExpert(d_model, dim_feedforward, dropout)
        att = CustomMultiheadAttention(d_model, num_heads, dropout=dropout, RoPE=RoPE)
        difatt = [DifferentialMultiheadAttention(d_model, num_heads, dropout=dropout, RoPE=RoPE, depth=d) for d in range(n_layers)]
        rate, pre_norm = 3, version_name == '3.2'
        ffs = [swiglu] * rate + [moelayer] * (n_layers - rate)
        enc_att = [att] * n_layers if version_name == '3.0' else difatt
        enc = nn.ModuleList([TransformerEncoderLayer(enc_att[i], ffs[i], pre_norm=pre_norm, norm=norm, dropout=dropout) for i in range(n_layers)])
        dec = nn.ModuleList([TransformerDecoderLayer(difatt[i], difatt[i], ffs[i], pre_norm=pre_norm, norm=norm, dropout=dropout)
                             for i in range(n_layers)])
        self.transformer = _Transformer(TransformerEncoderShorter(enc, norm), TransformerDecoderShorter(dec, norm))
        self.Wout = nn.Linear(d_model, CHORD_SIZE)
        self.softmax = nn.Softmax(dim=-1)


_ZooModel.forward = VideoMusicTransformer_V2.forward
_ZooModel.generate = VideoMusicTransformer_V2.generate


def _generate_cached(self, *args, **kwargs):
    """KV-cached, batched generation (cached_decode.generate_cached): same tokens as `generate` video by video for the models
    whose decoder is cacheable (V1 '1.1' / '1.3', V2 '2.0', the GQA + MoE shell), O(n) instead of O(n^2) per video."""
    from .cached_decode import generate_cached
    return generate_cached(self, *args, **kwargs)


_ZooModel.generate_cached = _generate_cached
