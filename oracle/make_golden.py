"""Regenerates tests/golden/*.pt from the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in a container that has the reference mounted:

    python oracle/make_golden.py [--only NAME] [--gen-videos 4]

The reference has no golden vectors of its own, so these
fixtures are the pin: inputs and weights are regenerated from the seeds stored in
each fixture (video2music_b200/synthetic.py), only the reference's OUTPUTS are
stored.  `weights_checksum` lets a consumer prove it rebuilt the same weights.
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference, reference_cwd  # noqa: E402
from video2music_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _save(name, obj):
    path = os.path.join(GOLD, name)
    torch.save(obj, path)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def _load_weights(module, seed, wout_gain=1.0):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = syn.fill_like_reference_init(shapes, seed=seed, wout_gain=wout_gain)
    module.load_state_dict(sd, strict=False)
    return sd


def _amt(ref, vf, chord_embed=False, seed=0, wout_gain=1.0, **kw):
    with contextlib.redirect_stdout(io.StringIO()), reference_cwd():
        m = ref.vmt.VideoMusicTransformer(total_vf_dim=vf, rpr=True, chord_embed=chord_embed, **kw).eval()
    sd = _load_weights(m, seed, wout_gain)
    return m, sd


def golden_forward(ref):
    """BASELINE config 1: base AMT forward, B=4, T=299, S=300, eval, fp32 CPU."""
    spec = dict(batch=4, tgt_len=299, src_len=300, motion_type=0, input_seed=1234, weight_seed=0)
    m, sd = _amt(ref, syn.vf_dim(0), seed=spec["weight_seed"])
    inp = syn.make_inputs(spec["batch"], spec["input_seed"], spec["tgt_len"], spec["src_len"], 0)
    with torch.no_grad():
        y = m(inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
              inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"])
    _save("amt_forward_cfg1.pt", dict(spec=spec, weights_checksum=syn.checksum(sd), logits=y.contiguous()))
    # a ragged/short case (T=37, S=120, B=3, motion_type 1 -> vf 1287) and the no-mask path
    spec2 = dict(batch=3, tgt_len=37, src_len=120, motion_type=1, input_seed=99, weight_seed=5)
    m2, sd2 = _amt(ref, syn.vf_dim(1), seed=5)
    inp = syn.make_inputs(3, 99, 37, 120, 1)
    with torch.no_grad():
        y2 = m2(inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
                inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"])
    _save("amt_forward_small.pt", dict(spec=spec2, weights_checksum=syn.checksum(sd2), logits=y2.contiguous()))


def _train_step_record(ref, spec, keep, pad_tail=False):
    """One training step of the unmodified reference (run_model_vevo.py:84-121): loss, every parameter's gradient norm
    and the gradients named in `keep`.  `pad_tail`: the last positions of each target row are CHORD_PAD (158), which
    CrossEntropyLoss(ignore_index) drops (train.py:222) -- the ragged-target case of the data loader."""
    with contextlib.redirect_stdout(io.StringIO()), reference_cwd():
        m = ref.vmt.VideoMusicTransformer(total_vf_dim=syn.vf_dim(spec["motion_type"]), rpr=True, dropout=0.0).train()
    sd = _load_weights(m, spec["weight_seed"])
    inp = syn.make_inputs(spec["batch"], spec["input_seed"], spec["tgt_len"], spec["src_len"], spec["motion_type"])
    if pad_tail:
        inp["tgt"] = syn.pad_targets(inp["tgt"], spec["input_seed"])
    y = m(inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
          inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"])
    ce = torch.nn.CrossEntropyLoss(ignore_index=158, label_smoothing=0.1)          # train.py:222
    bce = torch.nn.BCEWithLogitsLoss()                                             # train.py:233
    loss_chord = ce(y.permute(0, 2, 1), inp["tgt"])                                 # run_model_vevo.py:101
    loss_emotion = bce(y.permute(0, 2, 1), inp["tgt_emotion"].permute(0, 2, 1))     # :102
    total = 0.4 * loss_chord + 0.6 * loss_emotion                                   # :119, LOSS_LAMBDA
    total.backward()
    grads = {}
    norms = {}
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        norms[n] = float(p.grad.double().norm())
        if n in keep:
            grads[n] = p.grad.clone()
    return dict(spec=spec, weights_checksum=syn.checksum(sd), logits=y.detach().clone(),
                loss=float(total), loss_chord=float(loss_chord), loss_emotion=float(loss_emotion),
                grad_norms=norms, grads=grads, pad_tail=pad_tail)


TRAIN_KEEP = ("transformer.decoder.layers.0.self_attn.Er", "transformer.decoder.layers.5.self_attn.Er",
              "Wout.bias", "transformer.decoder.layers.3.norm2.weight", "embedding_root.weight",
              "transformer.encoder.layers.0.self_attn.in_proj_bias", "Linear_chord.bias")


def golden_train(ref):
    """One training step's loss and gradients (run_model_vevo.py:84-121), dropout 0, fp32 CPU; small shape."""
    spec = dict(batch=2, tgt_len=64, src_len=300, motion_type=0, input_seed=4321, weight_seed=3)
    _save("amt_train_step.pt", _train_step_record(ref, spec, TRAIN_KEEP))


def golden_train_full(ref):
    """BASELINE config 3 at the real sequence shape: B=4, T=299, S=300 (multi-tile attention backward, the whole Er band,
    the T=299 tails), once with PAD-free targets and once with ragged PAD tails.  Logits are stored for video 0 only."""
    keep = TRAIN_KEEP + ("transformer.decoder.layers.2.self_attn.in_proj_weight",
                         "transformer.decoder.layers.4.multihead_attn.in_proj_bias",
                         "transformer.decoder.layers.1.self_attn.Er")
    out = {}
    for name, pad in (("full", False), ("ragged", True)):
        spec = dict(batch=4, tgt_len=299, src_len=300, motion_type=0, input_seed=2025 + pad, weight_seed=11)
        r = _train_step_record(ref, spec, keep, pad_tail=pad)
        r["logits"] = r["logits"][:1].clone()
        # the 1536x512 in_proj gradient is 3 MB: keep every 8th row
        k = "transformer.decoder.layers.2.self_attn.in_proj_weight"
        r["grads"][k] = r["grads"][k][::8].clone()
        out[name] = r
        print(name, "loss", r["loss"], len(r["grad_norms"]), "gradient norms")
    _save("amt_train_step_full.pt", out)


def golden_generate(ref, n_videos):
    """BASELINE config 2 oracle: reference generate(beam=1) run video by video (batch 1)."""
    out = {}
    for chord_embed in (False, True):
        spec = dict(n_videos=n_videos, target_seq_length=300, motion_type=0, input_seed=2024, weight_seed=1,
                    chord_embed=chord_embed, primer=[1], primer_root=[1], primer_attr=[0], wout_gain=4.0)
        m, sd = _amt(ref, syn.vf_dim(0), chord_embed=chord_embed, seed=1, wout_gain=4.0)
        inp = syn.make_inputs(n_videos, 2024, 299, 300, 0)
        prim, pr, pa = (torch.tensor(spec[k]) for k in ("primer", "primer_root", "primer_attr"))
        seqs = []
        t0 = time.time()
        for b in range(n_videos):
            with torch.no_grad(), reference_cwd(), contextlib.redirect_stdout(io.StringIO()):
                g = m.generate(inp["feature_semantic_list"][b:b + 1], inp["feature_key"][b],
                               inp["feature_scene_offset"][b:b + 1], inp["feature_motion"][b:b + 1],
                               inp["feature_emotion"][b:b + 1], primer=prim, primer_root=pr, primer_attr=pa,
                               target_seq_length=300, beam=1, beam_chance=1.0)
            seqs.append(g[0].clone())
            print("generate chord_embed=%s video %d: %.1fs" % (chord_embed, b, time.time() - t0), flush=True)
        extra = {}
        if chord_embed:
            extra["chord_embedding_weight_checksum"] = float(m.chord_embedding_model.weight.double().abs().sum())
        out["chord_embed_%s" % chord_embed] = dict(spec=spec, weights_checksum=syn.checksum(sd),
                                                   tokens=torch.stack(seqs), seconds=time.time() - t0, **extra)
    _save("amt_generate_greedy.pt", out)


def golden_generate_primed(ref, n_videos=2):
    """Same loop with a 120-chord random primer (teacher-forced prefix makes every cached K/V row distinct)."""
    out = {}
    for chord_embed in (False, True):
        spec = dict(n_videos=n_videos, target_seq_length=300, motion_type=0, input_seed=777, weight_seed=2,
                    chord_embed=chord_embed, primer_len=120, wout_gain=4.0)
        m, sd = _amt(ref, syn.vf_dim(0), chord_embed=chord_embed, seed=2, wout_gain=4.0)
        inp = syn.make_inputs(n_videos, 777, 299, 300, 0)
        seqs = []
        t0 = time.time()
        for b in range(n_videos):
            prim, pr, pa = inp["x"][b, :120], inp["x_root"][b, :120], inp["x_attr"][b, :120]
            with torch.no_grad(), reference_cwd(), contextlib.redirect_stdout(io.StringIO()):
                g = m.generate(inp["feature_semantic_list"][b:b + 1], inp["feature_key"][b],
                               inp["feature_scene_offset"][b:b + 1], inp["feature_motion"][b:b + 1],
                               inp["feature_emotion"][b:b + 1], primer=prim, primer_root=pr, primer_attr=pa,
                               target_seq_length=300, beam=1, beam_chance=1.0)
            seqs.append(g[0].clone())
            print("primed generate chord_embed=%s video %d: %.1fs" % (chord_embed, b, time.time() - t0), flush=True)
        out["chord_embed_%s" % chord_embed] = dict(spec=spec, weights_checksum=syn.checksum(sd), tokens=torch.stack(seqs))
    _save("amt_generate_primed.pt", out)


def golden_rpr(ref):
    """MultiheadAttentionRPR module (rpr.py:112-424) on ragged shapes, incl. returned weights."""
    cases = []
    for (L, B, E, H, er_len, seed) in [(37, 3, 128, 4, 64, 11), (299, 1, 512, 8, 300, 12), (150, 2, 512, 8, 300, 13),
                                       (1, 2, 128, 2, 16, 14)]:
        torch.manual_seed(0)
        mod = ref.rpr.MultiheadAttentionRPR(E, H, dropout=0.0, er_len=er_len).eval()
        sd = _load_weights(mod, seed)
        g = syn._gen(seed, "x")
        x = syn.unit_uniform((L, B, E), g)
        mask = torch.triu(torch.full((L, L), float("-inf")), diagonal=1)
        x.requires_grad_(True)
        out, w = mod(x, x, x, attn_mask=mask)
        gy = syn.unit_uniform((L, B, E), syn._gen(seed, "gy"))
        (out * gy).sum().backward()
        cases.append(dict(spec=dict(L=L, B=B, E=E, H=H, er_len=er_len, seed=seed), weights_checksum=syn.checksum(sd),
                          out=out.detach().clone(), weights_mean=w.detach().clone() if L <= 64 else None,
                          grad_x=x.grad.clone(), grad_Er=mod.Er.grad.clone(),
                          grad_in_proj_weight_norm=float(mod.in_proj_weight.grad.double().norm()),
                          grad_in_proj_bias=mod.in_proj_bias.grad.clone()))
    # decoder layer (rpr.py:37-70)
    torch.manual_seed(0)
    layer = ref.rpr.TransformerDecoderLayerRPR(256, 4, 512, 0.0, er_len=80).eval()
    sd = _load_weights(layer, 21)
    tgt = syn.unit_uniform((50, 3, 256), syn._gen(21, "tgt"))
    mem = syn.unit_uniform((70, 3, 256), syn._gen(21, "mem"))
    mask = torch.triu(torch.full((50, 50), float("-inf")), diagonal=1)
    with torch.no_grad():
        y = layer(tgt, mem, tgt_mask=mask)
    _save("rpr_attention.pt", dict(cases=cases, layer=dict(spec=dict(T=50, S=70, B=3, E=256, H=4, ff=512, er_len=80, seed=21),
                                                          weights_checksum=syn.checksum(sd), out=y.clone())))


def golden_moe(ref):
    """MoELayer / SharedMoELayer (moe.py:150-302) with GLUExpert(512,1024), 6 experts, top-2, eval."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for shared in (False, True):
        torch.manual_seed(0)
        exp = ref.moe.GLUExpert(512, 1024, 0.0)
        cls = ref.moe.SharedMoELayer if shared else ref.moe.MoELayer
        mod = cls(exp, 512, n_experts=6, n_experts_per_token=2, dropout=0.0).eval()
        sd = _load_weights(mod, 31 + shared)
        x = syn.unit_uniform((75, 4, 512), syn._gen(31, "x"))
        with torch.no_grad():
            y = mod(x)
            logits = mod.gate(x)
            w, idx = torch.topk(logits, 2)
        top3 = torch.topk(logits, 3).values
        out["shared_%s" % shared] = dict(spec=dict(L=75, B=4, d=512, ff=1024, n_experts=6, k=2, seed=31 + shared, x_seed=31),
                                         weights_checksum=syn.checksum(sd), out=y.clone(), selected_experts=idx.clone(),
                                         gate_logits=logits.clone(), min_rank_gap=float((top3[..., :-1] - top3[..., 1:]).min()))
    _save("moe.pt", out)


def golden_gqa(ref):
    """scaled_dot_product_gqa (grouped_query_attention.py:19-170) and MultiheadGQA (:172-358)."""
    out = {}
    fn_cases = []
    for (b, n, s, hq, hk, d, causal, seed) in [(2, 33, 33, 8, 2, 64, True, 41), (3, 20, 45, 8, 4, 64, False, 42),
                                               (1, 300, 300, 8, 1, 64, True, 43)]:
        q = syn.unit_uniform((b, n, hq, d), syn._gen(seed, "q"))
        k = syn.unit_uniform((b, s, hk, d), syn._gen(seed, "k"))
        v = syn.unit_uniform((b, s, hk, d), syn._gen(seed, "v"))
        with torch.no_grad():
            o, _ = ref.gqa.scaled_dot_product_gqa(q, k, v, num_heads=hq, is_causal=True if causal else None)
        fn_cases.append(dict(spec=dict(b=b, n=n, s=s, hq=hq, hk=hk, d=d, causal=causal, seed=seed), out=o.clone()))
    out["function"] = fn_cases
    mod_cases = []
    for (L, S, B, seed) in [(40, 40, 1, 51), (24, 36, 3, 52)]:
        torch.manual_seed(0)
        mod = ref.gqa.MultiheadGQA(512, 8, 2, dropout=0.0).eval()
        sd = _load_weights(mod, seed)
        xq = syn.unit_uniform((L, B, 512), syn._gen(seed, "xq"))
        xk = syn.unit_uniform((S, B, 512), syn._gen(seed, "xk"))
        with torch.no_grad():
            y, _ = mod(xq, xk, xk)
        mod_cases.append(dict(spec=dict(L=L, S=S, B=B, E=512, hq=8, hk=2, seed=seed), weights_checksum=syn.checksum(sd),
                              out=y.clone()))
    out["module"] = mod_cases
    _save("gqa.pt", out)


VARIANT_CASES = [dict(name="post_ln_moe", shared=False, pre_norm=False, rms=False, T=24, S=40, B=1, seed=61),
                 dict(name="post_ln_sharedmoe_b2", shared=True, pre_norm=False, rms=False, T=24, S=40, B=2, seed=62),
                 dict(name="pre_rms_moe", shared=False, pre_norm=True, rms=True, T=17, S=33, B=1, seed=63)]


def build_variant(ct, gqa, moe, c, d=512, ff=1024, hq=8, hk=2, n_experts=6, k=2, n_layers=2):
    """BASELINE config 4 from the reference's own blocks: generic wrappers with
    att=MultiheadGQA(512, 8, kv_heads=2), ff=MoELayer|SharedMoELayer(GLUExpert(512,1024), 6 experts, top-2).
    `ct`, `gqa`, `moe` are the module namespaces (the reference's here, ours in the tests)."""
    import torch.nn as nn
    att = gqa.MultiheadGQA(d, hq, hk, dropout=0.0)
    cls = moe.SharedMoELayer if c["shared"] else moe.MoELayer
    ffl = cls(moe.GLUExpert(d, ff, 0.0), d, n_experts=n_experts, n_experts_per_token=k, dropout=0.0)
    norm = ct.RMSNorm(d) if c["rms"] else nn.LayerNorm(d)
    enc = ct.TransformerEncoder(ct.TransformerEncoderLayer(att, ffl, pre_norm=c["pre_norm"], norm=norm, dropout=0.0), n_layers, norm)
    dec = ct.TransformerDecoder(ct.TransformerDecoderLayer(att, att, ffl, pre_norm=c["pre_norm"], norm=norm, dropout=0.0), n_layers,
                                norm)
    return nn.ModuleDict(dict(enc=enc, dec=dec)).eval()


def golden_variant(ref):
    """GQA + MoE encoder/decoder stacks assembled from custom_transformer.py:1220-1401 wrappers."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for c in VARIANT_CASES:
        torch.manual_seed(0)
        net = build_variant(ref.custom_transformer, ref.gqa, ref.moe, c)
        sd = _load_weights(net, c["seed"])
        src = syn.unit_uniform((c["S"], c["B"], 512), syn._gen(c["seed"], "src"))
        tgt = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "tgt"))
        gaps = []

        def hook(mod, inp, outp):
            top3 = torch.topk(outp, 3).values
            gaps.append(float((top3[..., :-1] - top3[..., 1:]).min()))
        hs = [m.gate.register_forward_hook(hook) for m in net.modules() if isinstance(m, (ref.moe.MoELayer, ref.moe.SharedMoELayer))]
        mask = torch.triu(torch.full((c["T"], c["T"]), float("-inf")), diagonal=1)
        with torch.no_grad():
            mem = net["enc"](src)
            y = net["dec"](tgt, mem, tgt_mask=mask)          # MultiheadGQA ignores the mask (literal behaviour)
        for h in hs:
            h.remove()
        out[c["name"]] = dict(spec=dict(c), weights_checksum=syn.checksum(sd), memory=mem.clone(), out=y.clone(),
                              min_rank_gap=min(gaps))
        print(c["name"], "min gate gap %.2e" % min(gaps))
    _save("variant.pt", out)


MOE_TRAIN_KEEP = ("gate.weight", "gate.bias", "experts.0.gate.bias", "experts.2.linear2.bias", "experts.5.linear1.bias")


def golden_moe_train(ref):
    """Gradients of MoELayer / SharedMoELayer in train() mode (dropout 0) that torch autograd derives from moe.py:167-302:
    loss = sum(y * r); d loss / d x, d / d gate (through the softmax over the top-k logits), d / d experts."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for shared in (False, True):
        torch.manual_seed(0)
        exp = ref.moe.GLUExpert(512, 1024, 0.0)
        cls = ref.moe.SharedMoELayer if shared else ref.moe.MoELayer
        mod = cls(exp, 512, n_experts=6, n_experts_per_token=2, dropout=0.0).train()
        sd = _load_weights(mod, 71 + shared)
        x = syn.unit_uniform((23, 3, 512), syn._gen(71, "x")).requires_grad_(True)
        r = syn.unit_uniform((23, 3, 512), syn._gen(71, "r"))
        y = mod(x)
        loss = (y * r).sum()
        loss.backward()
        with torch.no_grad():
            logits = mod.gate(x)
            idx = torch.topk(logits, 2).indices
            top3 = torch.topk(logits, 3).values
        norms = {n: float(p.grad.double().norm()) for n, p in mod.named_parameters() if p.grad is not None}
        grads = {n: p.grad.clone() for n, p in mod.named_parameters() if n in MOE_TRAIN_KEEP}
        out["shared_%s" % shared] = dict(spec=dict(L=23, B=3, d=512, ff=1024, n_experts=6, k=2, seed=71 + shared, x_seed=71),
                                         weights_checksum=syn.checksum(sd), out=y.detach().clone(), loss=float(loss.detach()),
                                         selected_experts=idx.clone(), dx=x.grad.clone(), grad_norms=norms, grads=grads,
                                         min_rank_gap=float((top3[..., :-1] - top3[..., 1:]).min()))
    _save("moe_train.pt", out)


def golden_variant_train(ref):
    """One backward pass through the GQA + MoE stacks of golden_variant (train() mode, dropout 0): loss = sum(y * r)."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for c in VARIANT_CASES:
        torch.manual_seed(0)
        net = build_variant(ref.custom_transformer, ref.gqa, ref.moe, c).train()
        sd = _load_weights(net, c["seed"])
        src = syn.unit_uniform((c["S"], c["B"], 512), syn._gen(c["seed"], "src")).requires_grad_(True)
        tgt = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "tgt")).requires_grad_(True)
        r = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "r"))
        mask = torch.triu(torch.full((c["T"], c["T"]), float("-inf")), diagonal=1)
        y = net["dec"](tgt, net["enc"](src), tgt_mask=mask)
        loss = (y * r).sum()
        loss.backward()
        norms = {n: float(p.grad.double().norm()) for n, p in net.named_parameters() if p.grad is not None}
        keep = ("enc.layers.0.self_attn.k_proj.weight", "dec.layers.1.ff.gate.weight", "dec.norm.weight",
                "dec.layers.0.cross_attn.norm.bias", "enc.layers.1.ff.experts.3.linear2.bias")
        grads = {n: p.grad.clone() for n, p in net.named_parameters() if n in keep}
        out[c["name"]] = dict(spec=dict(c), weights_checksum=syn.checksum(sd), out=y.detach().clone(), loss=float(loss.detach()),
                              d_src=src.grad.clone(), d_tgt=tgt.grad.clone(), grad_norms=norms, grads=grads)
        print(c["name"], "loss %.6f, %d parameter gradients" % (float(loss), len(norms)))
    _save("variant_train.pt", out)


def golden_metrics(ref):
    """compute_vevo_accuracy / compute_hits_k of the reference (dataset/vevo_dataset.py:653-701) on seeded logits."""
    import importlib
    with reference_cwd():
        vd = importlib.import_module("dataset.vevo_dataset")
    cases = []
    for seed, pad_from in ((71, 299), (72, 200), (73, 0)):
        g = syn._gen(seed, "metrics")
        out = syn.unit_uniform((1, 299, 159), g) * 3.0
        tgt = (torch.rand((1, 299), generator=g) * 157).long()
        out[0, torch.arange(0, 299, 3), tgt[0, ::3]] += 4.0            # make a third of the rows correct
        tgt[0, pad_from:] = 158
        acc = vd.compute_vevo_accuracy(out, tgt)
        cases.append(dict(seed=seed, pad_from=pad_from, acc=float(acc), hits=[float(vd.compute_hits_k(out, tgt, k)) if pad_from > 0 else None
                                                                            for k in (1, 3, 5)]))
    # compute_vevo_correspondence (:747-810) on seeded emotion rows: quality columns ~ Bernoulli(0.3), some rows all-zero, some
    # flagged as padding, probabilities around the 0.8 threshold; the closed form of its JSON lookups is checked on the way
    import json, os
    with reference_cwd():
        inv = json.load(open(os.path.join("dataset", "vevo_meta", "chord_inv.json")))
        attr = json.load(open(os.path.join("dataset", "vevo_meta", "chord_attr.json")))
    for k, name in inv.items():
        parts = name.split(":")
        assert (1 if len(parts) == 1 else attr[parts[1]]) == (1 if int(k) == 0 else (int(k) - 1) % 13 + 1), (k, name)
    cor = []
    for seed, thr in ((81, 0.8), (82, 0.8), (83, 0.5), (84, 2.0)):
        out, emo, prob = correspondence_case(seed)
        with reference_cwd():
            v = vd.compute_vevo_correspondence(out, torch.zeros((1, 299), dtype=torch.long), emo, prob, thr)
        cor.append(dict(seed=seed, thr=thr, value=float(v)))
        print("correspondence seed %d thr %.1f: %.6f" % (seed, thr, float(v)))
    _save("metrics.pt", dict(cases=cases, correspondence=cor))


def correspondence_case(seed):
    """Seeded inputs of the correspondence goldens (also used by the tests): logits (1, 299, 159), tgt_emotion (1, 299, 159),
    tgt_emotion_prob (1, 299)."""
    g = syn._gen(seed, "correspondence")
    out = syn.unit_uniform((1, 299, 159), g) * 3.0
    emo = torch.zeros((1, 299, 159))
    emo[0, :, :14] = (torch.rand((299, 14), generator=g) < 0.3).float()
    emo[0, torch.arange(0, 299, 7), :14] = 0.0                       # rows without any quality
    emo[0, 250:, -1] = 1.0                                           # padding tail
    out[0, torch.arange(0, 299, 5), 157] += 9.0                      # END predictions: counted, never right
    out[0, torch.arange(1, 299, 11), 0] += 9.0                       # "N": quality 1 by the reference's literal rule
    prob = 0.6 + 0.4 * torch.rand((1, 299), generator=g)
    return out, emo, prob


def golden_custom_mha(ref):
    """CustomMultiheadAttention with RoPE (custom_transformer.py:51-321; V2 models: RotaryPositionalEmbeddings(d_model,
    max_sequence_video), video_music_transformer.py:379) and without, self (causal mask) and cross attention."""
    import importlib
    rot = importlib.import_module("model.rotate_operation")
    out = []
    for (L, S, B, rope, causal, self_att, seed) in [(40, 40, 2, True, True, True, 81), (24, 36, 3, True, False, False, 82),
                                                    (300, 300, 1, True, True, True, 83), (17, 17, 2, False, False, True, 84)]:
        torch.manual_seed(0)
        m = ref.custom_transformer.CustomMultiheadAttention(512, 8, 0.0, RoPE=rot.RotaryPositionalEmbeddings(512, 300) if rope else None).eval()
        sd = _load_weights(m, seed)
        xq = syn.unit_uniform((L, B, 512), syn._gen(seed, "xq"))
        xk = xq if self_att else syn.unit_uniform((S, B, 512), syn._gen(seed, "xk"))
        mask = torch.triu(torch.full((L, L), float("-inf")), diagonal=1) if causal else None
        with torch.no_grad():
            y, w = m(xq, xk, xk, attn_mask=mask)
        out.append(dict(spec=dict(L=L, S=S, B=B, rope=rope, causal=causal, self_att=self_att, seed=seed),
                        weights_checksum=syn.checksum(sd), y=y.clone(), w_mean=w.mean(dim=0).clone()))
    _save("custom_mha.pt", dict(cases=out))


def golden_v2(ref):
    """VideoMusicTransformer_V2 '2.2' (the shipped inference default, argument_generate_funcs.py:82) and '2.0': logits of an
    eval forward and a short generate(beam=1) of the unmodified reference."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for ver, seed in (("2.2", 91), ("2.0", 92), ("1.1", 93), ("1.3rms", 94), ("3.0", 95), ("3.1", 96), ("3.2", 97)):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()), reference_cwd():
            if ver.startswith("2"):
                m = ref.vmt.VideoMusicTransformer_V2(version_name=ver, total_vf_dim=syn.vf_dim(0), dropout=0.1).eval()
            elif ver.startswith("3"):
                m = ref.vmt.VideoMusicTransformer_V3(version_name=ver, total_vf_dim=syn.vf_dim(0), dropout=0.1).eval()
            else:
                m = ref.vmt.VideoMusicTransformer_V1(version_name=ver[:3], total_vf_dim=syn.vf_dim(0), dropout=0.1,
                                                     rms_norm=ver.endswith("rms")).eval()
        sd = _load_weights(m, seed)
        inp = syn.make_inputs(2, seed, 24, 40, 0)
        args = [inp[k] for k in ("x", "x_root", "x_attr", "feature_semantic_list", "feature_key", "feature_scene_offset",
                                 "feature_motion", "feature_emotion")]
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            y = m(*args)
            one = [t[:1] for t in args]
            prim = inp["x"][0, :3]
            with reference_cwd():
                g = m.generate(one[3], one[4][0], one[5], one[6], one[7], primer=prim, primer_root=inp["x_root"][0, :3],
                               primer_attr=inp["x_attr"][0, :3], target_seq_length=14, beam=1, beam_chance=1.0)
        top2 = torch.topk(torch.softmax(y, -1)[..., :157], 2).values
        out[ver] = dict(spec=dict(version=ver, seed=seed, B=2, T=24, S=40, n_params=sum(p.numel() for p in m.parameters()),
                                  n_keys=len(sd)), weights_checksum=syn.checksum(sd), logits=y.clone(), generated=g.clone(),
                        keys=sorted(sd.keys())[:5])
        print(ver, "params", out[ver]["spec"]["n_params"], "keys", len(sd), "generated", g.tolist(), "min top-2 gap %.2e" % float((top2[..., 0] - top2[..., 1]).min()))
    _save("v2.pt", out)


def golden_zoo_train(ref):
    """One training step (train() mode, dropout 0, torch autograd) of the unmodified reference's model zoo: V2 '2.2' (RoPE + GLU / SharedMoE
    feed-forwards, the shipped default), V2 '2.0' (position tables, top-k scheduler stepping), V1 '1.1' (MoE everywhere) and V3 '3.1'
    (differential attention): logits, loss = sum(y * r), every parameter-gradient norm and a few whole gradients."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    keep_of = {"2": ("transformer.decoder.layers.0.self_attn.in_proj_bias", "transformer.encoder.layers.3.ff.gate.weight", "Linear_chord.bias",
                     "embedding_attr.weight", "transformer.decoder.layers.3.ff.shared_expert.linear2.bias", "transformer.decoder.layers.1.norm2.weight"),
               "1": ("transformer.decoder.layers.1.cross_attn.in_proj_bias", "transformer.encoder.layers.0.ff.gate.weight", "embedding_attr.weight",
                     "Wout.bias", "transformer.decoder.layers.0.ff.experts.2.linear1.bias"),
               "3": ("transformer.decoder.layers.1.self_attn.lambda_q1", "transformer.decoder.layers.0.self_attn.subln.weight",
                     "transformer.decoder.layers.2.self_attn.lambda_k2", "Linear_vis.bias", "transformer.decoder.norm.weight")}
    for ver, seed, nl in (("2.2", 191, 4), ("2.0", 192, 4), ("1.1", 193, 2), ("3.1", 196, 3)):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()), reference_cwd():
            cls = {"2": ref.vmt.VideoMusicTransformer_V2, "1": ref.vmt.VideoMusicTransformer_V1, "3": ref.vmt.VideoMusicTransformer_V3}[ver[0]]
            m = cls(version_name=ver, n_layers=nl, total_vf_dim=syn.vf_dim(0), dropout=0.0).train()
        sd = _load_weights(m, seed)
        B, T, S = 2, 20, 24
        inp = syn.make_inputs(B, seed, T, S, 0)
        args = [inp[k] for k in ("x", "x_root", "x_attr", "feature_semantic_list", "feature_key", "feature_scene_offset",
                                 "feature_motion", "feature_emotion")]
        r = syn.unit_uniform((B, T, 159), syn._gen(seed, "r"))
        with contextlib.redirect_stdout(io.StringIO()):
            y = m(*args)
            loss = (y * r).sum()
            loss.backward()
        norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        keep = [n for n in keep_of[ver[0]] if n in norms]
        grads = {n: p.grad.clone() for n, p in m.named_parameters() if n in keep}
        out[ver] = dict(spec=dict(version=ver, seed=seed, n_layers=nl, B=B, T=T, S=S), weights_checksum=syn.checksum(sd), logits=y.detach().clone(),
                        loss=float(loss.detach()), grad_norms=norms, grads=grads)
        print(ver, "loss %.6f, %d parameter gradients, kept %s" % (float(loss), len(norms), keep))
    _save("zoo_train.pt", out)


def golden_regression(ref):
    """VideoRegression (video_regression.py:103-245) with the generate-time defaults of BASELINE config 5
    (argument_generate_funcs.py:87-91: n_layers 6, d_model 128, d_hidden 256, total_vf_dim 774), Mamba-family backbones."""
    import importlib
    import third_party.log_maxvio as lm
    lm.is_logging = False
    vr = importlib.import_module("model.video_regression")
    out = {}
    for reg, seed in (("mamba", 101), ("mamba+", 102), ("bimamba+", 103), ("sharedmoe_bimamba+", 104)):
        torch.manual_seed(0)
        m = vr.VideoRegression(n_layers=6, d_model=128, d_hidden=256, dropout=0.1, total_vf_dim=774, regModel=reg).eval()
        sd = _load_weights(m, seed)
        B, L = 2, 300
        sem = syn.unit_uniform((B, L, 768), syn._gen(seed, "sem"))
        emo = torch.softmax(syn.unit_uniform((B, L, 6), syn._gen(seed, "emo")), dim=-1)
        z = torch.zeros((B, L))
        with torch.no_grad():
            ln, inst = m(sem, z, z, emo)
        out[reg] = dict(spec=dict(reg=reg, seed=seed, B=B, L=L, n_keys=len(sd)), weights_checksum=syn.checksum(sd), ln=ln.clone(), inst=inst.clone())
        print(reg, "keys", len(sd), float(ln.abs().max()), float(inst.mean()))
    _save("regression.pt", out)


def golden_pscan(ref):
    """pscan forward/backward (pscan.py:154-226) incl. a non power-of-two length."""
    cases = []
    for (B, L, D, N, seed) in [(2, 300, 8, 16, 7), (1, 1024, 4, 16, 8), (3, 5, 2, 16, 9), (1, 1, 2, 16, 10)]:
        A = torch.rand((B, L, D, N), generator=syn._gen(seed, "A")) * 0.99
        X = syn.unit_uniform((B, L, D, N), syn._gen(seed, "X"))
        gH = syn.unit_uniform((B, L, D, N), syn._gen(seed, "gH"))
        A.requires_grad_(True)
        X.requires_grad_(True)
        if L >= 2:
            H = ref.pscan.pscan(A, X)
            (H * gH).sum().backward()
            cases.append(dict(spec=dict(B=B, L=L, D=D, N=N, seed=seed), H=H.detach().clone(), gA=A.grad.clone(), gX=X.grad.clone()))
        else:
            cases.append(dict(spec=dict(B=B, L=L, D=D, N=N, seed=seed), H=X.detach().clone(), gA=torch.zeros_like(A), gX=gH.clone()))
    _save("pscan.pt", dict(cases=cases))


def golden_mamba(ref):
    """MambaBlock (both versions), a 2-layer Mamba stack and a BiMambaEncoderLayer (mamba.py:259-351, bimamba.py:64-99),
    eval mode, fp32 CPU, regression-model sizes (d_model 128, d_inner 256, d_state 16, argument_generate_funcs.py:87-91)."""
    out = {}
    for name, ver in (("block_v0", 0), ("block_v1", 1)):
        cfg = ref.mamba.MambaConfig(d_model=128, n_layers=1, use_version=ver)
        m = ref.mamba.MambaBlock(cfg).eval()
        sd = _load_weights(m, 21 + ver)
        spec = dict(B=3, L=300, d_model=128, seed=31 + ver, weight_seed=21 + ver, use_version=ver)
        x = syn.unit_uniform((spec["B"], spec["L"], 128), syn._gen(spec["seed"], "x"))
        with torch.no_grad():
            y = m(x)
        out[name] = dict(spec=spec, weights_checksum=syn.checksum(sd), y=y.clone())
    cfg = ref.mamba.MambaConfig(d_model=128, n_layers=2)
    m = ref.mamba.Mamba(cfg).eval()
    sd = _load_weights(m, 23)
    spec = dict(B=2, L=77, d_model=128, seed=33, weight_seed=23, n_layers=2)
    x = syn.unit_uniform((spec["B"], spec["L"], 128), syn._gen(spec["seed"], "x"))
    with torch.no_grad():
        y = m(x)
    out["stack"] = dict(spec=spec, weights_checksum=syn.checksum(sd), y=y.clone())
    cfg = ref.mamba.MambaConfig(d_model=128, n_layers=1)
    m = ref.bimamba.BiMambaEncoderLayer(cfg, dim_feedforward=256, dropout=0.2).eval()
    sd = _load_weights(m, 24)
    spec = dict(B=2, L=300, d_model=128, d_ff=256, seed=34, weight_seed=24)
    x = syn.unit_uniform((spec["B"], spec["L"], 128), syn._gen(spec["seed"], "x"))
    with torch.no_grad():
        y = m(x)
    out["bimamba_layer"] = dict(spec=spec, weights_checksum=syn.checksum(sd), y=y.clone())
    # Bi-Mamba+ layers (bimamba.py:101-191): plain FFN post-norm, MoE feed-forward norm_first
    import third_party.log_maxvio as lm
    lm.is_logging = False
    for name, norm_first, use_moe in (("bimamba_v1_ffn", False, False), ("bimamba_v1_moe", True, True)):
        cfg = ref.mamba.MambaConfig(d_model=128, n_layers=1, use_version=1)
        moe = ref.moe.MoELayer(ref.moe.GLUExpert(128, 256, 0.0), 128, n_experts=6, n_experts_per_token=2, dropout=0.0) if use_moe else None
        m = ref.bimamba.BiMambaEncoderLayer_V1(cfg, dim_feedforward=256, dropout=0.2, moe_layer=moe, norm_first=norm_first).eval()
        sd = _load_weights(m, 25 + use_moe)
        spec = dict(B=2, L=150, d_model=128, d_ff=256, seed=35 + use_moe, weight_seed=25 + use_moe, norm_first=norm_first, moe=use_moe)
        x = syn.unit_uniform((spec["B"], spec["L"], 128), syn._gen(spec["seed"], "x"))
        with torch.no_grad():
            y = m(x)
        out[name] = dict(spec=spec, weights_checksum=syn.checksum(sd), y=y.clone())
    _save("mamba.pt", out)


def golden_mamba_step(ref):
    """Recurrent single-token inference (MambaBlock.step / ResidualBlock.step / Mamba.step, mamba.py:100-108,151-159,407-470):
    T tokens fed one by one from the empty cache (None, zeros); outputs of every step and the final cache.  Also stored: the
    same tokens through forward() -- for use_version 0 the two agree (the reference's own consistency), for use_version 1
    they differ because step() ignores the mamba+ gate (mamba.py:430 vs :283-287)."""
    out = {}
    for name, ver, layers in (("block_v0", 0, 0), ("block_v1", 1, 0), ("stack", 0, 2)):
        cfg = ref.mamba.MambaConfig(d_model=128, n_layers=max(layers, 1), use_version=ver)
        m = (ref.mamba.Mamba(cfg) if layers else ref.mamba.MambaBlock(cfg)).eval()
        sd = _load_weights(m, 41 + ver + layers)
        spec = dict(B=3, T=20, d_model=128, seed=51 + ver + layers, weight_seed=41 + ver + layers, use_version=ver, n_layers=layers)
        x = syn.unit_uniform((spec["B"], spec["T"], 128), syn._gen(spec["seed"], "x"))
        empty = lambda: (None, torch.zeros(spec["B"], cfg.d_inner, cfg.d_conv - 1))
        cache = [empty() for _ in range(layers)] if layers else empty()
        ys = []
        with torch.no_grad():
            for t in range(spec["T"]):
                y, cache = m.step(x[:, t], cache)
                ys.append(y.clone())
            full = m(x)
        last = cache[-1] if layers else cache
        out[name] = dict(spec=spec, weights_checksum=syn.checksum(sd), y=torch.stack(ys, 1), y_forward=full.clone(),
                         h=last[0].clone(), inputs=last[1].clone())
        print(name, "step vs forward max diff %.2e" % float((out[name]["y"] - full).abs().max()))
    _save("mamba_step.pt", out)


MAMBA_TRAIN_CASES = [dict(name="block_v0", kind="block", ver=0, B=2, L=130, wseed=91, seed=191),
                     dict(name="block_v1", kind="block", ver=1, B=2, L=77, wseed=92, seed=192),
                     dict(name="stack", kind="stack", ver=0, B=2, L=70, wseed=93, seed=193),
                     dict(name="bimamba_layer", kind="bi", ver=0, B=2, L=90, wseed=94, seed=194),
                     dict(name="bimamba_v1_ffn", kind="bi_v1", ver=1, B=2, L=66, wseed=95, seed=195, norm_first=False, moe=False),
                     dict(name="bimamba_v1_moe", kind="bi_v1", ver=1, B=1, L=100, wseed=96, seed=196, norm_first=True, moe=True)]


def build_mamba_case(mamba, bimamba, moe, c):
    """The module of one MAMBA_TRAIN_CASES entry from the given namespaces (the reference's here, ours in the tests);
    regression-model sizes (d_model 128, d_inner 256, d_state 16), dropout 0."""
    cfg = mamba.MambaConfig(d_model=128, n_layers=2 if c["kind"] == "stack" else 1, use_version=c["ver"])
    if c["kind"] == "block":
        return mamba.MambaBlock(cfg)
    if c["kind"] == "stack":
        return mamba.Mamba(cfg)
    if c["kind"] == "bi":
        return bimamba.BiMambaEncoderLayer(cfg, dim_feedforward=256, dropout=0.0)
    ml = moe.MoELayer(moe.GLUExpert(128, 256, 0.0), 128, n_experts=6, n_experts_per_token=2, dropout=0.0) if c["moe"] else None
    return bimamba.BiMambaEncoderLayer_V1(cfg, dim_feedforward=256, dropout=0.0, moe_layer=ml, norm_first=c["norm_first"])


def golden_mamba_train(ref):
    """Gradients of the Mamba-family blocks in train() mode (pscan path of the reference, mamba.py:333-351 + pscan.py autograd):
    loss = sum(y * r); every parameter gradient norm, a few full gradients and d loss / d x."""
    import third_party.log_maxvio as lm
    lm.is_logging = False
    out = {}
    for c in MAMBA_TRAIN_CASES:
        torch.manual_seed(0)
        m = build_mamba_case(ref.mamba, ref.bimamba, ref.moe, c).train()
        sd = _load_weights(m, c["wseed"])
        x = syn.unit_uniform((c["B"], c["L"], 128), syn._gen(c["seed"], "x")).requires_grad_(True)
        r = syn.unit_uniform((c["B"], c["L"], 128), syn._gen(c["seed"], "r"))
        y = m(x)
        loss = (y * r).sum()
        loss.backward()
        norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        keep = [n for n in norms if n.endswith(("A_log", ".D", "conv1d.weight", "dt_proj.bias", "x_proj.weight")) or n in ("A_log", "D")]
        grads = {n: m.get_parameter(n).grad.clone() for n in keep[:6]}
        out[c["name"]] = dict(spec=dict(c), weights_checksum=syn.checksum(sd), y=y.detach().clone(), loss=float(loss.detach()),
                              dx=x.grad.clone(), grad_norms=norms, grads=grads)
        print(c["name"], "loss %.6f, %d parameter gradients, kept %s" % (float(loss.detach()), len(norms), list(grads)))
    _save("mamba_train.pt", out)


REGRESSION_TRAIN_CASES = [("mamba+", 111), ("bimamba+", 112), ("sharedmoe_bimamba+", 113)]


def regression_train_targets(seed, B, L):
    """Synthetic (loudness, note density) and instrument targets of the regression training loss."""
    return syn.unit_uniform((B, L, 2), syn._gen(seed, "t_ln")), (syn.unit_uniform((B, L, 40), syn._gen(seed, "t_inst")) > 0.7).float()


def golden_regression_train(ref):
    """One backward pass of VideoRegression training (utilities/run_model_regression.py:39: MSE(ln_nd) + BCE(instruments)),
    2 layers, dropout 0, train() mode."""
    import importlib
    import third_party.log_maxvio as lm
    lm.is_logging = False
    vr = importlib.import_module("model.video_regression")
    out = {}
    for reg, seed in REGRESSION_TRAIN_CASES:
        torch.manual_seed(0)
        m = vr.VideoRegression(n_layers=2, d_model=128, d_hidden=256, dropout=0.0, total_vf_dim=774, regModel=reg).train()
        for mod in m.modules():                      # the experts are built with GLUExpert's default dropout 0.1 (video_regression.py:176)
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        sd = _load_weights(m, seed)
        B, L = 2, 120
        sem = syn.unit_uniform((B, L, 768), syn._gen(seed, "sem"))
        emo = torch.softmax(syn.unit_uniform((B, L, 6), syn._gen(seed, "emo")), dim=-1)
        t_ln, t_inst = regression_train_targets(seed, B, L)
        z = torch.zeros((B, L))
        ln, inst = m(sem, z, z, emo)
        loss = torch.nn.functional.mse_loss(ln, t_ln) + torch.nn.functional.binary_cross_entropy(inst, t_inst)
        loss.backward()
        norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        grads = {n: m.get_parameter(n).grad.clone() for n in ("regressor.weight", "classifier.0.bias", "in_proj.0.bias")}
        out[reg] = dict(spec=dict(reg=reg, seed=seed, B=B, L=L, n_layers=2), weights_checksum=syn.checksum(sd), ln=ln.detach().clone(),
                        inst=inst.detach().clone(), loss=float(loss.detach()), grad_norms=norms, grads=grads)
        print(reg, "loss %.6f, %d parameter gradients" % (float(loss.detach()), len(norms)))
    _save("regression_train.pt", out)


def golden_rpr_train(ref):
    """Gradients of TransformerDecoderLayerRPR and a 2-layer TransformerDecoderRPR (rpr.py:18-70) in train() mode, dropout 0:
    loss = sum(y * r)."""
    out = {}
    for name, n_layers in (("layer", 1), ("decoder", 2)):
        torch.manual_seed(0)
        layer = ref.rpr.TransformerDecoderLayerRPR(256, 4, 512, 0.0, er_len=80)
        m = (layer if n_layers == 1 else ref.rpr.TransformerDecoderRPR(layer, n_layers, torch.nn.LayerNorm(256))).train()
        seed = 121 + n_layers
        sd = _load_weights(m, seed)
        tgt = syn.unit_uniform((50, 3, 256), syn._gen(seed, "tgt")).requires_grad_(True)
        mem = syn.unit_uniform((70, 3, 256), syn._gen(seed, "mem")).requires_grad_(True)
        r = syn.unit_uniform((50, 3, 256), syn._gen(seed, "r"))
        mask = torch.triu(torch.full((50, 50), float("-inf")), diagonal=1)
        y = m(tgt, mem, tgt_mask=mask)
        (y * r).sum().backward()
        norms = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        keep = [n for n in norms if n.endswith(("self_attn.Er", "norm2.weight", "multihead_attn.in_proj_bias"))][:4]
        out[name] = dict(spec=dict(T=50, S=70, B=3, E=256, H=4, ff=512, er_len=80, seed=seed, n_layers=n_layers),
                         weights_checksum=syn.checksum(sd), out=y.detach().clone(), d_tgt=tgt.grad.clone(), d_mem=mem.grad.clone(),
                         grad_norms=norms, grads={n: m.get_parameter(n).grad.clone() for n in keep})
        print(name, len(norms), "parameter gradients, kept", keep)
    _save("rpr_train.pt", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--gen-videos", type=int, default=4)
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    jobs = dict(forward=lambda: golden_forward(ref), train=lambda: golden_train(ref), train_full=lambda: golden_train_full(ref), rpr=lambda: golden_rpr(ref),
                moe=lambda: golden_moe(ref), gqa=lambda: golden_gqa(ref), pscan=lambda: golden_pscan(ref),
                mamba=lambda: golden_mamba(ref), mamba_step=lambda: golden_mamba_step(ref), variant=lambda: golden_variant(ref), moe_train=lambda: golden_moe_train(ref),
                variant_train=lambda: golden_variant_train(ref), mamba_train=lambda: golden_mamba_train(ref), rpr_train=lambda: golden_rpr_train(ref), regression_train=lambda: golden_regression_train(ref), metrics=lambda: golden_metrics(ref), zoo_train=lambda: golden_zoo_train(ref), custom_mha=lambda: golden_custom_mha(ref), v2=lambda: golden_v2(ref), regression=lambda: golden_regression(ref),
                generate=lambda: golden_generate(ref, args.gen_videos),
                primed=lambda: golden_generate_primed(ref))
    for name, fn in jobs.items():
        if args.only and args.only != name:
            continue
        t0 = time.time()
        fn()
        print("%s done in %.1fs" % (name, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
