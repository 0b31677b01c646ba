"""CPU: the oracle (oracle/amt_oracle.py) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/*.pt, generator: oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from conftest import amt_state_dict, load_golden, rel_err, same_checksum
from oracle import amt_oracle as O
from video2music_b200 import synthetic as syn


def _inputs(spec):
    return syn.make_inputs(spec["batch"], spec["input_seed"], spec["tgt_len"], spec["src_len"], spec["motion_type"])


def _fwd(sd, inp, **kw):
    with torch.no_grad():
        return O.amt_forward(sd, inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
                             inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"], **kw)


def test_weights_regenerate_bit_identically():
    g = load_golden("amt_forward_cfg1.pt")
    _, sd = amt_state_dict(syn.vf_dim(0), g["spec"]["weight_seed"])
    assert same_checksum(syn.checksum({k: v for k, v in sd.items() if not k.endswith(".pe")}), g["weights_checksum"])


@pytest.mark.parametrize("name", ["amt_forward_small.pt", "amt_forward_cfg1.pt"])
def test_forward_matches_reference_golden(name):
    g = load_golden(name)
    spec = g["spec"]
    _, sd = amt_state_dict(syn.vf_dim(spec["motion_type"]), spec["weight_seed"])
    y = _fwd(sd, _inputs(spec))
    assert y.shape == g["logits"].shape
    assert rel_err(y, g["logits"]) < 2e-5


def test_skew_closed_form_equals_literal():
    torch.manual_seed(0)
    q = torch.randn(3, 17, 8)
    Er = torch.randn(17, 8)
    lit = O.skew_literal(torch.einsum("hld,md->hlm", q, Er))
    assert torch.allclose(lit, O.skew_closed_form(q, Er), atol=1e-5)


@pytest.mark.parametrize("chord_embed", [False, True])
def test_cached_greedy_equals_reference_generate(chord_embed):
    g = load_golden("amt_generate_greedy.pt")["chord_embed_%s" % chord_embed]
    spec = g["spec"]
    _, sd = amt_state_dict(syn.vf_dim(0), spec["weight_seed"], chord_embed=chord_embed, wout_gain=spec["wout_gain"])
    inp = syn.make_inputs(spec["n_videos"], spec["input_seed"], 299, 300, 0)
    prim, pr, pa = (torch.tensor(spec[k]) for k in ("primer", "primer_root", "primer_attr"))
    with torch.no_grad():
        gen = O.generate_greedy_cached(sd, inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"],
                                       inp["feature_motion"], inp["feature_emotion"], prim, pr, pa, 300,
                                       chord_embed=chord_embed)
    assert torch.equal(gen, g["tokens"])


@pytest.mark.parametrize("chord_embed", [False, True])
def test_cached_greedy_primed_equals_reference_generate(chord_embed):
    g = load_golden("amt_generate_primed.pt")["chord_embed_%s" % chord_embed]
    spec = g["spec"]
    _, sd = amt_state_dict(syn.vf_dim(0), spec["weight_seed"], chord_embed=chord_embed, wout_gain=spec["wout_gain"])
    inp = syn.make_inputs(spec["n_videos"], spec["input_seed"], 299, 300, 0)
    P = spec["primer_len"]
    with torch.no_grad():
        gen = O.generate_greedy_cached(sd, inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"],
                                       inp["feature_motion"], inp["feature_emotion"], inp["x"][:, :P], inp["x_root"][:, :P],
                                       inp["x_attr"][:, :P], 300, chord_embed=chord_embed)
    assert torch.equal(gen, g["tokens"])


def test_rpr_module_golden():
    g = load_golden("rpr_attention.pt")
    for case in g["cases"]:
        s = case["spec"]
        shapes = {"in_proj_weight": (3 * s["E"], s["E"]), "in_proj_bias": (3 * s["E"],), "out_proj.weight": (s["E"], s["E"]),
                  "out_proj.bias": (s["E"],), "Er": (s["er_len"], s["E"] // s["H"])}
        sd = syn.fill_like_reference_init(shapes, seed=s["seed"])
        assert same_checksum(syn.checksum(sd), case["weights_checksum"])
        x = syn.unit_uniform((s["L"], s["B"], s["E"]), syn._gen(s["seed"], "x"))
        mask = torch.triu(torch.full((s["L"], s["L"]), float("-inf")), diagonal=1)
        out, w = O.mha_forward(x, x, sd["in_proj_weight"], sd["in_proj_bias"], sd["out_proj.weight"], sd["out_proj.bias"],
                               s["H"], Er=sd["Er"], attn_mask=mask, need_weights=True)
        assert rel_err(out, case["out"]) < 2e-5
        if case["weights_mean"] is not None:
            assert rel_err(w, case["weights_mean"]) < 2e-5


def _moe_sd(spec, shared):
    shapes = {"gate.weight": (spec["n_experts"], spec["d"]), "gate.bias": (spec["n_experts"],)}
    names = ["experts.%d." % i for i in range(spec["n_experts"])] + (["shared_expert."] if shared else [])
    for p in names:
        shapes.update({p + "linear1.weight": (spec["ff"], spec["d"]), p + "linear1.bias": (spec["ff"],),
                       p + "gate.weight": (spec["ff"], spec["d"]), p + "gate.bias": (spec["ff"],),
                       p + "linear2.weight": (spec["d"], spec["ff"]), p + "linear2.bias": (spec["d"],)})
    return syn.fill_like_reference_init(shapes, seed=spec["seed"])


@pytest.mark.parametrize("shared", [False, True])
def test_moe_golden(shared):
    g = load_golden("moe.pt")["shared_%s" % shared]
    spec = g["spec"]
    sd = _moe_sd(spec, shared)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    x = syn.unit_uniform((spec["L"], spec["B"], spec["d"]), syn._gen(spec["x_seed"], "x"))
    out, idx, _ = O.moe_layer(x, sd, "", spec["n_experts"], spec["k"], shared=shared)
    assert torch.equal(idx, g["selected_experts"])
    assert rel_err(out, g["out"]) < 2e-5


def test_gqa_golden():
    g = load_golden("gqa.pt")
    for case in g["function"]:
        s = case["spec"]
        q = syn.unit_uniform((s["b"], s["n"], s["hq"], s["d"]), syn._gen(s["seed"], "q"))
        k = syn.unit_uniform((s["b"], s["s"], s["hk"], s["d"]), syn._gen(s["seed"], "k"))
        v = syn.unit_uniform((s["b"], s["s"], s["hk"], s["d"]), syn._gen(s["seed"], "v"))
        assert rel_err(O.sdp_gqa(q, k, v, is_causal=s["causal"]), case["out"]) < 2e-5
    for case in g["module"]:
        s = case["spec"]
        kv = s["E"] // s["hq"] * s["hk"]
        shapes = {"q_proj.weight": (s["E"], s["E"]), "q_proj.bias": (s["E"],), "k_proj.weight": (kv, s["E"]),
                  "k_proj.bias": (kv,), "v_proj.weight": (kv, s["E"]), "v_proj.bias": (kv,), "norm.weight": (s["E"],),
                  "norm.bias": (s["E"],), "out_proj.weight": (s["E"], s["E"]), "out_proj.bias": (s["E"],)}
        sd = syn.fill_like_reference_init(shapes, seed=s["seed"])
        assert same_checksum(syn.checksum(sd), case["weights_checksum"])
        xq = syn.unit_uniform((s["L"], s["B"], s["E"]), syn._gen(s["seed"], "xq"))
        xk = syn.unit_uniform((s["S"], s["B"], s["E"]), syn._gen(s["seed"], "xk"))
        y = O.mhgqa_forward(xq, xk, xk, sd, "", s["hq"], s["hk"])
        assert rel_err(y, case["out"]) < 2e-5


def _variant_net(c):
    import video2music_b200.custom_transformer as ct
    import video2music_b200.grouped_query_attention as gqa
    import video2music_b200.moe as moe
    from oracle.make_golden import build_variant
    net = build_variant(ct, gqa, moe, c)
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=c["seed"])
    return net, sd


@pytest.mark.parametrize("name", ["post_ln_moe", "post_ln_sharedmoe_b2", "pre_rms_moe"])
def test_variant_gqa_moe_stack_golden(name):
    """BASELINE config 4: the reference's generic wrappers around MultiheadGQA + (Shared)MoELayer, 2+2 layers."""
    g = load_golden("variant.pt")[name]
    c = g["spec"]
    _, sd = _variant_net(c)                                   # our modules have the reference's parameter names
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    src = syn.unit_uniform((c["S"], c["B"], 512), syn._gen(c["seed"], "src"))
    tgt = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "tgt"))
    mem, y = O.variant_stack_forward(sd, src, tgt, 2, 8, 2, 6, 2, c["shared"], c["pre_norm"], c["rms"])
    assert g["min_rank_gap"] > 1e-4
    assert rel_err(mem, g["memory"]) < 2e-5 and rel_err(y, g["out"]) < 2e-5


def _leaf_sd(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


def _check_grads(named_grads, g, tol):
    """All gradient norms within tol (relative) and the stored full gradients within tol of the reference's."""
    floor = 1e-3 * max(g["grad_norms"].values())     # k_proj.bias gradients are rounding noise (softmax is shift-invariant)
    for n, ref_norm in g["grad_norms"].items():
        got = float(named_grads[n].double().norm())
        assert abs(got - ref_norm) <= tol * max(ref_norm, floor), (n, got, ref_norm)
    for n, ref in g["grads"].items():
        assert rel_err(named_grads[n], ref) < tol, n


@pytest.mark.parametrize("shared", [False, True])
def test_moe_train_golden(shared):
    """Oracle MoE gradients (torch autograd over the restatement) == the reference's (moe.py:167-302 in train mode)."""
    g = load_golden("moe_train.pt")["shared_%s" % shared]
    spec = g["spec"]
    sd = _moe_sd(spec, shared)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    sd = _leaf_sd(sd)
    x = syn.unit_uniform((spec["L"], spec["B"], spec["d"]), syn._gen(spec["x_seed"], "x")).requires_grad_(True)
    r = syn.unit_uniform((spec["L"], spec["B"], spec["d"]), syn._gen(spec["x_seed"], "r"))
    out, idx, _ = O.moe_layer(x, sd, "", spec["n_experts"], spec["k"], shared=shared)
    (out * r).sum().backward()
    assert g["min_rank_gap"] > 1e-4 and torch.equal(idx, g["selected_experts"])
    assert rel_err(out, g["out"]) < 2e-5 and rel_err(x.grad, g["dx"]) < 2e-5
    _check_grads({k: v.grad for k, v in sd.items()}, g, 2e-5)


@pytest.mark.parametrize("name", ["post_ln_moe", "post_ln_sharedmoe_b2", "pre_rms_moe"])
def test_variant_train_golden(name):
    """BASELINE config 4 training: gradients of the oracle's GQA + MoE stack == the reference's autograd."""
    g = load_golden("variant_train.pt")[name]
    c = g["spec"]
    _, sd = _variant_net(c)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    sd = _leaf_sd(sd)
    src = syn.unit_uniform((c["S"], c["B"], 512), syn._gen(c["seed"], "src")).requires_grad_(True)
    tgt = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "tgt")).requires_grad_(True)
    r = syn.unit_uniform((c["T"], c["B"], 512), syn._gen(c["seed"], "r"))
    _, y = O.variant_stack_forward(sd, src, tgt, 2, 8, 2, 6, 2, c["shared"], c["pre_norm"], c["rms"])
    (y * r).sum().backward()
    assert rel_err(y, g["out"]) < 2e-5
    assert rel_err(src.grad, g["d_src"]) < 5e-5 and rel_err(tgt.grad, g["d_tgt"]) < 5e-5
    _check_grads({k: v.grad for k, v in sd.items() if v.grad is not None}, g, 5e-5)


def _mamba_train_module(c):
    import video2music_b200.mamba as mamba
    import video2music_b200.moe as moe
    from oracle.make_golden import build_mamba_case
    m = build_mamba_case(mamba, mamba, moe, c)
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=c["wseed"])
    return m, sd


def _mamba_oracle_forward(c, sd, x):
    if c["kind"] == "block":
        return O.mamba_block_forward(sd, "", x, dt_rank=8, use_version=c["ver"])
    if c["kind"] == "stack":
        return O.mamba_forward(sd, x, 2, dt_rank=8)
    if c["kind"] == "bi":
        return O.bimamba_layer_forward(sd, "", x, dt_rank=8)
    return O.bimamba_v1_layer_forward(sd, "", x, 8, c["norm_first"], dict(n_experts=6, k=2, shared=False) if c["moe"] else None)


@pytest.mark.parametrize("name", ["block_v0", "block_v1", "stack", "bimamba_layer", "bimamba_v1_ffn", "bimamba_v1_moe"])
def test_mamba_train_golden(name):
    """Oracle Mamba-family gradients (autograd over the restatement) == the reference's (pscan path, train mode)."""
    g = load_golden("mamba_train.pt")[name]
    c = g["spec"]
    _, sd = _mamba_train_module(c)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    sd = _leaf_sd(sd)
    x = syn.unit_uniform((c["B"], c["L"], 128), syn._gen(c["seed"], "x")).requires_grad_(True)
    r = syn.unit_uniform((c["B"], c["L"], 128), syn._gen(c["seed"], "r"))
    y = _mamba_oracle_forward(c, sd, x)
    (y * r).sum().backward()
    assert rel_err(y, g["y"]) < 2e-5 and rel_err(x.grad, g["dx"]) < 5e-5
    _check_grads({k: v.grad for k, v in sd.items() if v.grad is not None}, g, 5e-5)


def _regression_train_case(reg):
    from video2music_b200 import VideoRegression
    from oracle.make_golden import regression_train_targets
    g = load_golden("regression_train.pt")[reg]
    s = g["spec"]
    m = VideoRegression(n_layers=s["n_layers"], d_model=128, d_hidden=256, dropout=0.0, total_vf_dim=774, regModel=reg).train()
    for mod in m.modules():                          # GLUExpert's default dropout 0.1 (video_regression.py:176), zeroed as in the golden
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["seed"])
    sem = syn.unit_uniform((s["B"], s["L"], 768), syn._gen(s["seed"], "sem"))
    emo = torch.softmax(syn.unit_uniform((s["B"], s["L"], 6), syn._gen(s["seed"], "emo")), dim=-1)
    return g, m, sd, sem, emo, regression_train_targets(s["seed"], s["B"], s["L"])


def regression_loss(ln, inst, t_ln, t_inst):
    """utilities/run_model_regression.py:39."""
    return torch.nn.functional.mse_loss(ln, t_ln) + torch.nn.functional.binary_cross_entropy(inst, t_inst)


@pytest.mark.parametrize("reg", ["mamba+", "bimamba+", "sharedmoe_bimamba+"])
def test_video_regression_train_golden(reg):
    """Oracle gradients of one VideoRegression training step == the reference's."""
    g, m, sd, sem, emo, (t_ln, t_inst) = _regression_train_case(reg)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    sd = _leaf_sd(sd)
    ln, inst = O.video_regression_forward(sd, sem, emo, reg, g["spec"]["n_layers"], dt_rank=8)
    loss = regression_loss(ln, inst, t_ln, t_inst)
    loss.backward()
    assert rel_err(ln, g["ln"]) < 5e-5 and rel_err(inst, g["inst"]) < 5e-5 and abs(float(loss.detach()) - g["loss"]) < 1e-5 * abs(g["loss"])
    _check_grads({k: v.grad for k, v in sd.items() if v.grad is not None}, g, 1e-4)


def _rpr_train_module(name):
    from video2music_b200 import TransformerDecoderLayerRPR, TransformerDecoderRPR
    g = load_golden("rpr_train.pt")[name]
    s = g["spec"]
    layer = TransformerDecoderLayerRPR(s["E"], s["H"], s["ff"], 0.0, er_len=s["er_len"])
    m = layer if s["n_layers"] == 1 else TransformerDecoderRPR(layer, s["n_layers"], torch.nn.LayerNorm(s["E"]))
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["seed"])
    tgt = syn.unit_uniform((s["T"], s["B"], s["E"]), syn._gen(s["seed"], "tgt"))
    mem = syn.unit_uniform((s["S"], s["B"], s["E"]), syn._gen(s["seed"], "mem"))
    r = syn.unit_uniform((s["T"], s["B"], s["E"]), syn._gen(s["seed"], "r"))
    return g, m.train(), sd, tgt, mem, r


@pytest.mark.parametrize("name", ["layer", "decoder"])
def test_rpr_decoder_train_golden(name):
    """Oracle gradients of TransformerDecoderLayerRPR / a 2-layer TransformerDecoderRPR == the reference's autograd."""
    g, _, sd, tgt, mem, r = _rpr_train_module(name)
    s = g["spec"]
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    sd = _leaf_sd(sd)
    tgt.requires_grad_(True), mem.requires_grad_(True)
    mask = torch.triu(torch.full((s["T"], s["T"]), float("-inf")), diagonal=1)
    if s["n_layers"] == 1:
        y = O.decoder_layer_rpr(tgt, mem, sd, "", s["H"], mask)
    else:
        y = tgt
        for i in range(s["n_layers"]):
            y = O.decoder_layer_rpr(y, mem, sd, "layers.%d." % i, s["H"], mask)
        y = torch.nn.functional.layer_norm(y, (s["E"],), sd["norm.weight"], sd["norm.bias"])
    (y * r).sum().backward()
    assert rel_err(y, g["out"]) < 2e-5 and rel_err(tgt.grad, g["d_tgt"]) < 5e-5 and rel_err(mem.grad, g["d_mem"]) < 5e-5
    _check_grads({k: v.grad for k, v in sd.items() if v.grad is not None}, g, 5e-5)


def test_pscan_golden():
    for case in load_golden("pscan.pt")["cases"]:
        s = case["spec"]
        A = torch.rand((s["B"], s["L"], s["D"], s["N"]), generator=syn._gen(s["seed"], "A")) * 0.99
        X = syn.unit_uniform((s["B"], s["L"], s["D"], s["N"]), syn._gen(s["seed"], "X"))
        gH = syn.unit_uniform((s["B"], s["L"], s["D"], s["N"]), syn._gen(s["seed"], "gH"))
        H = O.pscan_forward(A, X)
        assert rel_err(H, case["H"]) < 1e-5
        gA, gX = O.pscan_backward(A, H, gH)
        assert rel_err(gX, case["gX"]) < 1e-5
        assert rel_err(gA, case["gA"]) < 1e-5 or float(case["gA"].abs().max()) == 0.0


def _mamba_sd(shapes, seed):
    return syn.fill_like_reference_init(shapes, seed=seed)


def _bimamba_v1(s):
    from video2music_b200 import BiMambaEncoderLayer_V1, GLUExpert, MambaConfig, MoELayer
    moe = MoELayer(GLUExpert(128, 256, 0.0), 128, n_experts=6, n_experts_per_token=2, dropout=0.0) if s["moe"] else None
    return BiMambaEncoderLayer_V1(MambaConfig(d_model=128, n_layers=1, use_version=1), dim_feedforward=s["d_ff"], moe_layer=moe,
                                  norm_first=s["norm_first"])


def test_mamba_oracle_matches_reference_golden():
    """oracle Mamba restatement (mamba.py:259-351, bimamba.py:64-99) vs outputs of the unmodified reference."""
    from video2music_b200.mamba import MambaConfig, MambaBlock, Mamba, BiMambaEncoderLayer
    g = load_golden("mamba.pt")
    for name in ("block_v0", "block_v1"):
        c = g[name]
        s = c["spec"]
        m = MambaBlock(MambaConfig(d_model=128, n_layers=1, use_version=s["use_version"]))
        sd = _mamba_sd({k: tuple(v.shape) for k, v in m.state_dict().items()}, s["weight_seed"])
        assert same_checksum(syn.checksum(sd), c["weights_checksum"])
        x = syn.unit_uniform((s["B"], s["L"], 128), syn._gen(s["seed"], "x"))
        y = O.mamba_block_forward(sd, "", x, dt_rank=8, use_version=s["use_version"])
        assert rel_err(y, c["y"]) < 2e-5, name
    c = g["stack"]
    s = c["spec"]
    m = Mamba(MambaConfig(d_model=128, n_layers=2))
    sd = _mamba_sd({k: tuple(v.shape) for k, v in m.state_dict().items()}, s["weight_seed"])
    assert same_checksum(syn.checksum(sd), c["weights_checksum"])
    x = syn.unit_uniform((s["B"], s["L"], 128), syn._gen(s["seed"], "x"))
    assert rel_err(O.mamba_forward(sd, x, 2, dt_rank=8), c["y"]) < 2e-5
    c = g["bimamba_layer"]
    s = c["spec"]
    m = BiMambaEncoderLayer(MambaConfig(d_model=128, n_layers=1), dim_feedforward=s["d_ff"])
    sd = _mamba_sd({k: tuple(v.shape) for k, v in m.state_dict().items()}, s["weight_seed"])
    assert same_checksum(syn.checksum(sd), c["weights_checksum"])
    x = syn.unit_uniform((s["B"], s["L"], 128), syn._gen(s["seed"], "x"))
    assert rel_err(O.bimamba_layer_forward(sd, "", x, dt_rank=8), c["y"]) < 2e-5
    for name in ("bimamba_v1_ffn", "bimamba_v1_moe"):
        c = g[name]
        s = c["spec"]
        m = _bimamba_v1(s)
        sd = _mamba_sd({k: tuple(v.shape) for k, v in m.state_dict().items()}, s["weight_seed"])
        assert same_checksum(syn.checksum(sd), c["weights_checksum"])
        x = syn.unit_uniform((s["B"], s["L"], 128), syn._gen(s["seed"], "x"))
        y = O.bimamba_v1_layer_forward(sd, "", x, 8, s["norm_first"], dict(n_experts=6, k=2, shared=False) if s["moe"] else None)
        assert rel_err(y, c["y"]) < 2e-5, name


def _metrics_case(c):
    g = syn._gen(c["seed"], "metrics")
    out = syn.unit_uniform((1, 299, 159), g) * 3.0
    tgt = (torch.rand((1, 299), generator=g) * 157).long()
    out[0, torch.arange(0, 299, 3), tgt[0, ::3]] += 4.0
    tgt[0, c["pad_from"]:] = 158
    return out, tgt


def test_metrics_oracle_matches_reference_golden():
    """compute_vevo_accuracy / compute_hits_k (dataset/vevo_dataset.py:653-701) restated vs the reference's values."""
    for c in load_golden("metrics.pt")["cases"]:
        out, tgt = _metrics_case(c)
        assert abs(O.vevo_accuracy(out, tgt) - c["acc"]) < 1e-6
        for k, h in zip((1, 3, 5), c["hits"]):
            if h is not None:
                assert abs(O.hits_k(out, tgt, k) - h) < 1e-6


def _correspondence_case(seed):
    """The seeded inputs oracle/make_golden.py::correspondence_case fed to the reference (kept in step with it)."""
    g = syn._gen(seed, "correspondence")
    out = syn.unit_uniform((1, 299, 159), g) * 3.0
    emo = torch.zeros((1, 299, 159))
    emo[0, :, :14] = (torch.rand((299, 14), generator=g) < 0.3).float()
    emo[0, torch.arange(0, 299, 7), :14] = 0.0
    emo[0, 250:, -1] = 1.0
    out[0, torch.arange(0, 299, 5), 157] += 9.0
    out[0, torch.arange(1, 299, 11), 0] += 9.0
    prob = 0.6 + 0.4 * torch.rand((1, 299), generator=g)
    return out, emo, prob


def test_correspondence_oracle_matches_reference_golden():
    """compute_vevo_correspondence (dataset/vevo_dataset.py:747-810) restated vs the values the reference's own function returned
    (its JSON chord tables included) on the seeded cases of oracle/make_golden.py."""
    correspondence_case = _correspondence_case
    cases = load_golden("metrics.pt")["correspondence"]
    assert len(cases) == 4
    seen = set()
    for c in cases:
        out, emo, prob = correspondence_case(c["seed"])
        v, pt, right = O.vevo_correspondence(out, emo, prob, c["thr"])
        assert abs(v - c["value"]) < 1e-6, (c, v)
        seen.add(v == -1)
    assert seen == {True, False}                       # both the "no position qualifies" (-1) and the regular branch are pinned


def _custom_mha_case(c):
    from video2music_b200 import CustomMultiheadAttention, RotaryPositionalEmbeddings
    s = c["spec"]
    m = CustomMultiheadAttention(512, 8, 0.0, RoPE=RotaryPositionalEmbeddings(512, 300) if s["rope"] else None).eval()
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["seed"])
    xq = syn.unit_uniform((s["L"], s["B"], 512), syn._gen(s["seed"], "xq"))
    xk = xq if s["self_att"] else syn.unit_uniform((s["S"], s["B"], 512), syn._gen(s["seed"], "xk"))
    return m, sd, xq, xk


def test_custom_mha_rope_oracle_matches_reference_golden():
    """nn.MultiheadAttention + the reference's literal RoPE reinterpretation (custom_transformer.py:1044-1053) restated."""
    for c in load_golden("custom_mha.pt")["cases"]:
        s = c["spec"]
        m, sd, xq, xk = _custom_mha_case(c)
        assert same_checksum(syn.checksum(sd), c["weights_checksum"])
        if s["rope"]:
            assert torch.equal(m.RoPE.cache, O.rope_cache(512, 300))
        y, w = O.custom_mha_forward(xq, xk, xk, sd, "", 8, O.rope_cache(512, 300) if s["rope"] else None, s["causal"])
        assert rel_err(y, c["y"]) < 2e-5 and rel_err(w.mean(dim=0), c["w_mean"]) < 2e-5


def _v2_case(ver):
    from video2music_b200 import VideoMusicTransformer_V1, VideoMusicTransformer_V2, VideoMusicTransformer_V3
    g = load_golden("v2.pt")[ver]
    s = g["spec"]
    if ver.startswith("2"):
        m = VideoMusicTransformer_V2(version_name=ver, total_vf_dim=syn.vf_dim(0), dropout=0.1).eval()
    elif ver.startswith("3"):
        m = VideoMusicTransformer_V3(version_name=ver, total_vf_dim=syn.vf_dim(0), dropout=0.1).eval()
    else:
        m = VideoMusicTransformer_V1(version_name=ver[:3], total_vf_dim=syn.vf_dim(0), dropout=0.1, rms_norm=ver.endswith("rms")).eval()
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["seed"])
    inp = syn.make_inputs(s["B"], s["seed"], s["T"], s["S"], 0)
    return g, m, sd, inp


@pytest.mark.parametrize("ver", ["2.2", "2.0", "1.1", "1.3rms", "3.0", "3.1", "3.2"])
def test_v2_model_oracle_matches_reference_golden(ver):
    """VideoMusicTransformer_V2 (video_music_transformer.py:317-520): same parameter set as the reference (count, names) and
    the oracle's restatement of its forward reproduces the reference's logits."""
    g, m, sd, inp = _v2_case(ver)
    assert sum(p.numel() for p in m.parameters()) == g["spec"]["n_params"] and len(sd) == g["spec"]["n_keys"]
    assert sorted(sd.keys())[:5] == g["keys"] and same_checksum(syn.checksum(sd), g["weights_checksum"])
    a = (sd, inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"],
         inp["feature_motion"], inp["feature_emotion"])
    if ver.startswith("2"):
        y = O.v2_forward(*a, version=ver)
    elif ver.startswith("3"):
        y = O.v3_forward(*a, version=ver)
    else:
        y = O.v1_forward(*a, version=ver[:3], rms=ver.endswith("rms"))
    assert rel_err(y, g["logits"]) < 5e-5


def _regression_case(reg):
    from video2music_b200 import VideoRegression
    g = load_golden("regression.pt")[reg]
    s = g["spec"]
    m = VideoRegression(n_layers=6, d_model=128, d_hidden=256, dropout=0.1, total_vf_dim=774, regModel=reg).eval()
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["seed"])
    sem = syn.unit_uniform((s["B"], s["L"], 768), syn._gen(s["seed"], "sem"))
    emo = torch.softmax(syn.unit_uniform((s["B"], s["L"], 6), syn._gen(s["seed"], "emo")), dim=-1)
    return g, m, sd, sem, emo


@pytest.mark.parametrize("reg", ["mamba", "mamba+", "bimamba+"])
def test_video_regression_oracle_matches_reference_golden(reg):
    """VideoRegression (video_regression.py:208-245) with the Mamba-family backbones of BASELINE config 5."""
    g, m, sd, sem, emo = _regression_case(reg)
    assert len(sd) == g["spec"]["n_keys"] and same_checksum(syn.checksum(sd), g["weights_checksum"])
    ln, inst = O.video_regression_forward(sd, sem, emo, reg, 6, dt_rank=8)
    assert rel_err(ln, g["ln"]) < 5e-5 and rel_err(inst, g["inst"]) < 5e-5


@pytest.mark.parametrize("name", ["full", "ragged"])
def test_amt_train_step_full_shape_golden(name):
    """BASELINE config 3 at the real shape (B=4, T=299, S=300): autograd over the oracle's forward + the loss of
    run_model_vevo.py:101-119 gives the reference's loss and gradients (PAD-free and ragged PAD-tail targets)."""
    g = load_golden("amt_train_step_full.pt")[name]
    s = g["spec"]
    _, sd = amt_state_dict(syn.vf_dim(s["motion_type"]), s["weight_seed"])
    assert same_checksum(syn.checksum({k: v for k, v in sd.items() if not k.endswith(".pe")}), g["weights_checksum"])
    inp = syn.make_inputs(s["batch"], s["input_seed"], s["tgt_len"], s["src_len"], s["motion_type"])
    if g["pad_tail"]:
        inp["tgt"] = syn.pad_targets(inp["tgt"], s["input_seed"])
        assert int((inp["tgt"] == syn.CHORD_PAD).sum()) > 0
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith(".pe")}
    full = dict(sd)
    full.update(leaves)
    y = O.amt_forward(full, inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
                      inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"])
    assert rel_err(y[:1], g["logits"]) < 5e-5
    ce = torch.nn.functional.cross_entropy(y.permute(0, 2, 1), inp["tgt"], ignore_index=158, label_smoothing=0.1)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(y, inp["tgt_emotion"])
    loss = 0.4 * ce + 0.6 * bce
    assert abs(float(loss) - g["loss"]) < 1e-5 * abs(g["loss"])
    loss.backward()
    for n, gn in g["grad_norms"].items():
        got = float(leaves[n].grad.double().norm())
        assert abs(got - gn) < 1e-3 * max(gn, 1e-9), n
    for n, gr in g["grads"].items():
        mine = leaves[n].grad
        if mine.shape != gr.shape:
            mine = mine[::8]
        assert rel_err(mine, gr) < 1e-3, n


@pytest.mark.parametrize("name", ["block_v0", "block_v1", "stack"])
def test_mamba_step_oracle_matches_reference_golden(name):
    """Recurrent single-token inference (MambaBlock.step / Mamba.step, mamba.py:100-108,407-470): T steps from the empty
    cache equal the reference's step outputs and final cache; for use_version 0 they also equal forward()."""
    from video2music_b200.mamba import MambaConfig, MambaBlock, Mamba
    g = load_golden("mamba_step.pt")[name]
    s = g["spec"]
    cfg = MambaConfig(d_model=128, n_layers=max(s["n_layers"], 1), use_version=s["use_version"])
    m = Mamba(cfg) if s["n_layers"] else MambaBlock(cfg)
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["weight_seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    x = syn.unit_uniform((s["B"], s["T"], 128), syn._gen(s["seed"], "x"))
    empty = lambda: (None, torch.zeros(s["B"], cfg.d_inner, cfg.d_conv - 1))
    ys = []
    with torch.no_grad():
        if s["n_layers"]:
            caches = [empty() for _ in range(s["n_layers"])]
            for t in range(s["T"]):
                y, caches = O.mamba_step(sd, x[:, t], caches, s["n_layers"], cfg.dt_rank, cfg.d_state)
                ys.append(y)
            last = caches[-1]
        else:
            cache = empty()
            for t in range(s["T"]):
                y, cache = O.mamba_block_step(sd, "", x[:, t], cache, cfg.dt_rank, cfg.d_state)
                ys.append(y)
            last = cache
    y = torch.stack(ys, 1)
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(last[0], g["h"]) < 1e-5 and rel_err(last[1], g["inputs"]) < 1e-6
    if s["use_version"] == 0:
        assert rel_err(y, g["y_forward"]) < 1e-5                  # step == forward (the reference's own consistency)
    else:
        assert rel_err(g["y"], g["y_forward"]) > 1e-2             # literal: step() ignores the mamba+ gate (mamba.py:430)


@pytest.mark.parametrize("ver", ["2.2", "2.0", "1.1", "3.1"])
def test_zoo_train_oracle_matches_reference_golden(ver):
    """torch autograd over the oracle's restatement of the model zoo == the reference's own training-step gradients
    (tests/golden/zoo_train.pt: train() mode, dropout 0): logits, loss, every parameter-gradient norm, the stored gradients."""
    g = load_golden("zoo_train.pt")[ver]
    c = g["spec"]
    from video2music_b200 import VideoMusicTransformer_V1, VideoMusicTransformer_V2, VideoMusicTransformer_V3
    cls = {"2": VideoMusicTransformer_V2, "1": VideoMusicTransformer_V1, "3": VideoMusicTransformer_V3}[ver[0]]
    m = cls(version_name=ver, n_layers=c["n_layers"], total_vf_dim=syn.vf_dim(0), dropout=0.0)
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=c["seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    inp = syn.make_inputs(c["B"], c["seed"], c["T"], c["S"], 0)
    a = (leaf, inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"],
         inp["feature_motion"], inp["feature_emotion"])
    if ver[0] == "2":
        y = O.v2_forward(*a, n_layers=c["n_layers"], version=ver, moe_k=6 if ver == "2.0" else 2)   # first scheduler step: k = 6
    elif ver[0] == "3":
        y = O.v3_forward(*a, n_layers=c["n_layers"], version=ver)
    else:
        y = O.v1_forward(*a, n_layers=c["n_layers"], version=ver)
    r = syn.unit_uniform((c["B"], c["T"], 159), syn._gen(c["seed"], "r"))
    loss = (y * r).sum()
    loss.backward()
    assert rel_err(y, g["logits"]) < 5e-5 and abs(float(loss) - g["loss"]) < 1e-4 * max(abs(g["loss"]), 1.0)
    grads = {n: v.grad for n, v in leaf.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    _check_grads(grads, g, 2e-4)
