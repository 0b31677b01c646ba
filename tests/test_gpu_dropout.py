"""Training-mode dropout outside the bf16 AMT path: the stateless mask (seed, row, column) of the GEMM epilogues is one
function shared by the fp32 GEMM epilogue, the fp32 attention forward / backward kernels, the element-wise dropout op and
the MoE / Mamba / regression modules.  The masks are recovered with `ops.dropout(ones)` and every result is compared with
torch computing the same expression under the SAME mask (the reference draws its masks from torch's Philox stream, so
individual draws are not comparable; what is pinned is the arithmetic around the mask)."""
import pytest
import torch

from conftest import rel_err
from video2music_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _u(shape, seed, name):
    return syn.unit_uniform(shape, syn._gen(seed, name))


def _mask(rows, cols, p, seed):
    """keep-mask * 1/(1-p) of the stateless hash at (row, col), through the op under test."""
    from video2music_b200 import ops
    return ops.dropout(torch.ones((rows, cols), device=DEV), p, seed)


def test_dropout_op_equals_the_fused_epilogue_masks():
    from video2music_b200 import ops
    from video2music_b200.autograd import DropoutFn
    M, N, K, p, seed = 301, 200, 64, 0.2, 777
    m = _mask(M, N, p, seed).cpu()
    scale = ops.drop_args(p, seed)[0]
    keep = float((m > 0).float().mean())
    assert abs(keep - (1 - 51 / 256)) < 0.01 and torch.allclose(m[m > 0], torch.tensor(scale))
    assert torch.equal(m, _mask(M, N, p, seed).cpu()) and not torch.equal(m, _mask(M, N, p, seed + 1).cpu())
    # the fp32 GEMM epilogue draws the same mask: zero weights, unit bias
    w = torch.zeros((N, K), device=DEV)
    z = ops.linear(torch.zeros((M, K), device=DEV), w, torch.ones((N,), device=DEV), dropout=(p, seed, False))
    assert torch.equal(z.cpu(), m)
    # ... and the bf16 tcgen05 epilogue (N multiple of 8)
    zb = ops.linear(torch.zeros((M, K), device=DEV, dtype=torch.bfloat16), w.bfloat16(), torch.ones((N,), device=DEV),
                    dropout=(p, seed, False), out_dtype=torch.float32)
    assert torch.equal((zb > 0).cpu(), m > 0)
    # autograd: the backward applies the same mask
    x = _u((M, N), 3, "x").to(DEV).requires_grad_(True)
    dy = _u((M, N), 3, "dy").to(DEV)
    y = DropoutFn.apply(x, p, seed)
    y.backward(dy)
    assert torch.equal(y.detach().cpu(), (x.detach().cpu() * m)) and torch.equal(x.grad.cpu(), dy.cpu() * m)


def test_linear_fp32_fused_dropout_forward_backward():
    """fp32 GEMM epilogue dropout (both placements) + dy_prep recomputing the mask, against torch with the same mask."""
    from video2music_b200 import ops
    from video2music_b200.autograd import LinearFn
    M, K, N, p, seed = 333, 96, 120, 0.1, 4242
    x, w, b, r, dy = _u((M, K), 41, "x"), _u((N, K), 41, "w") * 0.2, _u((N,), 41, "b") * 0.1, _u((M, N), 41, "r"), _u((M, N), 41, "dy")
    m = _mask(M, N, p, seed).cpu().double()
    for after_res, relu, with_res in ((False, False, True), (False, True, False), (True, False, True)):
        xf, wf, bf_, rf = (t.double().clone().requires_grad_(True) for t in (x, w, b, r))
        z = xf @ wf.t() + bf_
        if relu:
            z = z.relu()
        ref = (z + rf) * m if after_res else z * m + (rf if with_res else 0.0)
        ref.backward(dy.double())
        xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
        rd = r.to(DEV).requires_grad_(not after_res)
        y = LinearFn.apply(xd, wd, bd, wd, K, relu, 1.0, 0, rd if with_res else None, 0, torch.float32, (p, seed, after_res))
        y.backward(dy.to(DEV))
        assert rel_err(y, ref) < 1e-5
        assert rel_err(xd.grad, xf.grad) < 1e-5 and rel_err(wd.grad, wf.grad) < 1e-5 and rel_err(bd.grad, bf_.grad) < 1e-5
        if with_res and not after_res:
            assert rel_err(rd.grad, rf.grad) < 1e-6


@pytest.mark.parametrize("case", ["rpr_causal", "cross", "gqa"])
def test_attention_fp32_probability_dropout_forward_backward(case):
    """softmax -> dropout -> P V (rpr.py:407-414, F.multi_head_attention_forward, grouped_query_attention.py:152-153) on the
    fp32 kernels: output, returned (dropped) weights and dQ / dK / dV / dEr against torch autograd with the recovered mask."""
    from video2music_b200 import ops
    from video2music_b200 import autograd as ag
    from oracle import amt_oracle as O
    p, seed = 0.25, 99
    if case == "gqa":
        b, n, s, hq, hk, d = 2, 37, 45, 8, 2, 64
        q, k, v = _u((b, n, hq, d), 5, "q") - 0.5, _u((b, s, hk, d), 5, "k") - 0.5, _u((b, s, hk, d), 5, "v")
        dy = _u((n, b, hq, d), 5, "dy")
        qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
        out = ag.GqaAttnFn.apply(qd, kd, vd, False, d ** -0.5, (p, seed))
        out.backward(dy.to(DEV))
        m = _mask(b * hq * n, s, p, seed).cpu().double().view(b, hq, n, s)
        qr, kr, vr = (t.double().clone().requires_grad_(True) for t in (q, k, v))
        g = hq // hk
        kk = kr.permute(0, 2, 1, 3).repeat_interleave(g, dim=1)            # (b, hq, s, d): query head h*g+gi reads kv head h
        vv = vr.permute(0, 2, 1, 3).repeat_interleave(g, dim=1)
        P = torch.softmax(qr.permute(0, 2, 1, 3) @ kk.transpose(-1, -2) * d ** -0.5, -1) * m
        ref = (P @ vv).permute(2, 0, 1, 3)                                   # (n, b, hq, d)
        ref.backward(dy.double())
        assert rel_err(out, ref) < 5e-5
        for mine, theirs in ((qd.grad, qr.grad), (kd.grad, kr.grad), (vd.grad, vr.grad)):
            assert rel_err(mine, theirs) < 1e-4
        return
    B, H, E = 2, 4, 256
    L, S, causal, with_er = (41, 41, True, True) if case == "rpr_causal" else (29, 53, False, False)
    dh = E // H
    q, k, v = _u((L * B, E), 6, "q") - 0.5, _u((S * B, E), 6, "k") - 0.5, _u((S * B, E), 6, "v")
    er = (_u((64, dh), 6, "er") - 0.5) if with_er else None
    dy = _u((L * B, E), 6, "dy")
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    erd = er.to(DEV).requires_grad_(True) if with_er else None
    out, pw = ag.AttnRowsFn.apply(qd, kd, vd, erd, B, L, S, H, causal, (E, B * E), (E, B * E), True, (p, seed))
    out.backward(dy.to(DEV))
    m = _mask(B * H * L, S, p, seed).cpu().double().view(B, H, L, S)
    qr, kr, vr = (t.double().clone().requires_grad_(True) for t in (q, k, v))
    err = er.double().clone().requires_grad_(True) if with_er else None
    heads = lambda t, n: t.view(n, B, H, dh).permute(1, 2, 0, 3)              # rows (l, b) -> (B, H, n, dh)
    qh, kh, vh = heads(qr, L), heads(kr, S), heads(vr, S)
    sc = qh @ kh.transpose(-1, -2)
    if with_er:
        sc = sc + O.skew_closed_form(qh.reshape(B * H, L, dh), O.get_valid_embedding(err, L)).view(B, H, L, L)
    if causal:
        sc = sc + torch.triu(torch.full((L, S), float("-inf"), dtype=torch.float64), 1)
    P = torch.softmax(sc, -1) * m
    ref = (P @ vh).permute(2, 0, 1, 3).reshape(L * B, E)
    ref.backward(dy.double())
    assert rel_err(out, ref) < 5e-5                                          # fp32 kernels against float64 torch
    assert rel_err(pw.view(B, H, L, S), P) < 5e-5                           # the weights come back dropped, as in torch
    for mine, theirs in ((qd.grad, qr.grad), (kd.grad, kr.grad), (vd.grad, vr.grad)):
        assert rel_err(mine, theirs) < 1e-4
    if with_er:
        assert rel_err(erd.grad, err.grad) < 1e-4


@pytest.mark.parametrize("shared", [False, True])
def test_moe_layer_dropout_forward_backward_with_recovered_masks(shared):
    """(Shared)MoELayer in training mode with the reference's dropout sites (GLUExpert hidden rows, moe.py:48; expert output
    rows, moe.py:197): the forward and every gradient against torch autograd over the reference's loop with the masks our
    kernels drew (recovered per permuted row through the saved permutation)."""
    from video2music_b200 import GLUExpert, MoELayer, SharedMoELayer, ops
    d, ff, E, k, L, B, p_h, p_o = 64, 96, 6, 2, 23, 3, 0.1, 0.2
    cls = SharedMoELayer if shared else MoELayer
    mod = cls(GLUExpert(d, ff, p_h), d, n_experts=E, n_experts_per_token=k, dropout=p_o).train()
    mod.load_state_dict(syn.fill_like_reference_init({n: tuple(v.shape) for n, v in mod.state_dict().items()}, seed=17))
    mod = mod.to(DEV)
    x = (_u((L, B, d), 18, "x")).to(DEV).requires_grad_(True)
    r = _u((L, B, d), 18, "r").to(DEV)
    captured = {}
    orig = ops.moe_experts_fwd_saved

    def spy(*a, **kw):
        out, saved = orig(*a, **kw)
        captured["drops"], captured["perm"] = kw.get("drops"), saved[5].clone()
        return out, saved
    ops.moe_experts_fwd_saved = spy
    try:
        with ops.fixed_dropout_seed(5):
            y = mod(x)
            seeds_after = ops._drop_state["calls"]
    finally:
        ops.moe_experts_fwd_saved = orig
    (y * r).sum().backward()
    ph, sh, po, so = captured["drops"]
    assert (ph, po) == (p_h, p_o) and seeds_after == (3 if shared else 2)
    T = L * B
    perm = captured["perm"].cpu().long().view(T, k)                           # perm[t, r] = permuted row of copy (t, r)
    mh = _mask(T * k, ff, p_h, sh).cpu().double()
    mo = _mask(T * k, d, p_o, so).cpu().double()
    # torch restatement of moe.py:180-199 / 244-301 in float64 with those masks
    ref = cls(GLUExpert(d, ff, 0.0), d, n_experts=E, n_experts_per_token=k, dropout=0.0).double()
    ref.load_state_dict({n: v.double().cpu() for n, v in mod.state_dict().items()})
    xr = x.detach().cpu().double().requires_grad_(True)
    x2 = xr.view(T, d)
    logits = x2 @ ref.gate.weight.t() + ref.gate.bias
    wts, sel = torch.topk(logits, k)
    assert torch.equal(sel.view(L, B, k), mod.last_selected_experts.cpu())
    wts = torch.softmax(wts, -1)
    out = torch.zeros(T, d, dtype=torch.float64)
    silu = torch.nn.functional.silu
    for t in range(T):
        for rr in range(k):
            e = ref.experts[int(sel[t, rr])]
            row = int(perm[t, rr])
            h = (e.linear1(x2[t]) * silu(e.gate(x2[t]))) * mh[row]
            out[t] = out[t] + wts[t, rr] * (e.linear2(h) * mo[row])
    if shared:
        ms = _mask(T, ff, p_h, _third_seed(5)).cpu().double()                  # the shared expert's own hidden dropout
        se = ref.shared_expert
        out = out + (1.0 / k) * se.linear2((se.linear1(x2) * silu(se.gate(x2))) * ms)
    (out.view(L, B, d) * r.cpu().double()).sum().backward()
    assert rel_err(y, out.view(L, B, d)) < 2e-5
    assert rel_err(x.grad, xr.grad) < 5e-5
    mine = dict(mod.named_parameters())
    for n, pr in ref.named_parameters():
        if pr.grad is None:
            assert float(mine[n].grad.abs().max()) == 0.0, n
        else:
            assert rel_err(mine[n].grad, pr.grad) < 5e-5, n


def _third_seed(fixed):
    from video2music_b200 import ops
    with ops.fixed_dropout_seed(fixed):
        ops.next_dropout_seed(); ops.next_dropout_seed()
        return ops.next_dropout_seed()


def test_bimamba_and_regression_train_with_their_default_dropout():
    """BiMambaEncoderLayer (dropout after each branch and inside the FFN, bimamba.py:50-98) and
    VideoRegression(regModel='sharedmoe_bimamba+') with its default dropout 0.1 (video_regression.py:176,203): training-mode
    forwards are stochastic, reproducible under a fixed seed, eval is deterministic, gradients flow to every parameter and a
    few Adam steps reduce the loss."""
    from video2music_b200 import VideoRegression, ops
    from video2music_b200.mamba import BiMambaEncoderLayer, MambaConfig
    lay = BiMambaEncoderLayer(MambaConfig(d_model=64, n_layers=1), dim_feedforward=96, dropout=0.2).to(DEV).train()
    x = _u((3, 40, 64), 8, "x").to(DEV)
    with ops.fixed_dropout_seed(1):
        a = lay(x).detach()
    with ops.fixed_dropout_seed(1):
        b = lay(x).detach()
    with ops.fixed_dropout_seed(2):
        c = lay(x).detach()
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()
    lay.eval()
    with torch.no_grad():
        assert torch.equal(lay(x), lay(x))
    torch.manual_seed(3)
    reg = VideoRegression(n_layers=2, d_model=64, d_hidden=96, total_vf_dim=774, regModel="sharedmoe_bimamba+").to(DEV).train()
    assert reg.in_proj[1].p == 0.1
    sem, emo = _u((4, 50, 768), 9, "sem").to(DEV), torch.softmax(_u((4, 50, 6), 9, "emo"), -1).to(DEV)
    zz = torch.zeros(4, 50, device=DEV)
    tgt = _u((4, 50, 2), 9, "t").to(DEV)
    opt = torch.optim.Adam(reg.parameters(), lr=2e-3)
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        ln, inst = reg(sem, zz, zz, emo)
        loss = ((ln - tgt) ** 2).mean() + 0.1 * inst.mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses) and min(losses[-3:]) < losses[0]
    missing = [n for n, p_ in reg.named_parameters() if p_.grad is None and "bias" not in n.split(".")[-1]]
    assert not [n for n in missing if "experts" not in n], missing          # (an expert without tokens may have no gradient)


def test_amt_fp32_training_with_dropout():
    """The fp32 exact path now trains with the reference's dropout as well (fp32 GEMM epilogue + fp32 attention kernels)."""
    from video2music_b200 import VideoMusicTransformer
    m = VideoMusicTransformer(n_layers=2, total_vf_dim=syn.vf_dim(0), rpr=True, dropout=0.1, max_sequence_chord=64, max_sequence_video=48)
    m.load_state_dict(syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=1), strict=False)
    m = m.to(DEV).train()
    inp = {k: v.to(DEV) for k, v in syn.make_inputs(3, 99, 40, 48, 0).items()}
    args = [inp[k] for k in ("x", "x_root", "x_attr", "feature_semantic_list", "feature_key", "feature_scene_offset",
                             "feature_motion", "feature_emotion")]
    y1, y2 = m(*args), m(*args)
    assert torch.isfinite(y1).all() and not torch.equal(y1.detach(), y2.detach())
    y1.sum().backward()
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    with torch.no_grad():
        assert torch.equal(m(*args), m(*args))
