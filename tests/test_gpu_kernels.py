"""GPU parity tests of the individual kernels, called through the C ABI (ctypes) exactly like the product."""
import pytest
import torch

from conftest import load_golden, rel_err, same_checksum
from oracle import amt_oracle as O
from video2music_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _u(shape, seed, name="t"):
    return syn.unit_uniform(shape, syn._gen(seed, name))


# ---------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(1, 8, 4), (111, 159, 512), (300, 512, 776), (257, 1536, 512), (64, 512, 1287), (0, 16, 16)])
def test_gemm_f32(M, N, K):
    from video2music_b200 import ops
    a, w, b = _u((M, K), 1, "a"), _u((N, K), 1, "w") * 0.1, _u((N,), 1, "b")
    res = _u((max(M, 1), N), 1, "r")[:M]
    y = ops.linear(a.to(DEV), w.to(DEV), b.to(DEV), relu=True, alpha=0.5, alpha_cols=N // 2, residual=res.to(DEV))
    ref = a.double() @ w.double().T + b.double()
    ref[:, : N // 2] *= 0.5
    ref = torch.relu(ref) + res.double()
    assert y.shape == (M, N)
    if M:
        assert rel_err(y, ref) < 1e-5


def test_gemm_f32_rowscale_resmod_and_batch_invariance():
    from video2music_b200 import ops
    M, N, K, T = 90, 512, 512, 30
    a, w = _u((M, K), 2, "a").to(DEV), (_u((N, K), 2, "w") * 0.1).to(DEV)
    rs, cv, pe = _u((M,), 2, "rs").to(DEV), _u((N,), 2, "cv").to(DEV), _u((T, N), 2, "pe").to(DEV)
    y = ops.linear(a, w, None, row_scale=rs, col_vec=cv, residual=pe, res_mod=T)
    ref = a.double() @ w.double().T + rs.double()[:, None] * cv.double()[None] + pe.double().repeat(M // T, 1)
    assert rel_err(y, ref) < 1e-5
    y1 = ops.linear(a[7:8].contiguous(), w, None)
    yb = ops.linear(a, w, None)
    assert torch.equal(y1[0], yb[7])              # same bits whatever the batch


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 512), (300, 512, 776), (1000, 1536, 512), (19200, 512, 1024),
                                   (77, 159, 512), (64, 1024, 1288)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_gemm_bf16_tcgen05(M, N, K, out_bf16):
    from video2music_b200 import ops
    a = _u((M, K), 3, "a").to(DEV).to(torch.bfloat16)
    w = (_u((N, K), 3, "w") * 0.1).to(DEV).to(torch.bfloat16)
    b = _u((N,), 3, "b").to(DEV)
    res = _u((M, N), 3, "r").to(DEV)
    y = ops.linear(a, w, b, alpha=0.25, alpha_cols=64, relu=False, residual=res,
                   out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ref = a.double() @ w.double().T + b.double()
    ref[:, :64] *= 0.25
    ref = ref + res.double()
    tol = 1.2e-2 if out_bf16 else 2e-3
    assert rel_err(y.float(), ref) < tol, "tcgen05 GEMM mismatch"


def test_gemm_bf16_exact_small_integers():
    """Integer-valued inputs make the bf16 tensor-core result exact: catches any descriptor / swizzle slip."""
    from video2music_b200 import ops
    M, N, K = 256, 256, 192
    g = syn._gen(5, "ints")
    a = torch.randint(-3, 4, (M, K), generator=g).float()
    w = torch.randint(-3, 4, (N, K), generator=g).float()
    y = ops.linear(a.to(DEV).to(torch.bfloat16), w.to(DEV).to(torch.bfloat16), None, out_dtype=torch.float32)
    assert torch.equal(y.cpu(), a @ w.T)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 192), (300, 512, 130), (159, 512, 128), (1536, 512, 600)])
def test_gemm_bf16_transposed_operands_exact(M, N, K, a_mn, b_mn):
    """MN-major (transposed-storage) operands of the tcgen05 GEMM, as used by dX = dY W and dW = dY^T X:
    small-integer inputs make the result exact."""
    from video2music_b200 import ops
    g = syn._gen(15, "ints2")
    a = torch.randint(-3, 4, (M, K), generator=g).float()
    b = torch.randint(-3, 4, (N, K), generator=g).float()
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8)).to(DEV).to(torch.bfloat16)[:, :t.shape[1]]
    ad = pad(a.t().contiguous()) if a_mn else pad(a)
    bd = pad(b.t().contiguous()) if b_mn else pad(b)
    y = ops.linear_general(ad, bd, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, out_dtype=torch.float32)
    assert torch.equal(y.cpu(), a @ b.T)


def test_gemm_head_scatter():
    from video2music_b200 import ops
    B, S, H, dh, E = 3, 50, 4, 64, 256
    x, w, b = _u((B * S, E), 6, "x"), _u((2 * H * dh, E), 6, "w") * 0.1, _u((2 * H * dh,), 6, "b")
    ref = (x.double() @ w.double().T + b.double()).view(B, S, 2, H, dh).permute(2, 0, 3, 1, 4)
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1.2e-2)):
        out = torch.zeros((2, B, H, S, dh), device=DEV, dtype=dt)
        ops.linear(x.to(DEV).to(dt), w.to(DEV).to(dt), b.to(DEV), out=out,
                   head_scatter=dict(S=S, H=H, dh=dh, cap=S, pos0=0, part_stride=B * H * S * dh))
        assert rel_err(out.float(), ref) < tol


# ---------------------------------------------------------------- attention (fp32)
def _attn_ref(q, k, v, Er, causal):
    """q (B,H,L,dh) already scaled; oracle arithmetic of rpr.py:387-414 with the literal _skew."""
    B, H, L, dh = q.shape
    S = k.shape[2]
    w = torch.einsum("bhld,bhsd->bhls", q, k)
    if Er is not None:
        qe = torch.einsum("hld,md->hlm", q.reshape(B * H, L, dh), O.get_valid_embedding(Er, L))
        w = w + O.skew_literal(qe).view(B, H, L, L)
    if causal:
        w = w + torch.triu(torch.full((L, S), float("-inf")), diagonal=1 + S - L)
    p = torch.softmax(w, dim=-1)
    return torch.einsum("bhls,bhsd->bhld", p, v), p


@pytest.mark.parametrize("B,H,L,S,dh,er_len,causal", [(2, 8, 299, 299, 64, 300, True), (3, 4, 37, 37, 32, 64, True),
                                                      (2, 8, 299, 300, 64, 0, False), (1, 2, 1, 1, 64, 16, True),
                                                      (2, 4, 33, 33, 64, 40, False), (1, 8, 300, 300, 64, 300, True)])
def test_attention_f32(B, H, L, S, dh, er_len, causal):
    from video2music_b200 import ops
    q, k, v = _u((B, L, H, dh), 7, "q") * 0.3, _u((B, S, H, dh), 7, "k"), _u((B, S, H, dh), 7, "v")
    Er = _u((er_len, dh), 7, "er") if er_len else None
    ref, pref = _attn_ref(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), Er, causal)
    out = torch.empty((B, L, H, dh), device=DEV)
    p_out = torch.empty((B * H, L, S), device=DEV)
    ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh,
                  q_strides=(L * H * dh, H * dh), k_strides=(S * H * dh, H * dh), v_strides=(S * H * dh, H * dh),
                  o_strides=(L * H * dh, H * dh), causal=causal, Er=Er.to(DEV) if er_len else None, p_out=p_out)
    assert rel_err(out.permute(0, 2, 1, 3), ref) < 2e-5
    assert rel_err(p_out.view(B, H, L, S), pref) < 2e-5


@pytest.mark.parametrize("B,Hq,Hkv,L,S,er_len,causal", [
    (2, 8, 8, 299, 299, 300, True),      # decoder RPR self-attention, BASELINE shape
    (2, 8, 8, 300, 300, 0, False),       # encoder self-attention
    (2, 8, 8, 299, 300, 0, False),       # decoder cross-attention
    (3, 4, 4, 37, 37, 64, True),         # single tile, tiny band
    (2, 4, 4, 150, 150, 160, False),     # RPR without a mask
    (2, 8, 2, 130, 170, 0, False),       # grouped-query (Hkv < Hq), ragged
    (1, 2, 2, 1, 1, 16, True),
    # stock attentions -> the single-kernel tcgen05 backward (csrc/attn_bwd_tc5.cu): tails, exact tile multiples, the
    # largest supported lengths, odd / even numbers of query blocks, more (video, head) items than SMs
    (3, 4, 4, 37, 53, 0, False),
    (1, 2, 2, 64, 128, 0, False),
    (2, 4, 4, 65, 129, 0, False),
    (2, 4, 4, 320, 320, 0, False),
    (2, 2, 2, 257, 200, 0, False),
    (20, 8, 8, 100, 90, 0, False),
    (1, 1, 1, 1, 1, 0, False),
])
def test_attention_bwd_tensor_core(B, Hq, Hkv, L, S, er_len, causal):
    """mma.sync attention backward (csrc/attn_bwd_tc.cu) against torch autograd over the oracle's arithmetic (the
    reference's gradients come from autograd over the same ops, rpr.py:387-414) on the same bf16-rounded inputs."""
    from video2music_b200 import ops
    dh, bf = 64, torch.bfloat16
    q = (_u((B, L, Hq, dh), 11, "q") * 0.3).to(bf)
    k, v = _u((B, S, Hkv, dh), 11, "k").to(bf), _u((B, S, Hkv, dh), 11, "v").to(bf)
    dO = _u((B, L, Hq, dh), 11, "do").to(bf)
    Er = _u((er_len, dh), 11, "er").to(bf) if er_len else None
    g = Hq // Hkv
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    erf = Er.float().requires_grad_(True) if er_len else None
    ref, _ = _attn_ref(qf.permute(0, 2, 1, 3), kf.repeat_interleave(g, dim=2).permute(0, 2, 1, 3),
                       vf.repeat_interleave(g, dim=2).permute(0, 2, 1, 3), erf, causal)
    ref = ref.permute(0, 2, 1, 3)
    ref.backward(dO.float())
    qd, kd, vd, dOd = q.to(DEV), k.to(DEV), v.to(DEV), dO.to(DEV)
    out = torch.zeros((B, L, Hq, dh), device=DEV, dtype=bf)
    lse = torch.zeros((B * Hq, L), device=DEV)
    qs, ks = (L * Hq * dh, Hq * dh), (S * Hkv * dh, Hkv * dh)
    ops.attention(qd, kd, vd, out, B=B, Hq=Hq, Hkv=Hkv, Lq=L, Lk=S, dh=dh, q_strides=qs, k_strides=ks, v_strides=ks,
                  o_strides=qs, causal=causal, Er=Er.to(DEV) if er_len else None, lse=lse)
    dq = torch.full_like(qd, float("nan"))
    dk, dv = torch.full_like(kd, float("nan")), torch.full_like(vd, float("nan"))
    der = torch.zeros((er_len, dh), device=DEV) if er_len else None
    ops.attention_bwd(qd, kd, vd, out, dOd, lse, Er.to(DEV) if er_len else None, dq, dk, dv, der, B=B, Hq=Hq, Hkv=Hkv, Lq=L,
                      Lk=S, dh=dh, q_strides=qs, k_strides=ks, v_strides=ks, o_strides=qs, do_strides=qs, dq_strides=qs,
                      dkv_strides=ks, causal=causal, tensor_core=True)
    errs = {"dq": rel_err(dq.float().cpu(), qf.grad), "dk": rel_err(dk.float().cpu(), kf.grad),
            "dv": rel_err(dv.float().cpu(), vf.grad)}
    if er_len:
        errs["dEr"] = rel_err(der.cpu(), erf.grad)
    print("attention backward (tensor core) rel errs", {n: "%.2e" % e for n, e in errs.items()})
    for n, e in errs.items():
        assert e < 2e-2, n


@pytest.mark.parametrize("B,Hq,Hkv,L,S,er_len,causal,seq_first", [
    (2, 8, 8, 299, 299, 300, True, False),      # decoder RPR self-attention, BASELINE shape
    (1, 8, 8, 300, 300, 300, True, False),
    (2, 8, 8, 300, 300, 0, False, False),       # encoder self-attention
    (2, 8, 8, 299, 300, 0, False, False),       # decoder cross-attention
    (3, 4, 4, 37, 37, 64, True, False),         # single tile, tiny band
    (2, 4, 4, 200, 200, 300, True, True),       # sequence-first (L, B, E) layout of the module API, QE half B in use
    (2, 4, 4, 150, 150, 160, False, False),     # RPR without a mask (mask=False path)
    (2, 8, 2, 130, 170, 0, False, False),       # grouped-query (Hkv < Hq), ragged
    (2, 8, 8, 1, 1, 300, True, False),
])
def test_attention_bf16_tcgen05(B, Hq, Hkv, L, S, er_len, causal, seq_first):
    from video2music_b200 import ops
    dh = 64
    bf = torch.bfloat16
    q = (_u((B, L, Hq, dh), 9, "q") * 0.3).to(bf)
    k, v = _u((B, S, Hkv, dh), 9, "k").to(bf), _u((B, S, Hkv, dh), 9, "v").to(bf)
    Er = _u((er_len, dh), 9, "er").to(bf) if er_len else None
    g = Hq // Hkv
    kx = k.float().repeat_interleave(g, dim=2)
    vx = v.float().repeat_interleave(g, dim=2)
    ref, _ = _attn_ref(q.float().permute(0, 2, 1, 3), kx.permute(0, 2, 1, 3), vx.permute(0, 2, 1, 3),
                       Er.float() if er_len else None, causal)
    ref = ref.permute(0, 2, 1, 3)                                                  # (B, L, H, dh)
    if seq_first:
        qd, kd, vd = (t.permute(1, 0, 2, 3).contiguous().to(DEV) for t in (q, k, v))  # (L, B, H, dh)
        out = torch.zeros((L, B, Hq, dh), device=DEV, dtype=bf)
        qs, ks, os_ = (Hq * dh, B * Hq * dh), (Hkv * dh, B * Hkv * dh), (Hq * dh, B * Hq * dh)
    else:
        qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
        out = torch.zeros((B, L, Hq, dh), device=DEV, dtype=bf)
        qs, ks, os_ = (L * Hq * dh, Hq * dh), (S * Hkv * dh, Hkv * dh), (L * Hq * dh, Hq * dh)
    lse = torch.zeros((B * Hq, L), device=DEV)
    ops.attention(qd, kd, vd, out, B=B, Hq=Hq, Hkv=Hkv, Lq=L, Lk=S, dh=dh, q_strides=qs, k_strides=ks, v_strides=ks,
                  o_strides=os_, causal=causal, Er=Er.to(DEV) if er_len else None, lse=lse)
    got = out.float().cpu()
    if seq_first:
        got = got.permute(1, 0, 2, 3)
    err = rel_err(got, ref)
    print("bf16 attention rel err", err)
    assert err < 1.5e-2
    assert torch.isfinite(lse).all()


def test_rpr_module_golden_gpu():
    from video2music_b200 import MultiheadAttentionRPR
    g = load_golden("rpr_attention.pt")
    for case in g["cases"]:
        s = case["spec"]
        mod = MultiheadAttentionRPR(s["E"], s["H"], dropout=0.0, er_len=s["er_len"]).eval()
        shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        mod.load_state_dict(syn.fill_like_reference_init(shapes, seed=s["seed"]))
        mod = mod.to(DEV)
        x = _u((s["L"], s["B"], s["E"]), s["seed"], "x").to(DEV)
        mask = torch.triu(torch.full((s["L"], s["L"]), float("-inf"), device=DEV), diagonal=1)
        with torch.no_grad():
            out, w = mod(x, x, x, attn_mask=mask)
        assert rel_err(out, case["out"]) < 1e-4
        if case["weights_mean"] is not None:
            assert rel_err(w, case["weights_mean"]) < 1e-4


def test_decoder_layer_golden_gpu():
    from video2music_b200 import TransformerDecoderLayerRPR
    g = load_golden("rpr_attention.pt")["layer"]
    s = g["spec"]
    layer = TransformerDecoderLayerRPR(s["E"], s["H"], s["ff"], 0.0, er_len=s["er_len"]).eval()
    shapes = {k: tuple(v.shape) for k, v in layer.state_dict().items()}
    sd = syn.fill_like_reference_init(shapes, seed=s["seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    layer.load_state_dict(sd)
    layer = layer.to(DEV)
    tgt, mem = _u((s["T"], s["B"], s["E"]), s["seed"], "tgt").to(DEV), _u((s["S"], s["B"], s["E"]), s["seed"], "mem").to(DEV)
    mask = torch.triu(torch.full((s["T"], s["T"]), float("-inf"), device=DEV), diagonal=1)
    with torch.no_grad():
        y = layer(tgt, mem, tgt_mask=mask)
    assert rel_err(y, g["out"]) < 1e-4


# ---------------------------------------------------------------- layernorm
@pytest.mark.parametrize("M,D", [(1, 128), (77, 512), (300, 1024), (5, 96)])
def test_layernorm(M, D):
    from video2music_b200 import ops
    x, r, g, b = _u((M, D), 8, "x"), _u((M, D), 8, "r"), 1 + 0.1 * _u((D,), 8, "g"), _u((D,), 8, "b")
    y = ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), res=r.to(DEV))
    ref = torch.nn.functional.layer_norm((x + r).double(), (D,), g.double(), b.double(), 1e-5)
    assert rel_err(y, ref) < 1e-5
    yb = ops.layernorm(x.to(DEV).to(torch.bfloat16), g.to(DEV), b.to(DEV))
    ref = torch.nn.functional.layer_norm(x.to(torch.bfloat16).double(), (D,), g.double(), b.double(), 1e-5)
    assert rel_err(yb.float(), ref) < 1e-2


# ---------------------------------------------------------------- pscan
def test_pscan_golden_gpu():
    from video2music_b200 import pscan
    for case in load_golden("pscan.pt")["cases"]:
        s = case["spec"]
        shp = (s["B"], s["L"], s["D"], s["N"])
        A = (torch.rand(shp, generator=syn._gen(s["seed"], "A")) * 0.99).to(DEV).requires_grad_(True)
        X = _u(shp, s["seed"], "X").to(DEV).requires_grad_(True)
        gH = _u(shp, s["seed"], "gH").to(DEV)
        A0, X0 = A.detach().clone(), X.detach().clone()
        H = pscan(A, X)
        (H * gH).sum().backward()
        assert torch.equal(A.detach(), A0) and torch.equal(X.detach(), X0)       # inputs untouched
        assert rel_err(H, case["H"]) < 1e-5
        assert rel_err(X.grad, case["gX"]) < 1e-5
        if float(case["gA"].abs().max()) > 0:
            assert rel_err(A.grad, case["gA"]) < 1e-5


@pytest.mark.parametrize("B,L,D,N", [(64, 300, 256, 16), (1, 4096, 256, 16), (8, 4096, 256, 16), (2, 700, 4, 16), (2, 1500, 8, 16)])
def test_pscan_full_size_properties(B, L, D, N):
    """Full BASELINE sizes: recurrence residual H[t] - A[t]*H[t-1] - X[t] == 0 (size-independent property),
    a sampled channel against the sequential oracle, and linearity in X."""
    from video2music_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    A = (torch.rand((B, L, D, N), generator=g) * 0.99).to(DEV)
    X = syn.unit_uniform((B, L, D, N), g).to(DEV)
    H = ops.pscan_fwd(A, X)
    resid = H[:, 1:] - (A[:, 1:] * H[:, :-1] + X[:, 1:])
    assert float(resid.abs().max()) < 1e-4 * float(H.abs().max())
    assert torch.allclose(H[:, 0], X[:, 0])
    Hs = O.pscan_forward(A[:1, :, :2].cpu(), X[:1, :, :2].cpu())
    assert rel_err(H[:1, :, :2], Hs) < 1e-5
    H2 = ops.pscan_fwd(A, 2.0 * X)
    assert rel_err(H2, 2.0 * H) < 1e-6
    gH = syn.unit_uniform((B, L, D, N), g).to(DEV)
    gA, gX = ops.pscan_bwd(A, H, gH)
    gAs, gXs = O.pscan_backward(A[:1, :, :2].cpu(), Hs, gH[:1, :, :2].cpu())
    assert rel_err(gX[:1, :, :2], gXs) < 1e-5 and rel_err(gA[:1, :, :2], gAs) < 1e-5


@pytest.mark.parametrize("B,L,D,N", [(1, 1, 2, 16), (3, 7, 2, 16), (2, 8, 2, 16), (2, 9, 4, 16), (1, 129, 2, 16), (5, 257, 6, 16),
                                     (2, 300, 4, 16), (1, 1031, 2, 16), (148, 65, 4, 16), (37, 64, 16, 16)])
def test_pscan_edge_shapes_vs_oracle(B, L, D, N):
    """Chunk / item boundaries of the TMA streaming kernel: one step, a single partial chunk, L just over one / two chunks, a ragged
    last chunk, 32- and 64-channel items, more items than SMs and fewer -- the WHOLE tensors against the sequential oracle."""
    from video2music_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11 + L)
    A = (torch.rand((B, L, D, N), generator=g) * 0.99)
    X = syn.unit_uniform((B, L, D, N), g)
    gH = syn.unit_uniform((B, L, D, N), g)
    H = ops.pscan_fwd(A.to(DEV), X.to(DEV))
    Hs = O.pscan_forward(A, X)
    assert rel_err(H, Hs) < 1e-5
    gA, gX = ops.pscan_bwd(A.to(DEV), H, gH.to(DEV))
    gAs, gXs = O.pscan_backward(A, Hs, gH)
    assert rel_err(gX, gXs) < 1e-5
    assert rel_err(gA, gAs) < 1e-5 if L > 1 else float(gA.abs().max()) == 0.0


def test_pscan_unaligned_pointers_take_the_first_kernel():
    """Views whose storage offset is not a multiple of 16 bytes cannot be described by a tensor map: the register kernel runs."""
    from video2music_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    B, L, D, N = 2, 77, 4, 16
    n = B * L * D * N
    bufA, bufX = (torch.rand(n + 1, generator=g) * 0.99).to(DEV), syn.unit_uniform((n + 1,), g).to(DEV)
    A, X = bufA[1:].view(B, L, D, N), bufX[1:].view(B, L, D, N)
    assert A.data_ptr() % 16 != 0 and A.is_contiguous()
    H = ops.pscan_fwd(A, X)
    assert rel_err(H, O.pscan_forward(A.cpu(), X.cpu())) < 1e-5


# ---------------------------------------------------------------- MoE
@pytest.mark.parametrize("shared", [False, True])
def test_moe_golden_gpu(shared):
    from video2music_b200 import GLUExpert, MoELayer, SharedMoELayer
    g = load_golden("moe.pt")["shared_%s" % shared]
    s = g["spec"]
    cls = SharedMoELayer if shared else MoELayer
    mod = cls(GLUExpert(s["d"], s["ff"], 0.0), s["d"], n_experts=s["n_experts"], n_experts_per_token=s["k"], dropout=0.0).eval()
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = syn.fill_like_reference_init(shapes, seed=s["seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    mod.load_state_dict(sd)
    mod = mod.to(DEV)
    x = _u((s["L"], s["B"], s["d"]), s["x_seed"], "x").to(DEV)
    with torch.no_grad():
        y = mod(x)
    assert g["min_rank_gap"] > 1e-4                                        # no near-tie in the reference's own logits
    assert torch.equal(mod.last_selected_experts.cpu(), g["selected_experts"])   # routing indices bit-exact
    assert rel_err(y, g["out"]) < 1e-4


@pytest.mark.parametrize("shared", [False, True])
def test_moe_golden_gpu_bf16_tensor_core(shared):
    """Expert GEMMs on the grouped tcgen05 path: routing bit-exact (fp32 router), outputs within the bf16 tolerance."""
    from video2music_b200 import GLUExpert, MoELayer, SharedMoELayer
    g = load_golden("moe.pt")["shared_%s" % shared]
    s = g["spec"]
    cls = SharedMoELayer if shared else MoELayer
    mod = cls(GLUExpert(s["d"], s["ff"], 0.0), s["d"], n_experts=s["n_experts"], n_experts_per_token=s["k"], dropout=0.0).eval()
    mod.load_state_dict(syn.fill_like_reference_init({k: tuple(v.shape) for k, v in mod.state_dict().items()}, seed=s["seed"]))
    mod = mod.to(DEV)
    mod.compute_dtype = torch.bfloat16
    x = _u((s["L"], s["B"], s["d"]), s["x_seed"], "x").to(DEV)
    with torch.no_grad():
        y = mod(x)
    assert torch.equal(mod.last_selected_experts.cpu(), g["selected_experts"])
    assert rel_err(y, g["out"]) < 2e-2


def test_moe_bf16_skewed_groups():
    """Ragged groups on the tensor-core path (one empty expert, one holding most tokens) vs the fp32 path of the same module."""
    from video2music_b200 import GLUExpert, MoELayer
    E, k, d, ff, tokens = 6, 2, 256, 320, 1111
    mod = MoELayer(GLUExpert(d, ff, 0.0), d, n_experts=E, n_experts_per_token=k, dropout=0.0).eval()
    sd = syn.fill_like_reference_init({n: tuple(v.shape) for n, v in mod.state_dict().items()}, seed=78)
    sd["gate.bias"] = torch.tensor([3.0, -50.0, 1.0, 0.5, 0.0, 0.0])
    mod.load_state_dict(sd)
    mod = mod.to(DEV)
    x = _u((tokens, 1, d), 6, "x").to(DEV)
    with torch.no_grad():
        y32 = mod(x)
        mod.compute_dtype = torch.bfloat16
        y16 = mod(x)
    assert int((mod.last_selected_experts == 1).sum()) == 0
    assert rel_err(y16, y32) < 2e-2


@pytest.mark.parametrize("tokens,E,k,d,ff", [(203, 6, 2, 64, 80), (1, 4, 4, 32, 48), (1500, 8, 3, 128, 272), (300, 6, 2, 128, 257)])
def test_moe_dispatch_ragged_groups_vs_oracle(tokens, E, k, d, ff):
    """Skewed router (one expert takes most tokens, some take none): group sizes are neither equal nor tile multiples."""
    from video2music_b200 import GLUExpert, MoELayer
    mod = MoELayer(GLUExpert(d, ff, 0.0), d, n_experts=E, n_experts_per_token=k, dropout=0.0).eval()
    sd = syn.fill_like_reference_init({n: tuple(v.shape) for n, v in mod.state_dict().items()}, seed=77)
    sd["gate.bias"] = torch.tensor([3.0, -50.0, 1.0, 0.5] + [0.0] * (E - 4))[:E]      # expert 1 never chosen when k < E
    mod.load_state_dict(sd)
    mod = mod.to(DEV)
    x = _u((tokens, 1, d), 5, "x")
    with torch.no_grad():
        y = mod(x.to(DEV))
        y_again = mod(x.to(DEV))
    ref, idx, _ = O.moe_layer(x, sd, "", E, k)
    assert torch.equal(mod.last_selected_experts.cpu(), idx)
    assert rel_err(y, ref) < 1e-4
    assert torch.equal(y, y_again)                    # row order inside a group is arbitrary, the result is not
    if k < E:
        assert int((idx == 1).sum()) == 0
    # an in-place weight update must invalidate the stacked-weight cache
    with torch.no_grad():
        mod.experts[0].linear2.bias.add_(1.0)
        y2 = mod(x.to(DEV))
    sd["experts.0.linear2.bias"] = sd["experts.0.linear2.bias"] + 1.0
    assert rel_err(y2, O.moe_layer(x, sd, "", E, k)[0]) < 1e-4


def test_moe_route_ties_and_bias():
    from video2music_b200 import ops
    x = torch.zeros((5, 64), device=DEV)
    wg = torch.zeros((6, 64), device=DEV)
    bg = torch.tensor([0.5, 0.5, 0.1, 0.9, 0.9, 0.0], device=DEV)
    idx, w, hist, _ = ops.moe_route(x, wg, bg, 3)
    assert idx[0].tolist() == [3, 4, 0]                                     # lowest index first on exact ties
    assert hist.tolist() == [5, 0, 0, 5, 5, 0]
    assert torch.allclose(w.sum(-1), torch.ones(5, device=DEV))
    idx2, w2, _, _ = ops.moe_route(x, wg, bg, 2, sel_bias=torch.tensor([0., 0., 5., 0., 0., 0.], device=DEV))
    assert idx2[0].tolist() == [2, 3]
    assert torch.allclose(w2[0], torch.softmax(torch.tensor([0.1, 0.9]), 0).to(DEV))    # weights from un-biased logits


@pytest.mark.parametrize("name", ["post_ln_moe", "post_ln_sharedmoe_b2", "pre_rms_moe"])
def test_variant_gqa_moe_stack_golden_gpu(name):
    """BASELINE config 4 through our drop-in wrappers (custom_transformer.py) + MultiheadGQA + (Shared)MoELayer."""
    from test_oracle import _variant_net
    g = load_golden("variant.pt")[name]
    c = g["spec"]
    net, sd = _variant_net(c)
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    net.load_state_dict(sd)
    net = net.to(DEV)
    src = _u((c["S"], c["B"], 512), c["seed"], "src").to(DEV)
    tgt = _u((c["T"], c["B"], 512), c["seed"], "tgt").to(DEV)
    mask = torch.triu(torch.full((c["T"], c["T"]), float("-inf"), device=DEV), diagonal=1)
    with torch.no_grad():
        mem = net["enc"](src)
        y = net["dec"](tgt, mem, tgt_mask=mask)
    assert rel_err(mem, g["memory"]) < 1e-4 and rel_err(y, g["out"]) < 1e-4
    ref_mem, ref_y = O.variant_stack_forward(sd, src.cpu(), tgt.cpu(), 2, 8, 2, 6, 2, c["shared"], c["pre_norm"], c["rms"])
    assert rel_err(y, ref_y) < 1e-4


# ---------------------------------------------------------------- GQA
def test_gqa_golden_gpu():
    from video2music_b200 import MultiheadGQA, scaled_dot_product_gqa
    g = load_golden("gqa.pt")
    for case in g["function"]:
        s = case["spec"]
        q, k, v = (_u(shp, s["seed"], n).to(DEV) for shp, n in (((s["b"], s["n"], s["hq"], s["d"]), "q"),
                                                               ((s["b"], s["s"], s["hk"], s["d"]), "k"),
                                                               ((s["b"], s["s"], s["hk"], s["d"]), "v")))
        out, _ = scaled_dot_product_gqa(q, k, v, num_heads=s["hq"], is_causal=True if s["causal"] else None)
        assert rel_err(out, case["out"]) < 2e-5
    for case in g["module"]:
        s = case["spec"]
        mod = MultiheadGQA(s["E"], s["hq"], s["hk"], dropout=0.0).eval()
        shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        mod.load_state_dict(syn.fill_like_reference_init(shapes, seed=s["seed"]))
        mod = mod.to(DEV)
        xq, xk = _u((s["L"], s["B"], s["E"]), s["seed"], "xq").to(DEV), _u((s["S"], s["B"], s["E"]), s["seed"], "xk").to(DEV)
        with torch.no_grad():
            y, _ = mod(xq, xk, xk)
        assert rel_err(y, case["out"]) < 1e-4
    with pytest.raises(ValueError):
        scaled_dot_product_gqa(torch.zeros(1, 2, 3, 8, device=DEV), torch.zeros(1, 2, 2, 8, device=DEV),
                               torch.zeros(1, 2, 2, 8, device=DEV))


# ---------------------------------------------------------------- Mamba block (fused selective scan)
from test_oracle import _bimamba_v1  # noqa: E402


def test_mamba_golden_gpu():
    """MambaBlock (both versions), the 2-layer stack and the BiMamba layer on the GPU kernels vs the reference's outputs."""
    from video2music_b200.mamba import MambaConfig, MambaBlock, Mamba, BiMambaEncoderLayer
    g = load_golden("mamba.pt")
    cases = [("block_v0", lambda s: MambaBlock(MambaConfig(d_model=128, n_layers=1, use_version=0))),
             ("block_v1", lambda s: MambaBlock(MambaConfig(d_model=128, n_layers=1, use_version=1))),
             ("stack", lambda s: Mamba(MambaConfig(d_model=128, n_layers=2))),
             ("bimamba_layer", lambda s: BiMambaEncoderLayer(MambaConfig(d_model=128, n_layers=1), dim_feedforward=s["d_ff"])),
             ("bimamba_v1_ffn", _bimamba_v1), ("bimamba_v1_moe", _bimamba_v1)]       # Bi-Mamba+ (bimamba.py:101-191)
    for name, make in cases:
        c = g[name]
        s = c["spec"]
        m = make(s)
        sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["weight_seed"])
        assert same_checksum(syn.checksum(sd), c["weights_checksum"])
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        x = syn.unit_uniform((s["B"], s["L"], 128), syn._gen(s["seed"], "x")).to(DEV)
        with torch.no_grad():
            y = m(x)
        err = rel_err(y, c["y"])
        print("mamba %s rel err %.2e" % (name, err))
        assert err < 1e-4, name


@pytest.mark.parametrize("B,L", [(64, 300), (8, 4096)])
def test_selective_scan_fused_equals_pscan_composition(B, L):
    """BASELINE config 5 sizes: the fused scan (no (B, L, ED, N) tensors) equals exp/mul + the pscan kernel + contraction."""
    from video2music_b200 import ops
    ED, N = 256, 16
    gx = syn._gen(5, "scan")
    x = syn.unit_uniform((B * L, ED), gx).to(DEV)
    draw = (syn.unit_uniform((B * L, ED), gx) * 0.5 - 1.0).to(DEV)
    bias = (syn.unit_uniform((ED,), gx) * 0.1).to(DEV)
    A_log = torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(ED, 1).to(DEV)
    bc = syn.unit_uniform((B * L, 2 * N), gx).to(DEV)
    D = syn.unit_uniform((ED,), gx).to(DEV)
    y = ops.selective_scan(x, draw, bias, A_log, bc[:, :N], bc[:, N:], D, None, B, L)
    delta = torch.nn.functional.softplus(draw + bias).view(B, L, ED)
    A = -torch.exp(A_log)
    xs = x.view(B, L, ED)
    dA = torch.exp(delta.unsqueeze(-1) * A)
    BX = delta.unsqueeze(-1) * bc[:, :N].view(B, L, 1, N) * xs.unsqueeze(-1)
    hs = ops.pscan_fwd(dA, BX)
    ref = (hs * bc[:, N:].view(B, L, 1, N)).sum(-1) + D * xs
    assert rel_err(y.view(B, L, ED), ref) < 2e-5


@pytest.mark.parametrize("B,L,nvid", [(64, 300, 64), (8, 4096, 2)])
def test_mamba_block_full_size_vs_oracle(B, L, nvid):
    """BASELINE config 5 sizes against the ORACLE (O.mamba_block_forward, the sequential restatement of mamba.py:259-351), not
    against our own pscan kernel: the whole MambaBlock (in_proj, conv, SiLU, x_proj / dt_proj, fused selective scan, gate,
    out_proj) at (64, 300, 256, 16), and at the (8, 4096) stress shape for `nvid` of the videos (videos are independent;
    the oracle's (B, L, ED, N) tensors are 0.5 GB per video pair there)."""
    from video2music_b200.mamba import MambaConfig, MambaBlock
    for ver in (0, 1):
        cfg = MambaConfig(d_model=128, n_layers=1, use_version=ver)
        m = MambaBlock(cfg)
        sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=21 + ver)
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        x = syn.unit_uniform((B, L, 128), syn._gen(31, "x"))
        with torch.no_grad():
            y = m(x.to(DEV)).cpu()
            pick = torch.linspace(0, B - 1, nvid).round().long().unique()
            ref = O.mamba_block_forward(sd, "", x[pick], cfg.dt_rank, cfg.d_state, use_version=ver)
        err = rel_err(y[pick], ref)
        print("MambaBlock v%d (%d, %d): %d videos vs oracle, rel err %.2e" % (ver, B, L, len(pick), err))
        assert err < 1e-4


@pytest.mark.parametrize("B,L,nvid", [(64, 300, 8), (8, 4096, 1)])
def test_selective_scan_full_size_vs_oracle_scan(B, L, nvid):
    """The fused scan kernel alone at the config-5 sizes against the oracle's sequential scan (O.pscan_forward over
    exp(delta A), delta B x; mamba.py:343-352) on a sample of the videos."""
    from video2music_b200 import ops
    ED, N = 256, 16
    gx = syn._gen(6, "scan")
    x = syn.unit_uniform((B * L, ED), gx)
    draw = syn.unit_uniform((B * L, ED), gx) * 0.5 - 1.0
    bias = syn.unit_uniform((ED,), gx) * 0.1
    A_log = torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(ED, 1)
    bc = syn.unit_uniform((B * L, 2 * N), gx)
    D = syn.unit_uniform((ED,), gx)
    y = ops.selective_scan(x.to(DEV), draw.to(DEV), bias.to(DEV), A_log.to(DEV), bc[:, :N].to(DEV), bc[:, N:].to(DEV), D.to(DEV),
                           None, B, L).view(B, L, ED).cpu()
    pick = torch.linspace(0, B - 1, nvid).round().long().unique()
    xs = x.view(B, L, ED)[pick]
    delta = torch.nn.functional.softplus(draw + bias).view(B, L, ED)[pick]
    Bm, Cm = bc[:, :N].reshape(B, L, N)[pick], bc[:, N:].reshape(B, L, N)[pick]
    hs = O.pscan_forward(torch.exp(delta.unsqueeze(-1) * (-torch.exp(A_log))), delta.unsqueeze(-1) * Bm.unsqueeze(2) * xs.unsqueeze(-1))
    ref = (hs @ Cm.unsqueeze(-1)).squeeze(3) + D * xs
    err = rel_err(y[pick], ref)
    print("selective_scan (%d, %d): %d videos vs oracle scan, rel err %.2e" % (B, L, len(pick), err))
    assert err < 2e-5


@pytest.mark.parametrize("name", ["block_v0", "block_v1", "stack"])
def test_mamba_step_golden_gpu(name):
    """MambaBlock.step / Mamba.step on the GPU step kernels against the unmodified reference (mamba_step.pt): every step's
    output, the final cache, the caller's cache left untouched, and (use_version 0) agreement with our own forward()."""
    from video2music_b200.mamba import MambaConfig, MambaBlock, Mamba
    g = load_golden("mamba_step.pt")[name]
    s = g["spec"]
    cfg = MambaConfig(d_model=128, n_layers=max(s["n_layers"], 1), use_version=s["use_version"])
    m = Mamba(cfg) if s["n_layers"] else MambaBlock(cfg)
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=s["weight_seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = syn.unit_uniform((s["B"], s["T"], 128), syn._gen(s["seed"], "x")).to(DEV)
    empty = lambda: (None, torch.zeros(s["B"], cfg.d_inner, cfg.d_conv - 1, device=DEV))
    cache = [empty() for _ in range(s["n_layers"])] if s["n_layers"] else empty()
    ys = []
    with torch.no_grad():
        for t in range(s["T"]):
            before = cache
            keep = [c[1].clone() for c in cache] if s["n_layers"] else cache[1].clone()
            y, cache = m.step(x[:, t], cache)
            if s["n_layers"]:
                assert all(torch.equal(b[1], k_) for b, k_ in zip(before, keep))       # the caller's cache is not modified
            else:
                assert torch.equal(before[1], keep)
            ys.append(y)
        y = torch.stack(ys, 1)
        last = cache[-1] if s["n_layers"] else cache
        assert rel_err(y, g["y"]) < 1e-4 and rel_err(last[0], g["h"]) < 1e-4 and rel_err(last[1], g["inputs"]) < 1e-5
        if s["use_version"] == 0:
            assert rel_err(m(x), y) < 1e-4                      # recurrent steps == the chunked parallel scan of forward()


# ---------------------------------------------------------------- evaluation metrics
def test_amt_metrics_kernel_vs_reference_golden_and_oracle():
    from test_oracle import _metrics_case
    from video2music_b200 import ops
    for c in load_golden("metrics.pt")["cases"]:
        out, tgt = _metrics_case(c)
        cnt = ops.amt_metrics(out.to(DEV), tgt.to(DEV)).tolist()
        n_valid = int((tgt != 158).sum())
        assert cnt[0] == n_valid
        if n_valid == 0:
            continue
        assert abs(cnt[1] / n_valid - c["acc"]) < 1e-6
        for i, h in enumerate(c["hits"]):
            assert abs(cnt[2 + i] / n_valid - h) < 1e-6
    # a batch with exact ties: the lower class index wins, as torch.argmax / topk on equal values
    out = torch.zeros((2, 7, 159))
    tgt = torch.tensor([[0, 1, 2, 3, 4, 158, 158], [158, 0, 0, 0, 0, 0, 5]])
    cnt = ops.amt_metrics(out.to(DEV), tgt.to(DEV), ks=(1, 3, 5)).tolist()
    assert cnt == [11, 6, 6, 8, 10]
    assert abs(O.vevo_accuracy(out, tgt) - 6 / 11) < 1e-6


def test_amt_correspondence_kernel_vs_reference_golden_and_oracle():
    """Emotion correspondence of the evaluation loop (dataset/vevo_dataset.py:747-810) in one launch: equal to the reference's own
    values (golden) and to the oracle's counters; a batch of several videos is the sum over its rows."""
    from test_oracle import _correspondence_case as correspondence_case
    from video2music_b200 import ops
    outs, emos, probs, tot = [], [], [], [0, 0]
    for c in load_golden("metrics.pt")["correspondence"]:
        out, emo, prob = correspondence_case(c["seed"])
        v = ops.compute_vevo_correspondence(out.to(DEV), None, emo.to(DEV), prob.to(DEV), c["thr"])
        assert abs(v - c["value"]) < 1e-6, (c, v)
        _, pt, right = O.vevo_correspondence(out, emo, prob, c["thr"])
        assert ops.amt_correspondence(out.to(DEV), emo.to(DEV), prob.to(DEV), c["thr"]).tolist() == [pt, right]
        if c["thr"] == 0.8:
            outs.append(out); emos.append(emo); probs.append(prob)
            tot[0] += pt; tot[1] += right
    cnt = ops.amt_correspondence(torch.cat(outs).to(DEV), torch.cat(emos).to(DEV), torch.cat(probs).to(DEV), 0.8).tolist()
    assert cnt == tot and tot[0] > 0
    # arg-max ties: the lower chord id wins (torch.argmax), here chord 0 ("N" -> quality 1)
    out = torch.zeros((1, 3, 159)); emo = torch.zeros((1, 3, 159)); emo[0, :, 1] = 1.0; emo[0, 2, 1] = 0.0; emo[0, 2, 5] = 1.0
    assert ops.amt_correspondence(out.to(DEV), emo.to(DEV), torch.ones((1, 3), device=DEV), 0.8).tolist() == [3, 2]
    assert ops.compute_vevo_correspondence(out.to(DEV), None, emo[:, :0].to(DEV), torch.ones((1, 0), device=DEV), 0.8) == 1.0


# ---------------------------------------------------------------- kernels of one generation step (csrc/step_f32.cu)
@pytest.mark.parametrize("M,N,K", [(1, 159, 512), (3, 512, 512), (64, 1536, 512), (64, 512, 1024), (70, 128, 516), (130, 24, 64)])
def test_step_linear_vs_float64_and_batch_invariance(M, N, K):
    """The few-row linear layer of a KV-cached generation step: against a float64 reference (bias, ReLU, the rank-1
    row_scale x col_vec term of Linear_chord), ragged M / N / K tails, and bit-identical rows whatever the batch size."""
    from video2music_b200 import ops
    x, w, b = _u((M, K), 11, "x"), _u((N, K), 11, "w") * 0.1, _u((N,), 11, "b")
    rs, cv = _u((M,), 11, "rs"), _u((N,), 11, "cv")
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    ref = x.double() @ w.double().t() + b.double()
    y = ops.step_linear(xd, wd, bd)
    assert rel_err(y, ref.float()) < 2e-6
    y2 = ops.step_linear(xd, wd, bd, relu=True, row_scale=rs.to(DEV), col_vec=cv.to(DEV))
    assert rel_err(y2, torch.relu(ref + rs.double()[:, None] * cv.double()[None, :]).float()) < 2e-6
    assert rel_err(ops.step_linear(xd, wd, None), (ref - b.double()).float()) < 2e-6
    one = ops.step_linear(xd[M - 1:].contiguous(), wd, bd)
    assert torch.equal(one[0], y[M - 1])                                   # fixed summation order: no dependence on the row count
    wide = torch.cat([w, _u((N, 4), 12, "pad")], 1).to(DEV)               # leading dimension > K (k = K of a wider weight)
    assert torch.equal(ops.step_linear(xd, wide, bd, k=K), y)
    assert rel_err(y, ops.linear(xd, wd, bd)) < 2e-6                       # the tiled fp32 GEMM of the full forward


@pytest.mark.parametrize("Hq,Hkv,n", [(8, 8, 1), (8, 2, 77), (8, 2, 300), (4, 1, 129)])
def test_step_attention_vs_float64(Hq, Hkv, n):
    """One query row per (video, head) over the first n cached rows (n from a device word): against float64 softmax attention;
    rows beyond n are never read (NaN there), grouped heads share their kv head, slices of wider rows as caches."""
    from video2music_b200 import ops
    B, dh, cap = 5, 64, 304
    q = _u((B, Hq * dh), 21, "q")
    kv = _u((B, cap, 2 * Hkv * dh), 21, "kv")                              # K | V side by side: row stride 2 Hkv dh
    kv[:, n:] = float("nan")
    qd, kvd = q.to(DEV), kv.to(DEV)
    K, V = kvd[:, :, :Hkv * dh], kvd[:, :, Hkv * dh:]
    nd = torch.tensor([n], dtype=torch.int32, device=DEV)
    out = ops.step_attention(qd, K, V, Hq=Hq, Hkv=Hkv, dh=dh, n_max=cap, kv_strides=(K.stride(0), K.stride(1)), n_dev=nd, q_scale=0.125)
    qh = q.double().view(B, Hq, dh) * 0.125
    kh = kv[:, :n, :Hkv * dh].double().view(B, n, Hkv, dh).repeat_interleave(Hq // Hkv, dim=2)
    vh = kv[:, :n, Hkv * dh:].double().view(B, n, Hkv, dh).repeat_interleave(Hq // Hkv, dim=2)
    p = torch.softmax(torch.einsum("bhd,bnhd->bhn", qh, kh), dim=-1)
    ref = torch.einsum("bhn,bnhd->bhd", p, vh).reshape(B, Hq * dh)
    assert torch.isfinite(out).all() and rel_err(out, ref.float()) < 2e-6
    out_h = ops.step_attention(qd, K, V, Hq=Hq, Hkv=Hkv, dh=dh, n_max=n, kv_strides=(K.stride(0), K.stride(1)), q_scale=0.125)
    assert torch.equal(out, out_h)                                         # host-side count == device-side count


def test_moe_experts_small_batch_kernel_equals_tiled_kernel():
    """A generation step carries a few token copies per expert: moe_grouped_gemm then runs on the weight-streaming kernel of
    step_f32.cu (<= 256 copies); same outputs as the tiled kernel that the same tokens get inside a larger batch."""
    from video2music_b200 import GLUExpert, MoELayer
    E, k, d, ff = 6, 2, 512, 1024
    mod = MoELayer(GLUExpert(d, ff, 0.0), d, n_experts=E, n_experts_per_token=k, dropout=0.0).eval()
    sd = syn.fill_like_reference_init({n: tuple(v.shape) for n, v in mod.state_dict().items()}, seed=79)
    sd["gate.bias"] = torch.tensor([2.0, -50.0, 1.0, 0.5, 0.0, 0.0])      # one expert without tokens
    mod.load_state_dict(sd)
    mod = mod.to(DEV)
    x = _u((300, 1, d), 7, "x").to(DEV)
    with torch.no_grad():
        big = mod(x)                                                        # 600 copies: tiled kernel
        for m in (1, 64, 128):                                              # 2 .. 256 copies: step kernel
            small = mod(x[:m].contiguous())
            assert rel_err(small, big[:m]) < 1e-5, m


# ---------------------------------------------------------------- fused dropout (training)
def test_linear_fused_dropout_forward_backward():
    """Dropout fused into the GEMM epilogue (both placements) and re-derived by dy_prep in backward, against torch with the
    SAME mask (read back through a GEMM with zero weights and unit bias)."""
    from video2music_b200 import ops
    from video2music_b200.autograd import LinearFn
    bf = torch.bfloat16
    M, K, N, pdrop, seed = 333, 512, 384, 0.1, 12345
    x = _u((M, K), 41, "x").to(bf)
    w = (_u((N, K), 41, "w") * 0.05).to(bf)
    b = _u((N,), 41, "b") * 0.1
    r = _u((M, N), 41, "r").to(bf)
    dy = _u((M, N), 41, "dy").to(bf)
    xd, wd, bd, rd = x.to(DEV), w.to(DEV), b.to(DEV), r.to(DEV)
    ones = torch.ones((N,), device=DEV)
    m1 = ops.linear(torch.zeros_like(xd), wd, ones, dropout=(pdrop, seed, False)).float().cpu()
    mask = (m1 > 0).float()
    keep = float(mask.mean())
    assert abs(keep - (1 - pdrop)) < 0.01 and torch.allclose(m1[m1 > 0], torch.tensor(ops.drop_args(pdrop, seed)[0]), rtol=1e-2)
    m2 = ops.linear(torch.zeros_like(xd), wd, ones, dropout=(pdrop, seed, False)).float().cpu()
    m3 = ops.linear(torch.zeros_like(xd), wd, ones, dropout=(pdrop, seed + 1, False)).float().cpu()
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    scale = ops.drop_args(pdrop, seed)[0]            # the drop probability is quantised to 1/256
    # the three placements of the model: sub-layer output before the residual add, FFN hidden layer (ReLU, no residual),
    # positional-encoding dropout after the (constant) residual
    for after_res, relu, with_res in ((False, False, True), (False, True, False), (True, False, True)):
        xf, wf, bfp, rf = (t.float().clone().requires_grad_(True) for t in (x, w, b, r))    # fresh leaves every case
        z = xf @ wf.t() + bfp
        if relu:
            z = torch.relu(z)
        rterm = rf if with_res else 0.0
        ref = (z + rterm) * mask * scale if after_res else z * mask * scale + rterm
        ref.backward(dy.float())
        xg, wg, bg, rg = (t.detach().clone().requires_grad_(True) for t in (xd, wd.float(), bd, rd))
        res = None if not with_res else (rd if after_res else rg)
        y = LinearFn.apply(xg, wg, bg, wd, K, relu, 1.0, 0, res, 0, bf, (pdrop, seed, after_res))
        assert rel_err(y.float(), ref) < 2e-2
        y.backward(dy.to(DEV))
        assert rel_err(xg.grad.float(), xf.grad) < 2e-2 and rel_err(wg.grad, wf.grad) < 2e-2 and rel_err(bg.grad, bfp.grad) < 2e-2
        if with_res and not after_res:
            assert rel_err(rg.grad.float(), rf.grad) < 2e-2


@pytest.mark.parametrize("L,S,er_len,causal", [(150, 150, 160, True), (99, 140, 0, False), (299, 300, 0, False)])
def test_attention_probability_dropout_forward_backward(L, S, er_len, causal):
    """Dropout of the attention probabilities inside the fused forward kernel and the backward rows kernel, against torch
    autograd with the SAME mask (recovered by attending with q = k = 0 over one-hot V blocks)."""
    from video2music_b200 import ops
    B, H, dh, bf, pdrop, seed = 2, 4, 64, torch.bfloat16, 0.1, 777
    scale = ops.drop_args(pdrop, seed)[0]
    q = (_u((B, L, H, dh), 51, "q") * 0.3).to(bf)
    k, v = _u((B, S, H, dh), 51, "k").to(bf), _u((B, S, H, dh), 51, "v").to(bf)
    dO = _u((B, L, H, dh), 51, "do").to(bf)
    Er = _u((er_len, dh), 51, "er").to(bf) if er_len else None
    qs, ks = (L * H * dh, H * dh), (S * H * dh, H * dh)

    def fwd(qq, kk, vv, er, lse=None):
        out = torch.zeros((B, L, H, dh), device=DEV, dtype=bf)
        ops.attention(qq, kk, vv, out, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=qs, k_strides=ks, v_strides=ks, o_strides=qs,
                      causal=causal, Er=er, lse=lse, dropout=(pdrop, seed))
        return out
    # mask recovery: uniform probabilities, V = one-hot column blocks
    mask = torch.zeros((B, H, L, S))
    z_q, z_k = torch.zeros_like(q).to(DEV), torch.zeros_like(k).to(DEV)
    for blk in range((S + dh - 1) // dh):
        vv = torch.zeros((B, S, H, dh))
        for d in range(dh):
            if blk * dh + d < S:
                vv[:, blk * dh + d, :, d] = 1.0
        o = fwd(z_q, z_k, vv.to(bf).to(DEV), None).float().cpu()        # (B, L, H, dh): P_drop[i, blk*dh + d]
        n = min(dh, S - blk * dh)
        mask[:, :, :, blk * dh:blk * dh + n] = (o.permute(0, 2, 1, 3)[..., :n] > 0).float()
    vis = torch.ones((L, S)) if not causal else torch.tril(torch.ones((L, S)), diagonal=S - L)
    keep = float((mask * vis).sum() / (vis.sum() * B * H))
    assert abs(keep - (1 - pdrop)) < 0.01
    # reference with the explicit mask
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    erf = Er.float().requires_grad_(True) if er_len else None
    _, pref = _attn_ref(qf.permute(0, 2, 1, 3), kf.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), erf, causal)
    ref = torch.einsum("bhls,bhsd->bhld", pref * mask * scale, vf.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    ref.backward(dO.float())
    qd, kd, vd, dOd = q.to(DEV), k.to(DEV), v.to(DEV), dO.to(DEV)
    erd = Er.to(DEV) if er_len else None
    lse = torch.zeros((B * H, L), device=DEV)
    out = fwd(qd, kd, vd, erd, lse)
    assert rel_err(out.float(), ref) < 2e-2
    dq, dk, dv = torch.full_like(qd, float("nan")), torch.full_like(kd, float("nan")), torch.full_like(vd, float("nan"))
    der = torch.zeros((er_len, dh), device=DEV) if er_len else None
    ops.attention_bwd(qd, kd, vd, out, dOd, lse, erd, dq, dk, dv, der, B=B, Hq=H, Hkv=H, Lq=L, Lk=S, dh=dh, q_strides=qs,
                      k_strides=ks, v_strides=ks, o_strides=qs, do_strides=qs, dq_strides=qs, dkv_strides=ks, causal=causal,
                      tensor_core=True, dropout=(pdrop, seed))
    errs = {"dq": rel_err(dq.float().cpu(), qf.grad), "dk": rel_err(dk.float().cpu(), kf.grad), "dv": rel_err(dv.float().cpu(), vf.grad)}
    if er_len:
        errs["dEr"] = rel_err(der.cpu(), erf.grad)
    print("attention dropout rel errs", {n: "%.2e" % e for n, e in errs.items()})
    for n, e in errs.items():
        assert e < 2e-2, n


# ---------------------------------------------------------------- V2 / V3 attention (nn.MultiheadAttention + RoPE)
def test_custom_mha_rope_golden_gpu():
    from test_oracle import _custom_mha_case
    for c in load_golden("custom_mha.pt")["cases"]:
        s = c["spec"]
        m, sd, xq, xk = _custom_mha_case(c)
        m.load_state_dict(sd)
        m = m.to(DEV)
        mask = torch.triu(torch.full((s["L"], s["L"]), float("-inf"), device=DEV), diagonal=1) if s["causal"] else None
        xqd = xq.to(DEV)
        xkd = xqd if s["self_att"] else xk.to(DEV)
        with torch.no_grad():
            y, w = m(xqd, xkd, xkd, attn_mask=mask)
            y2, w2 = m(xqd, xkd, xkd, attn_mask=mask, need_weights=False)
        assert rel_err(y, c["y"]) < 1e-4 and rel_err(w.mean(dim=0), c["w_mean"]) < 1e-4, s
        assert w2 is None and torch.equal(y, y2)


# ---------------------------------------------------------------- VideoRegression (BASELINE config 5)
@pytest.mark.parametrize("reg", ["mamba", "mamba+", "bimamba+", "sharedmoe_bimamba+"])
def test_video_regression_golden_gpu(reg):
    """Loudness / note-density regressor and instrument classifier over the Mamba-family backbones (n_layers 6, d_model 128,
    d_hidden 256, total_vf_dim 774, L = 300) against the unmodified reference."""
    from test_oracle import _regression_case
    g, m, sd, sem, emo = _regression_case(reg)
    assert len(sd) == g["spec"]["n_keys"] and same_checksum(syn.checksum(sd), g["weights_checksum"])
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    z = torch.zeros(sem.shape[:2])
    with torch.no_grad():
        ln, inst = m(sem.to(DEV), z, z, emo.to(DEV))
    assert rel_err(ln, g["ln"]) < 1e-4 and rel_err(inst, g["inst"]) < 1e-4


@pytest.mark.gpu
def test_gemm_cta_pair_kernel_forced_for_every_shape():
    """The CTA-pair (cta_group::2) variant of the tcgen05 GEMM serves plain GEMMs with many rounds of tiles by default; the
    switch V2M_GEMM_PAIR=2 (read once per process, hence the subprocess) forces it for every eligible shape: all epilogues,
    K-major and MN-major B, M tails where the second CTA of the last pair owns no valid row -- against torch.matmul."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, V2M_GEMM_PAIR="2")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "pair_gemm_check.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
def test_gemm_large_m_default_dispatch_matches_matmul():
    """BASELINE config 3 shape (512 videos x 299 tokens): the default dispatch (CTA pairs at this size) against torch.matmul."""
    from video2music_b200 import ops
    g = torch.Generator().manual_seed(11)
    M, N, K = 512 * 299, 512, 512
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV).bfloat16()
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.linear(a, w, b, relu=True, out_dtype=torch.bfloat16)
    for lo in (0, M // 2 - 64, M - 4096):                  # slices: first tiles, a middle pair, the M tail
        ref = torch.relu(a[lo:lo + 4096].float() @ w.float().t() + b)
        assert rel_err(y[lo:lo + 4096].float(), ref) < 2e-2
