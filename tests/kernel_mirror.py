"""CPU mirrors of the CUDA ops used by the training paths of the variants.  TEST INFRASTRUCTURE ONLY.

The product has no CPU path (ops.* raise without a B200).  These torch restatements exist so that the `-m "not gpu"` suite can
exercise the HOST side of `video2music_b200.autograd` -- argument order, strides of the (L, B, E) / packed views, which gradient
goes to which parameter, the chunk / carry algebra -- against the reference's golden gradients without a GPU:
`install(monkeypatch)` swaps the entries of `video2music_b200.ops` for the duration of one test.  `selective_scan_bwd` and
`mamba_conv_silu_bwd` follow the recurrences of the CUDA kernels (csrc/mamba.cu) step by step, so they double as an executable
statement of what those kernels compute; the others are the plain definition of each op (the backward ones via torch autograd).
"""
import torch
import torch.nn.functional as F

MIRRORS = {}


def install(monkeypatch):
    """Replace video2music_b200.ops.<name> by its CPU mirror until the test ends."""
    from video2music_b200 import ops
    for name, fn in MIRRORS.items():
        monkeypatch.setattr(ops, name, fn)


MIRRORS['require_device'] = lambda t: None

def linear(x, w, bias=None, *, k=None, relu=False, alpha=1.0, alpha_cols=0, residual=None, res_mod=0, row_scale=None, col_vec=None,
           out_dtype=None, out=None, head_scatter=None, dropout=None):
    K = k if k is not None else min(x.shape[1], w.shape[1])
    y = x[:, :K].detach() @ w[:, :K].detach().t()
    if bias is not None: y = y + bias.detach()
    if alpha_cols: y[:, :alpha_cols] *= alpha
    if relu: y = y.relu()
    if residual is not None: y = y + residual.detach()
    return y
MIRRORS['linear'] = linear
def gemm_strided(a, a_rs, a_cs, w, w_rs, w_cs, M, N, K, out=None):
    A = torch.as_strided(a.detach(), (M, K), (a_rs, a_cs)); W = torch.as_strided(w.detach(), (N, K), (w_rs, w_cs))
    r = A @ W.t()
    if out is not None:
        out.copy_(r); return out
    return r
MIRRORS['gemm_strided'] = gemm_strided
def dy_prep(dy, y, relu, alpha, alpha_cols, out_dtype, want_dz=True, dropout=None, db_out=None):
    dz = dy.clone()
    if relu: dz = dz * (y > 0)
    if alpha_cols: dz[:, :alpha_cols] *= alpha
    db = dz.sum(0)
    if db_out is not None:
        db_out += db; db = db_out
    return (dz if want_dz else None), db
MIRRORS['dy_prep'] = dy_prep
def layernorm(x, gamma, beta, *, res=None, eps=1e-5, **kw):
    x = x.detach()
    if res is not None: x = x + res.detach()
    return F.layer_norm(x, (x.shape[-1],), gamma.detach(), beta.detach(), eps)
MIRRORS['layernorm'] = layernorm
def layernorm_bwd(x, gamma, dy, eps=1e-5, dg_out=None, db_out=None):
    x = x.detach().clone().requires_grad_(True); g = gamma.detach().clone().requires_grad_(True); b = torch.zeros_like(g).requires_grad_(True)
    with torch.enable_grad():
        F.layer_norm(x, (x.shape[-1],), g, b, eps).backward(dy)
    if dg_out is not None:
        dg_out += g.grad; db_out += b.grad
        return x.grad, dg_out, db_out
    return x.grad, g.grad, b.grad
MIRRORS['layernorm_bwd'] = layernorm_bwd
def rms(x, w, eps): 
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return y * w if w is not None else y
MIRRORS['rmsnorm'] = lambda x, w, eps=1e-5: rms(x.detach(), None if w is None else w.detach(), eps)
def rmsnorm_bwd(x, w, dy, eps=1e-5):
    x = x.detach().clone().requires_grad_(True); w2 = None if w is None else w.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        rms(x, w2, eps).backward(dy)
    return x.grad, None if w2 is None else w2.grad
MIRRORS['rmsnorm_bwd'] = rmsnorm_bwd
MIRRORS['axpy'] = lambda a, b, alpha: a.detach() + alpha * b.detach()
MIRRORS['swiglu'] = lambda a, g: a.detach() * F.silu(g.detach())
def swiglu_bwd(a, g, dh):
    a = a.detach().clone().requires_grad_(True); g = g.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (a * F.silu(g)).backward(dh)
    return torch.cat([a.grad, g.grad], 1)
MIRRORS['swiglu_bwd'] = swiglu_bwd
def moe_route(x, wg, bg, k, *, sel_bias=None, inv_t_pre=1.0, inv_t_post=1.0, want_logits=False):
    logits = (x @ wg.t() + bg) * inv_t_pre
    sel = logits + (sel_bias if sel_bias is not None else 0)
    idx = torch.topk(sel, k).indices
    w = torch.softmax(torch.gather(logits, -1, idx) * inv_t_post, -1)
    hist = torch.bincount(idx.flatten(), minlength=wg.shape[0]).int()
    return idx, w, hist, None
MIRRORS['moe_route'] = moe_route

def _glu(x, w1, b1, wg, bg, w2, b2):
    return (F.linear(x, w1, b1) * F.silu(F.linear(x, wg, bg))) @ w2.t() + b2
def moe_experts_fwd_saved(x, idx, w, hist, w1, b1, wg, bg, w2, b2, drops=None):
    T, k = idx.shape; E = w1.shape[0]
    order = torch.argsort(idx.flatten(), stable=True)          # row -> item
    perm = torch.empty_like(order); perm[order] = torch.arange(T * k)
    off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(hist.long(), 0)])
    xp = x[order // k]
    e_of_row = idx.flatten()[order]
    a = torch.stack([F.linear(xp[i], w1[e_of_row[i]], b1[e_of_row[i]]) for i in range(T * k)])
    g = torch.stack([F.linear(xp[i], wg[e_of_row[i]], bg[e_of_row[i]]) for i in range(T * k)])
    h = a * F.silu(g)
    yp = torch.stack([F.linear(h[i], w2[e_of_row[i]], b2[e_of_row[i]]) for i in range(T * k)])
    out = (w.unsqueeze(-1) * yp[perm].view(T, k, -1)).sum(1)
    return out, (xp, a, g, h, yp, perm, off)
MIRRORS['moe_experts_fwd_saved'] = moe_experts_fwd_saved
def moe_experts_bwd(dout, saved, idx, w, scale, w1g_t, w2_t, E, drops=None):
    xp, a, g, h, yp, perm, off = saved
    T, k = idx.shape; M = xp.shape[0]; ff = a.shape[1]
    dyp = torch.empty_like(yp); dw = torch.empty_like(w)
    for t in range(T):
        for r in range(k):
            row = perm[t * k + r]
            dyp[row] = w[t, r] * dout[t]; dw[t, r] = dout[t] @ yp[row]
    s = (w * dw).sum(-1, keepdim=True)
    dl = torch.zeros(T, E); dl.scatter_(1, idx, scale * w * (dw - s))
    e_of_row = torch.empty(M, dtype=torch.long)
    for e in range(E): e_of_row[off[e]:off[e + 1]] = e
    dW2 = torch.stack([dyp[off[e]:off[e + 1]].t() @ h[off[e]:off[e + 1]] for e in range(E)])
    db2 = torch.stack([dyp[off[e]:off[e + 1]].sum(0) for e in range(E)])
    dh = torch.stack([dyp[i] @ w2_t[e_of_row[i]].t() for i in range(M)])
    dag = swiglu_bwd(a, g, dh)
    dW1g = torch.stack([dag[off[e]:off[e + 1]].t() @ xp[off[e]:off[e + 1]] for e in range(E)])
    db1g = torch.stack([dag[off[e]:off[e + 1]].sum(0) for e in range(E)])
    dxp = torch.stack([dag[i] @ w1g_t[e_of_row[i]].t() for i in range(M)])
    dx = dxp[perm].view(T, k, -1).sum(1)
    return dx, dl, dW1g, db1g, dW2, db2
MIRRORS['moe_experts_bwd'] = moe_experts_bwd

def _srel(Q, Er):
    if Er is None: return 0
    L = Q.shape[0]; er_len = Er.shape[0]
    i = torch.arange(L)[:, None]; j = torch.arange(L)[None, :]
    idx = (er_len - 1 - (i - j)).clamp(0, er_len - 1)
    qe = Q @ Er.t()                                  # (L, er_len)
    return torch.where(j <= i, torch.gather(qe, 1, idx), torch.zeros(()))
def attention(q, k, v, out, *, B, Hq, Hkv, Lq, Lk, dh, q_strides, k_strides, v_strides, o_strides, causal, Er=None, q_scale=1.0, lse=None, p_out=None, dropout=None):
    G = Hq // Hkv
    for b in range(B):
        for h in range(Hq):
            Q = torch.as_strided(q.detach(), (Lq, dh), (q_strides[1], 1), q.storage_offset() + b * q_strides[0] + h * dh) * q_scale
            K = torch.as_strided(k.detach(), (Lk, dh), (k_strides[1], 1), k.storage_offset() + b * k_strides[0] + (h // G) * dh)
            V = torch.as_strided(v.detach(), (Lk, dh), (v_strides[1], 1), v.storage_offset() + b * v_strides[0] + (h // G) * dh)
            S = Q @ K.t() + _srel(Q, Er)
            if causal: S = S + torch.triu(torch.full((Lq, Lk), float("-inf")), 1)
            if lse is not None: lse[b * Hq + h] = torch.logsumexp(S, -1)
            if p_out is not None: p_out[b * Hq + h] = torch.softmax(S, -1)
            O = torch.softmax(S, -1) @ V
            torch.as_strided(out, (Lq, dh), (o_strides[1], 1), b * o_strides[0] + h * dh).copy_(O)
    return out
MIRRORS['attention'] = attention
def attention_bwd(q, k, v, o, dO, lse, Er, dq, dk, dv, dEr, *, B, Hq, Hkv, Lq, Lk, dh, q_strides, k_strides, v_strides, o_strides, do_strides, dq_strides, dkv_strides, causal, q_scale=1.0, tensor_core=False, dropout=None, dq_scale=1.0):
    G = Hq // Hkv
    for b in range(B):
        for h in range(Hq):
            Q = torch.as_strided(q.detach(), (Lq, dh), (q_strides[1], 1), b * q_strides[0] + h * dh).clone().requires_grad_(True)
            K = torch.as_strided(k.detach(), (Lk, dh), (k_strides[1], 1), b * k_strides[0] + (h // G) * dh).clone().requires_grad_(True)
            V = torch.as_strided(v.detach(), (Lk, dh), (v_strides[1], 1), b * v_strides[0] + (h // G) * dh).clone().requires_grad_(True)
            dOo = torch.as_strided(dO, (Lq, dh), (do_strides[1], 1), b * do_strides[0] + h * dh)
            with torch.enable_grad():
                Erl = None if Er is None else Er.detach().clone().requires_grad_(True)
                S = (Q * q_scale) @ K.t() + _srel(Q * q_scale, Erl)
                if causal: S = S + torch.triu(torch.full((Lq, Lk), float("-inf")), 1)
                (torch.softmax(S, -1) @ V).backward(dOo)
                if Er is not None: dEr.add_(Erl.grad)
            torch.as_strided(dq, (Lq, dh), (dq_strides[1], 1), b * dq_strides[0] + h * dh).copy_(Q.grad * dq_scale)
            torch.as_strided(dk, (Lk, dh), (dkv_strides[1], 1), b * dkv_strides[0] + (h // G) * dh).add_(K.grad)
            torch.as_strided(dv, (Lk, dh), (dkv_strides[1], 1), b * dkv_strides[0] + (h // G) * dh).add_(V.grad)
MIRRORS['attention_bwd'] = attention_bwd
MIRRORS['moe_experts'] = lambda *a: moe_experts_fwd_saved(*a)[0]

def mamba_conv_silu(xz, ED, w, bias, B, L):
    x = xz[:, :ED].detach().reshape(B, L, ED).transpose(1, 2)
    y = F.conv1d(x, w.detach().unsqueeze(1), None if bias is None else bias.detach(), padding=w.shape[1] - 1, groups=ED)[:, :, :L]
    return F.silu(y).transpose(1, 2).reshape(B * L, ED).contiguous()
MIRRORS['mamba_conv_silu'] = mamba_conv_silu

def _sp(x): return torch.where(x > 20, x, torch.log1p(torch.exp(x)))
def selective_scan(x, draw, dt_bias, A_log, Bm, Cm, D, z, B, L, plus=False):
    ED, N = A_log.shape
    x, draw, Bm, Cm, z = (t.detach().reshape(B, L, -1) for t in (x, draw, Bm, Cm, z))
    A = -torch.exp(A_log.detach()); dl = _sp(draw + dt_bias.detach())
    h = torch.zeros(B, ED, N); out = []
    for l in range(L):
        h = torch.exp(dl[:, l, :, None] * A) * h + (dl[:, l] * x[:, l])[:, :, None] * Bm[:, l, None, :]
        y = (h * Cm[:, l, None, :]).sum(-1) + D.detach() * x[:, l]
        s = F.silu(z[:, l]); o = y * s
        if plus: o = o + x[:, l] * (1 - torch.sigmoid(s))
        out.append(o)
    return torch.stack(out, 1).reshape(B * L, ED)
MIRRORS['selective_scan'] = selective_scan

SCAN_CHUNK = 64                                   # kScanChunk of csrc/mamba.cu


def selective_scan_bwd(x, draw, dt_bias, A_log, Bm, Cm, D, z, dout, dBm, dCm, dz, B, L, plus=False):
    """Mirror of selective_scan_bwd (csrc/mamba.cu), vectorised over (video, channel): forward pass 0 + carry -> state at every
    chunk start; backward pass 0 (chunk-local reverse sweep from gh = 0) + reverse carry -> dL/dh arriving at every chunk end;
    backward pass 1 per chunk: forward sweep storing the states, reverse sweep with dL/dh carried along."""
    ED, N = A_log.shape
    X, DR, BM, CM, Z, DO = (t.detach().reshape(B, L, -1) for t in (x, draw, Bm, Cm, z, dout))
    A = -torch.exp(A_log.detach()); db = dt_bias.detach(); Dc = D.detach()
    DL = _sp(DR + db)                                                            # (B, L, ED)
    n_chunks = (L + SCAN_CHUNK - 1) // SCAN_CHUNK
    bounds = [(ch * SCAN_CHUNK, min(L, ch * SCAN_CHUNK + SCAN_CHUNK)) for ch in range(n_chunks)]
    step = lambda h, l: torch.exp(DL[:, l, :, None] * A) * h + (DL[:, l] * X[:, l])[..., None] * BM[:, l, None, :]
    # forward pass 0 + carry: chunk-local end states from h = 0, then h_end[ch] = exp(A sum_delta[ch]) h_end[ch-1] + local[ch]
    h_start = [torch.zeros(B, ED, N)]
    for ch, (l0, l1) in enumerate(bounds[:-1]):
        h = torch.zeros(B, ED, N)
        for l in range(l0, l1):
            h = step(h, l)
        h_start.append(torch.exp(A * DL[:, l0:l1].sum(1)[..., None]) * h_start[ch] + h)
    DY = DO * F.silu(Z)
    # backward pass 0 + reverse carry: G[ch] = dL/dh arriving at the end of chunk ch from all later chunks
    G = [None] * n_chunks
    G[n_chunks - 1] = torch.zeros(B, ED, N)
    for ch in range(n_chunks - 1, 0, -1):
        l0, l1 = bounds[ch]
        g = torch.zeros(B, ED, N)
        for l in range(l1 - 1, l0 - 1, -1):
            g = (g + DY[:, l, :, None] * CM[:, l, None, :]) * torch.exp(DL[:, l, :, None] * A)
        G[ch - 1] = torch.exp(A * DL[:, l0:l1].sum(1)[..., None]) * G[ch] + g
    gA = torch.zeros(B, ED, N); gD = torch.zeros(B, ED); gdb = torch.zeros(B, ED)
    dx = torch.zeros(B, L, ED); ddraw = torch.zeros(B, L, ED); dzo = torch.zeros(B, L, ED)
    dB = torch.zeros(B, L, N); dC = torch.zeros(B, L, N)
    for ch, (l0, l1) in enumerate(bounds):                                       # backward pass 1
        Hs = {}
        h = h_start[ch]
        for l in range(l0, l1):
            h = step(h, l)
            Hs[l] = h
        gh = G[ch]
        for l in range(l1 - 1, l0 - 1, -1):
            xv = X[:, l]; raw = DR[:, l] + db; dl = DL[:, l]; g = DO[:, l]
            Bv = BM[:, l, None, :]; Cv = CM[:, l, None, :]
            y = Dc * xv + (h * Cv).sum(-1)
            zv = Z[:, l]; sg = torch.sigmoid(zv); sl = zv * sg; dsl = sg * (1 + zv * (1 - sg))
            dy = g * sl; dzv = g * y * dsl; dxv = torch.zeros_like(xv)
            if plus:
                s2 = torch.sigmoid(sl); dxv = g * (1 - s2); dzv = dzv - g * xv * s2 * (1 - s2) * dsl
            dxv = dxv + dy * Dc; gD += dy * xv
            hp = Hs[l - 1] if l > l0 else h_start[ch]
            a = torch.exp(dl[..., None] * A)
            dC[:, l] = (dy[..., None] * h).sum(1)
            gh = gh + dy[..., None] * Cv
            t = gh * hp
            ddl = (t * A * a).sum(-1) + (gh * Bv).sum(-1) * xv
            gA += t * dl[..., None] * a
            dB[:, l] = (gh * (dl * xv)[..., None]).sum(1)
            dxv = dxv + (gh * dl[..., None] * Bv).sum(-1)
            gh = gh * a; h = hp
            ddr = ddl * torch.where(raw > 20, torch.ones_like(raw), torch.sigmoid(raw)); gdb += ddr
            dx[:, l] = dxv; ddraw[:, l] = ddr; dzo[:, l] = dzv
    dBm += dB.reshape(B * L, N); dCm += dC.reshape(B * L, N); dz.copy_(dzo.reshape(B * L, ED))
    return dx.reshape(B * L, ED), ddraw.reshape(B * L, ED), (gA * A).sum(0), gD.sum(0), gdb.sum(0)
MIRRORS['selective_scan_bwd'] = selective_scan_bwd

def mamba_conv_silu_bwd(xz, ED, w, bias, dy, dxz, B, L):
    """Mirror of mamba_conv_silu_bwd_kernel: per chunk of SCAN_CHUNK steps, a KW-wide sliding window of inputs (primed with the
    KW-1 inputs before the chunk) and of pending input gradients, walked over the chunk plus a (KW-1)-step halo on the right."""
    KW = w.shape[1]; X = xz[:, :ED].detach().reshape(B, L, ED); DY = dy.detach().reshape(B, L, ED); w = w.detach()
    gw = [torch.zeros(B, ED) for _ in range(KW)]; gb = torch.zeros(B, ED)
    dx = torch.full((B, L, ED), float("nan")); bs = bias.detach() if bias is not None else 0
    for l0 in range(0, L, SCAN_CHUNK):
        l1 = min(L, l0 + SCAN_CHUNK); lh = min(L, l1 + KW - 1)
        win = [torch.zeros(B, ED)] + [X[:, l0 - KW + k] if l0 - KW + k >= 0 else torch.zeros(B, ED) for k in range(1, KW)]
        acc = [torch.zeros(B, ED) for _ in range(KW)]
        for l in range(l0, lh):
            win = win[1:] + [X[:, l]]; acc = acc[1:] + [torch.zeros(B, ED)]
            pre = bs + sum(w[:, k] * win[k] for k in range(KW))
            sg = torch.sigmoid(pre); dpre = DY[:, l] * sg * (1 + pre * (1 - sg))
            if l < l1:
                gb = gb + dpre
                gw = [gw[k] + dpre * win[k] for k in range(KW)]
            acc = [acc[k] + w[:, k] * dpre for k in range(KW)]
            p = l - (KW - 1)
            if l0 <= p < l1: dx[:, p] = acc[0]
        for k in range(1, KW):
            p = lh - 1 - (KW - 1) + k
            if l0 <= p < l1: dx[:, p] = acc[k]
    dxz[:, :ED] = dx.reshape(B * L, ED)
    return torch.stack([g.sum(0) for g in gw], 1), (gb.sum(0) if bias is not None else None)
MIRRORS['mamba_conv_silu_bwd'] = mamba_conv_silu_bwd
MIRRORS['sigmoid'] = lambda a: torch.sigmoid(a.detach())
MIRRORS['sigmoid_bwd'] = lambda dy, s: dy * s * (1 - s)
MIRRORS['dw_f32'] = lambda dz, x, K: dz.detach().t() @ x.detach()[:, :K]


# ---- inference ops of the KV-cached generation of the generic decoder stacks (video2music_b200/cached_decode.py): the plain definition of
# each op, so that the HOST side of the generation loop (position kept in a tensor, caches written with index_copy_, token bookkeeping
# with gather / scatter, primer handling, sampling constraints) runs on a CPU box
def concat_features(sem, scene, motion, emotion, out_dtype, ld_out):
    rows = sem.shape[0] * sem.shape[1]
    mo = motion.float().reshape(rows, -1)
    parts = [sem.float().reshape(rows, -1), scene.float().reshape(rows, 1), mo, emotion.float().reshape(rows, -1)]
    cat = torch.cat(parts, 1)
    out = torch.zeros((rows, ld_out), dtype=out_dtype)
    out[:, :cat.shape[1]] = cat
    return out
MIRRORS['concat_features'] = concat_features
def embed_sum(idx_a, table_a, idx_b, table_b, out_dtype):
    y = table_a.detach()[idx_a.reshape(-1)]
    if idx_b is not None: y = y + table_b.detach()[idx_b.reshape(-1)]
    return y.to(out_dtype)
MIRRORS['embed_sum'] = embed_sum
def step_linear(x, w, bias=None, *, k=None, relu=False, row_scale=None, col_vec=None):
    K = k if k is not None else min(x.shape[1], w.shape[1])
    y = x[:, :K] @ w[:, :K].t()
    if bias is not None: y = y + bias
    if row_scale is not None: y = y + row_scale[:, None] * col_vec[None, :]
    return y.relu() if relu else y
MIRRORS['step_linear'] = step_linear
def step_attention(q, K, V, *, Hq, Hkv, dh, n_max, kv_strides, n_dev=None, q_scale=1.0):
    B = q.shape[0]
    n = n_max if n_dev is None else min(n_max, int(n_dev))
    out = torch.empty((B, Hq * dh))
    for b in range(B):
        for h in range(Hq):
            off = K.storage_offset() + b * kv_strides[0] + (h // (Hq // Hkv)) * dh
            Kh = torch.as_strided(K, (n, dh), (kv_strides[1], 1), off)
            Vh = torch.as_strided(V, (n, dh), (kv_strides[1], 1), V.storage_offset() + b * kv_strides[0] + (h // (Hq // Hkv)) * dh)
            p = torch.softmax((q[b, h * dh:(h + 1) * dh] * q_scale) @ Kh.t(), -1)
            out[b, h * dh:(h + 1) * dh] = p @ Vh
    return out
MIRRORS['step_attention'] = step_attention


# ---- model zoo training (V1 / V2 / V3): RoPE with the literal reinterpretation, embedding backward
def rope_quirk(x, cache, B, H):
    """csrc/elementwise.cu rope_quirk_kernel: the (len*B, E) rows are read as [H][len][B][dh] and the cache [len][E/2][2] as
    [H][len][dh/2][2]; pair j of element (h', l', b') is rotated by cache entry (h'*len + l')*(dh/2) + j."""
    E = x.shape[1]
    length, dh2 = x.shape[0] // B, E // H // 2
    xp = x.detach().reshape(H, length, B, dh2, 2)
    cs = cache.reshape(-1, 2)[:H * length * dh2].reshape(H, length, 1, dh2, 2)
    y = torch.stack([xp[..., 0] * cs[..., 0] - xp[..., 1] * cs[..., 1], xp[..., 1] * cs[..., 0] + xp[..., 0] * cs[..., 1]], -1)
    return y.reshape(x.shape).contiguous()
MIRRORS['rope_quirk'] = rope_quirk
def embed_bwd(idx, d, n_rows_table, D, out=None):
    dt = torch.zeros((n_rows_table, D)) if out is None else out
    dt.index_add_(0, idx.reshape(-1), d[:, :D].float())
    return dt
MIRRORS['embed_bwd'] = embed_bwd
