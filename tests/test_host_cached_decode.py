"""Host side of the KV-cached generation of the generic decoder stacks (video2music_b200/cached_decode.py, BASELINE config 4) without
a GPU: `video2music_b200.ops` is replaced by the CPU mirrors of tests/kernel_mirror.py, so what runs here is the product's own
generation loop -- position kept in a tensor, caches written with index_copy_, keys counted by a device word, tokens / roots /
attributes moved with gather and scatter, primer positions, the sampling constraints, the per-model session cache -- against the
oracle's literal one-forward-per-token loop over the reference's blocks.  The kernels themselves are proven by the `-m gpu` tests
(tests/test_gpu_cached_decode.py, tests/test_gpu_kernels.py::test_step_*)."""
import pytest
import torch

import kernel_mirror
from oracle import amt_oracle as O
from video2music_b200 import synthetic as syn

KEYS = ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")


@pytest.fixture
def mirrored(monkeypatch):
    kernel_mirror.install(monkeypatch)


def _shell(shared, rms, pre_norm, seed):
    from video2music_b200.video_music_transformer_v2 import VideoMusicTransformer_GQA
    torch.manual_seed(0)
    m = VideoMusicTransformer_GQA(n_layers=2, total_vf_dim=syn.vf_dim(0), shared_moe=shared, rms_norm=rms, pre_norm=pre_norm, dropout=0.1).eval()
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed)
    sd["Wout.weight"] = sd["Wout.weight"] * 4.0                         # a decisive arg-max
    m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("shared,rms,pre_norm", [(False, False, False), (True, True, True)])
def test_cached_generation_host_logic_equals_oracle_literal_loop(mirrored, shared, rms, pre_norm):
    from video2music_b200.cached_decode import cacheable
    m, sd = _shell(shared, rms, pre_norm, 61 + shared)
    assert cacheable(m)
    B, T = 3, 12
    inp = syn.make_inputs(B, 77, 20, 24, 0)
    feats = [inp[k] for k in KEYS]
    prim, pr, pa = inp["x"][0, :3], inp["x_root"][0, :3], inp["x_attr"][0, :3]          # a 3-token primer: positions 0..2 only fill the caches
    out = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=T, beam=1, beam_chance=1.0, use_graph=False)
    assert out.shape == (B, T) and torch.equal(out[:, :3], prim.view(1, 3).expand(B, 3))
    again = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=T, beam=1, beam_chance=1.0, use_graph=False)
    assert torch.equal(out, again)                                       # the session (buffers, caches) is reused and fully re-initialised
    for b in range(B):
        cpu = [inp[k][b:b + 1] for k in KEYS]
        fwd = lambda xr, xa: O.gqa_moe_forward(sd, xr, xa, *cpu, n_layers=2, shared=shared, rms=rms, pre_norm=pre_norm)
        with torch.no_grad():
            orc = O.zoo_generate_greedy_literal(fwd, *cpu, prim, pr, pa, T)
        assert torch.equal(out[b:b + 1], orc), b
    flipped = m.generate_cached(*[t.flip(0) for t in feats], primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=T, beam=1,
                                beam_chance=1.0, use_graph=False)
    assert torch.equal(flipped, out.flip(0))                             # other videos through the same session


def test_cached_generation_sampling_branch_host_logic(mirrored):
    """beam=0: the no-"N" / no-three-equal-chords constraints, the root / attribute inputs of drawn chords, identical uniforms ->
    identical sequences, and a changed parameter invalidates the session."""
    m, _ = _shell(False, False, False, 5)
    inp = syn.make_inputs(2, 9, 8, 16, 0)
    feats = [inp[k] for k in KEYS]
    prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
    u = torch.rand((2, 30), generator=torch.Generator().manual_seed(3))
    a = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=30, beam=0, uniforms=u, use_graph=False)
    b = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=30, beam=0, uniforms=u, use_graph=False)
    assert torch.equal(a, b) and a.shape == (2, 30)
    assert int((a[:, 1:] == 0).sum()) == 0 and int(a.max()) < 157
    same3 = (a[:, 2:] == a[:, 1:-1]) & (a[:, 1:-1] == a[:, :-2])
    assert not bool(same3.any())
    from video2music_b200 import cached_decode
    n_before = len(cached_decode._SESSIONS[m])
    with torch.no_grad():
        m.Wout.bias.add_(1.0)                                            # in-place update bumps the version: snapshots are stale
    c = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=30, beam=0, uniforms=u, use_graph=False)
    assert len(cached_decode._SESSIONS[m]) == n_before == 1 and c.shape == (2, 30)
