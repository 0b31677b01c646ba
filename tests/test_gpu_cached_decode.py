"""KV-cached, batched generation of the MoE / GQA decoder stacks (BASELINE config 4 "generation",
video2music_b200/cached_decode.py): greedy tokens against the unmodified reference's generate() (V1 models, tests/golden/v2.pt),
against our literal one-forward-per-token loop, and -- for the GQA + MoE shell, which the reference does not ship as a class --
against the oracle's literal loop over the reference's blocks."""
import pytest
import torch

from conftest import load_golden, rel_err, same_checksum
from oracle import amt_oracle as O
from video2music_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")


def _load(m, seed):
    sd = syn.fill_like_reference_init({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed)
    m.load_state_dict(sd)
    return sd


@pytest.mark.parametrize("ver", ["1.1", "1.3rms"])
def test_cached_generate_v1_equals_reference_generate(ver):
    """V1 '1.1' (MoELayer) / '1.3' with RMSNorm (SharedMoELayer): stock attention honours the causal mask, so the literal
    reference is cacheable.  The cached path reproduces the reference's own generate(beam=1) tokens (golden, video 0) and, for
    all videos of a batch, our literal loop run video by video."""
    from video2music_b200 import VideoMusicTransformer_V1
    from video2music_b200.cached_decode import cacheable
    g = load_golden("v2.pt")[ver]
    s = g["spec"]
    m = VideoMusicTransformer_V1(version_name=ver[:3], total_vf_dim=syn.vf_dim(0), dropout=0.1, rms_norm=ver.endswith("rms")).eval()
    sd = _load(m, s["seed"])
    assert same_checksum(syn.checksum(sd), g["weights_checksum"])
    m = m.to(DEV)
    assert cacheable(m)
    inp = syn.make_inputs(2, s["seed"], 24, 40, 0)
    feats = [inp[k].to(DEV) for k in KEYS]
    prim, pr, pa = inp["x"][0, :3], inp["x_root"][0, :3], inp["x_attr"][0, :3]
    out = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=14, beam=1, beam_chance=1.0)
    assert out.shape == (2, 14)
    assert torch.equal(out[:1].cpu(), g["generated"])                  # the unmodified reference's tokens (it ran video 0)
    for b in range(2):
        one = [t[b:b + 1] for t in feats]
        lit = m.generate(one[0], one[1][0], one[2], one[3], one[4], primer=prim, primer_root=pr, primer_attr=pa,
                         target_seq_length=14, beam=1, beam_chance=1.0)
        assert torch.equal(out[b:b + 1], lit), b


@pytest.mark.parametrize("shared,rms,pre_norm", [(False, False, False), (True, True, True)])
def test_cached_generate_gqa_moe_equals_literal_loop_and_oracle(shared, rms, pre_norm):
    """The GQA (8 query / 2 kv heads) + MoE (6 experts, top-2) shell: full forward against the oracle's composition of the
    reference's blocks, cached batched greedy generation == our literal loop == the oracle's literal loop (fp32, bit-exact
    tokens), KV cache 4x smaller than with 8 kv heads."""
    from video2music_b200.video_music_transformer_v2 import VideoMusicTransformer_GQA
    from video2music_b200.cached_decode import CachedDecoder, cacheable
    torch.manual_seed(0)
    m = VideoMusicTransformer_GQA(n_layers=3, total_vf_dim=syn.vf_dim(0), shared_moe=shared, rms_norm=rms, pre_norm=pre_norm, dropout=0.1).eval()
    sd = _load(m, 61 + shared)
    sd["Wout.weight"] = sd["Wout.weight"] * 4.0                         # a decisive arg-max (as the AMT generation goldens)
    m.load_state_dict(sd)
    m = m.to(DEV)
    assert cacheable(m)
    B, T, S = 3, 20, 40
    inp = syn.make_inputs(B, 77, T, S, 0)
    args = [inp[k] for k in ("x", "x_root", "x_attr") + KEYS]
    with torch.no_grad():
        y = m(*[a.to(DEV) for a in args])
        for b in range(B):                                              # B = 1 slices: MultiheadGQA's `.view` mixes batch and time for B > 1
            yb = m(*[a[b:b + 1].to(DEV) for a in args])
            ref = O.gqa_moe_forward(sd, *[a[b:b + 1] for a in args[1:]], n_layers=3, shared=shared, rms=rms, pre_norm=pre_norm)
            assert rel_err(yb, ref) < 2e-4, b
    feats = [inp[k].to(DEV) for k in KEYS]
    prim, pr, pa = inp["x"][0, :2], inp["x_root"][0, :2], inp["x_attr"][0, :2]
    out = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=16, beam=1, beam_chance=1.0)
    eager = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=16, beam=1, beam_chance=1.0,
                              use_graph=False)
    assert torch.equal(out, eager)                                      # one captured CUDA graph per position == eager launches
    again = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=16, beam=1, beam_chance=1.0)
    assert torch.equal(out, again)                                      # second call: same session, graph replayed without capture
    other = [t.flip(0) for t in feats]                                  # other videos through the same captured graph
    out2 = m.generate_cached(*other, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=16, beam=1, beam_chance=1.0)
    assert torch.equal(out2, out.flip(0))
    for b in range(B):
        one = [t[b:b + 1] for t in feats]
        lit = m.generate(one[0], one[1][0], one[2], one[3], one[4], primer=prim, primer_root=pr, primer_attr=pa,
                         target_seq_length=16, beam=1, beam_chance=1.0)
        assert torch.equal(out[b:b + 1], lit), b
        cpu = [inp[k][b:b + 1] for k in KEYS]
        fwd = lambda xr, xa: O.gqa_moe_forward(sd, xr, xa, *cpu, n_layers=3, shared=shared, rms=rms, pre_norm=pre_norm)
        with torch.no_grad():
            orc = O.zoo_generate_greedy_literal(fwd, *cpu, prim, pr, pa, 16)
        assert torch.equal(out[b:b + 1].cpu(), orc), b
    # the cache of grouped-query attention holds kv_heads (2) of the 8 heads
    dec = CachedDecoder(m, torch.zeros(S, B, 512, device=DEV), 300)
    assert dec.K[0].shape == (B, 300, 2 * 64)


def test_cached_generate_sampling_branch_and_errors():
    """beam=0: constraints hold for every video (never "N", never three equal chords), the root / attribute inputs follow the
    drawn chords, and identical uniforms give identical sequences; a RoPE model is refused."""
    from video2music_b200 import VideoMusicTransformer_V1, VideoMusicTransformer_V2
    m = VideoMusicTransformer_V1(version_name="1.1", n_layers=2, total_vf_dim=syn.vf_dim(0)).eval()
    _load(m, 5)
    m = m.to(DEV)
    inp = syn.make_inputs(4, 9, 8, 30, 0)
    feats = [inp[k].to(DEV) for k in KEYS]
    prim, pr, pa = torch.tensor([1]), torch.tensor([1]), torch.tensor([0])
    u = torch.rand((4, 40), generator=torch.Generator().manual_seed(3)).to(DEV)
    a = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=40, beam=0, uniforms=u)
    b = m.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=40, beam=0, uniforms=u, use_graph=False)
    assert torch.equal(a, b) and a.shape == (4, 40)
    assert int((a[:, 1:] == 0).sum()) == 0 and int(a.max()) < 157
    same3 = (a[:, 2:] == a[:, 1:-1]) & (a[:, 1:-1] == a[:, :-2])
    assert not bool(same3.any())
    m2 = VideoMusicTransformer_V2(version_name="2.2", n_layers=4, total_vf_dim=syn.vf_dim(0)).eval().to(DEV)
    with pytest.raises(NotImplementedError):
        m2.generate_cached(*feats, primer=prim, primer_root=pr, primer_attr=pa, target_seq_length=8, beam=1, beam_chance=1.0)
