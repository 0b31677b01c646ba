"""CPU, only where /root/reference is mounted: the oracle against the LIVE unmodified reference
(imported through oracle/ref_shim.py).  Skipped on the GPU box, where the goldens stand in."""
import contextlib
import io

import pytest
import torch

from oracle import amt_oracle as O
from oracle.ref_shim import REFERENCE_ROOT, load_reference, reference_available, reference_cwd
from video2music_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


def _ref_model(ref, chord_embed=False, seed=0):
    with contextlib.redirect_stdout(io.StringIO()), reference_cwd():
        m = ref.vmt.VideoMusicTransformer(total_vf_dim=776, rpr=True, chord_embed=chord_embed).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(syn.fill_like_reference_init(shapes, seed=seed), strict=False)
    return m, {k: v.clone() for k, v in m.state_dict().items()}


def test_state_dict_keys_match_reference():
    from video2music_b200 import VideoMusicTransformer
    ref = load_reference()
    m, sd = _ref_model(ref)
    ours = VideoMusicTransformer(total_vf_dim=776, rpr=True)
    osd = ours.state_dict()
    assert set(osd) == set(sd)
    assert all(tuple(osd[k].shape) == tuple(sd[k].shape) for k in sd)
    ours.load_state_dict(m.state_dict())          # a reference checkpoint loads unmodified (train.py:180)
    assert torch.equal(ours.positional_encoding.pe, m.positional_encoding.pe)


def test_forward_live():
    ref = load_reference()
    m, sd = _ref_model(ref, seed=9)
    inp = syn.make_inputs(2, 5, 33, 300, 0)
    args = (inp["x"], inp["x_root"], inp["x_attr"], inp["feature_semantic_list"], inp["feature_key"],
            inp["feature_scene_offset"], inp["feature_motion"], inp["feature_emotion"])
    with torch.no_grad():
        y_ref = m(*args)
        y = O.amt_forward(sd, *args)
    assert float((y - y_ref).abs().max()) < 5e-5


@pytest.mark.parametrize("chord_embed", [False, True])
def test_generate_live(chord_embed):
    ref = load_reference()
    m, sd = _ref_model(ref, chord_embed=chord_embed, seed=4)
    inp = syn.make_inputs(1, 6, 299, 300, 0)
    prim, pr, pa = torch.tensor([1, 20, 33]), torch.tensor([1, 2, 3]), torch.tensor([0, 7, 6])
    with torch.no_grad(), reference_cwd(), contextlib.redirect_stdout(io.StringIO()):
        g_ref = m.generate(inp["feature_semantic_list"], inp["feature_key"][0], inp["feature_scene_offset"],
                           inp["feature_motion"], inp["feature_emotion"], primer=prim, primer_root=pr, primer_attr=pa,
                           target_seq_length=20, beam=1)
    with torch.no_grad():
        g_lit = O.generate_greedy_literal(sd, inp["feature_semantic_list"], inp["feature_key"][0], inp["feature_scene_offset"],
                                          inp["feature_motion"], inp["feature_emotion"], prim, pr, pa, 20, chord_embed=chord_embed)
        g_c = O.generate_greedy_cached(sd, inp["feature_semantic_list"], inp["feature_key"], inp["feature_scene_offset"],
                                       inp["feature_motion"], inp["feature_emotion"], prim, pr, pa, 20, chord_embed=chord_embed)
    assert torch.equal(g_ref, g_lit) and torch.equal(g_ref, g_c)


def test_chord_id_to_root_attr_closed_form():
    """The sampling branch maps the drawn chord id to (root, attr) through three JSON dictionaries
    (video_music_transformer.py:1052-1057,1105-1123); our kernels and the oracle use the closed form."""
    import json
    import os
    base = os.path.join(REFERENCE_ROOT, "dataset", "vevo_meta")
    inv = json.load(open(os.path.join(base, "chord_inv.json")))
    root = json.load(open(os.path.join(base, "chord_root.json")))
    attr = json.load(open(os.path.join(base, "chord_attr.json")))
    for c in range(157):
        parts = inv[str(c)].split(":")
        want = (root[parts[0]], 1) if len(parts) == 1 else (root[parts[0]], attr[parts[1]])
        got = (0, 1) if c == 0 else ((c - 1) // 13 + 1, (c - 1) % 13 + 1)
        assert got == want, (c, inv[str(c)])
