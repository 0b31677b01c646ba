"""Config-4 (GQA + MoE) cached generation, a few positions, eager launches: for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from video2music_b200 import synthetic as syn
from video2music_b200.video_music_transformer_v2 import VideoMusicTransformer_GQA
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gm = VideoMusicTransformer_GQA(n_layers=6, total_vf_dim=syn.vf_dim(0)).eval().to(dev)
gi = syn.make_inputs(64, 1234, 299, 300, 0)
feats = [gi[k].to(dev) for k in ("feature_semantic_list", "feature_key", "feature_scene_offset", "feature_motion", "feature_emotion")]
one = torch.tensor([1])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
out = gm.generate_cached(*feats, primer=one, primer_root=one, primer_attr=torch.tensor([0]), target_seq_length=n, beam=1, beam_chance=1.0, use_graph=False)
torch.cuda.synchronize()
print(out[0].tolist())
