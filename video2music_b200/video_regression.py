"""Drop-in for `VideoRegression` (model/video_regression.py:103-245) with the Mamba-family backbones of BASELINE config 5:
regModel in {"mamba", "mamba+", "bimamba", "bimamba+", "moe_bimamba+", "sharedmoe_bimamba+"} -- the loudness / note-density
regressor and instrument classifier that `video2music.py:613,651` runs next to the chord transformer.  Same constructor
signature and parameter names (`in_proj.0`, `model.*`, `regressor`, `classifier.0`).  The LSTM / GRU / CNN-GRU / minGRU /
moemamba backbones are outside the hot path (not built).  Trains in fp32 with dropout 0 (gradients through the autograd
Functions of `autograd.py`: every forward and backward kernel is ours)."""
import torch
import torch.nn as nn

from . import ops
from .mamba import BiMambaEncoder, Mamba, MambaConfig
from .moe import GLUExpert, MoELayer, SharedMoELayer

INSTRUMENT_SIZE = 40                                      # utilities/constants.py:84


class VideoRegression(nn.Module):
    def __init__(self, n_layers=2, d_model=64, d_hidden=1024, dropout=0.1, use_KAN=False, max_sequence_video=300, total_vf_dim=0,
                 regModel="bilstm", scene_embed=False, chord_embed=False):
        super().__init__()
        if use_KAN:
            raise NotImplementedError("KAN projections need efficient_kan (out of scope)")
        self.n_layers, self.d_model, self.d_hidden = n_layers, d_model, d_hidden
        self.dropout_layer = nn.Dropout(dropout)
        self.max_seq_video, self.total_vf_dim, self.regModel = max_sequence_video, total_vf_dim, regModel
        self.scene_embed, self.chord_embed = scene_embed, chord_embed
        if regModel == "mamba":                                                  # video_regression.py:142-145
            self.model = Mamba(MambaConfig(d_model=d_model, n_layers=n_layers, use_KAN=use_KAN, bias=True))
        elif regModel == "mamba+":
            self.model = Mamba(MambaConfig(d_model=d_model, n_layers=n_layers, use_KAN=use_KAN, bias=True, use_version=1))
        elif regModel in ("bimamba", "bimamba+", "moe_bimamba+", "sharedmoe_bimamba+"):      # :158-186
            cfg = MambaConfig(d_model=d_model, n_layers=1, dropout=dropout, use_KAN=use_KAN, bias=True,
                              use_version=0 if regModel == "bimamba" else 1)
            moe = None
            if regModel == "moe_bimamba+":
                moe = MoELayer(GLUExpert(d_model, d_model * 2 + 1), d_model, n_experts=6, n_experts_per_token=2, dropout=dropout)
            elif regModel == "sharedmoe_bimamba+":
                moe = SharedMoELayer(GLUExpert(d_model, d_model * 2 + 1), d_model, n_experts=6, n_experts_per_token=2, dropout=dropout)
            self.model = BiMambaEncoder(cfg, d_hidden, n_encoder_layers=n_layers, dropout=dropout, moe_layer=moe)
        else:
            raise NotImplementedError("regModel %r: only the Mamba-family backbones are built" % (regModel,))
        self.in_proj = nn.Sequential(nn.Linear(total_vf_dim, d_model), nn.Dropout(dropout))
        self.regressor = nn.Linear(d_model, 2)
        self.classifier = nn.Sequential(nn.Linear(d_model, INSTRUMENT_SIZE), nn.Sigmoid())

    def get_feature(self, feature_semantic_list, feature_scene_offset, feature_motion, feature_emotion):
        dev = self.regressor.weight.device
        vf = torch.cat([feature_semantic_list.float().to(dev), feature_emotion.float().to(dev)], dim=-1)   # :211-213
        B, L, F = vf.shape
        from . import autograd as ag
        if ag.tracking(vf, self) or ag.has_dropout(self):
            x = ag.drop_rows(ag.linear_fn(vf.reshape(B * L, F).contiguous(), self.in_proj[0]), self.in_proj[1], self.training)   # :203
            return self.model(x.view(B, L, self.d_model))
        x = ops.linear(vf.reshape(B * L, F).contiguous(), self.in_proj[0].weight.detach(), self.in_proj[0].bias.detach())
        return self.model(x.view(B, L, self.d_model))

    def forward(self, feature_semantic_list, feature_scene_offset, feature_motion, feature_emotion):
        out = self.get_feature(feature_semantic_list, feature_scene_offset, feature_motion, feature_emotion)
        B, L, E = out.shape
        from . import autograd as ag
        if ag.tracking(out, self):
            o2 = ag.rows_f32(out)
            return (ag.linear_fn(o2, self.regressor).view(B, L, 2),
                    ag.SigmoidFn.apply(ag.linear_fn(o2, self.classifier[0])).view(B, L, INSTRUMENT_SIZE))
        o2 = out.reshape(B * L, E).float().contiguous()
        ln = ops.linear(o2, self.regressor.weight.detach(), self.regressor.bias.detach()).view(B, L, 2)
        inst = ops.sigmoid(ops.linear(o2, self.classifier[0].weight.detach(), self.classifier[0].bias.detach())).view(B, L, INSTRUMENT_SIZE)
        return ln, inst
