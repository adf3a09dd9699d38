"""`SSTBEVBackbone` -- row N1 of SURVEY.md section 8f: the dense 2-D backbone right behind `spatial_features`.

Mirrors pcdet/models/backbones_2d/sst_bev_backbone.py:6-44 (constructor `(model_cfg, **kwargs)`, `num_bev_features`,
`forward(data_dict)` reading `spatial_features` and writing `spatial_features_2d`, parameters
`conv_layer.{i}.{0,1}.*`): a stack of Conv2d(3x3, bias=False) -> BatchNorm2d(eps 1e-3, momentum 0.01) -> ReLU with
additive shortcuts on the layers listed in CONV_SHORTCUT (t_mae.yaml:197-206).

As for the dense decoder (row A13), the convolutions stay on cuDNN -- dense contractions it already maps to the Blackwell
tensor cores -- and run channels-last; BatchNorm + ReLU run on the library's row kernels (`tmae_bn_*` for fp32 maps,
`tmae_bn_bf16_*` for the bf16 maps the throughput-mode decoder returns), which read the convolution output once for the
statistics and once to normalise, and recompute the ReLU mask in the backward pass instead of keeping it.
There is no CPU path: CPU tensors raise.
"""
import torch
import torch.nn as nn

from .backbone import _bn2d_relu_cat
from .vfe import bn_relu


class SSTBEVBackbone(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        cin = model_cfg["NUM_FILTER"]
        self.conv_shortcut = list(model_cfg["CONV_SHORTCUT"])
        layers = []
        for kw in model_cfg["CONV_KWARGS"]:
            kw = dict(kw)
            layers.append(nn.Sequential(nn.Conv2d(cin, **kw, bias=False),
                                        nn.BatchNorm2d(kw["out_channels"], eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
            cin = kw["out_channels"]
        self.conv_layer = nn.ModuleList(layers)
        self.num_bev_features = cin

    @staticmethod
    def _conv_bn_relu(block, x):
        conv, bn = block[0], block[1]
        if x.dtype == torch.bfloat16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                v = conv(x)
            return _bn2d_relu_cat([bn], [v])                       # (B, C, Y, X) view of a channels-last bf16 buffer
        v = conv(x.contiguous(memory_format=torch.channels_last))  # cuDNN, fp32 (TF32 only if the caller allows it)
        B, C, Y, X = v.shape
        rows = v.permute(0, 2, 3, 1).reshape(B * Y * X, C)         # a view of the channels-last map
        return bn_relu(rows, bn).view(B, Y, X, C).permute(0, 3, 1, 2)

    def forward(self, data_dict):
        out = data_dict["spatial_features"]
        if not out.is_cuda:
            raise RuntimeError("tmae_b200 modules need CUDA tensors (there is no CPU path)")
        for i, block in enumerate(self.conv_layer):
            t = self._conv_bn_relu(block, out)
            out = t + out if (t.shape == out.shape and i in self.conv_shortcut) else t
        data_dict["spatial_features_2d"] = out
        return data_dict
