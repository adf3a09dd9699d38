"""Plain-dict model configs equal to cfg.MODEL.{VFE,BACKBONE_3D} of the reference YAMLs
(tools/cfgs/once_models/t_mae_ssl.yaml:44-176 -> 'pretrain', t_mae.yaml:58-195 -> 'finetune').
pcdet's EasyDict configs work too: the modules only use item access and .get()."""


def _levels(spec):
    d = {str(i): {"max_tokens": t, "drop_range": [lo, hi]} for i, (t, lo, hi) in enumerate(spec)}
    return {"train": d, "test": {k: dict(v) for k, v in d.items()}}


def model_cfg(kind):
    if kind == "pretrain":
        lv = [(16, 0, 16), (32, 16, 32), (64, 32, 100000)]
    elif kind == "finetune":
        lv = [(8, 0, 8), (16, 8, 16), (32, 16, 32), (48, 32, 48), (64, 48, 100000)]
    else:
        raise ValueError(kind)

    def block(name, stride, C, FF):
        return {"NAME": name,
                "PREPROCESS": {"WINDOW_SHAPE": [8, 8, 1], "DROP_INFO": _levels(lv), "SHUFFLE_VOXELS": False,
                               "POS_TEMPERATURE": 1000, "NORMALIZE_POS": False},
                "ENCODER": {"NUM_BLOCKS": 2, "STRIDE": stride, "D_MODEL": C, "NHEAD": 8, "DIM_FEEDFORWARD": FF,
                            "DROPOUT": 0.0, "ACTIVATION": "gelu", "LAYER_CFG": {"cosine": True, "tau_min": 0.01}}}

    bb = {"NAME": "SiamWCA_MAE" if kind == "pretrain" else "SiamWCA",
          "SST_BLOCK_LIST": [block("sst_block_x1", 1, 128, 256), block("sst_block_x2", 2, 256, 512),
                             block("sst_block_x4", 2, 256, 512)],
          "FEATURES_SOURCE": ["x_conv1", "x_conv2", "x_conv3"],
          "FUSE_LAYER": {"x_conv1": {"UPSAMPLE_STRIDE": 1, "NUM_FILTER": 128, "NUM_UPSAMPLE_FILTER": 128},
                         "x_conv2": {"UPSAMPLE_STRIDE": 2, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128},
                         "x_conv3": {"UPSAMPLE_STRIDE": 4, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128}}}
    if kind == "pretrain":
        bb["MASK_CONFIG"] = {"RATIO": 0.75, "NUM_PRD_POINTS": 16, "NUM_GT_POINTS": 64, "USE_GROUND_MASK": False,
                             "DIS_THRESH": 0.3, "NUM_ABOVE_GROUND": 0}
    vfe = {"NAME": "TemporalDynVFE", "TYPE": "mean", "WITH_DISTANCE": False, "USE_ABSLOTE_XYZ": True,
           "USE_CLUSTER_XYZ": True, "MLPS": [[64, 128]], "FT": kind == "finetune"}
    cfg = {"VFE": vfe, "BACKBONE_3D": bb}
    if kind == "finetune":   # t_mae.yaml:197-206
        conv = lambda d: {"out_channels": 128, "kernel_size": 3, "dilation": d, "padding": d, "stride": 1}  # noqa: E731
        cfg["BACKBONE_2D"] = {"NAME": "SSTBEVBackbone", "NUM_FILTER": 128, "CONV_KWARGS": [conv(1), conv(1), conv(2), conv(1)],
                              "CONV_SHORTCUT": [0, 1, 2]}
    return cfg
