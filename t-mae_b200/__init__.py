"""tmae_b200: B200-native (sm_100a) implementation of T-MAE's sparse-window voxel-encoder hot path.

Host side mirrors the pcdet module API (vfe / backbone_3d `forward(batch_dict)`, `get_loss()`);
all arithmetic on the path runs in hand-written CUDA kernels behind the C-ABI library
`csrc/libtmae_sm100.so` (see include/tmae_sm100.h).  There is no CPU fallback: using the
compute modules without the built library, or with CPU tensors, raises.

    from tmae_b200 import build_model
    vfe, backbone = build_model("pretrain", grid_size, voxel_size, point_cloud_range)

`vfe_registry` / `backbone_registry` have the shape of pcdet's `vfe.__all__` / `backbones_3d.__all__`
(pcdet/models/backbones_3d/vfe/__init__.py:8-14, pcdet/models/backbones_3d/__init__.py:7-12); `backbone_2d_registry`
that of `backbones_2d.__all__` (pcdet/models/backbones_2d/__init__.py) for `SSTBEVBackbone`.
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: `import tmae_b200.synth` must work without torch/CUDA
    if name in ("DynVFE", "TemporalDynVFE", "vfe_registry"):
        from . import vfe
        reg = {"DynVFE": vfe.DynVFE, "TemporalDynVFE": vfe.TemporalDynVFE}
        return reg if name == "vfe_registry" else reg[name]
    if name in ("SiamWCA", "SiamWCA_MAE", "backbone_registry"):
        from . import backbone
        reg = {"SiamWCA": backbone.SiamWCA, "SiamWCA_MAE": backbone.SiamWCA_MAE}
        return reg if name == "backbone_registry" else reg[name]
    if name in ("SSTBEVBackbone", "backbone_2d_registry"):
        from . import bev
        return {"SSTBEVBackbone": bev.SSTBEVBackbone} if name == "backbone_2d_registry" else bev.SSTBEVBackbone
    if name == "build_model":
        return _build_model
    raise AttributeError(name)


def _build_model(kind, grid_size, voxel_size, point_cloud_range, num_point_features=5, cfg=None):
    """(vfe, backbone_3d) for kind in {'pretrain', 'finetune'} the way Detector3DTemplate.build_vfe /
    build_backbone_3d do (pcdet/models/detectors/detector3d_template.py:70-100)."""
    from . import backbone, config, vfe
    cfg = cfg or config.model_cfg(kind)
    v = vfe.TemporalDynVFE(cfg["VFE"], num_point_features, voxel_size, point_cloud_range, grid_size)
    cls = {"SiamWCA": backbone.SiamWCA, "SiamWCA_MAE": backbone.SiamWCA_MAE}[cfg["BACKBONE_3D"]["NAME"]]
    return v, cls(cfg["BACKBONE_3D"], v.get_output_feature_dim(), grid_size, voxel_size, point_cloud_range)
