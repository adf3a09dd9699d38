"""TEST INFRASTRUCTURE ONLY -- tier-1 oracle: the reference's own hot-path modules, loaded
read-only from /root/reference with leaf shims.

Nothing is copied out of the reference: its files are imported from where they lie.  The
reference root is searched in $TMAE_REF, /root/reference; when none exists (the GPU box)
`available()` is False and tier-1 tests skip -- tier 2 (`oracle/restated.py`), pinned to
tier 1 by the committed golden tensors under tests/golden/, takes over.
"""
import os
import sys
import types

import numpy as np
import yaml

from . import shims

_ROOTS = [os.environ.get("TMAE_REF", ""), "/root/reference"]
_loaded = {}


def ref_root():
    for r in _ROOTS:
        if r and os.path.isdir(os.path.join(r, "pcdet", "models", "backbones_3d")):
            return r
    return None


def available():
    return ref_root() is not None


def load():
    """Returns a namespace with the reference classes: TemporalDynVFE, DynVFE, SiamWCA_MAE,
    SiamWCA, SSTInputLayer, sst_utils, sst_ops_utils, common_utils, cosine_msa."""
    if _loaded:
        return _loaded["ns"]
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (tier-1 oracle unavailable)")
    # bare parent packages: their __init__.py never runs (pcdet/__init__.py:4 needs version.py,
    # pcdet/models/__init__.py pulls every compiled op)
    for name, rel in [("pcdet", "pcdet"), ("pcdet.models", "pcdet/models"), ("pcdet.ops", "pcdet/ops"),
                      ("pcdet.utils", "pcdet/utils"), ("pcdet.models.model_utils", "pcdet/models/model_utils"),
                      ("pcdet.ops.sst_ops", "pcdet/ops/sst_ops")]:
        if name in sys.modules and not getattr(sys.modules[name], "_tmae_bare", False):
            raise RuntimeError("a real pcdet is already imported")
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(root, rel)]
        m._tmae_bare = True
        sys.modules[name] = m
    for name, m in shims.make_modules().items():
        sys.modules[name] = m
    import importlib
    ns = types.SimpleNamespace()
    ns.TemporalDynVFE = importlib.import_module("pcdet.models.backbones_3d.vfe.temporal_dyn_vfe").TemporalDynVFE
    ns.DynVFE = importlib.import_module("pcdet.models.backbones_3d.vfe.dyn_vfe").DynVFE
    ns.SiamWCA_MAE = importlib.import_module("pcdet.models.backbones_3d.SiamWCA_MAE").SiamWCA_MAE
    siam = importlib.import_module("pcdet.models.backbones_3d.SiamWCA")
    ns.SiamWCA = siam.SiamWCA
    ns.SSTInputLayer_Temporal = siam.SSTInputLayer_Temporal
    spt = importlib.import_module("pcdet.models.backbones_3d.spt_backbone")
    ns.SSTInputLayer, ns.SSTBlockV1 = spt.SSTInputLayer, spt.SSTBlockV1
    ns.sst_utils = importlib.import_module("pcdet.models.model_utils.sst_utils")
    ns.sst_ops_utils = importlib.import_module("pcdet.ops.sst_ops.sst_ops_utils")
    ns.common_utils = importlib.import_module("pcdet.utils.common_utils")
    ns.cosine_msa = importlib.import_module("pcdet.models.model_utils.cosine_msa")
    ns.root = root
    _loaded["ns"] = ns
    return ns


def load_cfg(kind):
    """Fresh EasyDict of cfg.MODEL from the reference's YAML (a fresh copy per model instance:
    WCABlock.__init__ mutates ENCODER.NUM_BLOCKS, SiamWCA.py:294-296).  kind: 'pretrain'|'finetune'."""
    root = ref_root()
    fn = {"pretrain": "t_mae_ssl.yaml", "finetune": "t_mae.yaml"}[kind]
    with open(os.path.join(root, "tools", "cfgs", "once_models", fn)) as f:
        y = yaml.safe_load(f)
    return shims.EasyDict(y["MODEL"])


def build(kind, grid_size, voxel_size, pc_range, num_point_features=5, seed=0):
    """(vfe, backbone) reference modules under torch.manual_seed(seed)."""
    import torch
    ns = load()
    cfg = load_cfg(kind)
    torch.manual_seed(seed)
    vfe = ns.TemporalDynVFE(cfg.VFE, num_point_features, voxel_size, pc_range, np.asarray(grid_size))
    cls = ns.SiamWCA_MAE if kind == "pretrain" else ns.SiamWCA
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):  # WCABlock prints a warning per block
        bb = cls(cfg.BACKBONE_3D, vfe.get_output_feature_dim(), np.asarray(grid_size), voxel_size, pc_range)
    return vfe, bb


def build_bev(seed=0):
    """The reference's own SSTBEVBackbone (pcdet/models/backbones_2d/sst_bev_backbone.py, loaded by file path: it imports
    numpy and torch only) with cfg.MODEL.BACKBONE_2D of t_mae.yaml."""
    import importlib.util
    import torch
    root = ref_root()
    spec = importlib.util.spec_from_file_location("_tmae_ref_sst_bev", os.path.join(root, "pcdet", "models", "backbones_2d", "sst_bev_backbone.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(seed)
    return mod.SSTBEVBackbone(load_cfg("finetune").BACKBONE_2D)
