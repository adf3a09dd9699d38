"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/tmae_sm100.h declares
(no compute calls without a GPU), and the host package mirrors the reference's interface."""
import os
import subprocess

import pytest

import tmae_b200
from tmae_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` (build) first"
    decls = _lib.parse_header()
    assert len(decls) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(decls) <= exported, sorted(set(decls) - exported)
    assert {e for e in exported if e.startswith("tmae_")} <= set(decls), "exported but undeclared symbols"


def test_library_loads_and_reports_version():
    L = _lib.lib()
    assert L.version() == 1
    assert L.last_error_string() in (b"",) or isinstance(L.last_error_string(), bytes)


def test_error_path_returns_code_not_exit():
    """Bad arguments come back as a negative code + message (the reference's op calls exit(-1), sst_ops.cpp:7-19)."""
    L = _lib.lib()
    with pytest.raises(RuntimeError, match="point_stride"):
        L.vfe_point_features(None, 0, 99, None, None, None, None, None, None, None)


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_module_interface_mirrors_pcdet():
    import inspect
    from tmae_b200 import backbone, config, vfe
    for cls in (vfe.TemporalDynVFE, vfe.DynVFE):
        assert list(inspect.signature(cls.__init__).parameters)[1:6] == ["model_cfg", "num_point_features", "voxel_size", "point_cloud_range", "grid_size"]
        assert hasattr(cls, "get_output_feature_dim")
    for cls in (backbone.SiamWCA, backbone.SiamWCA_MAE):
        assert list(inspect.signature(cls.__init__).parameters)[1:6] == ["model_cfg", "input_channels", "grid_size", "voxel_size", "point_cloud_range"]
    assert hasattr(backbone.SiamWCA_MAE, "get_loss")
    assert set(tmae_b200.vfe_registry) == {"DynVFE", "TemporalDynVFE"} and set(tmae_b200.backbone_registry) == {"SiamWCA", "SiamWCA_MAE"}
    v, b = tmae_b200.build_model("pretrain", [96, 96, 1], [0.32, 0.32, 8.0], [-15.36, -15.36, -5, 15.36, 15.36, 3])
    assert v.get_output_feature_dim() == 128 and b.num_point_features == 128
    keys = set(b.state_dict())
    for k in ("sst_blocks.1.conv_down.0.weight", "sst_blocks.0.encoder_blocks.1.encoder_list.0.win_attn.self_attn.in_proj_weight",
              "wca_blocks.2.encoder_blocks.0.encoder_list.1.win_attn.cross_attn.tau", "decoder_deblocks.2.0.weight",
              "decoder_conv_out.1.running_var", "decoder_pred.bias", "sst_blocks.2.conv_out.1.num_batches_tracked"):
        assert k in keys, k
    assert tuple(b.state_dict()["sst_blocks.1.conv_down.0.weight"].shape) == (256, 3, 3, 128)
    assert set(v.state_dict()) == {f"dvfe_mlps.0.{i}.weight" for i in (0, 1, 3, 4)} | {f"dvfe_mlps.0.{i}.{n}" for i in (1, 4) for n in ("bias", "running_mean", "running_var", "num_batches_tracked")}


def test_state_dict_interchanges_with_oracle():
    from oracle import cases, restated
    S = cases.SMALL
    for kind in ("pretrain", "finetune"):
        v, b = tmae_b200.build_model(kind, S["grid"], S["voxel"], S["range"])
        ov, ob = restated.build(kind, S["grid"], S["voxel"], S["range"])
        v.load_state_dict(ov.state_dict(), strict=True)
        b.load_state_dict(ob.state_dict(), strict=True)


def test_no_oracle_import_in_product():
    root = os.path.dirname(_lib.__file__)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_ops_refuse_to_run_without_cuda():
    import torch
    from tmae_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.add_pos(torch.zeros(4, 128), torch.zeros(4, dtype=torch.uint8), torch.zeros(64, 128))


def test_record_all_reaches_every_tensor_whatever_the_nesting():
    """ops.record_all (side-stream tables -> main stream): no depth limit, cycles tolerated, modules and ctypes structs skipped.
    Host logic only: tensors are stand-ins that claim to be CUDA tensors and log the call."""
    import ctypes

    import torch

    from tmae_b200 import ops

    seen = []

    class Fake(torch.Tensor):
        @property
        def is_cuda(self):
            return True

        def record_stream(self, stream):
            seen.append((id(self), stream))

    def fake():
        return torch.zeros(1).as_subclass(Fake)

    class Slotted:
        __slots__ = ("a", "b", "tcache")

    class Plain:
        pass

    class Struct(ctypes.Structure):
        _fields_ = [("p", ctypes.c_void_p)]

    leaves = [fake() for _ in range(6)]
    s = Slotted()
    s.a, s.b, s.tcache = leaves[0], [leaves[1], (leaves[2],)], {"k": Struct()}
    p = Plain()
    p.part, p.self_ref, p.mod = s, p, torch.nn.Linear(2, 2)       # a cycle and a module: neither may be walked into
    obj = leaves[3]
    for _ in range(12):                                           # far below the old depth limit of 6
        obj = {"x": [obj]}
    ops.record_all((obj, [[[[[[[[p]]]]]]]], leaves[4], None, "text", 3), "STREAM")
    assert {i for i, _ in seen} == {id(t) for t in leaves[:5]} and all(st == "STREAM" for _, st in seen)
    assert id(leaves[5]) not in {i for i, _ in seen}
