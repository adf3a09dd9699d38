"""Batch assembly in front of the VFE: oracle pins on the CPU, CUDA parity on the GPU.

The path is float64 affine maps + comparisons + a stable compaction, so the bar is BIT-EXACT float32 output rows (the
kernel evaluates the two length-3 / length-4 dot products as the fused multiply-add chain numpy's BLAS produces; on the
golden inputs the float64 intermediates agree bit for bit).
"""
import os

import numpy as np
import pytest
import torch

from oracle import assemble_ref, cases

import tmae_b200  # noqa: F401
from tmae_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assemble.npz")
KINDS = [("once", True), ("once", False), ("waymo", True), ("waymo", False)]


def _golden(kind, align):
    g = np.load(GOLDEN)
    return cases.raw_samples(int(g["seed"]), int(g["n"]), kind), g[f"{kind}_{int(align)}_points"], g[f"{kind}_{int(align)}_points_prev"]


# ------------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("kind,align", KINDS)
def test_tier2_equals_golden(kind, align):
    samples, gp, gq = _golden(kind, align)
    p, q = assemble_ref.assemble(samples, synth.SHAPES[kind]["range"], align_two_frames=align)
    assert p.dtype == np.float32 and np.array_equal(p, gp, equal_nan=True)
    assert np.array_equal(q, gq, equal_nan=True)


@pytest.mark.skipif(not assemble_ref.tier1_available(), reason="reference tree not present")
def test_tier1_equals_tier2_live():
    samples = cases.raw_samples(905, 5000, "once")
    rng = synth.SHAPES["once"]["range"]
    a, b = assemble_ref.assemble_tier1(samples, rng), assemble_ref.assemble(samples, rng)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_known_answers():
    """Hand-checkable cases: the ego test is strict, the crop is a closed interval on the float32 range, a pure-yaw pose
    pair maps a point as expected, all-zero poses leave the points alone."""
    rng = [-10.0, -10.0, -5.0, 10.0, 10.0, 3.0]
    pts = np.array([[2.0, 0.0, 0, 1], [1.5, -1.5, 0, 2], [9.0, -9.0, 0, 3], [10.000001, 0.0, 0, 4], [3.0, 4.0, 1, 5]], np.float32)
    half = np.sqrt(0.5)
    s = dict(points=pts, points_prev=pts.copy(), pose=np.array([0, 0, 0, 1, 1.0, 0, 0]), pose_prev=np.array([0, 0, half, half, 0, 0, 0.0]))
    p, q = assemble_ref.assemble([s], rng)
    assert p[:, 4].tolist() == [1, 3, 5] and (p[:, 0] == 0).all()
    # prev -> global: +90 deg yaw (x, y) -> (-y, x); global -> cur: subtract (1, 0, 0)
    assert q[:, 4].tolist() == [1, 3, 5]
    np.testing.assert_allclose(q[:, 1:4], [[-1, 2, 0], [8, 9, 0], [-5, 3, 1]], atol=1e-6)
    s0 = dict(s, pose=np.zeros(7), pose_prev=np.zeros(7))
    p0, q0 = assemble_ref.assemble([s0], rng)
    assert np.array_equal(p0, q0)


def test_pose_affines_host_logic():
    """The product's host-side pose -> affine conversion applied with numpy reproduces convert_prv_frame_to_cur bit for
    bit, and the zero-pose rule yields flag 0."""
    from tmae_b200.assemble import pose_affines
    for s in cases.raw_samples(31, 800, "once"):
        xf, fl = pose_affines(s["pose_prev"], s["pose"])
        x = s["points_prev"][:, :3]
        if fl[0]:
            x = np.dot(x, xf[0, :, :3].T) + xf[0, :, 3]
        if fl[1]:
            x = np.dot(np.concatenate([x, np.ones((x.shape[0], 1))], -1), np.vstack([xf[1], [0, 0, 0, 1]]).T)[:, :3]
        ref = assemble_ref.convert_prv_frame_to_cur(s["points_prev"], s["pose_prev"], s["pose"])[:, :3]
        assert np.array_equal(x, ref, equal_nan=True)
        assert fl.tolist() == [int(np.any(s["pose_prev"])), int(np.any(s["pose"]))]
    with pytest.raises(ValueError):
        pose_affines(np.array([0, 0, 0, 0, 1.0, 2, 3]), np.zeros(7))   # zero-norm quaternion, like scipy in the reference


def test_oracle_properties():
    """Size-independent properties of the path (the GPU tests check the same on the CUDA output through the oracle):
    every kept point is outside the ego box in its RAW coordinates and inside the crop after alignment; the order of the
    kept points is the input order; samples are independent (assembling a sub-batch gives the same rows up to the sample
    index); a second pass with zero poses and no ego box is the identity."""
    kind = "once"
    rng = synth.SHAPES[kind]["range"]
    samples = cases.raw_samples(77, 2500, kind)
    p, q = assemble_ref.assemble(samples, rng)
    r32 = np.asarray(rng, np.float32)
    for out in (p, q):
        ok = np.isfinite(out[:, 1])
        assert ok.all()
        assert ((out[:, 1] >= r32[0]) & (out[:, 1] <= r32[3]) & (out[:, 2] >= r32[1]) & (out[:, 2] <= r32[4])).all()
        assert (np.diff(out[:, 0]) >= 0).all()                       # samples back to back
    for b, s in enumerate(samples):                                  # current frame: rows are a subsequence of the raw rows
        mine = p[p[:, 0] == b][:, 1:]
        raw = s["points"]
        keep = ~((np.abs(raw[:, 0]) < 2) & (np.abs(raw[:, 1]) < 2)) & assemble_ref.mask_points_by_range(raw, r32)
        assert np.array_equal(mine, raw[keep], equal_nan=True)
    sub_p, sub_q = assemble_ref.assemble(samples[2:3], rng)          # sample independence
    assert np.array_equal(sub_p[:, 1:], p[p[:, 0] == 2][:, 1:]) and np.array_equal(sub_q[:, 1:], q[q[:, 0] == 2][:, 1:])
    again = [dict(points=p[p[:, 0] == b][:, 1:], points_prev=q[q[:, 0] == b][:, 1:], pose=np.zeros(7), pose_prev=np.zeros(7))
             for b in range(len(samples))]
    p2, q2 = assemble_ref.assemble(again, rng, ego_radius=0)
    assert np.array_equal(p2, p) and np.array_equal(q2, q)


def test_assembler_refuses_cpu():
    from tmae_b200.assemble import FrameAssembler
    with pytest.raises(RuntimeError):
        FrameAssembler(synth.ONCE["range"], device="cpu")


# ------------------------------------------------------------------------------------------------- GPU: parity
gpu = pytest.mark.gpu


def _assembler(kind, align=True):
    from tmae_b200.assemble import FrameAssembler
    return FrameAssembler(synth.SHAPES[kind]["range"], align_two_frames=align)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@gpu
@pytest.mark.parametrize("kind,align", KINDS)
def test_cuda_equals_golden(kind, align):
    samples, gp, gq = _golden(kind, align)
    bd = _assembler(kind, align)(samples)
    assert bd["batch_size"] == len(samples)
    assert _same(bd["points"].cpu().numpy(), gp), "points differ from the reference's output"
    assert _same(bd["points_prev"].cpu().numpy(), gq), "points_prev differ from the reference's output"


@gpu
@pytest.mark.parametrize("kind,n,batch", [("once", 60000, 4), ("waymo", 180000, 2)])
def test_cuda_equals_oracle_full_size(kind, n, batch):
    samples = [synth.raw_scan_pair(400 + i, n, kind) for i in range(batch)]
    p, q = assemble_ref.assemble(samples, synth.SHAPES[kind]["range"])
    bd = _assembler(kind)(samples)
    assert _same(bd["points"].cpu().numpy(), p)
    assert _same(bd["points_prev"].cpu().numpy(), q)


@gpu
def test_ragged_and_empty_samples():
    """Empty samples in front, in the middle and at the end; a sample whose points are all dropped; sizes straddling the
    512-point tile of the compaction."""
    rng = synth.ONCE["range"]
    base = synth.raw_scan_pair(77, 5000)
    sizes = [0, 511, 0, 1025, 1, 2049, 0]
    samples = []
    for i, k in enumerate(sizes):
        s = {key: (v[i * 7:i * 7 + k].copy() if isinstance(v, np.ndarray) and v.ndim == 2 else v) for key, v in base.items()}
        samples.append(s)
    samples[4]["points"][:, :2] = 0.5          # all ego points
    samples[4]["points_prev"][:, 0] = 500.0    # all out of range
    p, q = assemble_ref.assemble(samples, rng)
    bd = _assembler("once")(samples)
    assert _same(bd["points"].cpu().numpy(), p) and _same(bd["points_prev"].cpu().numpy(), q)
    empty = [dict(samples[0])]
    bd = _assembler("once")(empty)
    assert bd["points"].shape == (0, 5) and bd["points_prev"].shape == (0, 5)


@gpu
def test_no_sync_tail_is_dropped_by_the_voxeliser():
    """sync=False returns all rows, the tail holding sentinel rows: the VFE's own range test removes them, so voxel
    coordinates, kept points and features equal those of the exact-count call -- with no host read in the assembler."""
    from tmae_b200 import ops
    kind = "once"
    S = synth.SHAPES[kind]
    samples = [synth.raw_scan_pair(500 + i, 20000, kind) for i in range(2)]
    asm = _assembler(kind)
    exact, padded = asm(samples, sync=True), asm(samples, sync=False)
    grid = [int(v) for v in synth.grid_size(S)]
    for key in ("points", "points_prev"):
        n_exact = exact[key].shape[0]
        assert padded[key].shape[0] == sum(s[key].shape[0] for s in samples) > n_exact
        assert torch.equal(padded[key][:n_exact], exact[key])
        tail = padded[key][n_exact:]
        assert (tail[:, 1:4] == 1.0e6).all() and (tail[:, 0] == 0).all() and (tail[:, 4:] == 0).all()
        a = ops.voxelize(exact[key].contiguous(), S["range"], S["voxel"], grid, 2)
        b = ops.voxelize(padded[key], S["range"], S["voxel"], grid, 2)
        assert torch.equal(a["counts"], b["counts"])
        nk, nv = int(a["counts"][0]), int(a["counts"][1])
        assert torch.equal(a["voxel_coords"][:nv], b["voxel_coords"][:nv])
        assert torch.equal(a["points"][:nk], b["points"][:nk])
        assert torch.equal(a["voxel_mean"][:nv], b["voxel_mean"][:nv])


@gpu
def test_assembler_on_main_feeds_the_side_stream_prepass():
    """The assembler produces its tensors on the CURRENT stream; the VFE's side-stream pre-pass must wait for them
    (batch_dict["inputs_ready_event"]).  The main stream is kept busy in front of the assembler so that an unsynchronised side stream
    would read the buffers before they are written."""
    from tmae_b200 import ops
    DEV = "cuda"
    kind = "once"
    S = synth.SHAPES[kind]
    grid = [int(v) for v in synth.grid_size(S)]
    vfe, _ = tmae_b200.build_model("pretrain", grid, S["voxel"], S["range"], num_point_features=5)
    vfe.to(DEV).eval()
    samples = [synth.raw_scan_pair(700 + i, 30000, kind) for i in range(2)]
    asm = _assembler(kind)
    with torch.no_grad():
        ref = vfe(asm(samples, sync=False))
        torch.cuda.synchronize()
        side = ops.side_stream(DEV)
        spin = torch.empty(64 << 20, device=DEV)
        for _ in range(3):
            for _ in range(20):
                spin.add_(1.0)                 # ~10 ms of main-stream work in front of the assembler's kernels
            bd = asm(samples, sync=False)
            assert "inputs_ready_event" in bd
            bd["side_stream"] = side
            out = vfe(bd)
            torch.cuda.synchronize()
            assert torch.equal(out["voxel_coords"], ref["voxel_coords"]) and torch.equal(out["voxel_coords_prev"], ref["voxel_coords_prev"])
            assert torch.equal(out["voxel_features"], ref["voxel_features"])


@gpu
def test_assembled_batch_through_the_vfe_matches_the_oracle_chain():
    """assemble -> TemporalDynVFE on the GPU against oracle assemble -> tier-2 VFE on the CPU."""
    from oracle import restated
    S = cases.SMALL
    samples = cases.raw_samples(61, 4000, "once")
    p, q = assemble_ref.assemble(samples, S["range"])
    from tmae_b200.assemble import FrameAssembler
    bd = FrameAssembler(S["range"])(samples)
    vfe, _ = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
    ovfe, _ = restated.build("pretrain", S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(ovfe)
    vfe.cuda()
    out = vfe(bd)
    ref = ovfe(dict(points=torch.from_numpy(p), points_prev=torch.from_numpy(q), batch_size=len(samples)))
    for key in ("voxel_coords", "voxel_coords_prev"):
        assert torch.equal(out[key].cpu(), ref[key])
    for key in ("voxel_features", "voxel_features_prev"):
        torch.testing.assert_close(out[key].detach().cpu(), ref[key].detach(), rtol=1e-5, atol=1e-5)
