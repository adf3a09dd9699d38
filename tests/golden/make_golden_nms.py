"""Golden vectors of row N4 made with the REFERENCE's own rotated-IoU implementation (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp, compiled from
/root/reference by oracle/build_ref.py): the IoU matrix of seeded box sets and the keep list of the reference's greedy sweep
(iou3d_nms.cpp:116-135) driven by that matrix.   python tests/golden/make_golden_nms.py  ->  tests/golden/nms.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401  (libtorch must be loaded before the reference .so)
from oracle import head_ref  # noqa: E402

assert head_ref.ref_available(), "run `python oracle/build_ref.py` where /root/reference exists first"
out = {}
for seed, n in ((0, 64), (1, 200), (2, 37)):
    b = head_ref.random_boxes(seed, n)
    iou = head_ref.ref_iou(b, b)
    keep = []
    for i in range(n):                       # boxes are taken as already score-sorted
        if not any(iou[t, i] > 0.5 for t in keep):
            keep.append(i)
    out[f"boxes{seed}"], out[f"iou{seed}"], out[f"keep{seed}"] = b, iou, np.asarray(keep, np.int64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nms.npz"), **out)
print({k: v.shape for k, v in out.items()})
