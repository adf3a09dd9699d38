"""TEST INFRASTRUCTURE ONLY.  Builds the checkers of row N4:
  oracle/_build/libnms_oracle.so  <- oracle/nms_ref.c (the in-repo C restatement), always;
  oracle/_ref/libiou3d_ref.so     <- the REFERENCE's own CPU rotated-IoU (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp), compiled with g++ directly
                                     from /root/reference (only when that tree exists: it does not on the GPU box, where the prebuilt file
                                     travels with the snapshot) plus oracle/ref_glue.cpp.  Nothing is copied out of the reference.
    python oracle/build_ref.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp"
ORACLE_SO = os.path.join(HERE, "_build", "libnms_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libiou3d_ref.so")


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(verbose=False):
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    src = os.path.join(HERE, "nms_ref.c")
    if _newer(ORACLE_SO, [src]):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", ORACLE_SO, "-lm"], check=True)
    out = {"oracle": ORACLE_SO, "reference": None}
    if os.path.exists(REF_SRC):
        glue = os.path.join(HERE, "ref_glue.cpp")
        if _newer(REF_SO, [REF_SRC, glue]):
            import torch
            from torch.utils import cpp_extension as ce
            os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
            import sysconfig
            inc = [f"-I{p}" for p in ce.include_paths()] + ["-I/usr/local/cuda/include", f"-I{os.path.dirname(REF_SRC)}",
                                                             f"-I{sysconfig.get_paths()['include']}"]
            lib = os.path.join(os.path.dirname(torch.__file__), "lib")
            cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-ffp-contract=off", "-w", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))] + inc + \
                  [REF_SRC, glue, "-o", REF_SO, f"-L{lib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10", f"-Wl,-rpath,{lib}"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write("oracle/build_ref.py: the reference's iou3d_cpu.cpp did not compile here (recorded in DESIGN.md):\n" + r.stderr[-2000:] + "\n")
                return out
        out["reference"] = REF_SO
    elif os.path.exists(REF_SO):
        out["reference"] = REF_SO
    return out


if __name__ == "__main__":
    print(build(True))
