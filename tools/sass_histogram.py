"""SASS opcode histogram of the library's kernels (read here, no GPU):  python tools/sass_histogram.py [kernel-substring ...]

`cuobjdump -sass` of csrc/libtmae_sm100.so; per kernel the instruction count, the tensor-core / TMA / TMEM mnemonics that prove the
Blackwell path (UTCHMMA/UTCQMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) and the ten most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "t-mae_b200", "csrc", "libtmae_sm100.so")
PROOF = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS", "HMMA", "FFMA", "MUFU")


def main():
    subs = sys.argv[1:]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    name, hist = None, None
    kernels = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("tmae::", "")
            hist = collections.Counter()
            kernels.append((name, hist))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and hist is not None:
            hist[m.group(1)] += 1
    for name, hist in sorted(kernels):
        if subs and not any(s in name for s in subs):
            continue
        total = sum(hist.values())
        proof = "  ".join(f"{k}={hist[k]}" for k in PROOF if hist.get(k))
        top = " ".join(f"{k}:{v}" for k, v in hist.most_common(10))
        print(f"{name}\n    {total} instructions | {proof}\n    top: {top}")


if __name__ == "__main__":
    main()
