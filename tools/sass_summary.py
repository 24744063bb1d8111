#!/usr/bin/env python
"""Opcode histogram of the shipped library (and the tensor-core probe) -- the SASS evidence behind DESIGN.md's claims:
TMA (UTMALDG), mbarriers (SYNCS), packed-byte ALU ops (VABSDIFF4, IDP.4A, LOP3, IADD3/IMAD.IADD), FP64 (DFMA) and, in
tools/tc_probe, tcgen05 (UTCIMMA / STTM / LDTM).   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "SYNCS", "VABSDIFF4", "IDP.4A", "LOP3", "IADD3", "IMAD.IADD", "IMAD", "DFMA", "DMUL", "F2I", "I2F", "LDS", "STS",
        "LDG", "STG", "ATOMS", "BAR", "SHFL", "POPC", "UTCIMMA", "UTCBAR", "STTM", "LDTM", "UTCATOMSWS", "HMMA", "IMMA"]


def dump(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(.*", "", fn)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    return hist


def main():
    for path in [os.path.join(ROOT, "vcs_h264_b200", "libvcs_b200.so"), os.path.join(ROOT, "build", "tc_probe")]:
        if not os.path.exists(path):
            continue
        print(f"== {os.path.relpath(path, ROOT)}  (cuobjdump -sass, sm_100a) ==")
        tot = collections.Counter()
        for fn, h in dump(path).items():
            n = sum(h.values())
            if n < 200 and "tc_probe" not in path:
                continue
            sel = []
            for k in KEYS:
                c = sum(v for op, v in h.items() if op == k or op.startswith(k + ".") or (k == "IMAD" and op.startswith("IMAD") and not op.startswith("IMAD.IADD")))
                if c:
                    sel.append(f"{k} {c}")
            for op, v in h.items():
                tot[op.split(".")[0] if not op.startswith(("IDP", "IMAD.IADD", "VABSDIFF4")) else ".".join(op.split(".")[:2])] += v
            print(f"{fn[:110]}\n    instructions {n}: " + ", ".join(sel))
        print("  whole file: " + ", ".join(f"{k} {v}" for k, v in tot.most_common(24)))
        print()


if __name__ == "__main__":
    main()
