#!/usr/bin/env python
"""Fingerprint of the SASS of every kernel in libmergespmv.so: instruction count and a hash of the
instruction stream that ignores operand order inside an instruction (nvcc swaps the operands of
commutative adds when code is moved between inline functions).  Used to show that the kernels the
bench runs by default are still the ones the committed measurements were taken with
(profiles/sass_fingerprint_r01.json, tests/test_abi.py).

    python tools/sass_fingerprint.py [lib.so | dump.sass] > fingerprint.json
"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fingerprint(path):
    if path.endswith(".so"):
        text = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    out, name = {}, None
    for line in text.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
            ins = re.sub(r"`\(\.L_x_\d+\)", "L", ins)  # local labels are renumbered when kernels are added
            toks = re.findall(r"[^\s,]+", ins)
            out[name].append(" ".join(toks[:1] + sorted(toks[1:])) if toks else "")
    return {k: {"instructions": len(v), "sha1": hashlib.sha1("\n".join(v).encode()).hexdigest()} for k, v in out.items()}


if __name__ == "__main__":
    p = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "merge-spmv_b200", "libmergespmv.so")
    print(json.dumps(fingerprint(p), indent=1, sort_keys=True))
