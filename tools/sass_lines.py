#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel of libmergespmv.so (built with
-lineinfo).  Works without a GPU: cuobjdump -xelf + nvdisasm --print-line-info.

    python tools/sass_lines.py spmv_pipe_kernelINS_7PipeCfgId        # substring of the mangled kernel name
    python tools/sass_lines.py spmv_pipe_kernelINS_7PipeCfgIf --by-file

Straight-line code like the consumer loop of the pipe kernel executes most lines once per thread and tile, so the static count
is a first estimate of the dynamic one; lines inside uniform branches that are not taken (edge
cases of the staging helpers, the duplicated L1 / no-L1 gather loop) must be discounted by hand.
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble(lib):
    d = tempfile.mkdtemp(prefix="sass_lines_")
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    out = []
    for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):  # one per translation unit
        r = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True)
        if r.returncode == 0:
            out += r.stdout.split("\n")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "merge-spmv_b200", "libmergespmv.so"))
    ap.add_argument("--by-file", action="store_true")
    ap.add_argument("--ops", action="store_true", help="opcode histogram instead of source lines")
    args = ap.parse_args()
    lines = disassemble(args.lib)
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and args.kernel in l]
    if not starts:
        raise SystemExit(f"no kernel matching {args.kernel!r}")
    start = starts[0]
    end = next(i for i in range(start + 1, len(lines)) if lines[i].strip().startswith(".section") or i == len(lines) - 1)
    cur, cnt, ops = None, collections.Counter(), collections.Counter()
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", l)
        if m:
            cnt[cur] += 1
            op = re.sub(r"^@!?U?P\w+\s+", "", m.group(1)).split()[0].split(".")[0]
            ops[op] += 1
    print(f"{lines[start].rstrip(':')}\n{sum(cnt.values())} instructions")
    if args.ops:
        for op, c in ops.most_common():
            print(f"{c:5d} {op}")
        return
    if args.by_file:
        byf = collections.Counter()
        for (f, _), c in cnt.items():
            byf[f] += c
        for f, c in byf.most_common():
            print(f"{c:5d} {f}")
        return
    src = {}
    for (f, ln), c in sorted((k, v) for k, v in cnt.items() if k):
        if f not in src:
            p = os.path.join(ROOT, "merge-spmv_b200", "csrc", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else None
        text = src[f][ln - 1].strip()[:100] if src[f] and ln - 1 < len(src[f]) else ""
        print(f"{c:5d} {f}:{ln}: {text}")


if __name__ == "__main__":
    main()
