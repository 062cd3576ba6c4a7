#!/usr/bin/env python
"""Static SASS instruction counts of a kernel per source line of clip.cu (innermost clip.cu frame of
every instruction's inline chain, from `nvdisasm -c -gi`).  Usage:
    python tools/sass_lines.py [libtess_b200.so] [kernel substring] [lo hi]
Prints the per-line counts (or the sum over [lo, hi]) — a quick way to see what a source change does to
the instruction stream before spending GPU time."""
import os
import re
import subprocess
import sys
import tempfile


def per_line(so, kernel="SmallCfgELb0", src="clip.cu"):
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(td) if f.startswith(src.split(".")[0] + ".") and f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-c", "-gi", os.path.join(td, cub)], capture_output=True, text=True).stdout
    counts, total, inside, cur = {}, 0, False, None
    pend = []  # the location comment lines that precede an instruction (innermost frame first)
    for ln in txt.split("\n"):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pend.append(m)
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            if pend:
                cur = None
                for m in pend:  # first frame that lies in the source file
                    if m.group(1).endswith(src):
                        cur = int(m.group(2))
                        break
                    if m.group(3) and m.group(3).endswith(src):
                        cur = int(m.group(4))
                        break
                pend = []
            total += 1
            counts[cur] = counts.get(cur, 0) + 1
    return counts, total


if __name__ == "__main__":
    so = sys.argv[1] if len(sys.argv) > 1 else "the-tessellator_b200/libtess_b200.so"
    kern = sys.argv[2] if len(sys.argv) > 2 else "SmallCfgELb0"
    c, t = per_line(so, kern)
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3]), int(sys.argv[4])
        print(sum(v for k, v in c.items() if k is not None and lo <= k <= hi), "of", t)
    else:
        print("total", t)
        for k in sorted(c, key=lambda x: (x is None, x)):
            print(k, c[k])
