#!/usr/bin/env python
"""Static SASS histogram per source line of one kernel (no GPU needed; needs -lineinfo).

    python profiles/sass_lines.py <file.o|.cubin> <kernel-name-substring> [--top N] [--by file|line|region]

Extracts the cubin (cuobjdump -xelf), disassembles with `nvdisasm -g` and attributes every SASS instruction to the
innermost `//## File "...", line N` marker in front of it.  Prints FP64-pipe / other instruction counts per line so that a
restructuring of the sweep kernel can be judged before GPU time is spent (the plane loop is almost straight-line code, so the
static count per plane tracks the dynamic one).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def cubin_of(path):
    if path.endswith(".cubin"):
        return path
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, check=True, capture_output=True)
    c = [f for f in os.listdir(d) if f.endswith(".cubin")]
    return os.path.join(d, c[0])


def main():
    path, filt = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    by = sys.argv[sys.argv.index("--by") + 1] if "--by" in sys.argv else "line"
    out = subprocess.run(["nvdisasm", "-g", cubin_of(path)], capture_output=True, text=True).stdout
    cur_fn, cur_line = None, ("?", 0)
    hist = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur_fn = m.group(1)
            continue
        if cur_fn is None or filt not in cur_fn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            if op == "MUFU" and "64H" in m.group(2):
                op = "MUFU64"
            key = cur_line if by == "line" else (cur_line[0], 0)
            hist[key][op] += 1
            total[op] += 1
    n = sum(total.values())
    nf = sum(v for k, v in total.items() if k in FP64 or k == "MUFU64")
    print(f"kernel filter {filt!r}: {n} SASS instr, FP64 pipe {nf} ({100.0 * nf / max(n, 1):.1f} %)")
    print("  " + "  ".join(f"{k} {v}" for k, v in total.most_common(14)))
    rows = []
    for key, c in hist.items():
        t = sum(c.values())
        f = sum(v for k, v in c.items() if k in FP64 or k == "MUFU64")
        rows.append((t, f, key, c))
    rows.sort(key=lambda r: -r[0])
    for t, f, key, c in rows[:top]:
        print(f"{key[0]}:{key[1]:<5d} total {t:5d}  fp64 {f:5d}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(6)))


if __name__ == "__main__":
    main()
