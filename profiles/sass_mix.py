#!/usr/bin/env python
"""Static SASS instruction mix of the kernels in an object / shared library (no GPU needed).

    python profiles/sass_mix.py ramses_b200/csrc/sweep_inst_3d_exact.o [name-filter]

For every kernel whose (demangled) name contains the filter: SASS instruction count, FP64 pipe instructions (DFMA, DMUL,
DADD, DSETP, MUFU.RCP64H/RSQ64H), their share, loads/stores, and the ten most frequent opcodes.  Used between GPU sessions to
judge a code change before spending GPU time: the dynamic mix of the sweep kernel measured by ncu (profiles/r1_ncu_full_*.txt:
FP64 = 50 % of executed warp instructions, issue-slot activity = FP64-pipe activity = 46.6 %) means that non-FP64 instructions
cost exactly as much issue bandwidth as FP64 ones -- trimming either half raises the roof by the same amount.
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, ops = None, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, ops
            name, ops = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            op = m.group(1)
            if op == "MUFU" and ("64H" in m.group(2)):
                op = "MUFU.64H"
            ops[op] += 1
    if name:
        yield name, ops


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    fp64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.64H")
    mem = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "LDC", "LDCU")
    for name, ops in kernels(path):
        d = demangle(name)
        if flt and flt not in d:
            continue
        tot = sum(ops.values())
        nf = sum(ops[o] for o in fp64)
        print(f"{d[:150]}\n  {tot} SASS instr, FP64 pipe {nf} ({100.0 * nf / max(tot, 1):.1f} %), "
              f"memory {sum(ops[o] for o in mem)}, local {ops['LDL'] + ops['STL']}, branches {ops['BRA']}")
        print("  " + "  ".join(f"{o} {c}" for o, c in ops.most_common(10)))


if __name__ == "__main__":
    main()
