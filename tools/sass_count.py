#!/usr/bin/env python
"""Instruction mix of the loops of one kernel in libldpc535.so (cuobjdump -sass), to state
per-iteration instruction counts in DESIGN.md / bench.py without running anything.

    python tools/sass_count.py 'decode_warp_kernelILi0ELi6ELi3ELb0' [binary]     # min-sum warp kernel

Prints every backward branch (= loop) with its address range, instruction count and opcode
histogram; fp64-pipe instructions (D*) are summed separately.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gr-ldpc_ece535a_b200", "libldpc535.so")


def main():
    pat = re.compile(sys.argv[1])
    so = sys.argv[2] if len(sys.argv) > 2 else SO           # optional: another binary (a microbenchmark)
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    on, ins = False, []
    for line in sass.splitlines():
        if "Function :" in line:
            if on:
                break
            on = bool(pat.search(line))
            if on:
                print(line.strip())
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
    print("%d instructions" % len(ins))
    for a, t in ins:
        if t.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [x for b, x in ins if lo <= b <= a]
                c = collections.Counter(x.split()[0].split(".")[0] for x in body)
                fp64 = sum(v for k, v in c.items() if k.startswith("D") and k not in ("DEPBAR",))
                print("loop 0x%04x..0x%04x: %d instructions, %d on the fp64 pipe, %d MUFU" % (lo, a, len(body), fp64, c.get("MUFU", 0)))
                print("   ", ", ".join("%s %d" % kv for kv in c.most_common(14)))


if __name__ == "__main__":
    main()
