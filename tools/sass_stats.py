#!/usr/bin/env python3
"""Static SASS statistics of one kernel of a built library: instruction count, opcode
histogram, registers.  `python tools/sass_stats.py <lib.so> <substring of the mangled name>`.
A CPU-side check before spending GPU time (cuobjdump needs no GPU)."""
import collections
import re
import subprocess
import sys


def kernel_sass(lib, pattern):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    blocks = out.split("Function : ")
    return [(b.split("\n", 1)[0].strip(), b) for b in blocks[1:] if pattern in b.split("\n", 1)[0]]


def stats(body):
    ops = collections.Counter()
    for line in body.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    return ops


def main():
    lib, pattern = sys.argv[1], sys.argv[2]
    for name, body in kernel_sass(lib, pattern):
        ops = stats(body)
        total = sum(ops.values())
        fp = sum(v for k, v in ops.items() if k in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "MUFU"))
        print("%s\n  %d instructions, %d floating point; top: %s" % (
            name, total, fp, ", ".join("%s %d" % kv for kv in ops.most_common(24))))


if __name__ == "__main__":
    main()
