#!/usr/bin/env python
"""SASS evidence for profiles/: per hot kernel of libdvbs2_cuda.so, the instruction count and the opcode histogram
(cuobjdump -sass), plus the mnemonics the design relies on (packed-halfword VIADD.16x2 / VIMNMX.S16x2, MUFU.EX2 / MUFU.LG2,
mbarrier SYNCS, VOTE) and the ones it deliberately does not use (UTMALDG / UTMASTG / UBLKCP: no TMA).
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dvbs2_b200", "libdvbs2_cuda.so")
HOT = ["ldpc_decode_kernelIfLi25ELi2ELb0", "ldpc_decode_kernelIfLi9ELi2ELb0", "ldpc_decode_kernelIfLi25ELi1ELb0",
       "ldpc_decode_kernelIsLi25ELi4ELb1", "ldpc_decode_kernelIsLi9ELi4ELb1", "ldpc_decode_coop_kernelIfLi25E", "ldpc_decode_coop_kernelIsLi25E",
       "ldpc_onchip_kernelILi25ELi3ELb1", "demap_to_var_kernelILi2EfLi2ELi0", "demap_to_var_kernelILi3EfLi2ELi0",
       "demap_to_var_kernelILi2EfLi2ELi2", "bch_remainder_kernel", "bch_correct_kernel", "unpack_bytes_kernel"]
WATCH = ["VIADD.16x2", "VIMNMX.S16x2", "VIADDMNMX.S16x2", "VIMNMX3", "FMNMX", "MUFU.EX2", "MUFU.LG2", "SYNCS", "VOTE", "REDUX", "LDG.E.64",
         "STG.E.64", "LDS", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "UTCHMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s+Function : ", sass)[1:]
    print("cuobjdump -sass dvbs2_b200/libdvbs2_cuda.so (sm_100a), %d kernels in the library\n" % len(funcs))
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if not any(h in name for h in HOT):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Za-z0-9_.]+)", f, flags=re.M)
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"dvbs2::\(anonymous namespace\)::", "", dem)
        hist = collections.Counter(o.split(".")[0] for o in ops)
        print("## %s\n   %d instructions; top opcodes: %s" % (dem, len(ops), ", ".join("%s %d" % kv for kv in hist.most_common(12))))
        full = collections.Counter(ops)
        seen = []
        for w in WATCH:
            n = sum(c for o, c in full.items() if o.startswith(w) or (w in ("LDG.E.64", "STG.E.64") and o.startswith(w[:3]) and ".64" in o))
            seen.append("%s %d" % (w, n))
        print("   watched: " + ", ".join(seen) + "\n")


if __name__ == "__main__":
    main()
