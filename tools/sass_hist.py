#!/usr/bin/env python3
"""Opcode histogram of the filter path's kernels in the built library (cuobjdump -sass), whole kernel and hot loop,
plus the register-bank issue model of the hot loop (tools/sass_bank_model.py).  Output is committed under profiles/
with every kernel change so that the FFMA2 / FMNMX3 / UBLKCP claims can be checked without rebuilding.

    python tools/sass_hist.py [path/to/librma_b200.so] > profiles/r02_fsearch_sass_hist.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rma_net_b200", "lib", "librma_b200.so")
KERNELS = ["chamfer_fsearch_kernel", "chamfer_fresolve_kernel", "chamfer_fprep_kernel", "chamfer_cd_backward_kernel",
           "chamfer_search_kernel"]


def functions(text):
    out, name, buf = [], None, []
    for line in text.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                out.append((name, buf))
            name, buf = m.group(1), []
        elif name:
            buf.append(line)
    if name:
        out.append((name, buf))
    return out


def parse(lines):
    ins = []
    for l in lines:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            t = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            ins.append((int(m.group(1), 16), t.split()[0], t))
    return ins


def hot_loop(ins):
    best = None
    for addr, op, text in ins:
        m = re.search(r"0x([0-9a-f]+)\s*$", text)
        if op.startswith("BRA") and m and int(m.group(1), 16) < addr:
            t = int(m.group(1), 16)
            n = sum(1 for a, o, _ in ins if t <= a <= addr and o.startswith("FFMA"))
            if n and (best is None or n > best[0] or (n == best[0] and addr - t < best[2] - best[1])):
                best = (n, t, addr)
    return best


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print(f"# {os.path.relpath(LIB, ROOT)}   (cuobjdump -sass; sm_100a)")
    for name, lines in functions(sass):
        if not any(k in name for k in KERNELS):
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        ins = parse(lines)
        c = Counter(o for _, o, _ in ins)
        print(f"\n== {dem}\n   {len(ins)} instructions: " + ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
        tma = {k: v for k, v in c.items() if k.startswith(("UBLKCP", "SYNCS", "UTMA", "UTC", "LDTM", "HMMA", "QMMA"))}
        print(f"   TMA / mbarrier / tensor opcodes: {tma if tma else 'none'}")
        hl = hot_loop(ins)
        if hl and ("fsearch" in name or "chamfer_search" in name):
            n, lo, hi = hl
            loop = [i for i in ins if lo <= i[0] <= hi]
            lc = Counter(o for _, o, _ in loop)
            print(f"   hot loop {lo:04x}..{hi:04x}: {len(loop)} instructions: " + ", ".join(f"{k} {v}" for k, v in lc.most_common()))
            model = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_bank_model.py"), "--from", f"{lo:x}",
                                    "--to", f"{hi:x}", "--pairs",
                                    # filter loop: 3 FFMA2 per 2 pairs; exact loop: 1 FMUL2 (dx*dx) per 2 pairs
                                    str(2 * lc["FMUL2"] if lc.get("FMUL2") else 2 * lc.get("FFMA2", 192) // 3)],
                                   input="\n".join(lines), capture_output=True, text=True).stdout
            print("   register-bank issue model of the loop (rt = max(pipe, distinct even regs, distinct odd regs), reuse latch):")
            for l in model.splitlines():
                print("     " + l)


if __name__ == "__main__":
    main()
