"""Per-kernel SASS opcode histogram of the shipped library (`cuobjdump -sass libpalhist.so`): the evidence for which
hardware paths each kernel uses — UTCHMMA (tcgen05.mma kind::f16), LDTM/STTM (tcgen05.ld/st), UBLKCP / UBLKRED
(cp.async.bulk / cp.reduce.async.bulk), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32x2), FHFMA + F2FP (fp16 split),
MUFU, MATCH (warp match), LDG.E.128 / STG.E.128 (128-bit global accesses).

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "palette_and_histo_gan_b200", "libpalhist.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKRED", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FHFMA", "F2FP",
        "MUFU", "MATCH", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "ATOMS", "REDUX", "DADD", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if "." in k:  # e.g. LDG.E.128: the width is the last modifier, cache hints may sit in between
                    head, width = k.rsplit(".", 1)
                    hit = (op + ".").startswith(head + ".") and ("." + width + ".") in (op + ".")
                else:
                    hit = op == k or op.startswith(k + ".")
                if hit:
                    cur[k] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (static instruction counts, sm_100a)")
    for name, c in kernels.items():
        nice = re.sub(r"\(.*", "", demangle(name))
        counts = ", ".join(f"{k} {c[k]}" for k in KEYS if c[k])
        print(f"{nice}: {c['_total']} instructions; {counts}")


if __name__ == "__main__":
    sys.exit(main())
