"""Per-kernel count of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05 MMA
(UTCHMMA / UTCQMMA ...), tensor-memory loads and stores (LDTM / STTM), TMA tensor loads (UTMALDG), bulk copies (UBLKCP),
mbarrier traffic (SYNCS), packed fp32 (FFMA2 / FMUL2), elect, and the 128-bit global / shared accesses of the reduction kernels.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt        (needs no GPU: cuobjdump on the in-tree library)"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aocb200", "libaocb200.so")
COLS = ["UTC.MMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCBAR", "FFMA2/FMUL2", "ELECT", "LDG.128", "STG.128",
        "LDS.128", "SHFL", "RED/ATOM", "instrs"]


def classify(op):
    out = []
    if re.match(r"UTC[A-Z]*MMA", op):
        out.append("UTC.MMA")
    if op.startswith("LDTM"):
        out.append("LDTM")
    if op.startswith("STTM"):
        out.append("STTM")
    if op.startswith("UTMALDG"):
        out.append("UTMALDG")
    if op.startswith("UBLKCP"):
        out.append("UBLKCP")
    if op.startswith("SYNCS"):
        out.append("SYNCS")
    if op.startswith("UTCBAR"):
        out.append("UTCBAR")
    if op.startswith("FFMA2") or op.startswith("FMUL2"):
        out.append("FFMA2/FMUL2")
    if op.startswith("ELECT"):
        out.append("ELECT")
    if op.startswith("LDG") and ".128" in op:
        out.append("LDG.128")
    if op.startswith("STG") and ".128" in op:
        out.append("STG.128")
    if op.startswith("LDS") and ".128" in op:
        out.append("LDS.128")
    if op.startswith("SHFL"):
        out.append("SHFL")
    if op.startswith("RED") or op.startswith("ATOM"):
        out.append("RED/ATOM")
    return out


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True,
                           text=True).stdout.split("\n")
    tab = OrderedDict()
    cur = None
    it = iter(names)
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", next(it)).replace("void ", "").replace("aoc::", "")
            tab[cur] = dict.fromkeys(COLS, 0)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            tab[cur]["instrs"] += 1
            for c in classify(m.group(1)):
                tab[cur][c] += 1
    print("# cuobjdump -sass aocb200/libaocb200.so (sm_100a), instruction counts per kernel; UTC.MMA = tcgen05.mma, LDTM / STTM ="
          " tcgen05.ld / st,\n# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR = tcgen05.commit")
    print("%-44s " % "kernel" + " ".join("%8s" % c[:8] for c in COLS))
    tot = dict.fromkeys(COLS, 0)
    for k, v in sorted(tab.items(), key=lambda kv: (-kv[1]["UTC.MMA"], -kv[1]["instrs"])):
        print("%-44s " % k[:44] + " ".join("%8d" % v[c] for c in COLS))
        for c in COLS:
            tot[c] += v[c]
    print("%-44s " % ("total (%d kernels)" % len(tab)) + " ".join("%8d" % tot[c] for c in COLS))


if __name__ == "__main__":
    main()
