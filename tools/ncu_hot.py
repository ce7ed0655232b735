"""Top sampled SASS instructions of one ncu capture (warp-state sampling), with a few instructions of context, and the
sample share of 100-instruction regions.    python tools/ncu_hot.py <file.ncu-rep> [top]"""
import csv
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr, data = rows[h], rows[h + 1:]
    isamp, isrc = hdr.index("# Samples"), hdr.index("Source")
    stall = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    n = [int(r[isamp] or 0) for r in data]
    tot = sum(n)
    print(rows[0][1] if len(rows[0]) > 1 else "", "samples", tot, "instructions", len(data))
    print("-- regions of 100 instructions: first instruction, share of samples")
    for k in range(0, len(data), 100):
        s = sum(n[k:k + 100])
        if s * 50 >= tot:
            print("%6d %5.1f%%  %s" % (k, 100.0 * s / tot, data[k][isrc][:70]))
    print("-- top instructions")
    for i in sorted(range(len(data)), key=lambda i: -n[i])[:top]:
        why = sorted(((int(data[i][j] or 0), hdr[j][6:]) for j in stall), reverse=True)[:2]
        print("%6d %5.1f%%  %-60s %s" % (i, 100.0 * n[i] / tot, data[i][isrc][:60], " ".join("%s=%d" % (w, c) for c, w in why if c)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
