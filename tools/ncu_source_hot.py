"""Hottest SASS instructions (by warp-stall samples) from `ncu --page source --csv`."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = rows[2:]
    si = hdr.index("# Samples")
    ei = hdr.index("Instructions Executed")
    src = hdr.index("Source")
    wi = hdr.index("L1 Wavefronts Shared")
    tot = sum(int(r[si]) for r in data if r[si].isdigit())
    tote = sum(int(r[ei]) for r in data if r[ei].isdigit())
    totw = sum(int(r[wi]) for r in data if r[wi].isdigit())
    print("total samples", tot, "instructions", tote, "smem wavefronts", totw, "sass lines", len(data))
    # cumulative by region: print instructions in program order with samples share, compressed
    ranked = sorted(range(len(data)), key=lambda i: -int(data[i][si]) if data[i][si].isdigit() else 0)[:top]
    for i in sorted(ranked):
        r = data[i]
        print(f"{i:5d} {100*int(r[si])/tot:5.1f}%  exec={int(r[ei]):>11d} smemwf={r[wi]:>10s} {r[src].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
