"""Per CUDA source line: instructions executed and stall samples, from
`ncu -i rep --page source --csv --print-source cuda,sass` (needs -lineinfo + --import-source on)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    file = None
    agg = []
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":  # a CUDA line with its aggregated metrics
            try:
                agg.append((file, int(r[0]), r[1].strip()[:100], int(r[si]), int(r[ei])))
            except ValueError:
                pass
    tot_s = sum(a[3] for a in agg) or 1
    tot_e = sum(a[4] for a in agg) or 1
    print(f"total samples {tot_s} instructions {tot_e}")
    for a in sorted(agg, key=lambda a: -a[4])[:top]:
        print(f"{a[0]:14s}:{a[1]:4d} inst {100*a[4]/tot_e:5.1f}%  samples {100*a[3]/tot_s:5.1f}%  {a[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
