"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, optionally, every launch."""
import collections
import csv
import sys


def main(path, detail_pat=None):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.defaultdict(lambda: [0, 0.0])
    detail = []
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
        if detail_pat and detail_pat in name:
            detail.append((name, r[gi], v))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':64s} {'n':>5s} {'us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} {v[0]:5d} {v[1]:10.1f} {v[1] / tot:6.3f}")
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for d in detail:
        print(f"  {d[0][:40]:40s} grid {d[1]:18s} {d[2]:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
