"""Aggregate one training step out of an ncu launch list (gpu__time_duration.sum CSV): per (kernel, grid) totals of the
step between the last two adam_ema launches."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    L = [(r[ki].split('(')[0].replace('void ', '').replace('dmu::', '').replace('__nv_bfloat16', 'bf16'), r[gi], float(r[vi].replace(',', '')) / 1e3)
         for r in rows[1:]]
    idx = [i for i, x in enumerate(L) if 'adam_ema' in x[0]]
    step = L[idx[-2] + 1:idx[-1] + 1]
    print(f"{len(step)} launches, {sum(x[2] for x in step):.1f} us (cold-cache, serialised)")
    fam = collections.defaultdict(lambda: [0, 0.0])
    for n, g, t in step:
        fam[n][0] += 1
        fam[n][1] += t
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:52]:52s} n={v[0]:3d} tot={v[1]:7.1f}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, t in step:
        agg[(n, g)][0] += 1
        agg[(n, g)][1] += t
    print("by (kernel, grid):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {k[0][:44]:44s} {k[1]:16s} n={v[0]:3d} tot={v[1]:7.1f} avg={v[1] / v[0]:6.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
