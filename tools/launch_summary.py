#!/usr/bin/env python
"""Per-kernel totals of the LAST run_sweeps call in an ncu launch list (gpu__time_duration.sum, csv).
usage: tools/launch_summary.py launches.csv [marker-kernel]   (a step starts at the all-cells sweep kernel; the step's first
launches -- lag snapshot, cell records -- sit in front of it, so the cut is made at the last `cellrec_kernel` or, failing
that, at the marker)."""
import csv, sys, collections

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("ssw::", "")
    if name.startswith("cub::"):
        name = name.split("<")[0]
    rows.append((name, us))
names = [n for n, _ in rows]
# the last step: from the last ionization_time_kernel-terminated step backwards
ends = [i for i, n in enumerate(names) if n.startswith("levels_kernel") or n.startswith("peer_hist_push")]
if len(ends) >= 2:
    lo, hi = ends[-2] + 1, ends[-1] + 1
else:
    lo, hi = 0, len(rows)
step = rows[lo:hi]
tot = collections.OrderedDict()
cnt = collections.Counter()
for n, us in step:
    tot[n] = tot.get(n, 0.0) + us
    cnt[n] += 1
total = sum(tot.values())
print(f"# launches {lo}..{hi - 1} of {len(rows)}: the last run_sweeps call (per-launch times under ncu are cold-cache and serialised)")
for n, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{n:62s} n={cnt[n]:3d} total={us:10.2f} us  {100 * us / total:5.1f}%")
print(f"{'sum':62s} n={len(step):3d} total={total:10.2f} us")
if "--seq" in sys.argv:
    for n, us in step:
        print(f"    {us:9.2f}  {n}")
