#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): key raw metrics + per-instruction hot spots.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"), file=out)
    for k in WANT:
        if k in d:
            print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}", file=out)
    for k in hdr:
        if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith("_per_issue_active.ratio") or \
           (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
            try:
                if float(d[k]) > 0.3:
                    print(f"  {k:70s} {d[k]:>16s}", file=out)
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = None
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        hi = i
        break
if hi is not None:
    h = rows[hi]
    ix = {k: j for j, k in enumerate(h)}
    body = []
    for r in rows[hi + 1:]:   # a report with several launches repeats the header block: keep the first launch
        if len(r) != len(h) or r == h:
            if body:
                break
            continue
        body.append(r)
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body) or 1
    print(f"-- per-instruction (>=1% of {tot} samples, or memory instructions) --", file=out)
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    for r in body:
        s = int(r[ix["# Samples"]] or 0)
        th = int(r[ix["L2 Theoretical Sectors Global"]] or 0)
        if s >= 0.01 * tot or th > 0:
            top = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:2]
            tops = " ".join(f"{k[6:]}={v}" for v, k in top if v)
            print(f"  {r[ix['Source']].strip()[:58]:58s} {100*s/tot:5.1f}% sect={th:>10d} excess={r[ix['L2 Theoretical Sectors Global Excessive']]:>10s} {tops}", file=out)
