#!/usr/bin/env python
"""Extract the reference's hard-coded direction tables into a data file.

Reads  /root/reference/src/sweep/direction/healpix.rs (DIRECTION_BINS_{16,21,32,64,84})
and writes subsweep_b200/data/direction_bins.json: {"16": [[x,y,z],...], ...}.
The numbers are written with repr() so the f64 values are identical to the Rust
literals (6 significant digits, NOT re-normalised -- src/sweep/direction/mod.rs:58-75).
Only runs where /root/reference is mounted; the JSON is committed.
"""
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/sweep/direction/healpix.rs")
text = src.read_text()
tables = {}
for m in re.finditer(r"DIRECTION_BINS_(\d+):\s*\[\[f64;\s*3\];\s*(\d+)\]\s*=\s*\[(.*?)\];", text, re.S):
    n, n2, body = int(m.group(1)), int(m.group(2)), m.group(3)
    rows = re.findall(r"\[\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*,?\s*\]", body)
    assert n == n2 == len(rows), (n, n2, len(rows))
    tables[str(n)] = [[float(a), float(b), float(c)] for a, b, c in rows]
tables["1"] = [[1.0, 0.0, 0.0]]  # src/sweep/direction/mod.rs:60
out = Path(__file__).resolve().parent.parent / "subsweep_b200" / "data" / "direction_bins.json"
out.write_text(json.dumps(tables, separators=(",", ":")))
print("wrote", out, {k: len(v) for k, v in tables.items()})
