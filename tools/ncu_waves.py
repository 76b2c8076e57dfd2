"""Grid size against resident slots for every kernel of an `ncu --set full --page raw --csv` export: a persistent grid that is not a
whole number of resident waves pays a full wave time for its last, partly filled wave.
usage: ncu_waves.py raw.csv[.gz] [min_ms]"""
import csv
import gzip
import math
import re
import sys

path = sys.argv[1]
min_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.03
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(fh) if r]
while rows and rows[0][0] != "ID":
    rows.pop(0)
head, units = rows[0], rows[1]


def col(name):
    return head.index(name)


limits = [col(f"launch__occupancy_limit_{k}") for k in ("registers", "shared_mem", "warps", "blocks")]
c_grid, c_dur, c_name, c_sm = col("launch__grid_size"), col("gpu__time_duration.sum"), col("Kernel Name"), col("launch__sm_count") if "launch__sm_count" in head else None
seen = set()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[c_name])[-60:]
    if name in seen:
        continue
    seen.add(name)
    dur = float(r[c_dur]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[c_dur], 1.0)
    if dur < min_ms:
        continue
    sms = int(float(r[c_sm])) if c_sm is not None else 148
    grid, occ = float(r[c_grid]), min(float(r[i]) for i in limits)
    waves = grid / (sms * occ)
    print(f"{name:60s} grid {int(grid):6d}  {int(occ):2d} CTAs/SM  {waves:5.2f} waves  tail x{math.ceil(waves - 1e-9) / waves:4.2f}  {dur:6.3f} ms")
