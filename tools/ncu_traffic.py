"""Per-kernel DRAM traffic / duration summary of an `ncu --set full` capture, as JSON for bench.py's roofline.traffic.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv         (or a .csv.gz committed under profiles/)
    python tools/ncu_traffic.py raw.csv[.gz] profiles/r02_ncu_traffic.json [--commit <sha>]

Output: {"source": <csv>, "commit": <sha>, "kernels": {<kernel name>: {"launches": n, "dram_bytes": mean bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), "dram_read": ..., "dram_write": ..., "duration_us": mean gpu__time_duration,
"regs": launch__registers_per_thread, "tensor_pct": sm__pipe_tensor..., "grid": ...}}}.  Kernel names are the demangled
function names without their argument lists; template arguments are kept (they select the window shape).
"""
import csv
import gzip
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short_name(full: str) -> str:
    name = re.sub(r"\((int|bool|unsigned|long)\)", "", full)          # the source page prints template arguments with casts
    name = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()
    name = re.sub(r"\b\w+::", "", name)                                  # drop namespaces
    return re.sub(r"\s+", "", name)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else None
    fh = gzip.open(src, "rt") if src.endswith(".gz") else open(src)
    rows = [r for r in csv.reader(fh) if r]
    while rows and rows[0][0] != "ID":
        rows.pop(0)                                   # ncu banner lines
    head, units, body = rows[0], rows[1], rows[2:]

    def col(pattern):
        for i, h in enumerate(head):
            if re.search(pattern, h):
                return i
        return None

    c_name = head.index("Kernel Name")
    c_grid = head.index("Grid Size")
    cols = {"dram_read": col(r"dram__bytes_read\.sum$"), "dram_write": col(r"dram__bytes_write\.sum$"),
            "duration_us": col(r"gpu__time_duration\.sum$"), "regs": col(r"launch__registers_per_thread$"),
            "tensor_pct": col(r"sm__pipe_tensor_cycles_active.*pct_of_peak_sustained_(active|elapsed)$"),
            "warps_active_pct": col(r"sm__warps_active\.avg\.pct_of_peak_sustained_active$"),
            "issue_pct": col(r"sm__inst_issued.*pct_of_peak_sustained_active$|smsp__issue_active\.avg\.pct_of_peak_sustained_active$")}
    agg = {}
    for r in body:
        k = short_name(r[c_name])
        a = agg.setdefault(k, {"launches": 0, "grid": r[c_grid]})
        a["launches"] += 1
        for key, c in cols.items():
            if c is None or c >= len(r) or r[c] == "":
                continue
            try:
                v = float(r[c].replace(",", "")) * (UNIT.get(units[c], 1.0) if key in ("dram_read", "dram_write", "duration_us") else 1.0)
            except ValueError:
                continue
            a[key] = a.get(key, 0.0) + v
    for a in agg.values():
        n = a["launches"]
        for key in cols:
            if key in a:
                a[key] /= n
        if "dram_read" in a and "dram_write" in a:
            a["dram_bytes"] = a["dram_read"] + a["dram_write"]
    json.dump({"source": src, "commit": commit, "kernels": agg}, open(dst, "w"), indent=1, sort_keys=True)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1].get("duration_us", 0)):
        print(f"{k[:70]:70s} n={a['launches']:3d} {a.get('duration_us', 0):9.1f} us  {a.get('dram_bytes', 0) / 1e9:7.3f} GB  regs {a.get('regs', 0):.0f}")


if __name__ == "__main__":
    main()
