"""Per-source-line instruction / stall-sample totals from `ncu --page source --print-source cuda,sass --csv`.
usage: ncu_lines.py file.csv kernel-substring [top]"""
import csv, sys, collections
path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
cur_file, cur_fn, agg, active = None, None, collections.OrderedDict(), False
tot_i = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; active = want in cur_fn; continue
    if r[0] == "Line No": H = r; continue
    if not active or r[0] in ("",): continue
    try:
        ln = int(r[0]); ie = int(r[7]); sm = int(r[6])
    except Exception: continue
    key = (cur_file.split("/")[-1], ln, r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0]); a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
print(f"total inst {tot_i:.3e} samples {tot_s}")
for (f, ln, src), (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ie/tot_i*100:5.1f}%i {sm/max(tot_s,1)*100:5.1f}%s  {f}:{ln}  {src}")
