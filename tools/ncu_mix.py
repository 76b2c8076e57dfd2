"""Summarise an `ncu --page raw --csv` + `--page source --csv` pair: per-kernel headline metrics and SASS opcode mix."""
import csv, re, collections, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(open(raw)))
cols = list(rows[0].keys())
stall = [c for c in cols if c.startswith('smsp__average_warps_issue_stalled') and c.endswith('per_issue_active.ratio')]
def f(r, k):
    try: return float(r[k].replace(',', ''))
    except Exception: return float('nan')
for r in rows[1:]:
    st = sorted(((f(r, k), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in stall), reverse=True)[:6]
    print(re.sub(r'\(.*', '', r['Kernel Name'])[-40:], r['Grid Size'], f"t={f(r,'gpu__time_duration.sum'):.3f}ms tensor%={f(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} warps%={f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} issue%={f(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} inst={f(r,'smsp__inst_executed.sum'):.3e} regs={r['launch__registers_per_thread']} dram={(f(r,'dram__bytes_read.sum')+f(r,'dram__bytes_write.sum')):.3e}{r and ''} |", ", ".join(f"{n}:{v:.2f}" for v, n in st))
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if len(r) > 2 and r[0] == 'Address' and r[1] == 'Source']
seen = set()
for k, start in enumerate(hdr):
    H = rows[start]; cs = H.index('Source'); ce = H.index('Instructions Executed'); cm = H.index('# Samples')
    data = []
    for r in rows[start + 1:]:
        if len(r) <= ce or r[0] == 'Address': break
        try: e = int(float(r[ce].replace(',', '') or 0)); sm = int(float(r[cm].replace(',', '') or 0))
        except Exception: continue
        data.append((e, sm, r[cs]))
    tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
    if tot in seen: continue
    seen.add(tot)
    op = collections.Counter(); ops = collections.Counter()
    for e, sm, s in data:
        m = re.sub(r'^@!?U?P\d+\s+', '', s.strip()).split()[0].split('.')[0] if s.strip() else '?'
        op[m] += e; ops[m] += sm
    print(f"--- kernel {k}: total inst {tot:.3e}")
    print("   " + "  ".join(f"{a}:{v/tot*100:.1f}%({ops[a]/max(ts,1)*100:.0f}%s)" for a, v in op.most_common(18)))
