"""Summarise an `ncu --page source --csv` export: instructions / smem wavefronts / stall samples per opcode
and per code region (regions are delimited by BAR.SYNC). Usage: ncu_source_summary.py file.csv [rows]"""
import csv, sys, re
from collections import defaultdict
path = sys.argv[1]; rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    try: return float(r[ix[name]].replace(',', ''))
    except Exception: return 0.0
ops = defaultdict(lambda: [0, 0, 0, 0])
regions = []; cur = dict(n=0, inst=0, wf=0, wf_ideal=0, samples=0, start=0)
stall_names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls = defaultdict(float)
for k, r in enumerate(rows[2:]):
    src = r[ix['Source']].strip()
    op = re.sub(r'^@!?U?P\d+\s+', '', src).split()[0] if src else '?'
    inst = num(r, 'Instructions Executed'); wf = num(r, 'L1 Wavefronts Shared'); wfi = num(r, 'L1 Wavefronts Shared Ideal'); smp = num(r, '# Samples')
    base = op.split('.')[0]
    key = base if base not in ('LDS', 'STS', 'LDG', 'STG') else '.'.join(op.split('.')[:3])
    o = ops[key]; o[0] += inst; o[1] += wf; o[2] += smp; o[3] += wfi
    cur['inst'] += inst; cur['wf'] += wf; cur['wf_ideal'] += wfi; cur['samples'] += smp; cur['n'] += 1
    for s in stall_names: stalls[s] += num(r, s)
    if base in ('BAR',) or src.startswith('SYNCS') :
        cur['end'] = k; cur['endop'] = src[:40]; regions.append(cur); cur = dict(n=0, inst=0, wf=0, wf_ideal=0, samples=0, start=k + 1)
cur['end'] = len(rows) - 3; cur['endop'] = 'END'; regions.append(cur)
tot_inst = sum(o[0] for o in ops.values()); tot_smp = sum(o[2] for o in ops.values())
print(f"total warp-instructions {tot_inst:.0f}  per row {tot_inst/rows_n:.0f}; samples {tot_smp:.0f}")
print("-- by opcode (inst/row, smem wavefronts/row (ideal), % of samples)")
for k, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"  {k:22s} {o[0]/rows_n:9.1f} {o[1]/rows_n:9.1f} ({o[3]/rows_n:7.1f}) {100*o[2]/max(tot_smp,1):6.1f}%")
print("-- by region (between barriers): sass lines, inst/row, wavefronts/row (ideal), %samples, ends with")
for rg in regions:
    if rg['inst'] / rows_n < 5 and rg['samples'] < 0.005 * tot_smp: continue
    print(f"  [{rg['start']:5d}-{rg['end']:5d}] {rg['inst']/rows_n:9.1f} {rg['wf']/rows_n:9.1f} ({rg['wf_ideal']/rows_n:7.1f}) {100*rg['samples']/max(tot_smp,1):6.1f}%  {rg['endop']}")
print("-- stall reasons (% of samples)")
ts = sum(stalls.values())
for s, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {100*v/max(ts,1):6.1f}%")
