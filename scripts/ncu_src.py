"""Summarise an `ncu --page source --csv` dump: hot instructions and the share of stall samples between barriers."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
hdr = rows[1]
i_src = hdr.index('Source'); i_s = hdr.index('Warp Stall Sampling (All Samples)'); i_ex = hdr.index('Instructions Executed')
i_wf = hdr.index('L1 Wavefronts Shared'); i_wfi = hdr.index('L1 Wavefronts Shared Ideal')
def I(x):
    try: return int(x)
    except ValueError: return 0
data = [r for r in rows[2:] if len(r) > i_wfi]
tot = sum(I(r[i_s]) for r in data)
print('total samples', tot, 'ninstr', len(data))
for k, r in enumerate(data):
    s = I(r[i_s])
    if s > tot * thr:
        print(k, r[i_src].strip()[:70], s, r[i_ex], r[i_wf], r[i_wfi])
print('--- regions between barriers')
cum = last = ex = wf = 0
for k, r in enumerate(data):
    cum += I(r[i_s]); ex += I(r[i_ex]); wf += I(r[i_wf])
    if 'BAR' in r[i_src] or 'EXIT' in r[i_src]:
        print(k, r[i_src].strip()[:40], 'samples', cum - last, f'{(cum-last)/max(tot,1):.3f}', 'warp-inst', ex, 'smem wf', wf)
        last = cum; ex = 0; wf = 0
