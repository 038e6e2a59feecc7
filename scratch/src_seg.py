import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; data = rows[2:]
ix = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ix[k]])
    except: return 0.0
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
seg = []; cur = dict(start=0, n=0, ex=0, samp=0, wf=0, wfx=0, st=collections.Counter(), ops=collections.Counter())
TILES = float(sys.argv[2]) if len(sys.argv) > 2 else 48128.0
for i, r in enumerate(data):
    s = r[ix['Source']].strip()
    op = s.split()[0] if not s.startswith('@') else s.split()[1]
    cur['n'] += 1; cur['ex'] += f(r, 'Instructions Executed'); cur['samp'] += f(r, '# Samples')
    cur['wf'] += f(r, 'L1 Wavefronts Shared'); cur['wfx'] += f(r, 'L1 Wavefronts Shared Excessive')
    for k in stalls: cur['st'][k] += f(r, k)
    cur['ops'][op.split('.')[0]] += f(r, 'Instructions Executed')
    if op.startswith('BAR') or op.startswith('BRA') and False:
        cur['end'] = i; seg.append(cur); cur = dict(start=i + 1, n=0, ex=0, samp=0, wf=0, wfx=0, st=collections.Counter(), ops=collections.Counter())
cur['end'] = len(data) - 1; seg.append(cur)
tot = sum(s['samp'] for s in seg)
for s in seg:
    if s['samp'] < 0.002 * tot: continue
    top = ', '.join(f"{k[6:]}={v / s['samp']:.2f}" for k, v in s['st'].most_common(6))
    ops = ', '.join(f"{k}={v / TILES:.0f}" for k, v in s['ops'].most_common(10))
    print(f"seg [{s['start']}-{s['end']}] sass={s['n']} exec/tile={s['ex'] / TILES:.0f} samples={100 * s['samp'] / tot:.1f}% wf/tile={s['wf'] / TILES:.0f} (excess {s['wfx'] / TILES:.0f})\n   stalls: {top}\n   ops/tile: {ops}")
