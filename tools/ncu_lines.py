"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: python tools/ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == 'Line No')
ia = hdr.index('Address'); isamp = hdr.index('# Samples'); iins = hdr.index('Instructions Executed')
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
cur = None; src = {}
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows:
    if not r or r[0] in ('File Path', 'Function Name', 'Line No'): continue
    if r[0] != '':
        cur = r[0]
        # source text may contain commas that split the row: join until the address column pattern
        src[cur] = ','.join(r[1:len(r) - (len(hdr) - 2)]) if len(r) > len(hdr) else r[1]
        continue
    if len(r) < len(hdr) or cur is None: continue
    try:
        s = int(r[isamp]); n = int(r[iins])
    except ValueError:
        continue
    a = agg[cur]; a[0] += s; a[1] += n
    for h, i in stall_cols.items():
        try: a[2][h] += int(r[i])
        except ValueError: pass
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}  instructions {tot_i}")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join(f"{k[6:]}={v}" for k, v in a[2].most_common(3))
    print(f"{100*a[0]/max(tot_s,1):5.1f}% smp {100*a[1]/max(tot_i,1):5.1f}% ins  L{line:>4}  {src.get(line,'')[:80]:80s} {st}")
