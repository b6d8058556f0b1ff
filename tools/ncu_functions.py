"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by the FUNCTION of a .cu file the lines belong to.
usage: python tools/ncu_functions.py dump.csv [path/to/file.cu]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == 'Line No')
isamp = hdr.index('# Samples'); iins = hdr.index('Instructions Executed')
import os
src=open(sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'partitionedls.jl_b200', 'csrc', 'nnls5.cu')).read().splitlines()
# function start lines
import re
funcs=[]
for i,l in enumerate(src,1):
    m=re.match(r'^(?:template.*\n)?__device__ .*? (\w+)\(', l) or re.match(r'^__global__ .*? (\w+)\(', l)
    if m: funcs.append((i,m.group(1)))
def fn(line):
    name='?'
    for s,n in funcs:
        if s<=line: name=n
    return name
agg=collections.defaultdict(lambda:[0,0]); cur=None
for r in rows:
    if not r or r[0] in ('File Path','Function Name','Line No'): continue
    if r[0]!='':
        cur=r[0]; continue
    if len(r)<len(hdr) or cur is None: continue
    try: s=int(r[isamp]); n=int(r[iins])
    except ValueError: continue
    try: ln=int(cur)
    except: continue
    # only lines of nnls5.cu are meaningful; header-inlined lines get attributed by number anyway
    a=agg[fn(ln)]; a[0]+=s; a[1]+=n
ts=sum(a[0] for a in agg.values()); ti=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(f"{k:20s} {100*a[0]/ts:5.1f}% smp {100*a[1]/ti:5.1f}% ins")
