#!/usr/bin/env python
"""Counts generic (LD / ST / ATOM / RED), global (LDG / STG / ATOMG / REDG) and local (LDL / STL)
memory instructions per kernel in the built objects (cuobjdump -sass sperr_b200/build/*.cu.o).
A kernel is listed when it still has generic accesses; --all lists every kernel. No GPU needed.
Generic accesses of a descriptor pointer are what DESIGN.md section 9 (first bullet) is about."""
import subprocess,re,sys,collections,glob
for o in sorted(glob.glob('sperr_b200/build/*.cu.o')):
    out=subprocess.run(['cuobjdump','-sass',o],capture_output=True,text=True).stdout
    fn=None; cnt=collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m=re.search(r'Function : (\S+)',line)
        if m: fn=m.group(1); continue
        m=re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)',line)
        if m and fn:
            op=m.group(1)
            if op in('LD','ST','LDG','STG','ATOM','ATOMG','RED','REDG','LDL','STL'): cnt[fn][op]+=1
    for fn,c in cnt.items():
        if 'cub' in fn or 'thrust' in fn: continue
        if c['LD']+c['ST']+c['ATOM']+c['RED']==0 and '--all' not in sys.argv: continue
        name=subprocess.run(['c++filt',fn],capture_output=True,text=True).stdout.strip()[:70]
        print(o.split('/')[-1][:-5].ljust(14), name.ljust(72), dict(c))
