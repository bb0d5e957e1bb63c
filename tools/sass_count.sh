#!/bin/bash
# static instruction count of the decoder's VN and CN phases (between barriers) for bp_regular_kernel<6,3,384,MINB>
SO=${1:-kmldpc_b200/lib/libkmldpc_b200.so}; MINB=${2:-3}; SUF=${3:-ELb1ELi2E}
cuobjdump -sass $SO 2>/dev/null | awk -v pat="bp_regular_kernelILi6ELi3ELi384ELi${MINB}${SUF}" '/Function : /{p = index($0, pat) > 0} p{print}' | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+//; s/\s*\/\*.*$//' > /tmp/bp_sass.txt
python3 - <<'PY'
import collections
L=[l.strip() for l in open('/tmp/bp_sass.txt')]
bars=[i for i,l in enumerate(L) if l.startswith('BAR')]
print("total", len(L), "bars at", bars)
def hist(a,b,name):
    c=collections.Counter()
    for l in L[a:b]:
        t=l.split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]; c[op]+=1
    n=b-a
    print(f"{name}: {n} instr;", " ".join(f"{k}:{v}" for k,v in c.most_common(16)))
    return n
if len(bars)>=4:
    vn=hist(bars[1]+1,bars[2],"VN phase (18 edges)")
    cn=hist(bars[2]+1,bars[3],"CN phase (18 edges)")
    print("per edge-iteration:", (vn+cn)/18)
PY
