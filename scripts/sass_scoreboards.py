import re, sys, subprocess
so, pat = sys.argv[1], sys.argv[2]
txt=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout.split('\n')
# isolate function
out=[];f=False
for l in txt:
    if 'Function :' in l:
        f = pat in l
    if f: out.append(l)
ins=[];i=0
while i<len(out):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/',out[i])
    if m and i+1<len(out):
        m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',out[i+1])
        if m2:
            hi=int(m2.group(1),16); ctrl=(hi>>41)&0x7fffff
            ins.append((int(m.group(1),16),m.group(2).strip(),ctrl&0xf,(ctrl>>5)&7,(ctrl>>8)&7,(ctrl>>11)&0x3f))
            i+=2; continue
    i+=1
print(len(ins),"instructions")
import collections
sb=collections.Counter(a[3] for a in ins if a[1].split()[0].startswith("LDG") or (a[1].startswith("@") and a[1].split()[1].startswith("LDG")))
print("LDG write barriers:",dict(sb))
# print sequence of events: LDG batches (compressed) and waits on those barriers
bars=set(sb.keys())
last=None;cnt=0
for k,a in enumerate(ins):
    op=a[1].split()[1] if a[1].startswith("@") else a[1].split()[0]
    isld=op.startswith("LDG")
    waits=[b for b in bars if b<6 and (a[5]>>b)&1]
    if isld:
        cnt+=1; continue
    if cnt:
        print(f"   ... {cnt} LDG issued (up to idx {k-1})"); cnt=0
    if waits:
        print(f"WAIT{waits} idx {k} {a[0]:05x} {a[1][:60]}")
print("---- compact timeline")
events=[]
for k,a in enumerate(ins):
    op=a[1].split()[1] if a[1].startswith("@") else a[1].split()[0]
    if op.startswith("LDG"): events.append(("L",k))
    elif (a[5]>>5)&1: events.append(("W",k,a[1][:50]))
    elif op in ("BRA","EXIT","CALL.REL.NOINC","RET.REL.NODEC","WARPSYNC","BSYNC"): pass
i=0
while i<len(events):
    e=events[i]
    if e[0]=="L":
        j=i
        while j+1<len(events) and events[j+1][0]=="L": j+=1
        print(f"  LDG x{j-i+1:3d}  idx {e[1]}..{events[j][1]}")
        i=j+1
    else:
        print(f"  WAIT5      idx {e[1]}  {e[2]}")
        i+=1
