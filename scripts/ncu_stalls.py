import csv, collections, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
for h in hdr_idx:
    hdr=rows[h]; name=rows[h-1][1][:90]
    body=[]
    for r in rows[h+1:]:
        if not r or r[0] in("Kernel Name","Address"): break
        body.append(r)
    cols={n:i for i,n in enumerate(hdr)}
    reasons=[n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
    tot=collections.Counter(); tab=collections.defaultdict(collections.Counter)
    for r in body:
        src=r[cols["Source"]].split()
        op=(src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        for n in reasons:
            v=int(r[cols[n]] or 0); tot[n]+=v; tab[op][n]+=v
    T=sum(tot.values())
    print(name, "instr:",len(body), "samples:",T)
    print("  ", {k.replace("stall_",""):f"{100*v/T:.1f}%" for k,v in tot.most_common(10)})
    for op in ["FADD","IADD3","LOP3","LDS","LDG","STS","STG","FMUL","FFMA"]:
        print("   ",op, {k.replace("stall_",""):v for k,v in tab[op].most_common(4)})
