NP=6;EPW=5
pairs=[(i,j) for i in range(3) for j in range(i,3)]
pidx={pr:i for i,pr in enumerate(pairs)}
def wf64(addrs):
    tot=0
    for half in (0,1):
        banks={}
        for lane,a in addrs:
            if lane//16==half: banks.setdefault(a%16,set()).add(a)
        tot+=max((len(s) for s in banks.values()),default=0)
    return tot
def wf128(addrs):  # addr in doubles, even; quarter-warps, 8 units of 16B
    tot=0
    for qt in range(4):
        banks={}
        for lane,a in addrs:
            if lane//8==qt: banks.setdefault((a//2)%8,set()).add(a)
        tot+=max((len(s) for s in banks.values()),default=0)
    return tot
def sim(ELSM,BS,RSTR):
    s1=0
    for a in range(8):
        for c in range(4):
            addrs=[(lane,(lane//NP)*ELSM+(lane%NP)*BS+a*RSTR+2*c) for lane in range(30)]
            s1+=wf128(addrs)
    s2=0
    for a in range(8):
        for dr in range(3):
            addrs=[]
            for lane in range(24):
                k=lane//3;dc=lane%3
                if dr<=dc: ad=pidx[(dr,dc)]*BS+a*RSTR+k
                else: ad=pidx[(dc,dr)]*BS+k*RSTR+a
                addrs.append((lane,ad))
            s2+=wf64(addrs)
    return s1/EPW, s2
res=[]
for ELSM in range(682,712,2):
    for BS in range(80,100,2):
        for RSTR in (10,):
            if 5*BS+8*RSTR+24>648: continue
            a,b=sim(ELSM,BS,RSTR); res.append((a+b,ELSM,BS,RSTR,a,b))
res.sort(); print(res[:10]); print([r for r in res if r[1]==694][:5])
