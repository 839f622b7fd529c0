import itertools, sys
ND=3;NNPE=8;NF=3;NQ=8;NP=6;EPW=5;NDF=9
pairs=[(i,j) for i in range(3) for j in range(i,3)]
def sym(i,j):
    if i>j: i,j=j,i
    return i*NDF-(i*(i-1))//2+(j-i)
def wf(addrs):
    tot=0
    for half in (0,1):
        banks={}
        for lane,a in addrs:
            if lane//16==half: banks.setdefault(a%16,set()).add(a)
        tot+=max((len(s) for s in banks.values()),default=0)
    return tot
def sim(ELSM,SLOT,RS,aperm=None):
    # phase K loads per warp per q (q only shifts by SLOT*q: include all q)
    ldK=0; ideal=0
    for q in range(NQ):
        for j1 in range(3):
            for j2 in range(3):
                addrs=[]
                for lane in range(30):
                    el=lane//NP;t=lane%NP;d1,d2=pairs[t]
                    idx=sym(d1*3+j1,d2*3+j2)
                    if aperm: idx=aperm[idx]
                    addrs.append((lane,el*ELSM+q*SLOT+24+idx))
                ldK+=wf(addrs); ideal+=2
        for a in range(8):
            for k in range(3):
                addrs=[(lane,(lane//NP)*ELSM+q*SLOT+a*3+k) for lane in range(30)]
                ldK+=wf(addrs); ideal+=2  # broadcast: could be 1+1
    # phase G stores: round 1 lanes t=0..5 -> q=t ; round 2 q=t+6 for t<2
    stG=0
    for rnd in (0,1):
        for idx in range(78):
            addrs=[]
            for lane in range(30):
                el=lane//NP;t=lane%NP;q=t+6*rnd
                if q>=NQ: continue
                addrs.append((lane,el*ELSM+q*SLOT+idx))
            stG+=wf(addrs)
    # S1 stores
    st1=0
    for a in range(8):
        for b in range(8):
            for mirror in (0,1):
                addrs=[]
                for lane in range(30):
                    el=lane//NP;t=lane%NP;d1,d2=pairs[t]
                    if mirror:
                        if d1==d2: continue
                        addrs.append((lane,el*ELSM+(b*3+d2)*RS+a*3+d1))
                    else: addrs.append((lane,el*ELSM+(a*3+d1)*RS+b*3+d2))
                st1+=wf(addrs)
    return ldK/EPW, stG/EPW, st1/EPW
print('base  ELSM=694 SLOT=81 RS=24 ->', sim(694,81,24))
best=[]
for ELSM in range(682,712,2):
    for SLOT in range(78,84):
        if SLOT*8>ELSM-34: continue
        r=sim(ELSM,SLOT,24)
        best.append((sum(r),ELSM,SLOT,r))
best.sort()
for b in best[:8]: print(b)

print("---- no-mirror staging")
def sim2(ELSM,RS):
    st1=0
    for a in range(8):
        for b in range(8):
            addrs=[]
            for lane in range(30):
                el=lane//NP;t=lane%NP;d1,d2=pairs[t]
                addrs.append((lane,el*ELSM+(a*3+d1)*RS+b*3+d2))
            st1+=wf(addrs)
    ld2=0
    for row in range(24):
        dr=row%3
        addrs=[]
        for lane in range(24):
            dc=lane%3
            addrs.append((lane, row*RS+lane if dr<=dc else lane*RS+row))
        ld2+=wf(addrs)
    return st1/EPW, ld2   # ld2 per element
for RS in (24,25,26):
    for ELSM in range(682,712,2):
        print(RS,ELSM,sim2(ELSM,RS))
