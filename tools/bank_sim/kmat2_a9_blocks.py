import itertools
NP=6;EPW=5;NDF=9;NQ=8
pairs=[(i,j) for i in range(3) for j in range(i,3)]
def wf(addrs):
    tot=0
    for half in (0,1):
        banks={}
        for lane,a in addrs:
            if lane//16==half: banks.setdefault(a%16,set()).add(a)
        tot+=max((len(s) for s in banks.values()),default=0)
    return tot
def local(p,j1,j2):
    d1,d2=pairs[p]
    if d1==d2:
        a,b=min(j1,j2),max(j1,j2)
        return a*3-(a*(a-1))//2+(b-a)
    return j1*3+j2
size=[6 if pairs[p][0]==pairs[p][1] else 9 for p in range(6)]
def cost(order,ELSM,SLOT):
    base={};o=0
    for p in order: base[p]=o;o+=size[p]
    tot=0
    for q in range(NQ):
      for j1 in range(3):
        for j2 in range(3):
            addrs=[(lane,(lane//NP)*ELSM+q*SLOT+24+base[lane%NP]+local(lane%NP,j1,j2)) for lane in range(30)]
            tot+=wf(addrs)
    return tot/EPW
res=[]
for order in itertools.permutations(range(6)):
    for ELSM in (694,):
        for SLOT in (81,):
            res.append((cost(order,ELSM,SLOT),order,ELSM,SLOT))
res.sort()
print(res[:5]); print('ideal',9*2*8/5)
res=[]
for order in itertools.permutations(range(6)):
    for ELSM in range(682,712,2):
        for SLOT in (79,81,83):
            if SLOT*8+34>ELSM: continue
            res.append((cost(order,ELSM,SLOT),order,ELSM,SLOT))
res.sort()
print(res[:5])
