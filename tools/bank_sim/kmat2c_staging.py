import itertools
pairs=[(i,j) for i in range(3) for j in range(i,3)]
def wf(addrs):
    # addrs: list of (lane, addr in doubles); 64-bit access: two half-warps, 16 banks of 8B
    tot=0
    for half in (0,1):
        banks={}
        for lane,a in addrs:
            if lane//16==half:
                banks.setdefault(a%16,set()).add(a)
        tot+=max((len(s) for s in banks.values()),default=0)
    return tot
def sim(ELSM,RS,mapping):
    total=0
    for w in range(3):
        for a in range(8):
            for bl in range(4):
                for mirror in (0,1):
                    addrs=[]
                    for lane in range(32):
                        tid=w*32+lane
                        el=tid//12; r=tid%12
                        if mapping=='th': t=r//2; h=r%2
                        else: t=r%6; h=r//6
                        d1,d2=pairs[t]; b=4*h+bl
                        if mirror:
                            if d1==d2: continue
                            addr=el*ELSM+(b*3+d2)*RS+a*3+d1
                        else:
                            addr=el*ELSM+(a*3+d1)*RS+b*3+d2
                        addrs.append((lane,addr))
                    total+=wf(addrs)
    return total
for ELSM in (658,660,662,666,670):
  for RS in (24,25):
    for m in ('th','ht'):
        print(ELSM,RS,m,sim(ELSM,RS,m))
