"""numpy emulation of `emit_long_kernel` (csrc/shuffle.cu) with the kernel's own window sizing (1.40 c + 1024 words, doubled
when it turns out too short), flag encoding (bit 0 surely accepted, bit 1 surely rejected), convergence test (U = L up to
and including the epoch's last accept) and no-progress test (totals unchanged) — checked against numpy's shuffle
(permutations and stream position) for n = 16 385 ... 65 536 before the kernel first ran on a GPU. The block-wide scans
are plain cumulative sums here; the per-thread segmentation of the kernel does not change any value."""
import numpy as np, sys
LW_MAX=65536
def excl(x):
    c=np.cumsum(x,dtype=np.int64); return c-x
def emit_long(words, n, P):
    n_words=len(words); pos=0; jout=np.zeros(P*(n-1),dtype=np.int64); step_base=0; rounds_hist=[]
    for p in range(P):
        i_hi=n-1
        while i_hi>=1:
            b=int(i_hi).bit_length(); M=(1<<b)-1; i_lo=max(1<<(b-1),1); c=i_hi-i_lo+1
            W=min(LW_MAX,int(np.float32(1.40)*np.float32(c))+1024)
            done=False
            while not done:
                if pos+W>n_words: W=n_words-pos
                if W<c: return None,-1,rounds_hist
                v=(words[pos:pos+W]&np.uint32(M)).astype(np.int64)
                f=np.where(v<=i_hi,0,2).astype(np.int64)
                prevA=prevU=-1; rnd=0
                while True:
                    a=f&1; u=(f>>1)^1
                    bA=excl(a); bU=excl(u); totA=int(a.sum()); totU=int(u.sum())
                    done_at=-1; conv=0
                    if totA>=c:
                        inclA=bA+a
                        done_at=int(np.searchsorted(inclA,c))  # first t with inclA>=c ; need a[t]==1 there
                        assert a[done_at]==1 and inclA[done_at]==c
                        conv = int((bU[done_at]+u[done_at])==c)
                    if done_at>=0 and conv:
                        acc=np.nonzero(a[:done_at+1])[0]
                        assert len(acc)==c
                        jout[step_base+(n-1-i_hi):step_base+(n-1-i_hi)+c]=v[acc]
                        pos+=done_at+1; done=True; rounds_hist.append(rnd); break
                    if totA==prevA and totU==prevU:
                        if W>=LW_MAX or pos+W>=n_words: return None,(-3 if W>=LW_MAX else -1),rounds_hist
                        W=min(LW_MAX,2*W); break
                    prevA,prevU=totA,totU
                    nf=(v<=i_hi-bU).astype(np.int64)|((v>i_hi-bA).astype(np.int64)<<1)
                    f=nf; rnd+=1
            i_hi=i_lo-1
        step_base+=n-1
    return jout,pos,rounds_hist

for n,P in ((50000,3),(20000,5),(16385,4),(65536,2),(40000,3)):
    seed=7
    ref=np.random.RandomState(seed); raw=np.random.RandomState(seed)
    words=raw._bit_generator.random_raw(int(P*n*2.0)+100000).astype(np.uint32)
    jout,pos,rh=emit_long(words,n,P)
    assert jout is not None, pos
    x=np.arange(n); xr=np.arange(n)
    ok=True
    for p in range(P):
        ref.shuffle(xr)
        js=jout[p*(n-1):(p+1)*(n-1)]
        for k,i in enumerate(range(n-1,0,-1)):
            j=js[k]; x[i],x[j]=x[j],x[i]
        ok=ok and np.array_equal(x,xr)
    chk=np.random.RandomState(seed); chk._bit_generator.random_raw(pos)
    print(n,P,ok, chk.randint(0,1<<30)==ref.randint(0,1<<30), 'max rounds',max(rh),'mean',np.mean(rh))
# short stream -> -1
words=np.random.RandomState(1)._bit_generator.random_raw(30000).astype(np.uint32)
print('short stream ->', emit_long(words,20000,2)[1])
