import numpy as np, sys
sys.path.insert(0,'.')
from hicpeaks_b200.synth import synth_chromosome
n=6000; band=500
inp=synth_chromosome(n,band,5,maxww=10,seed=17)
num=inp['num']
M=np.zeros((n,n+2*num),dtype=np.int32)  # M[r, c] stored as row r, col offset c-r+num? use dense (r, d)
raw=np.zeros((num,n),dtype=np.int64)
for d in range(num): raw[d,:n-d]=inp['Diags'][d]
# dense matrix band as dict: X[r, c] = raw[c-r, r]
def shift(a,b):
    out=np.zeros_like(raw)
    dd=b-a
    dlo,dhi=max(0,-dd),min(num,num-dd); rlo,rhi=max(0,-a),min(n,n-a)
    out[dlo:dhi,rlo:rhi]=raw[dlo+dd:dhi+dd,rlo+a:rhi+a]
    return out
p=2
reads=np.zeros_like(raw); lvl=np.full(raw.shape,99)
pix=(raw!=0); pix[:5]=False; pix[band+1:]=False
for w in range(5,11):
    for a in range(1,w+1):
        for b in range(-w,0):
            g=max(a,-b)
            if g<=p: continue
            if w>5 and g<w: continue
            reads+=shift(a,b)
    ok=(reads>=16)&(lvl==99)&pix
    lvl[ok]=w
    print(w, ok.sum()/pix.sum())
lvl[pix&(lvl==99)]=8   # frozen at 8: never-resolved need up to 8 anyway? (they run to last step)
lvl=np.where(pix,np.minimum(lvl,8),0)
cost={0:0,5:105,6:160,7:225,8:300}
costarr=np.vectorize(cost.get)(lvl)
ideal=costarr.sum()
dense=(5<=np.arange(num))[:,None]&(np.arange(num)<=band)[:,None]
# warp-task shapes: (rows R consecutive, diagonals D consecutive) approx
for R,D in [(128,2),(128,4),(32,1),(32,4),(16,8),(4,2),(4,4),(8,8)]:
    tot=0
    L=lvl[5:band+1]
    nd=(L.shape[0]//D)*D; nr=(n//R)*R
    blk=L[:nd,:nr].reshape(nd//D,D,nr//R,R).max(axis=(1,3))
    tot=np.vectorize(cost.get)(blk).sum()*R*D
    print(f"task {R}x{D}: work/ideal = {tot/ideal:.2f}")
