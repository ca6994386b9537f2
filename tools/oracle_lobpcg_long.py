import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import oracle
N=int(sys.argv[1]); steps=[int(s) for s in sys.argv[2:]]
O=oracle.laplace_matrix(np.float32,N,3,base=1)
X0=np.random.default_rng(1234321).random((O.n,16),dtype=np.float32)
for st in steps:
    t=time.time()
    try:
        r=oracle.lobpcg(O,False,X0,maxiter=st,tol=0.0)
        print(N,st,'lam0',r.lam[0],'resnorm max',np.max(r.residual_norms),'min',np.min(r.residual_norms),f'{time.time()-t:.1f}s',flush=True)
    except Exception as e:
        print(N,st,'EXC',repr(e)[:100],flush=True)
