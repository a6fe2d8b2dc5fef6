import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
from sparsergps_b200 import _lib as L
from sparsergps_b200.context import Context
ctx=Context(0)
for m in (256,384,512):
    rng=np.random.default_rng(m)
    X=rng.normal(size=(m,8))
    D=((X[:,None,:]-X[None,:,:])**2).sum(-1)
    A=np.asfortranarray(np.exp(-0.5*D)+1e-2*np.eye(m)+0.1*(X@X.T)/8)
    Lo=np.empty((m,m),order='F'); Ai=np.empty((m,m),order='F')
    logdet,info=L.cd(),L.ci()
    st=ctx._lib.srgp_test_chol_inverse(ctx.handle,m,L.ptr(A),L.ptr(Lo),L.ptr(Ai),C.byref(logdet),C.byref(info),0,None)
    Lref=np.linalg.cholesky(A)
    print("m",m,"st",st,"info",info.value)
    nb=(m+127)//128
    for bi in range(nb):
        for bj in range(bi+1):
            blk=np.s_[bi*128:(bi+1)*128, bj*128:(bj+1)*128]
            a=np.tril(Lo)[blk]; b=Lref[blk]
            err=np.nanmax(np.abs(a-b)) if np.isfinite(a).any() else np.nan
            print("  block",bi,bj,"maxerr %.3e"%err, "nan" if np.isnan(a).any() else "")
