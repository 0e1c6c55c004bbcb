import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/conservation-fem_b200")
import numpy as np
from cfem_b200 import Context, meshes, _lib as L
n=int(sys.argv[1]) if len(sys.argv)>1 else 1024
x,c=meshes.rectangle(n,n)
ctx=Context((x,c))
v=np.random.default_rng(0).normal(size=ctx.n)
ctx.state_set(uh=v,u_n=v,u_old=v,u_oo=v,RH=v,h=np.abs(v)+0.1)
for name,k in (("spmv",0),("asm_res",1),("asm_jac",2),("eps",3),("rv_rhs",4)):
    for flux in ((1,2) if k else (1,)):
        ms,by=ctx.time_kernel(k,flux,50)
        print(f"{os.environ.get('CFEM_SPMV','stream'):8s} n={n} {name:8s} flux={flux} {ms*1e3:8.1f} us  {by/ms/1e6:8.1f} GB/s algorithmic")
