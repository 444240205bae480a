import sys, time, ctypes
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import saamge_b200 as sab
n=int(sys.argv[1]); levels=int(sys.argv[2])
p=sab.default_params(num_levels=levels, first_elems_per_agg=52, elems_per_agg=64, partition_kind=2, block=(16,16,16))
t=time.time(); pr=sab.Problem(3,n,order=2,coef_kind=1); na=pr.partition(p); print("inputs %.1fs AEs %d ND %d"%(time.time()-t,na,pr.scalar("ND")),flush=True)
I=pr.get("AE_to_dof.I"); nn=np.diff(I); print("AE n: min %d mean %.0f max %d"%(nn.min(),nn.mean(),nn.max()))
h=sab.host_lib(); h.sa_drv_gpu_profile.argtypes=[ctypes.c_int,ctypes.c_char_p,ctypes.c_int]
t=time.time(); H=sab.ml_build(pr,p); print("ml_build %.2fs"%(time.time()-t),flush=True)
buf=ctypes.create_string_buffer(8192); h.sa_drv_gpu_profile(-1,buf,8192); print("PROF", buf.value.decode().replace("\n","; "))
print({k:round(v,3) for k,v in H.times().items() if v>0.01})
t=time.time(); it=sab.ml_pcg(H); print("pcg iters",it,"%.3fs"%(time.time()-t),"res",H.scalar("pcg.final_res_norm"))
