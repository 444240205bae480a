import sys, time
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import saamge_b200 as sab
n=int(sys.argv[1]); levels=int(sys.argv[2]); epa=int(sys.argv[3]); nupro=int(sys.argv[4]) if len(sys.argv)>4 else 0
p=sab.default_params(num_levels=levels, first_elems_per_agg=52, elems_per_agg=epa, partition_kind=2, block=(32,32,32), first_nu_pro=nupro, nu_pro=nupro)
t=time.time(); pr=sab.Problem(3,n,coef_kind=1); na=pr.partition(p); print("host inputs %.1fs AEs %d"%(time.time()-t,na), flush=True)
import ctypes
h=sab.host_lib(); h.sa_drv_gpu_profile.argtypes=[ctypes.c_int,ctypes.c_char_p,ctypes.c_int]
h.sa_drv_problem_pin.restype=ctypes.c_double; h.sa_drv_problem_pin.argtypes=[ctypes.c_void_p,ctypes.c_int]
if not os.environ.get('SA_NO_PIN'): print('pin %.3fs'%h.sa_drv_problem_pin(pr.handle,0))
if os.environ.get('SA_PROBE_PROFILE'): h.sa_drv_gpu_profile(1,None,0)
t=time.time(); H=sab.ml_build(pr,p); print("ml_build %.2fs"%(time.time()-t), flush=True)
buf=ctypes.create_string_buffer(8192); h.sa_drv_gpu_profile(-1,buf,8192); print("PROF", buf.value.decode().replace("\n","; "))
g=sab.gpu_lib(); clk=(ctypes.c_double*8)(); g.sa_gpu_debug_phase_clocks(clk); print("PHASE Mcycles", [round(x/1e6,1) for x in clk])
ts=(ctypes.c_double*16)(); g.sa_gpu_debug_ts_clocks(ts); print("TS Mcycles qr,gram,symm,Xtf,S,Z,syr2k,band | s2 ticks, s2 Mcyc, s2 steps:", [round(x/1e6,1) for x in ts[:8]], ts[8], round(ts[9]/1e6,1), ts[10])
tm=H.times(); print({k:round(v,3) for k,v in tm.items()})
for l in range(levels-1):
    print("level",l,"ND",H.scalar("ND",l),"nparts",H.scalar("nparts",l),"mises",H.scalar("num_mises",l))
t=time.time(); it=sab.ml_pcg(H); print("pcg iters",it,"%.3fs"%(time.time()-t),"res",H.scalar("pcg.final_res_norm"))
# device-scalar PCG of dist.cu on one rank (no exchange) + operator sizes per level
from saamge_b200.dist_solve import DistSolver
S=DistSolver(H,None); b=pr.get("b"); S.pcg(b,maxiter=2)
t=time.time(); x,it2,brr=S.pcg(b); print("dist-path pcg (1 rank) iters",it2,"wall %.3fs device %.4fs -> %.2f ms/iteration"%(time.time()-t,S.solve_seconds,1e3*S.solve_seconds/max(1,abs(it2))))
print("levels (rows, nnzA, nnzP)", S.level_info(), "bytes/iteration %.3e -> %.0f GB/s"%(S.bytes_per_iteration(), S.bytes_per_iteration()*abs(it2)/S.solve_seconds/1e9))
