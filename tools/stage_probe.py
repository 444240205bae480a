import sys, time, ctypes
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import saamge_b200 as sab
n=int(sys.argv[1]); epa=int(sys.argv[2]); kind=int(sys.argv[3]) if len(sys.argv)>3 else 0
p=sab.default_params(num_levels=2, first_elems_per_agg=epa, elems_per_agg=64, partition_kind=kind, block=((32,32,32) if kind==2 else (4,4,4)))
t=time.time(); pr=sab.Problem(3,n,coef_kind=1); print("gen %.2fs"%(time.time()-t))
t=time.time(); na=pr.partition(p); print("partition %.2fs AEs %d ND %d mises %d"%(time.time()-t,na,pr.scalar("ND"),pr.scalar("num_mises")))
h=sab.host_lib()
h.sa_drv_gpu_profile.argtypes=[ctypes.c_int,ctypes.c_char_p,ctypes.c_int]
B=h.sa_drv_bench_create(pr.handle,ctypes.byref(p),0)
h.sa_drv_gpu_profile(1,None,0)
for it in range(3):
    ms=h.sa_drv_bench_step(B,0,0,na)
    buf=ctypes.create_string_buffer(4096); h.sa_drv_gpu_profile(-1,buf,4096)
    ph=[h.sa_drv_bench_scalar(B,b"phase.%d"%i) for i in range(8)]
    print("resident step %.3f ms -> %.0f AE/s"%(ms, na/ms*1e3), buf.value.decode().replace("\n","; "), "phase Mcycles", [round(x/1e6,1) for x in ph])
h.sa_drv_gpu_profile(0,None,0)
for it in range(2):
    ms=h.sa_drv_bench_step(B,0,0,na); print("resident(noprof) %.3f ms -> %.0f AE/s"%(ms, na/ms*1e3))
F=h.sa_drv_bench_scalar(B,b"flops"); print("alg flops %.3e -> %.3f TFLOP/s; sum_m %d"%(F, F/ms/1e9, h.sa_drv_bench_scalar(B,b"sum_m")))
for it in range(6):
    ms=h.sa_drv_bench_step(B,1,0,na); print("e2e %.3f ms -> %.0f AE/s h2d %.1f MB d2h %.1f MB pinned %d"%(ms, na/ms*1e3, h.sa_drv_bench_scalar(B,b"h2d_bytes")/1e6, h.sa_drv_bench_scalar(B,b"d2h_bytes")/1e6, h.sa_drv_bench_scalar(B,b"pinned")))
    print("   breakdown upload %.1f compute %.1f readback %.1f ms"%tuple(h.sa_drv_bench_scalar(B,k) for k in (b"e2e.upload_ms",b"e2e.compute_ms",b"e2e.readback_ms")))
