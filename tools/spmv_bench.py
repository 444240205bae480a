import sys, ctypes
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import saamge_b200 as sab
p=sab.default_params(num_levels=2, first_elems_per_agg=64, elems_per_agg=64, partition_kind=1, block=(4,4,4), coarse_block=4)
pr=sab.Problem(3,128,coef_kind=1); pr.partition(p)
from saamge_b200 import cabi
ctx=cabi.Context(0); lev=cabi.Level(ctx,pr)
lev.build_Dinv_neg()
g=sab.gpu_lib(); g.sa_gpu_bench_spmv.restype=ctypes.c_double; g.sa_gpu_bench_smoother.restype=ctypes.c_double
nnz=pr.scalar("nnz"); nd=pr.scalar("ND")
for it in range(3):
    ms=g.sa_gpu_bench_spmv(lev.h,0,100); ms2=g.sa_gpu_bench_smoother(lev.h,100)
    b1=12.0*nnz+20.0*nd; b2=b1+24.0*nd
    print("spmv %.4f ms %.0f GB/s (%.1f%%)   smoother %.4f ms %.0f GB/s (%.1f%%)"%(ms,b1/ms/1e6,b1/ms/1e6/65.578,ms2,b2/ms2/1e6,b2/ms2/1e6/65.578))
