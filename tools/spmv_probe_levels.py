import sys, ctypes
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import saamge_b200 as sab
from saamge_b200 import cabi
n=int(sys.argv[1])
p=sab.default_params(num_levels=2, first_elems_per_agg=64, partition_kind=1, block=(4,4,4))
pr=sab.Problem(3,n,coef_kind=1); pr.partition(p)
ctx=cabi.Context(0); lev=cabi.Level(ctx,pr)
lev.build_Dinv_neg()
nnz=pr.scalar("nnz"); nd=pr.scalar("ND")
gl=sab.gpu_lib()
ms=gl.sa_gpu_bench_spmv(lev.h,0,50); ms2=gl.sa_gpu_bench_smoother(lev.h,50)
b=12*nnz+20*nd
print("n=%d spmv %.4f ms %.0f GB/s (%.1f%% of 6557.8) smoother %.4f ms %.0f GB/s"%(n,ms,b/ms/1e6,100*b/ms/1e6/6557.8,ms2,(b+24*nd)/ms2/1e6))
