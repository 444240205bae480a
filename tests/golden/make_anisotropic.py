"""Converts the reference's algebraic test matrix (amg/data/anisotropic.mat.00000, the input of
the `algebraic` CTest, amg/test/CMakeLists.txt:72-78) into a compressed CSR fixture.  Run in the
build container (the reference tree is not available on the GPU box):
    python tests/golden/make_anisotropic.py /root/reference/amg/data/anisotropic.mat.00000"""
import os
import sys

import numpy as np
import scipy.sparse as sp

src = sys.argv[1]
with open(src) as f:
    r0, r1, c0, c1 = [int(t) for t in f.readline().split()]
d = np.loadtxt(src, skiprows=1)
A = sp.coo_matrix((d[:, 2], (d[:, 0].astype(int), d[:, 1].astype(int))), shape=(r1 + 1, c1 + 1)).tocsr()
A.sum_duplicates()
A.sort_indices()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anisotropic_mat.npz")
np.savez_compressed(out, indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), data=A.data,
                    shape=np.array(A.shape))
print(out, A.shape, A.nnz)
