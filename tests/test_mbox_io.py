"""Binary on-disk formats of the reference (mbox_read_* / mbox_write_*, amg/src/mbox.cpp:310-483):
the files written by the host mirror have exactly the byte layout upstream SAAMGE dumps use
(native int32 sizes, int32 I / J, float64 data, no other header) and read back unchanged."""
import ctypes
import os
import struct

import numpy as np

import saamge_b200 as sab

ip = ctypes.POINTER(ctypes.c_int)
dp = ctypes.POINTER(ctypes.c_double)


def _csr():
    rng = np.random.default_rng(5)
    h, w = 7, 5
    dense = rng.standard_normal((h, w)) * (rng.random((h, w)) < 0.4)
    I = np.zeros(h + 1, dtype=np.int32)
    J, A = [], []
    for i in range(h):
        for j in range(w):
            if dense[i, j] != 0.0:
                J.append(j)
                A.append(dense[i, j])
        I[i + 1] = len(J)
    return h, w, I, np.asarray(J, dtype=np.int32), np.asarray(A, dtype=np.float64)


def test_sparse_matrix_layout_and_round_trip(tmp_path):
    hl = sab.host_lib()
    h, w, I, J, A = _csr()
    fn = str(tmp_path / "m.spm").encode()
    assert hl.sa_drv_mbox_write_sparse(fn, h, w, I.ctypes.data_as(ip), J.ctypes.data_as(ip), A.ctypes.data_as(dp)) == 0
    raw = open(fn, "rb").read()
    # amg/src/mbox.cpp:378-395: size, width, j_size, I, J, data
    expect = struct.pack("=iii", h, w, len(J)) + I.tobytes() + J.tobytes() + A.tobytes()
    assert raw == expect
    hh, ww, nz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    I2, J2, A2 = np.zeros_like(I), np.zeros_like(J), np.zeros_like(A)
    hl.sa_drv_mbox_read_sparse(fn, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(nz), I2.ctypes.data_as(ip),
                               J2.ctypes.data_as(ip), A2.ctypes.data_as(dp))
    assert (hh.value, ww.value, nz.value) == (h, w, len(J))
    assert np.array_equal(I, I2) and np.array_equal(J, J2) and np.array_equal(A, A2)


def test_table_layout_and_round_trip(tmp_path):
    hl = sab.host_lib()
    h, _w, I, J, _A = _csr()
    fn = str(tmp_path / "t.tbl").encode()
    hl.sa_drv_mbox_write_table(fn, h, I.ctypes.data_as(ip), J.ctypes.data_as(ip))
    # amg/src/mbox.cpp:331-344: size, j_size, I, J
    assert open(fn, "rb").read() == struct.pack("=ii", h, len(J)) + I.tobytes() + J.tobytes()
    nr, nc = ctypes.c_int(), ctypes.c_int()
    I2, J2 = np.zeros_like(I), np.zeros_like(J)
    hl.sa_drv_mbox_read_table(fn, ctypes.byref(nr), ctypes.byref(nc), I2.ctypes.data_as(ip), J2.ctypes.data_as(ip))
    assert (nr.value, nc.value) == (h, len(J)) and np.array_equal(I, I2) and np.array_equal(J, J2)


def test_dense_matrix_array_layout_and_round_trip(tmp_path):
    hl = sab.host_lib()
    rng = np.random.default_rng(6)
    hs = np.asarray([3, 1, 4], dtype=np.int32)
    ws = np.asarray([2, 5, 4], dtype=np.int32)
    mats = [np.asfortranarray(rng.standard_normal((a, b))) for a, b in zip(hs, ws)]
    data = np.concatenate([m.flatten(order="F") for m in mats])
    fn = str(tmp_path / "d.arr").encode()
    hl.sa_drv_mbox_write_dense_arr(fn, 3, hs.ctypes.data_as(ip), ws.ctypes.data_as(ip), data.ctypes.data_as(dp))
    # amg/src/mbox.cpp:469-481 + 438-446: n, then (height, width, column-major data) per matrix
    expect = struct.pack("=i", 3)
    for m in mats:
        expect += struct.pack("=ii", *m.shape) + m.flatten(order="F").tobytes()
    assert open(fn, "rb").read() == expect
    hs2, ws2, d2 = np.zeros(3, dtype=np.int32), np.zeros(3, dtype=np.int32), np.zeros_like(data)
    n = hl.sa_drv_mbox_read_dense_arr(fn, 3, hs2.ctypes.data_as(ip), ws2.ctypes.data_as(ip), d2.ctypes.data_as(dp))
    assert n == 3 and np.array_equal(hs, hs2) and np.array_equal(ws, ws2) and np.array_equal(data, d2)
    # a single matrix file has no leading count (amg/src/mbox.cpp:438-446)
    fn1 = str(tmp_path / "d.one").encode()
    hl.sa_drv_mbox_write_dense_arr(fn1, 1, hs.ctypes.data_as(ip), ws.ctypes.data_as(ip), data.ctypes.data_as(dp))
    assert open(fn1, "rb").read() == struct.pack("=ii", 3, 2) + mats[0].flatten(order="F").tobytes()
    assert os.path.getsize(fn1) == 8 + 8 * 6
