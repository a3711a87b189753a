import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec
ctx = vrec.Context(0)
lib = ctx.lib
lib.vrec_debug_tc_matmul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
SW = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def mm(A, B):
    A = np.ascontiguousarray(A, dtype=np.float16); B = np.ascontiguousarray(B, dtype=np.float16)
    out = np.zeros((128, 128), dtype=np.float32)
    rc = lib.vrec_debug_tc_matmul(ctx._h, A.ctypes.data, B.ctypes.data, out.ctypes.data, SW)
    assert rc == 0
    return out
ones = np.ones((128, 128))
# 1. output mapping: A = row index in every k? use A[i][k] = (k==0)*i, B[j][k] = (k==0)*1  -> C[i][j] = i
A = np.zeros((128, 128)); A[:, 0] = np.arange(128); B = np.zeros((128, 128)); B[:, 0] = 1
c = mm(A, B); print("T1 C[i][j]==i ?", np.array_equal(c, np.arange(128)[:, None] * np.ones((1, 128))), c[:4, :4], c[125:, :3])
# 2. B row mapping: A[:,0]=1, B[j][0]=j -> C[i][j]=j
A = np.zeros((128, 128)); A[:, 0] = 1; B = np.zeros((128, 128)); B[:, 0] = np.arange(128)
c = mm(A, B); print("T2 C[i][j]==j ?", np.array_equal(c, np.ones((128, 1)) * np.arange(128)[None, :]), c[:3, :10])
# 3. k pairing: A one-hot at k0 (all rows), B[j][k] = k -> C[i][j] = k0 expected
for k0 in (0, 1, 7, 8, 9, 15, 16, 17, 31, 64, 127):
    A = np.zeros((128, 128)); A[:, k0] = 1; B = np.ones((128, 1)) * np.arange(128)[None, :]
    c = mm(A, B); u = np.unique(c)
    print(f"T3 k0={k0}: unique C values {u[:8]} (expect {k0})")
# 4. A row mapping with k: A[i][k] = (k==k0)*i
for k0 in (0, 9, 100):
    A = np.zeros((128, 128)); A[:, k0] = np.arange(128); B = np.zeros((128, 128)); B[:, k0] = 1
    c = mm(A, B); print(f"T4 k0={k0}: C[:,0] == arange ? {np.array_equal(c[:,0], np.arange(128))}", c[:10, 0])

rng = np.random.default_rng(0)
A = rng.random((128, 128)).astype(np.float16); B = (rng.random((128, 128)) * 0.5).astype(np.float16)
c = mm(A, B); ref = A.astype(np.float64) @ B.astype(np.float64).T
print("random max err", np.abs(c - ref).max())
