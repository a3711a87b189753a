import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec
ctx = vrec.Context(0)
f = ctx.lib.vrec_debug_tc_mma_rate
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
out = (C.c_int64 * 2)()
for sw in (0, 1):
    for reps in (1, 100, 1000):
        assert f(ctx._h, sw, reps, out) == 0
        print(f"swizzled={sw} reps={reps}: issue {out[0] / (8 * reps):.0f} cycles/MMA, complete {out[1] / (8 * reps):.0f} cycles/MMA")
