"""L2 -> shared memory streaming bandwidth of 148 CTAs reading one common 256 MB buffer with bulk copies.
Needs a library built with the micro-benchmarks: make -C locations-recommender_b200/csrc EXTRA=-DVREC_WITH_MICROBENCH=1"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec  # noqa: E402

ctx = vrec.Context(0)
lib = ctx.lib
lib.vrec_debug_tc_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
ntiles = 7813
for ctas in (148, 74, 37):
    for stages in (3, 4, 6):
        for stagger in (1, 67):
            ms = C.c_double(0)
            rc = lib.vrec_debug_tc_stream(ctx._h, ctas, ntiles, stages, stagger, C.byref(ms))
            assert rc == 0
            gb = ctas * ntiles * 32768 / 1e9
            print(f"ctas={ctas} stages={stages} stagger={stagger}: {ms.value:.2f} ms  {gb / ms.value * 1e3:.0f} GB/s "
                  f"({gb / ms.value * 1e3 / ctas:.1f} GB/s per CTA)", flush=True)
