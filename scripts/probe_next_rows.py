"""Device time of the §8(f) kernels on the 512^3 half-grown lattice: grain clustering (N1) and the
defect-mask refresh (N3).  Wall time around synchronous C-ABI calls (each returns after a stream sync)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cetkmc
from cetkmc import _synth
from cetkmc._config import rate_params

L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
packed, th, ph, T = _synth.half_grown(L)
ctx = cetkmc.Context(L=L)
ctx.set_rate_params(rate_params(0.1))
ctx.upload_packed(packed); ctx.upload(theta=th, phi=ph, T=T)
N = L ** 3
occ = int(np.count_nonzero(packed & 15))
lib = cetkmc._lib.lib()
import ctypes as C
for rep in range(3):
    n = C.c_int64(0)
    t0 = time.perf_counter()
    cetkmc._lib.check(lib.cet_grains_label(ctx._h, 0.5, C.byref(n)), "label")
    t1 = time.perf_counter()
    g = ctx.grains(0.5)            # label again + stats + sort on the host
    t2 = time.perf_counter()
    print(f"grains L={L}: label {1e3*(t1-t0):.2f} ms ({N/(t1-t0):.3e} sites/s, {occ} occupied, {n.value} grains); "
          f"label+stats+host sort {1e3*(t2-t1):.2f} ms", flush=True)
n_c = int(ctx.counts()[3])
draws = np.random.default_rng(0).random(n_c)
for rep in range(3):
    t0 = time.perf_counter()
    a = ctx.defects_refresh(seed=1, epoch=rep)
    t1 = time.perf_counter()
    b = ctx.defects_refresh(draws=draws)
    t2 = time.perf_counter()
    print(f"defects L={L}: philox {1e3*(t1-t0):.2f} ms ({N*2/(t1-t0)/1e9:.0f} GB/s of vox in+out), "
          f"ordered stream ({n_c} draws, incl. their upload) {1e3*(t2-t1):.2f} ms; masks {a[1]} / {b[1]}", flush=True)
t0 = time.perf_counter(); c = ctx.counts(); t1 = time.perf_counter()
print(f"counts: {1e3*(t1-t0):.2f} ms")
ctx.close()
