"""Minimal driver for ncu captures of the dense rate kernel / refresh on the 512^3 half-grown lattice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cetkmc
from cetkmc import _synth
from cetkmc._config import rate_params

L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
packed, th, ph, T = _synth.half_grown(L)
ctx = cetkmc.Context(L=L)
ctx.set_rate_params(rate_params(0.1))
ctx.upload_packed(packed); ctx.upload(theta=th, phi=ph, T=T)
for _ in range(3):
    ctx.rates_build()
if sweeps:
    sp = cetkmc._lib.SweepParams()
    sp.seed, sp.events_per_sweep, sp.p_max, sp.defect_fraction, sp.thermal_every = 1, 0.005 * L ** 3, 0.1, 0.0, 0
    ctx.sweep_run(sweeps, sp, None)
ctx.sync()
ctx.close()
