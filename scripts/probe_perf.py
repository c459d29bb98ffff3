"""Quick device-time probe of the three dense kernels (CUDA events on the context stream)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cetkmc
from cetkmc import _synth
from cetkmc._config import rate_params, thermal_params

def main():
    Ls = [int(x) for x in sys.argv[1:]] or [256]
    for L in Ls:
        t0 = time.time()
        packed, th, ph, T = _synth.half_grown(L)
        print(f"L={L}: synth {time.time()-t0:.1f}s fill={np.count_nonzero(packed & 15)/packed.size:.3f}", flush=True)
        ctx = cetkmc.Context(L=L)
        ctx.set_rate_params(rate_params(0.1))
        ctx.upload_packed(packed); ctx.upload(theta=th, phi=ph, T=T)
        N = L ** 3
        tp = thermal_params(1e-6, nan_to_num=True)
        for name, fn, bytes_per_site in (("thermal_cet", lambda: ctx.thermal_cet(tp), 16),
                                         ("rates_build", lambda: ctx.rates_build(), 41)):
            for _ in range(3): fn()
            ctx.profile_enable(True)
            ctx.timer_begin()
            n = 10
            for _ in range(n): fn()
            ms = ctx.timer_end_ms() / n
            ctx.profile_enable(False)
            kms = ctx.profile_read("rates" if name == "rates_build" else "thermal")[0] / n
            print(f"  {name}: {ms:.3f} ms (kernel {kms:.3f} ms)  {N/ms*1e3:.3e} sites/s  {N*bytes_per_site/kms/1e6:.0f} GB/s algorithmic", flush=True)
        ctx.upload(T=T)
        sp = cetkmc._lib.SweepParams()
        for eps, pmax in ((0.02, 0.25), (0.005, 0.1), (0.001, 0.05)):
            sp.seed, sp.events_per_sweep, sp.p_max, sp.defect_fraction, sp.thermal_every = 1, eps * N, pmax, 0.0, 0
            r = ctx.sweep_run(3, sp, None)
            ctx.profile_enable(True)
            ctx.timer_begin()
            r = ctx.sweep_run(10, sp, None)
            ms = ctx.timer_end_ms() / 10
            ctx.profile_enable(False)
            kb = {k: ctx.profile_read(k)[0] / 10 for k in ("decide", "pick", "apply", "refresh", "rates")}
            print(f"  sweep eps={eps} pmax={pmax}: {ms:.3f} ms  {N/ms*1e3:.3e} site-updates/s  applied/sweep={r['events_applied']/10:.0f} fired/sweep={r['events_fired']/10:.0f} ovf={r['overflow']} "
                  + " ".join(f"{k}={v:.3f}" for k, v in kb.items()), flush=True)
        # fresh lattice (mostly empty): the HBM-bound regime
        st = np.zeros((L, L, L), np.uint8); st[:, :, 0] = (np.random.default_rng(0).random((L, L)) < 0.02)
        ctx.upload_packed(st); ctx.upload(theta=np.zeros((L, L, L)), phi=np.zeros((L, L, L)))
        for _ in range(3): ctx.rates_build()
        ctx.profile_enable(True)
        ctx.timer_begin()
        for _ in range(10): ctx.rates_build()
        ms = ctx.timer_end_ms() / 10
        ctx.profile_enable(False)
        kms = ctx.profile_read("rates")[0] / 10
        print(f"  rates_build(fresh): {ms:.3f} ms (kernel {kms:.3f} ms)  {N/ms*1e3:.3e} sites/s  {N*41/kms/1e6:.0f} GB/s algorithmic (41 B/site)", flush=True)
        ctx.close()

main()
