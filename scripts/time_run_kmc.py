"""Wall time of the exact drop-in `run_kmc` (configs[0]/[1] of BASELINE.json) on the GPU."""
import os, sys, time, tempfile, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cetkmc
from cetkmc import kmc_simulation as ks

def main():
    cases = [dict(L=30, n_steps=2000, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.1),
             dict(L=64, n_steps=1000, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.1),
             dict(L=128, n_steps=400, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.1)]
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        ks.run_kmc(L=12, n_steps=50, output_prefix="warm")           # context / library warm-up
        for kw in cases:
            buf = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(buf):
                ks.run_kmc(output_prefix="t", **kw)
            dt = time.perf_counter() - t0
            print(f"run_kmc L={kw['L']} steps={kw['n_steps']}: {dt:.2f} s wall, {kw['n_steps']/dt:.1f} steps/s "
                  f"({dt/kw['n_steps']*1e3:.2f} ms/step incl. metrics every 200 steps)", flush=True)
main()
