"""N4 / configs[4]: the sweep's per-case metrics are laid out so that the reference's unmodified
plot_cet.py loads them (plot_cet.py:26 globs outputs/impurity_c_*/metrics_*.csv)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

REF_PLOT = os.path.join(os.environ.get("CETKMC_REFERENCE_DIR", "/root/reference"), "plot_cet.py")
COLS = ("Step", "Time", "AspectRatio", "EquiaxedFraction", "NucleationDensity", "DefectDensity", "AvgGrainSize",
        "GrainCount", "W_Count", "Re_Count", "C_Count", "NucleationCount", "G_over_R", "G_phys", "R_phys",
        "G_over_R_phys", "CET_Class", "CET_Detected")


def _csv_from_fixture(name, path):
    import pandas as pd
    g = golden(name)
    pd.DataFrame({c: g[f"csv_{c}"] for c in COLS}).to_csv(path, index=False)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(REF_PLOT), reason="reference not present (build container only)")
def test_unmodified_plot_cet_reads_the_tree(tmp_path):
    """Three reference-format metrics files -> write_plot_cet_tree -> the UNMODIFIED plot_cet.py, run
    from the tree's root with a do-nothing matplotlib, discovers all three series and finishes."""
    from cetkmc import campaign
    cases = []
    for q, name in enumerate(("traj_L30_c00.npz", "traj_L30_c01.npz", "traj_L30_c02.npz")):
        p = tmp_path / f"case{q}.csv"
        _csv_from_fixture(name, p)
        cases.append((q, str(p)))
    root = campaign.write_plot_cet_tree(cases, str(tmp_path / "plot_cet"))
    # pandas' own .plot needs the real matplotlib: make it a no-op, then run the script unchanged
    code = ("import runpy, pandas.plotting._core as pc\n"
            "pc.PlotAccessor.__call__ = lambda self, *a, **k: None\n"
            f"runpy.run_path({REF_PLOT!r}, run_name='__main__')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "mplstub"))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "'0% C'" in out.stdout and "'1% C'" in out.stdout and "'2% C'" in out.stdout
    assert "Analysis complete" in out.stdout


@pytest.mark.gpu
def test_gr_sweep_emits_the_tree(cet, tmp_path, monkeypatch):
    import pandas as pd
    from cetkmc import campaign
    monkeypatch.chdir(tmp_path)
    rows = campaign.run_gr_sweep([2800.0, 3000.0], [5e12], L=16, n_sweeps=41, n_seeds=6, output_root="gr", metrics_every=20)
    assert len(rows) == 2
    campaign.merge_cet_map("gr")
    for q in range(2):
        df = pd.read_csv(tmp_path / "outputs" / "gr" / "plot_cet" / "outputs" / f"impurity_c_{q}" / f"metrics_{q}.csv")
        assert list(df.columns) == list(COLS) and len(df) >= 2
        assert np.all(np.diff(df["Step"]) > 0)
