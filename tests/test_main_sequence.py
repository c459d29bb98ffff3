"""BASELINE configs[0]: the main.py default run (main.py:17-108) — L = LATTICE_SIZE = 30,
defect_fraction = DEFECT_PROB = 3e-3, n_seeds = N_SEEDS = 20, carbon levels 0 / 0.1 / 0.2 — through the
drop-in modules.

The Python reference cannot travel to the GPU box, so the unmodified main.py is pinned in two halves:
  * here (build container, reference present): its call sequence is read from the unmodified source
    with `ast` and compared with the sequence the GPU test drives (same functions, same keyword
    arguments, same carbon levels, same output prefixes);
  * on the GPU: that sequence runs through the drop-in modules and every per-level result is compared
    with the fixture the UNMODIFIED reference produced for exactly those arguments
    (tests/golden/traj_L30_c*.npz; N_STEPS shortened from 20 000 to 2 001, oracle/gen_golden.py).
"""
import ast
import os

import numpy as np
import pytest

from conftest import golden

CARBON_LEVELS = [0.0, 0.1, 0.2]                                   # main.py:23
RUN_KMC_KWARGS = dict(L="LATTICE_SIZE", n_steps="N_STEPS", temp="T_SUB", defect_fraction="DEFECT_PROB",
                      n_seeds="N_SEEDS", impurity_c="c_level", output_prefix="prefix")      # main.py:56-64
INIT_KWARGS = dict(lattice_size="LATTICE_SIZE", n_seeds="N_SEEDS", T_sub="T_SUB", random_seed=42,
                   impurity_c="c_level")                                                      # main.py:44-50
N_STEPS_TEST = 2001
REF_MAIN = os.path.join(os.environ.get("CETKMC_REFERENCE_DIR", "/root/reference"), "main.py")


def _kw(call):
    return {k.arg: (k.value.id if isinstance(k.value, ast.Name) else ast.literal_eval(k.value)) for k in call.keywords}


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(REF_MAIN), reason="reference not present (build container only)")
def test_sequence_matches_unmodified_main_py():
    tree = ast.parse(open(REF_MAIN).read())
    calls = {n.func.id: n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)}
    assert _kw(calls["run_kmc"]) == RUN_KMC_KWARGS
    assert _kw(calls["initialize_lattice"]) == INIT_KWARGS
    levels = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "carbon_levels"]
    assert ast.literal_eval(levels[0].value) == CARBON_LEVELS
    src = open(REF_MAIN).read()
    assert 'prefix = f"impurity_c_{int(c_level*100)}"' in src and 'csv_path = f"outputs/{prefix}/metrics.csv"' in src
    # the constants the fixtures were generated with are the reference's (and the product's defaults)
    ns = {}
    exec(compile(open(os.path.join(os.path.dirname(REF_MAIN), "constants.py")).read(), "constants.py", "exec"), ns)
    assert (ns["LATTICE_SIZE"], ns["N_SEEDS"], ns["DEFECT_PROB"], ns["T_SUB"], ns["N_STEPS"]) == (30, 20, 3e-3, 2800, 20000)
    for c in CARBON_LEVELS:
        kw = ast.literal_eval(str(golden(f"traj_L30_c{int(c * 10):02d}.npz")["kwargs"]))
        assert kw == dict(L=30, n_steps=N_STEPS_TEST, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=c)


def test_product_defaults_are_the_reference_constants():
    from cetkmc._config import constants as K
    assert (K.LATTICE_SIZE, K.N_SEEDS, K.DEFECT_PROB, K.T_SUB, K.N_STEPS, K.METRIC_UPDATE_STEP) == (30, 20, 3e-3, 2800, 20000, 200)


@pytest.mark.gpu
def test_main_py_sequence_on_the_gpu(cet, tmp_path, monkeypatch):
    """main.py:33-79 with the drop-in modules: per carbon level initialise + save the lattice, run_kmc,
    read the CSV back, classify — compared with the unmodified reference's own runs."""
    import pandas as pd
    from cetkmc import campaign, kmc_simulation as ks, metrics as M
    from cetkmc._config import constants as K
    monkeypatch.chdir(tmp_path)
    for c_level in CARBON_LEVELS:
        prefix = f"impurity_c_{int(c_level*100)}"
        output_dir = f"outputs/{prefix}"
        os.makedirs(f"{output_dir}/microstructures", exist_ok=True)
        state, theta, phi, T, atom_type = ks.initialize_lattice(lattice_size=K.LATTICE_SIZE, n_seeds=K.N_SEEDS, T_sub=K.T_SUB,
                                                                random_seed=42, impurity_c=c_level)
        campaign.save_lattice(state, theta, phi, T, atom_type, prefix=f"{output_dir}/init")
        state, atom_type, total_time, theta, phi = ks.run_kmc(L=K.LATTICE_SIZE, n_steps=N_STEPS_TEST, temp=K.T_SUB,
                                                              defect_fraction=K.DEFECT_PROB, n_seeds=K.N_SEEDS,
                                                              impurity_c=c_level, output_prefix=prefix)
        g = golden(f"traj_L30_c{int(c_level * 10):02d}.npz")
        np.testing.assert_array_equal(state, g["state"])
        np.testing.assert_array_equal(atom_type, g["atom_type"])
        np.testing.assert_array_equal(theta, g["theta"])
        np.testing.assert_array_equal(phi, g["phi"])
        assert total_time == float(g["total_time"])
        df = pd.read_csv(f"outputs/{prefix}/metrics.csv")
        assert len(df) == len(g["csv_Step"]) == 11
        for col in df.columns:
            want = g[f"csv_{col}"]
            if col == "CET_Class":
                assert df[col].tolist() == want.tolist()
            else:
                np.testing.assert_allclose(df[col].to_numpy().astype(float), want.astype(float), rtol=1e-12, err_msg=col)
        final = df.iloc[-1].to_dict()
        assert bool(M.detect_CET_transition(final)) == bool(g["csv_CET_Class"][-1] == "Equiaxed")
        # the file name plot_cet.py globs (plot_cet.py:26) and the five-file snapshot (lattice_init.py:98-105)
        assert os.path.isfile(f"outputs/{prefix}/metrics_{int(c_level*100)}.csv")
        assert all(os.path.isfile(f"{output_dir}/init_{n}.npy") for n in
                   ("state", "orientation_theta", "orientation_phi", "temperature", "atom_type"))
