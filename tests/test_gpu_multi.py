"""GPU, >= 2 devices: the z-slab decomposed sublattice run reproduces the single-GPU run of the
same global lattice bit for bit (scripts/check_slabs.py under torch.distributed.run)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [["--L", "48", "--sweeps", "10"], ["--L", "40", "--n0", "80", "--sweeps", "16", "--eps", "0.004"],
                                   ["--L", "64", "--sweeps", "12"],                       # rows a tensor map can describe: the TMA dense kernel on the slabs
                                   ["--L", "64", "--sweeps", "12", "--flags", "32"],      # TMA tile kernel also for the refresh, on the slabs
                                   ["--L", "48", "--sweeps", "10", "--flags", "2"]])      # first-design gather path + full-plane exchange
def test_slabs_match_single_gpu(cet, extra):
    n = min(cet.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "check_slabs.py")] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SLAB CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_distributed_grains_match_single_gpu(cet):
    """Observables on the slab path: grains of a lattice split over the GPUs (local labelling + seam merge,
    metrics.grains_distributed) equal the single-context clustering exactly (scripts/check_grains_slabs.py)."""
    n = min(cet.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29300 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "check_grains_slabs.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "GRAIN SLAB CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_campaign_driver_on_slabs_matches_single_gpu(cet):
    """run_cet_sublattice(world > 1): same lattice, same clock and the same 18-column metrics.csv as the
    single-GPU run (scripts/check_campaign_slabs.py)."""
    n = min(cet.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29000 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "check_campaign_slabs.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CAMPAIGN SLAB CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
