"""Multi-rank host logic on CPU: two gloo ranks run tests/par_worker.py (SharingMap numbering,
box-decomposed topology, Assemble / IgnoreNonLocalRange / comm package vs the single-domain oracle)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_multi_rank_host_logic_gloo(nranks):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nranks, "--master-addr", "127.0.0.1",
           "--master-port", str(29533 + nranks), os.path.join(ROOT, "tests", "par_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PAR_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
