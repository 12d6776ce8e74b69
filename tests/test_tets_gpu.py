"""Tetrahedral meshes on the GPU (BASELINE configs[0], examples/MultigridTest0Form.cpp): DeRhamSequence::Coarsen() of
all four forms on a refined tetrahedral mesh against the oracle (integer tables bit-exact, values 1e-12), and the
driver's solver -- H1 Laplacian A = D_0^T M_1 D_0, PCG preconditioned by a 3-level AMGe with l1-Gauss-Seidel smoothers
(examples/testing_helpers/Create0FormParameterList.hpp, coarse solver PCG-GS) -- against the oracle's residual history."""
import os

import numpy as np
import pytest

from parelag_b200 import api
from oracle import amge, drivers, solve as orc, tets
from tests.test_coarsen_gpu import compare_levels

pytestmark = pytest.mark.gpu
MESH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cube456.npz")


@pytest.fixture(scope="module")
def sess():
    return api.session()


def test_coarsen_refined_kuhn_cubes_three_levels(sess):
    V, T, B, A = tets.cube_tets(2)
    mesh, seqs = tets.build_hierarchy(tets.TetMesh(V, T, B, A), 2, 3)
    S = api.Sequence.tet(V, T, B, A, 2, 3)
    compare_levels(S, seqs)
    for l in (1, 2):
        assert S.stat(l, "trace_null_1") == 0 and S.stat(l, "trace_null_2") == 0     # nested Whitney spaces
    S.free()


def test_cube456_h1_pcg_amge_history(sess):
    """configs[0] one refinement level down: cube456 refined twice, 3 levels"""
    V, T, B, A = tets.load_npz(MESH)
    mesh, seqs = tets.build_hierarchy(tets.TetMesh(V, T, B, A), 2, 3)
    S = api.Sequence.tet(V, T, B, A, 2, 3)
    compare_levels(S, seqs)
    ess = np.ones(6, dtype=np.int32)
    Ao, marker = drivers.system_matrix(seqs[0], 0, ess)
    Ad = S.assemble_system(sess, 0, 0, ess)
    Ag = Ad.to_scipy()
    assert abs(Ag - Ao).max() <= 1e-12 * abs(Ao).max()
    rng = np.random.default_rng(2)
    b = rng.standard_normal(Ao.shape[0]); b[marker] = 0.0
    for ordering in ("natural", "multicolor"):
        H = drivers.amge_pcg_solver(seqs, 0, ess, Ao, ordering=ordering)
        xo, ito, convo, histo = orc.pcg(Ao, H.mult, b, rtol=1e-6, atol=1e-6, max_iter=100)
        solver = api.Solver(api.library_xml(drivers.library_entries(0, ordering=ordering)), "PCG-AMGe", Ao, S, 0, 0, ess)
        x = solver.mult(b)
        hist, it, conv = solver.history()
        assert conv and convo and abs(it - ito) <= 1, (ordering, it, ito)
        m = min(len(hist), len(histo))
        assert np.max(np.abs(hist[:m] - np.array(histo[:m])) / np.abs(np.array(histo[:m]))) < 1e-9, ordering
        assert np.linalg.norm(x - xo) <= 1e-7 * np.linalg.norm(xo)
        solver.free()
    S.free()
