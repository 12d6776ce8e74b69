"""Mixed (Darcy) path: MfemBlockOperator, Block Jacobi / Block GS / Block LDU preconditioners with the
"DIAGONAL" Schur complement, blocked AMGe hierarchy (Forms = 2 3) and GMRES, driven through the
ParameterList API and compared with the oracle restatement on the same mesh (configs "MultigridTestDarcy")."""
import numpy as np
import pytest
import scipy.sparse as sp

from parelag_b200 import api
from oracle import amge, drivers, solve as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    return api.session()


@pytest.fixture(scope="module")
def hier():
    return amge.build_hierarchy((8, 8, 8), 3, jstart=2)


def test_darcy_blocks_match_oracle(sess, hier):
    mesh, seqs = hier
    S = api.Sequence.hex((8, 8, 8), 3, jstart=2)
    M, B, Bt = S.assemble_darcy(sess, 0)
    Mo, Bo = drivers.darcy_blocks(seqs[0])
    for X, Y in ((M.to_scipy(), Mo), (B.to_scipy(), Bo), (Bt.to_scipy(), sp.csr_matrix(Bo.T))):
        # values only: the device product keeps the explicit zeros of the dense element blocks
        Y = sp.csr_matrix(Y)
        assert X.shape == Y.shape and abs(X - Y).max() <= 1e-13 * abs(Y).max()
    S.free()


@pytest.mark.parametrize("block,amge_prec", [("Block Jacobi", True), ("Block GS", True), ("Block LDU", False), ("Block Jacobi", False)])
def test_gmres_block_preconditioners(sess, hier, block, amge_prec):
    mesh, seqs = hier
    S = api.Sequence.hex((8, 8, 8), 3, jstart=2)
    M, B, Bt = S.assemble_darcy(sess, 0)
    nu, npr = M.info()[0], B.info()[0]
    rng = np.random.default_rng(5)
    b = np.concatenate([rng.standard_normal(nu), rng.standard_normal(npr)])
    A0, prec = drivers.darcy_solver(seqs, block=block, amge=amge_prec)
    xo, ito, convo, histo = orc.gmres(A0.mult, prec, b, rtol=1e-6, atol=1e-6, max_iter=300, restart=50)
    xml = api.library_xml(drivers.darcy_library_entries(block=block, amge=amge_prec))
    solver = api.BlockSolver(xml, "GMRES-AMGe-Blk", [[M, Bt], [B, None]], S, 0, [2, 3])
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and convo, (conv, convo, it, ito)
    assert abs(it - ito) <= 1, (it, ito)
    m = min(len(hist), len(histo))
    ho = np.array(histo[:m])
    sel = ho > 1e-8 * ho[0]
    rel = np.abs(hist[:m] - ho)[sel] / ho[sel]
    assert rel.max() < 1e-8, rel
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    # the computed solution solves the saddle-point system
    assert np.linalg.norm(A0.mult(x) - b) <= 5e-2 * np.linalg.norm(b)
    solver.free(); S.free()
