"""GPU tests written after the last GPU call of round 2 (collected last, not yet run on a GPU): the inputs of the
reference's own unit test src/linalg/unit_test/block2x2_test.cpp through the product's block preconditioners (the oracle
half and the residual the unit test prints are checked on the CPU), and the "Hypre Jacobi" diagonal scaling."""
import numpy as np
import pytest
import scipy.sparse as sp

from parelag_b200 import api, capi
from oracle import solve as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    return api.session()


@pytest.mark.parametrize("block", ["Block GS", "Block Jacobi"])
def test_reference_block2x2_unit_test_inputs(sess, block):
    """The inputs of the reference's own unit test src/linalg/unit_test/block2x2_test.cpp:16-33,96-131,160-204 (6x6 block
    matrix, right-hand side, initial guess; "BlkGS" / "BlkJacobi" with hypre "Gauss-Seidel" inverses, "S Type" Diagonal,
    iterative_mode = true, forms {0, 0}).  The test prints the result without an expected value; here the product is
    compared with the oracle's restatement on exactly those inputs."""
    M = sp.csr_matrix(np.array([[2., 1, 0, 0], [3, 4, 3, 0], [0, 3, 4, 3], [0, 0, 1, 2]]))
    Bt = sp.csr_matrix(np.array([[1., 0], [1, 1], [1, 1], [0, 1]]))
    B = sp.csr_matrix(np.array([[2., 1, 2, 0], [0, 2, 1, 2]]))
    Cm = sp.csr_matrix(np.array([[3., 0], [1, 3]]))
    b = np.array([2., 1, 3, 1, 4, 1])
    x0 = np.array([1., 2, 3, 4, 5, 6])
    A0 = orc.BlockOp([[M, Bt], [B, Cm]])
    assert np.allclose(b - A0.mult(x0), [-7, -30, -38, -16, -21, -37])        # the residual the unit test prints first
    negS = sp.csr_matrix(orc.schur_complement(M, Bt, B, Cm, 1.0, "DIAGONAL") * (-1.0))
    g0, g1 = orc.Smoother(M, type=6), orc.Smoother(negS, type=6)
    inv = [lambda r: g0.apply(r, np.zeros_like(r), False), lambda r: g1.apply(r, np.zeros_like(r), False)]
    Po = (orc.BlockGS if block == "Block GS" else orc.BlockJacobi)(A0, inv)
    xo = Po.apply(b, x0, True)
    lib = {"Gauss-Seidel": ("Hypre", {"Type": "Gauss-Seidel"}),
           "Blk": (block, {"A00 Inverse": "Gauss-Seidel", "A11 Inverse": "Gauss-Seidel", "S Type": "Diagonal"})}
    dev = [[capi.Mat.from_scipy(sess, M), capi.Mat.from_scipy(sess, Bt)], [capi.Mat.from_scipy(sess, B), capi.Mat.from_scipy(sess, Cm)]]
    solver = api.BlockSolver(api.library_xml(lib), "Blk", dev, None, 0, [0, 0])
    x = solver.mult(b, x0=x0)
    assert np.abs(x - xo).max() <= 1e-13 * np.abs(xo).max(), (x, xo)
    solver.free()


def test_hypre_jacobi_is_diagonal_scaling(sess):
    """"Hypre Jacobi" (type 413, ParELAG_HypreSmootherFactory.cpp:27-31 -> mfem::HypreDiagScale): x = diag(A)^{-1} b, the
    initial guess is ignored"""
    from tests.util import random_spd
    A = random_spd(200, 0.05, 3)
    b = np.random.default_rng(4).standard_normal(200)
    solver = api.Solver(api.library_xml({"J": ("Hypre", {"Type": "Hypre Jacobi"})}), "J", A)
    want = b / A.diagonal()
    assert np.abs(solver.mult(b) - want).max() <= 1e-15 * np.abs(want).max()
    assert np.abs(solver.mult(b, x0=np.ones(200)) - want).max() <= 1e-15 * np.abs(want).max()
    solver.free()


def test_product_check_invariants_after_coarsen(sess):
    """DeRhamSequence::CheckInvariants (CheckCoarseMassMatrix, CheckD, CheckDP) on hierarchies the product coarsened itself"""
    from oracle import amge
    S = api.Sequence.hex((8, 8, 8), 3)
    for l in range(2):
        assert S.check_invariants(l) <= 1e-10
    S.free()
    X = amge.DeformedHexMesh(4, 4, 4, deform=amge.weak_scaling_deformation).vertex_coords()
    S = api.Sequence.hex((4, 4, 4), 3, jstart=1, coords=X)
    for l in range(2):
        assert S.check_invariants(l) <= 1e-9
    S.free()
