"""Mixed (Darcy) path: MfemBlockOperator, Block Jacobi / Block GS / Block LDU preconditioners with the
"DIAGONAL" Schur complement, blocked AMGe hierarchy (Forms = 2 3) and GMRES, driven through the
ParameterList API and compared with the oracle restatement on the same mesh (configs "MultigridTestDarcy")."""
import numpy as np
import pytest
import scipy.sparse as sp

from parelag_b200 import api, capi
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


def test_spe10_shaped_darcy_block_ldu(sess):
    """configs "MultigridTestSPE10-shaped": anisotropic cells (20 x 10 x 2) and a synthetic lognormal
    permeability (4 decades) in the H(div) mass matrix; GMRES + Block LDU with AMGe-GS on A00
    (examples/example_parameterlists/spe10_example_parameters.xml, BoomerAMG replaced by l1-GS)."""
    dims, L = (8, 8, 4), (160.0, 80.0, 8.0)
    rng = np.random.default_rng(13)
    kinv = 10.0 ** (-np.clip(rng.normal(-1.0, 1.5, size=dims[0] * dims[1] * dims[2]), -4.0, 2.0))   # 1/k per cell
    mesh, seqs = amge.build_hierarchy(dims, 2, L=L, beta=kinv, jstart=2)
    S = api.Sequence.hex(dims, 2, L=L, beta=kinv, jstart=2)
    # coarse spaces under 4-decade coefficients: P of H(div) and L2 vs the oracle
    for j in (2, 3):
        P, Po = S.get_csr(0, "P", j), sp.csr_matrix(seqs[0].P[j])
        assert abs(P - Po).max() <= 1e-10 * abs(Po).max()
    M, B, Bt = S.assemble_darcy(sess, 0)
    nu, npr = M.info()[0], B.info()[0]
    b = np.concatenate([np.zeros(nu), np.ones(npr)])            # unit source
    Mo, Bo = drivers.darcy_blocks(seqs[0])
    A0 = orc.BlockOp([[Mo, sp.csr_matrix(Bo.T)], [Bo, None]])

    def amge_gs(Mx):
        H = drivers.amge_pcg_solver(seqs, 2, np.zeros(6, dtype=np.int32), Mx, hiptmair=False)
        return H.mult
    negS = sp.csr_matrix(orc.schur_complement(Mo, sp.csr_matrix(Bo.T), Bo, None, 1.0, "DIAGONAL") * (-1.0))
    Sg = orc.Smoother(negS, type=2)
    inv = amge_gs(Mo)
    ldu = orc.BlockLDU(A0, inv, inv, inv, lambda r: Sg.apply(r, np.zeros_like(r), False), 0.775)
    xo, ito, convo, histo = orc.gmres(A0.mult, lambda r: ldu.apply(r, np.zeros_like(r), False), b, rtol=1e-6, atol=1e-6,
                                      max_iter=300, restart=50)
    lib = drivers.library_entries(2)                            # Gauss-Seidel, PCG-GS (+ unused Hiptmair entries)
    lib["AMGe-GS"] = ("AMGe", {"Maximum levels": -1, "Forms": [2], "PreSmoother": "Gauss-Seidel", "PostSmoother": "Gauss-Seidel",
                               "Coarse solver": "PCG-GS2", "Cycle type": "V-cycle"})
    lib["PCG-GS2"] = ("Krylov", {"Solver name": "PCG", "Preconditioner": "Gauss-Seidel", "Print level": -1, "Maximum iterations": 3,
                                 "Relative tolerance": 1e-4, "Absolute tolerance": 1e-4})
    lib["Block-LDU-AMGe-GS"] = ("Block LDU", {"Damping Factor": 0.775, "A00_1 Inverse": "AMGe-GS", "A00_2 Inverse": "AMGe-GS",
                                              "A00_3 Inverse": "AMGe-GS", "Alpha": 1.0, "S Type": "Diagonal", "S Inverse": "Gauss-Seidel"})
    lib["GMRES-Block-LDU-AMGe-GS"] = ("Krylov", {"Solver name": "GMRES", "Preconditioner": "Block-LDU-AMGe-GS", "Print level": -1,
                                                 "Maximum iterations": 300, "Relative tolerance": 1e-6, "Absolute tolerance": 1e-6,
                                                 "Restart size": 50})
    solver = api.BlockSolver(api.library_xml(lib), "GMRES-Block-LDU-AMGe-GS", [[M, Bt], [B, None]], S, 0, [2, 3],
                             ess_attr=np.zeros((2, 6), dtype=np.int32))
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and convo and abs(it - ito) <= 1, (it, ito)
    m = min(len(hist), len(histo))
    ho = np.array(histo[:m])
    sel = ho > 1e-7 * ho[0]
    assert (np.abs(hist[:m] - ho)[sel] / ho[sel]).max() < 1e-7
    solver.free(); S.free()


def test_mass_schur_complement(sess, hier):
    """"S Type" = "MASS" (SchurComplementFactory.cpp:43-50): the second diagonal operator of the block preconditioner is
    sequence.ComputeTrueM(forms.front()) -- DofHandler-assembled mass matrix of that form -- negated by the default
    "Use Negative S".  2x2 system on the L2 space (forms 3, 3) so that the block sizes match."""
    mesh, seqs = hier
    S = api.Sequence.hex((8, 8, 8), 3, jstart=2)
    W = sp.csr_matrix(seqs[0].mass_operator(3))
    n = W.shape[0]
    rng = np.random.default_rng(41)
    K = sp.csr_matrix(sp.diags(rng.uniform(1.0, 2.0, n)) @ W)
    blocks = [[K, sp.csr_matrix(0.1 * W)], [sp.csr_matrix(0.1 * W), sp.csr_matrix(3.0 * W)]]
    A0 = orc.BlockOp(blocks)
    J0, J1 = orc.Smoother(K, type=1), orc.Smoother(sp.csr_matrix(-W), type=1)
    bj = orc.BlockJacobi(A0, [lambda r: J0.apply(r, np.zeros_like(r), False), lambda r: J1.apply(r, np.zeros_like(r), False)])
    b = rng.standard_normal(2 * n)
    lib = {"J": ("Hypre", {"Type": "L1 Jacobi", "Sweeps": 1}),
           "Blk": ("Block Jacobi", {"A00 Inverse": "J", "A11 Inverse": "J", "S Type": "Mass"})}
    dev = [[capi.Mat.from_scipy(sess, blocks[i][j]) for j in range(2)] for i in range(2)]
    solver = api.BlockSolver(api.library_xml(lib), "Blk", dev, S, 0, [3, 3])
    x = solver.mult(b)
    xo = bj.apply(b, np.zeros_like(b), False)
    assert np.linalg.norm(x - xo) <= 1e-13 * np.linalg.norm(xo)
    assert np.linalg.norm(x[n:] + orc.Smoother(W, type=1).apply(b[n:], np.zeros(n), False)) <= 1e-13 * np.linalg.norm(x)   # -W, not 3W
    solver.free(); S.free()


def test_block_ldu_with_amge_on_both_blocks(sess):
    """The solver bench.py --config darcy / spe10 runs (spe10_example_parameters.xml with the BoomerAMG inverses replaced):
    Block LDU, A00 inverses = AMGe V-cycle on M (Forms 2), S inverse = AMGe V-cycle on the Schur complement (Forms 3,
    piecewise-constant interpolation): GMRES history vs the oracle; mesh-robust iteration counts (21 / 25 / 29 at 8^3 / 16^3 /
    24^3 in the oracle, where the blocked AMGe with Block Jacobi needs 147 / > 300)."""
    dims = (8, 8, 8)
    mesh, seqs = amge.build_hierarchy(dims, 3, jstart=2)
    S = api.Sequence.hex(dims, 3, jstart=2)
    M, B, Bt = S.assemble_darcy(sess, 0)
    nu, npr = M.info()[0], B.info()[0]
    rng = np.random.default_rng(5)
    b = rng.standard_normal(nu + npr)
    Mo, Bo = drivers.darcy_blocks(seqs[0])
    A0 = orc.BlockOp([[Mo, sp.csr_matrix(Bo.T)], [Bo, None]])
    noess = np.zeros(6, dtype=np.int32)
    negS = sp.csr_matrix(orc.schur_complement(Mo, sp.csr_matrix(Bo.T), Bo, None, 1.0, "DIAGONAL") * (-1.0))
    invM = drivers.amge_pcg_solver(seqs, 2, noess, Mo, hiptmair=False).mult
    invS = drivers.amge_pcg_solver(seqs, 3, noess, negS, hiptmair=False).mult
    ldu = orc.BlockLDU(A0, invM, invM, invM, invS, 0.775)
    xo, ito, convo, histo = orc.gmres(A0.mult, lambda r: ldu.apply(r, np.zeros_like(r), False), b, rtol=1e-6, atol=1e-6,
                                      max_iter=300, restart=50)
    import bench
    lib = bench.library_darcy_ldu("natural")
    solver = api.BlockSolver(api.library_xml(lib), "GMRES with Block LDU", [[M, Bt], [B, None]], S, 0, [2, 3],
                             ess_attr=np.zeros((2, 6), dtype=np.int32))
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and convo and abs(it - ito) <= 1 and it < 40, (it, ito)
    m = min(len(hist), len(histo))
    ho = np.array(histo[:m])
    sel = ho > 1e-7 * ho[0]
    assert (np.abs(hist[:m] - ho)[sel] / ho[sel]).max() < 1e-7
    solver.free(); S.free()
