"""DeRhamSequence::Coarsen() on the GPU (batched per-agglomerate kernels + host integer
tables) against the CPU oracle on the same mesh and coefficients.

Contract: integer outputs (coarse dof counts and numbering, entity->dof tables, sparsity
patterns of P and of the coarse D) bit-exact; floating point values to 1e-12 relative
(to the largest entry of the matrix)."""
import numpy as np
import pytest

from parelag_b200 import api
from oracle import amge, drivers

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def sess():
    return api.session()


def same_pattern(A, B):
    B = B.tocsr()
    return A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)


def close(A, B, tol=TOL):
    scale = max(abs(B).max(), 1e-300)
    return abs(A - B).max() <= tol * scale


def null_mask(dofh):
    """coarse dofs selected by an SVD of a target residual (NullSpace dofs, DofHandler.cpp:694-760)"""
    m = np.zeros(dofh.ndofs, dtype=bool)
    for d, t in (dofh.dof_type or {}).items():
        m[d] = (t == amge.NULLSPACE)
    return m


def close_split(A, B, tol, null_tol, row_null, col_null):
    """entries that touch a NullSpace dof (row or column) at null_tol, all others at tol; both relative to
    the largest entry of the matrix"""
    scale = max(abs(B).max(), 1e-300)
    E = abs(A - B).tocoo()
    loose = np.zeros(E.nnz, dtype=bool)
    if row_null is not None:
        loose |= row_null[E.row]
    if col_null is not None:
        loose |= col_null[E.col]
    e_strict = E.data[~loose].max() if (~loose).any() else 0.0
    e_loose = E.data[loose].max() if loose.any() else 0.0
    return e_strict <= tol * scale and e_loose <= null_tol * scale, (e_strict / scale, e_loose / scale)


def compare_levels(S, seqs, tol=TOL, null_tol=None):
    """null_tol: tolerance of the entries that belong to NullSpace coarse dofs.  Their basis functions are left
    singular vectors of a target RESIDUAL (targets minus what the PV/RangeT part already represents,
    DeRhamSequence.cpp:2480-2509); when that residual is small against the data (sigma * |T| << 1), the vector is
    determined only to eps / (sigma * |T|): two backward-stable implementations of the local solve (dsytrf + dgesvd in
    the oracle, pivoted LU + Jacobi SVD in the kernels) legitimately differ by that much.  Everything else keeps tol."""
    null_tol = tol if null_tol is None else null_tol
    for l in range(len(seqs) - 1):
        f, c = seqs[l], seqs[l + 1]
        for j in range(f.jstart, 4):
            # integer parity
            for cd in range(4 - j):
                Eg, Eo = S.get_csr(l + 1, "ED", j, cd), c.dof[j].entity_dof[cd]
                assert same_pattern(Eg, Eo), ("entity_dof", l, j, cd)
            nul = null_mask(c.dof[j])
            fnul = null_mask(f.dof[j]) if l > 0 else None
            P, Po = S.get_csr(l, "P", j), f.P[j]
            assert same_pattern(P, Po), ("P pattern", l, j)
            ok, err = close_split(P, Po, tol, null_tol, fnul, nul)
            assert ok, ("P values", l, j, err)
            if j < 3:
                D, Do = S.get_csr(l + 1, "D", j), c.D[j]
                assert same_pattern(D, Do), ("D pattern", l, j)
                ok, err = close_split(D, Do, tol, null_tol, null_mask(c.dof[j + 1]), nul)
                assert ok, ("D values", l, j, err)
            any_null = nul.any() or (fnul is not None and fnul.any())
            for cd in range(4 - j):
                Mg, Mo = S.get_csr(l + 1, "Me", j, cd), c.M[(j, cd)]
                assert Mg.shape == Mo.shape
                assert close(Mg, Mo, null_tol if any_null else tol), ("coarse mass", l, j, cd, abs(Mg - Mo).max())
            Tg, To = S.get_targets(l + 1, j), c.targets[j]
            assert np.abs(Tg - To).max() <= max(1e-11, null_tol if any_null else 0.0) * max(np.abs(To).max(), 1.0), ("targets", l, j)
            assert np.array_equal(S.get_bdr_mask(l + 1, j), drivers.bdr_mask(c.dof[j]))


def test_coarsen_uniform_three_levels(sess):
    dims = (8, 8, 8)
    mesh, seqs = amge.build_hierarchy(dims, 3)
    S = api.Sequence.hex(dims, 3)
    compare_levels(S, seqs)
    # lowest-order targets on a uniform mesh: one coarse dof per coarse entity, no NullSpace dofs
    assert S.stat(1, "facet_ext_2_0_null") == 0 and S.stat(1, "trace_null_2") == 0
    S.free()


def test_coarsen_anisotropic_variable_coefficients_nullspace_dofs(sess):
    """Variable coefficients make the constants leave the span of the M-harmonic
    extensions: NullSpace dofs appear (SVD rank decisions and singular vectors matter)."""
    dims, L = (4, 4, 4), (1.0, 2.0, 0.5)
    rng = np.random.default_rng(7)
    nel = 64
    alpha, beta = rng.uniform(0.5, 2.0, nel), 10.0 ** rng.uniform(-2, 2, nel)
    mesh, seqs = amge.build_hierarchy(dims, 2, L=L, alpha=alpha, beta=beta)
    S = api.Sequence.hex(dims, 2, L=L, alpha=alpha, beta=beta)
    assert seqs[0].stats[("facet_ext", 2)][1] > 0          # the case really has NullSpace dofs
    assert S.stat(1, "facet_ext_2_0_null") == seqs[0].stats[("facet_ext", 2)][1]
    compare_levels(S, seqs, tol=1e-10)
    S.free()


def test_coarsen_jform_start_and_rectangular_domain(sess):
    dims, L = (8, 4, 4), (2.0, 1.0, 1.0)
    mesh, seqs = amge.build_hierarchy(dims, 3, L=L, jstart=1)
    S = api.Sequence.hex(dims, 3, L=L, jstart=1)
    compare_levels(S, seqs)
    S.free()


def test_coarsen_ragged_cartesian_grid(sess):
    """A grid that is not a multiple of the coarsening ratio (the shape of the 60 x 220 x 85 SPE10 grid): logical
    Cartesian agglomeration with thin last blocks (LogicalPartitioner.hpp:46-103), anisotropic cells, lognormal
    coefficient in the H(div) mass matrix."""
    dims, L = (6, 5, 3), (120.0, 50.0, 6.0)
    rng = np.random.default_rng(19)
    kinv = 10.0 ** (-np.clip(rng.normal(-1.0, 1.5, size=dims[0] * dims[1] * dims[2]), -4.0, 2.0))
    mesh, seqs = amge.build_hierarchy(dims, 3, L=L, beta=kinv, jstart=2)
    S = api.Sequence.hex(dims, 3, L=L, beta=kinv, jstart=2)
    compare_levels(S, seqs, tol=1e-10, null_tol=1e-8)
    S.free()


def test_invariants_on_device_result(sess):
    """DeRhamSequence::CheckInvariants on the product's own output."""
    dims = (8, 8, 8)
    S = api.Sequence.hex(dims, 2)
    for j in range(4):
        P = S.get_csr(0, "P", j)
        Mf, Mc = S.get_csr(0, "M", j), S.get_csr(1, "M", j)
        assert abs(Mc - P.T @ Mf @ P).max() <= 1e-12 * abs(Mc).max()          # M_c = P^T M_f P
        if j < 3:
            Df, Dc, Pn = S.get_csr(0, "D", j), S.get_csr(1, "D", j), S.get_csr(0, "P", j + 1)
            assert abs(Df @ P - Pn @ Dc).max() <= 1e-12                        # D_f P_j = P_{j+1} D_c
        if j < 2:
            assert abs(S.get_csr(1, "D", j + 1) @ S.get_csr(1, "D", j)).max() <= 1e-12   # D D = 0
        T_f, T_c = S.get_targets(0, j), S.get_targets(1, j)
        assert np.abs(P @ T_c - T_f).max() <= 1e-12 * max(np.abs(T_f).max(), 1.0)   # targets reproduced
    S.free()


def test_solver_on_device_built_hierarchy(sess):
    """End to end with the product's own hierarchy: PCG-AMGe(Hiptmair) for H(div)."""
    from oracle import solve as orc
    dims = (8, 8, 8)
    mesh, seqs = amge.build_hierarchy(dims, 3)
    S = api.Sequence.hex(dims, 3)
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], 2, ess)
    rng = np.random.default_rng(1)
    b = rng.standard_normal(A.shape[0]); b[marker] = 0
    H = drivers.amge_pcg_solver(seqs, 2, ess, A)
    xo, ito, convo, histo = orc.pcg(A, H.mult, b, rtol=1e-6, atol=1e-6, max_iter=100)
    solver = api.Solver(api.library_xml(drivers.library_entries(2)), "PCG-AMGe", A, S, 0, 2, ess)
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and abs(it - ito) <= 1
    m = min(len(hist), len(histo))
    assert np.max(np.abs(hist[:m] - np.array(histo[:m])) / np.abs(np.array(histo[:m]))) < 1e-9
    solver.free(); S.free()
