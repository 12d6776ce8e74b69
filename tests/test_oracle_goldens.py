"""Pin the CPU oracle to the reference's own golden values (CPU-only tests).

The reference publishes numbering- and sign-invariant upscaling errors for its
coarse spaces (testsuite/CMakeLists.txt:114-118,156-161,171-176: UpscalingGeneralForm
--form F --nref_parallel 1 on the default cube of 2x2x2 hexahedra).  The oracle's
restatement of DeRhamSequence::Coarsen must reproduce them to the 5 printed digits,
and satisfy the identities of DeRhamSequence::CheckInvariants
(src/amge/DeRhamSequence.cpp:694-970)."""
import numpy as np
import pytest

from oracle import amge, solve as orc

import json
import os

_REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_upscaling_norms.json")))
GOLD = {f: (_REF["form%d" % f]["u_error"], _REF["form%d" % f]["du_error"]) for f in (0, 1, 2)}


@pytest.mark.parametrize("form", [0, 1, 2])
def test_upscaling_goldens(form):
    e_l2, e_en, seqs = amge.upscaling_errors(form, nref=1)
    assert "%.4e" % e_l2 == GOLD[form][0]
    assert "%.4e" % e_en == GOLD[form][1]


def test_check_invariants_three_levels():
    mesh, seqs = amge.build_hierarchy((8, 8, 8), 3)
    for s in seqs[:-1]:
        inv = amge.check_invariants(s)
        for k, v in inv.items():
            assert v < 1e-11, (k, v)
    # lowest-order targets on a structured mesh: exactly one coarse dof per coarse entity
    c = seqs[1]
    assert [c.dof[j].ndofs for j in range(4)] == [5 ** 3, 3 * 4 * 25, 3 * 5 * 16, 64]
    c = seqs[2]
    assert [c.dof[j].ndofs for j in range(4)] == [27, 54, 36, 8]


def test_anisotropic_cells_and_weights_keep_invariants():
    rng = np.random.default_rng(7)
    nel = 4 * 4 * 4
    mesh, seqs = amge.build_hierarchy((4, 4, 4), 2, L=(1.0, 2.0, 0.5),
                                      alpha=rng.uniform(0.5, 2.0, nel), beta=10.0 ** rng.uniform(-2, 2, nel))
    inv = amge.check_invariants(seqs[0])
    for k, v in inv.items():
        # variable coefficients: the constant targets are no longer in the span of the
        # M-harmonic extensions, so NullSpace dofs appear and Pi still reproduces them
        assert v < 1e-9, (k, v)


def test_topology_coarsening_counts_and_exactness():
    mesh = amge.HexMesh(4, 6, 2)
    topo = mesh.topology()
    assert abs(topo.B[0] @ topo.B[1]).max() == 0 and abs(topo.B[1] @ topo.B[2]).max() == 0
    ct = topo.coarsen(amge.refined_partition(mesh.dims))
    ref = amge.HexMesh(2, 3, 1)
    assert ct.n == [ref.nel, sum(ref.nf), sum(ref.ne), ref.nv]
    assert abs(ct.B[0] @ ct.B[1]).max() == 0 and abs(ct.B[1] @ ct.B[2]).max() == 0
    # every coarse boundary facet keeps exactly one attribute
    assert np.all(np.diff(ct.facet_bdr.indptr) <= 1)
    assert ct.facet_bdr.nnz == 2 * (2 * 3 + 3 * 1 + 2 * 1)


def test_solve_oracle_vcycle_converges_on_amge_hierarchy():
    """PCG + AMGe V-cycle (l1-GS smoother, PCG-GS coarse solver) on the H1 problem."""
    mesh, seqs = amge.build_hierarchy((8, 8, 8), 3)
    ess = np.ones(6, dtype=int)
    f = seqs[0]
    A = (f.D[0].T @ f.mass_operator(1) @ f.D[0]).tocsr()
    marker = f.dof[0].mark_bdr_dofs(ess)
    import scipy.sparse as sp
    keep = sp.diags((~marker).astype(float))
    A = (keep @ A @ keep + sp.diags(marker.astype(float))).tocsr()
    Ps = [seqs[l].get_P(0, ess) for l in range(2)]
    H = orc.build_hierarchy(A, Ps, lambda l, Al: orc.Smoother(Al, type=2),
                            lambda Ac: (lambda b, x: orc.pcg(Ac, lambda r: orc.Smoother(Ac, type=2).apply(r, 0 * r, False),
                                                             b, rtol=1e-4, atol=1e-4, max_iter=3)[0]))
    b = np.where(marker, 0.0, 1.0) * f.mass_operator(0).diagonal()
    x, it, conv, hist = orc.pcg(A, H.mult, b, rtol=1e-8, atol=0.0, max_iter=50)
    assert conv and it <= 12
    assert np.linalg.norm(b - A @ x) <= 1e-6 * np.linalg.norm(b)


def test_check_invariants_on_trilinear_hexahedra():
    """DeRhamSequence::CheckInvariants identities (M_c = P^T M P, D D = 0, D P = P D_c, Pi P = I, targets reproduced) on
    the three-level hierarchy of the moved-vertex mesh of examples/3DHdivWeakScaling.cpp, all four forms"""
    mesh, seqs = amge.build_hierarchy((4, 4, 4), 3, jstart=0, deform=amge.weak_scaling_deformation)
    for s in seqs[:-1]:
        for k, v in amge.check_invariants(s).items():
            assert v < 1e-9, (k, v)
