"""Host integer tables of the product (C++: amge_topology.hpp / amge_dofs.hpp /
amge_hex.hpp) against the oracle -- bit-exact, no GPU needed."""
import numpy as np
import pytest

from parelag_b200 import api
from oracle import amge


def same(A, B):
    A.sort_indices(); B = B.tocsr(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data))


# the last case is large enough (> 4096 rows per table) for the row-parallel host SpGEMM to use several threads
# (6, 5, 3), (15, 11, 7): not multiples of two -- logical Cartesian agglomeration with ragged last blocks
# (LogicalPartitioner.hpp:46-103; the shape of the 60 x 220 x 85 SPE10 grid, whose third level is 15 x 55 x 22)
@pytest.mark.parametrize("dims,nlev", [((4, 4, 4), 3), ((4, 6, 2), 2), ((8, 4, 4), 3), ((24, 16, 16), 3), ((6, 5, 3), 3),
                                       ((15, 11, 7), 4)])
def test_topology_hierarchy_bit_exact(dims, nlev):
    S = api.Sequence.hex(dims, nlev, svd_tol=-1.0)
    mesh = amge.HexMesh(*dims)
    topo = mesh.topology()
    d = dims
    for l in range(nlev):
        for c in range(3):
            assert same(S.get_csr(l, "B", c), topo.B[c]), (l, c)
        assert same(S.get_csr(l, "FB"), topo.facet_bdr)
        if l + 1 < nlev:
            ctopo = topo.coarsen(amge.refined_partition(d))
            for c in range(4):
                assert same(S.get_csr(l, "AE", c), topo.AE_entity[c]), (l, c)
            topo = ctopo
            d = amge.coarse_dims(d)
    S.free()


def test_fine_sequence_matrices_bit_exact():
    dims, L = (4, 2, 6), (1.0, 2.0, 0.75)
    rng = np.random.default_rng(3)
    nel = 4 * 2 * 6
    alpha, beta = rng.uniform(0.5, 2, nel), rng.uniform(0.1, 10, nel)
    S = api.Sequence.hex(dims, 1, L=L, alpha=alpha, beta=beta, svd_tol=-1.0)
    mesh = amge.HexMesh(*dims, L=L)
    seq = amge.fine_sequence(mesh, alpha=alpha, beta=beta)
    for j in range(3):
        assert same(S.get_csr(0, "D", j), seq.D[j])
    for j in range(4):
        for c in range(4 - j):
            Me = S.get_csr(0, "Me", j, c)
            assert abs(Me - seq.M[(j, c)]).max() <= 1e-15 * abs(seq.M[(j, c)]).max(), (j, c)
            assert same(S.get_csr(0, "ED", j, c), seq.dof[j].entity_dof[c])
        M = S.get_csr(0, "M", j)
        Mo = seq.mass_operator(j)
        assert abs(M - Mo).max() <= 1e-14 * abs(Mo).max()
        assert np.array_equal(S.get_targets(0, j), seq.targets[j])
        from oracle import drivers
        assert np.array_equal(S.get_bdr_mask(0, j), drivers.bdr_mask(seq.dof[j]))
    S.free()


def test_deformed_fine_sequence_matches_oracle():
    """Trilinear hexahedra (the 3DHdivWeakScaling vertex map, examples/3DHdivWeakScaling.cpp:148-158): the host-built
    H(div)-L2 part of the fine sequence -- cell volumes, RT0 mass matrices by mfem's 3-point Gauss rule, facet trace
    masses, D_2 = flux / volume, flux targets, boundary masks -- against the oracle, which reproduces the reference's
    golden on this mesh (tests/test_goldens_cpu.py).  Tolerance 1e-14 relative: same formulas, different summation order."""
    dims = (4, 4, 4)
    rng = np.random.default_rng(5)
    nel = 64
    alpha, beta = rng.uniform(0.5, 2, nel), rng.uniform(0.1, 10, nel)
    mesh = amge.DeformedHexMesh(*dims, deform=amge.weak_scaling_deformation)
    seq = amge.fine_sequence(mesh, alpha=alpha, beta=beta, jstart=2)
    S = api.Sequence.hex(dims, 1, alpha=alpha, beta=beta, jstart=2, svd_tol=-1.0, coords=mesh.vertex_coords())
    for j in (2,):
        D, Do = S.get_csr(0, "D", j), seq.D[j].tocsr()
        Do.sort_indices()
        assert np.array_equal(D.indptr, Do.indptr) and np.array_equal(D.indices, Do.indices)
        assert abs(D - Do).max() <= 1e-14 * abs(Do).max()
    for (j, c) in ((3, 0), (2, 0), (2, 1)):
        Me, Mo = S.get_csr(0, "Me", j, c), seq.M[(j, c)]
        assert Me.shape == Mo.shape and abs(Me - Mo).max() <= 1e-14 * abs(Mo).max(), (j, c)
    for j in (2, 3):
        M, Mo = S.get_csr(0, "M", j), seq.mass_operator(j)
        assert abs(M - Mo).max() <= 1e-14 * abs(Mo).max()
        assert np.abs(S.get_targets(0, j) - seq.targets[j]).max() <= 1e-14 * np.abs(seq.targets[j]).max()
        from oracle import drivers
        assert np.array_equal(S.get_bdr_mask(0, j), drivers.bdr_mask(seq.dof[j]))
    S.free()
    # with jformStart = 1 also the Nedelec part: element / facet / edge mass matrices, circulation targets
    seq = amge.fine_sequence(mesh, jstart=1)
    S = api.Sequence.hex(dims, 1, jstart=1, svd_tol=-1.0, coords=mesh.vertex_coords())
    for (j, c) in ((1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (3, 0)):
        Me, Mo = S.get_csr(0, "Me", j, c), seq.M[(j, c)]
        assert Me.shape == Mo.shape and abs(Me - Mo).max() <= 1e-13 * abs(Mo).max(), (j, c)
    M, Mo = S.get_csr(0, "M", 1), seq.mass_operator(1)
    assert abs(M - Mo).max() <= 1e-13 * abs(Mo).max()
    assert np.abs(S.get_targets(0, 1) - seq.targets[1]).max() <= 1e-14 * np.abs(seq.targets[1]).max()
    D1, D1o = S.get_csr(0, "D", 1), seq.D[1].tocsr()
    assert abs(D1 - D1o).max() == 0
    # the RT0 mass matrices of moved cells couple all six faces (no longer block diagonal per axis)
    blk = S.get_csr(0, "Me", 2, 0)[:6, :6].toarray()
    assert abs(blk[0, 2]) > 1e-3 * abs(blk[0, 0])
    S.free()
    # all four forms (jformStart = 0): H1 cell / facet / edge / vertex mass matrices, targets 1, z, y, x
    seq = amge.fine_sequence(mesh, jstart=0)
    S = api.Sequence.hex(dims, 1, jstart=0, svd_tol=-1.0, coords=mesh.vertex_coords())
    for (j, c) in ((0, 0), (0, 1), (0, 2), (0, 3)):
        Me, Mo = S.get_csr(0, "Me", j, c), seq.M[(j, c)]
        assert Me.shape == Mo.shape and abs(Me - Mo).max() <= 1e-13 * abs(Mo).max(), (j, c)
    assert np.abs(S.get_targets(0, 0) - seq.targets[0]).max() <= 1e-14
    S.free()


def test_identity_vertex_map_reproduces_the_closed_form_pools():
    """Quadrature path with unmoved vertices == closed-form path (anisotropic box): checks every local ordering and
    orientation convention of the trilinear-hexahedra code against the structured one, in the product and in the oracle."""
    dims, L = (4, 2, 6), (1.0, 2.0, 0.75)
    ref = api.Sequence.hex(dims, 1, L=L, svd_tol=-1.0)
    X = amge.HexMesh(*dims, L=L).vertex_coords()
    S = api.Sequence.hex(dims, 1, svd_tol=-1.0, coords=X)
    so = amge.fine_sequence(amge.DeformedHexMesh(*dims, deform=lambda Y: Y, L=L), jstart=0)
    sr = amge.fine_sequence(amge.HexMesh(*dims, L=L), jstart=0)
    for j in range(4):
        for c in range(4 - j):
            A, B = S.get_csr(0, "Me", j, c), ref.get_csr(0, "Me", j, c)
            assert abs(A - B).max() <= 1e-13 * abs(B).max(), (j, c)
            assert abs(so.M[(j, c)] - sr.M[(j, c)]).max() <= 1e-13 * abs(sr.M[(j, c)]).max(), (j, c)
        assert np.abs(S.get_targets(0, j) - ref.get_targets(0, j)).max() <= 1e-14
        assert np.abs(so.targets[j] - sr.targets[j]).max() <= 1e-14
    for j in range(3):
        assert abs(S.get_csr(0, "D", j) - ref.get_csr(0, "D", j)).max() <= 1e-13 * abs(ref.get_csr(0, "D", j)).max()
    S.free(); ref.free()


def test_host_tables_do_not_depend_on_the_thread_count():
    """hostcsr::Mult / DofAgglomeration / pool fill run under OpenMP with per-thread buffers stitched in row
    order: the bit-exact table tests above must also pass single-threaded and with more threads than cores."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for nt in ("1", "5"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.join(root, "tests", "test_topology_cpu.py"),
                            "-k", "bit_exact"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "OMP_NUM_THREADS=%s\n%s" % (nt, r.stdout[-2000:])
