"""DeRhamSequence::CheckInvariants of the product (DeRhamSequence.cpp:694-970: CheckD, CheckDP, CheckCoarseMassMatrix) on the
CPU: on sequences that hold supplied operators (the oracle's P and D of a three-level hierarchy, then deliberately broken
copies) and on the fine levels the product builds itself (hexahedra, trilinear hexahedra, tetrahedra)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import amge, tets
from parelag_b200 import api, capi


@pytest.fixture(scope="module")
def hierarchy():
    return amge.build_hierarchy((4, 4, 4), 3)[1]


def supplied(seqs, edit=None):
    S = api.Sequence(4, len(seqs))
    for l, s in enumerate(seqs):
        for j in range(3):
            D = s.D[j].tocsr().copy()
            if edit:
                D = edit("D", l, j, D)
            S.set_D(l, j, D)
        if l + 1 < len(seqs):
            for j in range(4):
                P = s.P[j].tocsr().copy()
                if edit:
                    P = edit("P", l, j, P)
                S.set_P(l, j, P)
    return S


def test_supplied_operators_pass_and_broken_ones_are_named(hierarchy):
    S = supplied(hierarchy)
    for l in range(3):
        assert S.check_invariants(l) <= 1e-12
    S.free()

    def scale_one_P_entry(kind, l, j, M):
        if (kind, l, j) == ("P", 0, 1):
            M.data[7] += 1e-3
        return M
    S = supplied(hierarchy, scale_one_P_entry)
    with pytest.raises(capi.PEError, match=r"D_\{0,fine\}\*P_0|D_\{1,fine\}\*P_1"):
        S.check_invariants(0)
    assert S.check_invariants(1) <= 1e-12           # the next level is untouched
    S.free()

    def break_the_complex(kind, l, j, M):
        if (kind, l, j) == ("D", 1, 1):
            M.data[3] = 5.0
        return M
    S = supplied(hierarchy, break_the_complex)
    with pytest.raises(capi.PEError, match=r"\|\|D_[12] \* D_[01]\|\| ="):
        S.check_invariants(1)
    S.free()

    def empty_D(kind, l, j, M):
        return sp.csr_matrix(M.shape) if (kind, l, j) == ("D", 2, 0) else M
    S = supplied(hierarchy, empty_D)
    with pytest.raises(capi.PEError, match=r"nnz\(D_0\) = 0"):
        S.check_invariants(2)
    S.free()


def test_fine_levels_built_by_the_product_form_complexes():
    S = api.Sequence.hex((4, 3, 5), 1, L=(1.0, 2.0, 0.5), svd_tol=-1.0)
    assert S.check_invariants(0) == 0.0
    S.free()
    X = amge.DeformedHexMesh(4, 4, 4, deform=amge.weak_scaling_deformation).vertex_coords()
    S = api.Sequence.hex((4, 4, 4), 1, svd_tol=-1.0, coords=X)
    assert S.check_invariants(0) <= 1e-12           # D_2 = flux / volume: rounding only
    S.free()
    S = api.Sequence.tet(*tets.cube_tets(2), 1, 1, svd_tol=-1.0)
    assert S.check_invariants(0) <= 1e-12
    S.free()
