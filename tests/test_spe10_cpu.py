"""The SPE10 permeability data set (BASELINE configs[3], examples/MultigridTestSPE10.cpp): the product's reader and
point evaluation (parelag_b200/src/spe10.hpp, mirror of src/SPE10/InversePermeabilityFunction.cpp) against the oracle's
restatement on a synthetic file in the data set's format, and the diagonal tensor coefficient in the RT0 element mass
matrices of the fine sequence.  CPU only.  (The real data/spe_perm.dat is not in the reference tree.)"""
import numpy as np
import pytest

from oracle import amge, spe10
from parelag_b200 import api


@pytest.fixture(scope="module")
def perm_file(tmp_path_factory):
    rng = np.random.default_rng(10)
    # lognormal, K_x = K_y != K_z like the original data set
    kh = np.exp(rng.normal(0.0, 2.0, size=(spe10.FULL[2], spe10.FULL[1], spe10.FULL[0])))
    kz = kh * np.exp(rng.normal(-2.0, 1.0, size=kh.shape))
    path = tmp_path_factory.mktemp("spe10") / "spe_perm.dat"
    spe10.write_permeability_file(path, np.stack([kh, kh, kz]))
    return path, np.stack([kh, kh, kz])


@pytest.mark.parametrize("N", [(60, 220, 85), (12, 20, 7)])
def test_reader_and_point_evaluation(perm_file, N):
    path, K = perm_file
    h = (20.0, 10.0, 2.0)
    ip = spe10.read_permeability_file(path, *N)
    # the stored array is 1 / K of the Nx x Ny x Nz corner, component after component
    assert np.array_equal(ip, np.concatenate([1.0 / K[c, :N[2], :N[1], :N[0]].ravel() for c in range(3)]))
    api.spe10_read(path, N, h)
    assert np.array_equal(api.spe10_data(), ip)                     # text -> double -> reciprocal: bit for bit
    rng = np.random.default_rng(1)
    x = rng.uniform(0.0, 1.0, size=(500, 3)) * (np.array(N) * np.array(h))
    # cell boundaries and the far corner of the domain (the reference divides by 1 + 3e-16 before the floor)
    x[:50] = np.floor(x[:50] / np.array(h)) * np.array(h)
    x[0] = np.array(N) * np.array(h) * (1.0 - 1e-16)
    x = x[(spe10.cell_of(x, N, h)[0] >= 0) & (spe10.cell_of(x, N, h)[2] >= 0)]
    got, want = api.spe10_inverse_permeability(x), spe10.inverse_permeability(ip, x, N, h)
    assert np.array_equal(got, want)
    # the data set's x and z axes run against the mesh axes
    i, j, k = spe10.cell_of(np.array([[0.5 * h[0], 0.5 * h[1], 0.5 * h[2]]]), N, h)
    assert (int(i[0]), int(j[0]), int(k[0])) == (N[0] - 1, 0, N[2] - 1)


def test_tensor_coefficient_in_the_fine_sequence(perm_file):
    path, K = perm_file
    N, h = (6, 10, 4), (20.0, 10.0, 2.0)
    ip = spe10.read_permeability_file(path, *N)
    kinv = spe10.element_inverse_permeability(ip, N, N, h)
    mesh = amge.HexMesh(*N, L=tuple(n * s for n, s in zip(N, h)))
    seq = amge.fine_sequence(mesh, beta=kinv, jstart=2)
    api.spe10_read(path, N, h)
    for S in (api.Sequence.spe10(N, h, 1, jstart=2, svd_tol=-1.0),
              api.Sequence.hex_tensor(N, 1, kinv, L=tuple(n * s for n, s in zip(N, h)), jstart=2, svd_tol=-1.0)):
        Me, Mo = S.get_csr(0, "Me", 2, 0), seq.M[(2, 0)]
        assert Me.shape == Mo.shape and abs(Me - Mo).max() <= 1e-15 * abs(Mo).max()
        M, Mass = S.get_csr(0, "M", 2), seq.mass_operator(2)
        assert abs(M - Mass).max() <= 1e-14 * abs(Mass).max()
        S.free()
    # an isotropic tensor is the scalar coefficient
    b = np.linspace(0.5, 2.0, mesh.nel)
    A = api.Sequence.hex_tensor(N, 1, np.stack([b, b, b], axis=1), jstart=2, svd_tol=-1.0)
    B = api.Sequence.hex(N, 1, beta=b, jstart=2, svd_tol=-1.0)
    assert abs(A.get_csr(0, "Me", 2, 0) - B.get_csr(0, "Me", 2, 0)).max() == 0
    A.free(); B.free()
