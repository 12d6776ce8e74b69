"""Oracle (TEST INFRASTRUCTURE ONLY): the SPE10 permeability data set as the reference reads and evaluates it
(src/SPE10/InversePermeabilityFunction.cpp; used by examples/MultigridTestSPE10.cpp:85-187,377-395 -- BASELINE configs[3]).
The data file itself (data/spe_perm.dat) is not part of the reference tree and cannot be fetched here; the tests write
synthetic files in its format.

File layout (InversePermeabilityFunction.cpp:57-100): three blocks K_x, K_y, K_z, each 60 x 220 x 85 numbers, x fastest,
then y, then z, whitespace separated.  The reader cuts the Nx x Ny x Nz corner out of every block and stores 1 / K,
index Nx Ny k + Nx j + i + component Nx Ny Nz.
Evaluation (:141-178): i = Nx - 1 - floor(x / hx / (1 + 3e-16)), j = floor(y / hy / (1 + 3e-16)),
k = Nz - 1 - floor(z / hz / (1 + 3e-16)): the data set's x and z axes run against the mesh axes."""
import numpy as np

FULL = (60, 220, 85)


def write_permeability_file(path, K):
    """K: array (3, 85, 220, 60) = (component, z, y, x) of permeabilities; six numbers per line like the original file"""
    K = np.asarray(K, dtype=np.float64)
    assert K.shape == (3, FULL[2], FULL[1], FULL[0])
    flat = K.reshape(-1, 6)
    with open(path, "w") as f:
        np.savetxt(f, flat, fmt="%.17g")


def read_permeability_file(path, Nx=60, Ny=220, Nz=85):
    """ReadPermeabilityFile: the stored inverse permeability, flat array of 3 Nx Ny Nz numbers"""
    vals = np.array(open(path).read().split(), dtype=np.float64)
    n_full = FULL[0] * FULL[1] * FULL[2]
    out = []
    pos = 0
    for comp in range(3):
        # the reader skips the rest of K_x and K_y, and stops after the needed part of K_z
        block = vals[comp * n_full:(comp + 1) * n_full] if comp < 2 else vals[2 * n_full:2 * n_full + FULL[0] * FULL[1] * Nz]
        nz_have = len(block) // (FULL[0] * FULL[1])
        assert nz_have >= Nz, "file ends early"
        B = block[:nz_have * FULL[0] * FULL[1]].reshape(nz_have, FULL[1], FULL[0])
        out.append(1.0 / B[:Nz, :Ny, :Nx].ravel())
    return np.concatenate(out)


def cell_of(x, N, h):
    """the (i, j, k) of the data cell InversePermeability reads at the points x (n, 3)"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    g = 1.0 + 3e-16
    i = N[0] - 1 - np.floor(x[:, 0] / h[0] / g).astype(np.int64)
    j = np.floor(x[:, 1] / h[1] / g).astype(np.int64)
    k = N[2] - 1 - np.floor(x[:, 2] / h[2] / g).astype(np.int64)
    return i, j, k


def inverse_permeability(ip, x, N, h):
    """InversePermeability at the points x: (n, 3) array (1/K_x, 1/K_y, 1/K_z)"""
    i, j, k = cell_of(x, N, h)
    n = N[0] * N[1] * N[2]
    c = N[1] * N[0] * k + N[0] * j + i
    return np.stack([ip[c], ip[c + n], ip[c + 2 * n]], axis=1)


def element_inverse_permeability(ip, dims, N, h):
    """the tensor coefficient of every cell of the driver's mesh (dims cells of size h, x fastest): value at the cell centre"""
    nx, ny, nz = dims
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    c = np.stack([(i.ravel() + 0.5) * h[0], (j.ravel() + 0.5) * h[1], (k.ravel() + 0.5) * h[2]], axis=1)
    return inverse_permeability(ip, c, N, h)
