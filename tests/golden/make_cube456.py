"""Writes tests/golden/cube456.npz: the input mesh of BASELINE configs[0] (the reference's meshes/cube456.mesh, NETGEN
neutral format: 141 vertices, 456 tetrahedra, 206 boundary triangles with attributes 1-6) as arrays, so that the GPU box
-- where /root/reference does not exist -- can run the configuration.  Run here:
    python tests/golden/make_cube456.py /root/reference/meshes/cube456.mesh"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tets   # noqa: E402

if __name__ == "__main__":
    V, T, B, A = tets.read_netgen_neutral(sys.argv[1])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cube456.npz")
    np.savez_compressed(out, vertices=V, tets=T.astype(np.int32), bdr_triangles=B.astype(np.int32), bdr_attributes=A.astype(np.int32))
    print(out, V.shape, T.shape, B.shape, np.bincount(A))
