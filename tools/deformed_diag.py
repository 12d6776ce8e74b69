"""Diagnostic for the trilinear-hexahedra Coarsen() check: per level and form, pattern equality and
value differences product vs oracle, raw and after aligning the sign of every coarse dof (column of P)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parelag_b200 import api
from oracle import amge

api.session()
dims, nlev = (4, 4, 4), 3
mesh, seqs = amge.build_hierarchy(dims, nlev, jstart=2, deform=amge.weak_scaling_deformation)
S = api.Sequence.hex(dims, nlev, jstart=2, coords=mesh.vertex_coords())
for l in range(nlev - 1):
    f, c = seqs[l], seqs[l + 1]
    for j in range(f.jstart, 4):
        for cd in range(4 - j):
            Eg, Eo = S.get_csr(l + 1, "ED", j, cd), c.dof[j].entity_dof[cd].tocsr()
            same = Eg.shape == Eo.shape and np.array_equal(Eg.indptr, Eo.indptr) and np.array_equal(Eg.indices, Eo.indices)
            print("L%d form %d codim %d entity_dof same=%s shape %s vs %s" % (l, j, cd, same, Eg.shape, Eo.shape))
        P, Po = S.get_csr(l, "P", j), f.P[j].tocsr()
        print("L%d form %d P shapes %s %s" % (l, j, P.shape, Po.shape))
        if P.shape != Po.shape:
            continue
        same = np.array_equal(P.indptr, Po.indptr) and np.array_equal(P.indices, Po.indices)
        d = abs(P - Po)
        Pd, Pod = P.toarray(), Po.toarray()
        sg = np.ones(P.shape[1])
        bad = []
        for k in range(P.shape[1]):
            a, b = Pd[:, k], Pod[:, k]
            if np.abs(a + b).max() < np.abs(a - b).max():
                sg[k] = -1
            e = np.abs(sg[k] * a - b).max()
            if e > 1e-10 * np.abs(Pod).max():
                bad.append((k, e))
        print("   pattern same=%s raw maxdiff %.3e  flipped columns %d  after sign alignment: %d bad columns %s"
              % (same, d.max(), int((sg < 0).sum()), len(bad), bad[:8]))
S.free()
