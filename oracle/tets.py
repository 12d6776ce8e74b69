"""Oracle (TEST INFRASTRUCTURE ONLY): lowest-order de Rham sequence on TETRAHEDRAL meshes -- the fine level of
BASELINE configs[0] (examples/MultigridTest0Form.cpp:147-212 on meshes/cube456.mesh: DeRhamSequence3D_FE,
src/amge/DeRhamSequenceFE.cpp:633-722, uniform refinement, MFEMRefinedMeshPartitioner) -- for oracle/amge.py's
geometry-agnostic coarsening.

What is restated here:
  * the NETGEN neutral mesh format of meshes/cube456.mesh (mfem::Mesh reader),
  * uniform (red) refinement of tetrahedra; the children of element e are numbered 8e .. 8e+7, which is the numbering
    src/partitioning/MFEMRefinedMeshPartitioner.cpp:48-66 assumes for MFEM >= 4.1 (partition = element / 8),
  * the topology tables (signed incidence B_0 element-facet with outward orientation, B_1 facet-ridge, B_2 ridge-peak;
    every facet / ridge oriented by ascending vertex number), boundary attributes,
  * Whitney forms: H1 vertex values, Nedelec edge circulations, Raviart-Thomas face fluxes, L2 cell values with
    D_2 = net outward flux / volume (the convention of oracle/amge.py on hexahedra), exact element / facet / ridge mass
    matrices in closed form from the barycentric gradients, order-0 upscaling targets and PV-trace geometry.

MFEM's own element numbering after refinement and its RT dof scaling on tetrahedra are not reproducible offline
(SURVEY 8c): integer parity on tetrahedra is product <-> oracle, and the reference-pinned quantities are the
CheckInvariants identities (DeRhamSequence.cpp:694-970), which do not depend on numbering or scaling.
"""
import numpy as np
import scipy.sparse as sp

from . import amge


def read_netgen_neutral(path):
    """vertices (nv,3), tets (ne,4) 0-based, boundary triangles (nb,3) 0-based, boundary attributes (nb,) 1-based"""
    tok = open(path).read().split()
    assert tok[0] == "NETGEN_Neutral_Format"
    p = 1
    nv = int(tok[p]); p += 1
    V = np.array(tok[p:p + 3 * nv], dtype=np.float64).reshape(nv, 3); p += 3 * nv
    ne = int(tok[p]); p += 1
    E = np.array(tok[p:p + 5 * ne], dtype=np.int64).reshape(ne, 5); p += 5 * ne
    nb = int(tok[p]); p += 1
    B = np.array(tok[p:p + 4 * nb], dtype=np.int64).reshape(nb, 4)
    return V, E[:, 1:] - 1, B[:, 1:] - 1, B[:, 0].copy()


def write_mfem_mesh(path, V, T, B, A, element_attribute=1):
    """MFEM mesh v1.0 (0-based vertex numbers; geometry type 4 = tetrahedron, 2 = triangle), with the comment block MFEM writes"""
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\n#\n# MFEM Geometry Types (see mesh/geom.hpp):\n#\n# POINT       = 0\n# SEGMENT     = 1\n"
                "# TRIANGLE    = 2\n# SQUARE      = 3\n# TETRAHEDRON = 4\n# CUBE        = 5\n#\n\ndimension\n3\n\nelements\n%d\n" % len(T))
        for t in T:
            f.write("%d 4 %d %d %d %d\n" % ((element_attribute,) + tuple(int(v) for v in t)))
        f.write("\nboundary\n%d\n" % len(B))
        for a, t in zip(A, B):
            f.write("%d 2 %d %d %d\n" % ((int(a),) + tuple(int(v) for v in t)))
        f.write("\nvertices\n%d\n3\n" % len(V))
        for x in V:
            f.write("%.17g %.17g %.17g\n" % tuple(x))


def read_mfem_mesh(path):
    """MFEM mesh v1.0, straight-sided tetrahedra: the same four arrays as read_netgen_neutral"""
    lines = open(path).read().split("\n")
    assert lines[0].startswith("MFEM mesh v1.0")
    tok = " ".join(l.split("#")[0] for l in lines[1:]).split()
    p = 0
    assert tok[p] == "dimension" and tok[p + 1] == "3"; p += 2
    assert tok[p] == "elements"; ne = int(tok[p + 1]); p += 2
    E = np.array(tok[p:p + 6 * ne], dtype=np.int64).reshape(ne, 6); p += 6 * ne
    assert np.all(E[:, 1] == 4)
    assert tok[p] == "boundary"; nb = int(tok[p + 1]); p += 2
    Bd = np.array(tok[p:p + 5 * nb], dtype=np.int64).reshape(nb, 5); p += 5 * nb
    assert np.all(Bd[:, 1] == 2)
    assert tok[p] == "vertices"; nv = int(tok[p + 1]); assert tok[p + 2] == "3"; p += 3
    V = np.array(tok[p:p + 3 * nv], dtype=np.float64).reshape(nv, 3)
    return V, E[:, 2:], Bd[:, 2:], Bd[:, 0].copy()


def load_npz(path):
    """the arrays of tests/golden/cube456.npz (written by tests/golden/make_cube456.py from the reference's mesh file)"""
    d = np.load(path)
    return d["vertices"], d["tets"].astype(np.int64), d["bdr_triangles"].astype(np.int64), d["bdr_attributes"].astype(np.int64)


def write_netgen_neutral(path, V, T, B, A, element_attribute=1):
    """the inverse of read_netgen_neutral (1-based vertex numbers)"""
    with open(path, "w") as f:
        f.write("NETGEN_Neutral_Format\n%d\n" % len(V))
        for x in V:
            f.write("  %.17g  %.17g  %.17g\n" % tuple(x))
        f.write("%d\n" % len(T))
        for t in T:
            f.write("   %d  %d %d %d %d\n" % ((element_attribute,) + tuple(int(v) + 1 for v in t)))
        f.write("%d\n" % len(B))
        for a, t in zip(A, B):
            f.write("   %d  %d %d %d\n" % ((int(a),) + tuple(int(v) + 1 for v in t)))


def cube_tets(n=1):
    """n x n x n cubes of the unit cube, each cut into 6 tetrahedra around the main diagonal (Kuhn); boundary
    attributes as mfem::Mesh::Make3D (z=0:1, y=0:2, x=1:3, y=1:4, x=0:5, z=1:6)"""
    g = np.arange(n + 1) / n
    k, j, i = np.meshgrid(g, g, g, indexing="ij")
    V = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    vid = lambda a, b, c: a + (n + 1) * (b + (n + 1) * c)
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    T = []
    for c in range(n):
        for b in range(n):
            for a in range(n):
                for pm in perms:
                    p = [a, b, c]
                    t = [vid(*p)]
                    for ax in pm:
                        p[ax] += 1
                        t.append(vid(*p))
                    T.append(t)
    T = np.array(T, dtype=np.int64)
    faces = {}
    for t in T:
        for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
            key = tuple(sorted(int(t[q]) for q in f))
            faces[key] = faces.get(key, 0) + 1
    Bt, Ba = [], []
    for key, cnt in faces.items():
        if cnt != 1:
            continue
        X = V[list(key)]
        for ax, val, attr in ((2, 0.0, 1), (1, 0.0, 2), (0, 1.0, 3), (1, 1.0, 4), (0, 0.0, 5), (2, 1.0, 6)):
            if np.all(np.abs(X[:, ax] - val) < 1e-12):
                Bt.append(key); Ba.append(attr)
                break
    return V, T, np.array(Bt, dtype=np.int64), np.array(Ba, dtype=np.int64)


class TetMesh:
    def __init__(self, V, T, Btri, Battr):
        self.V = np.asarray(V, dtype=np.float64)
        self.T = np.sort(np.asarray(T, dtype=np.int64), axis=1)        # local order = ascending vertex number
        self.Btri = np.sort(np.asarray(Btri, dtype=np.int64), axis=1)
        self.Battr = np.asarray(Battr, dtype=np.int64)
        self.nel, self.nv = len(self.T), len(self.V)
        self._build_entities()

    # ---- entities: faces / edges numbered in lexicographic order of their (sorted) vertex tuples
    def _build_entities(self):
        T = self.T
        fl = np.concatenate([T[:, [0, 1, 2]], T[:, [0, 1, 3]], T[:, [0, 2, 3]], T[:, [1, 2, 3]]])
        self.F, inv = np.unique(fl, axis=0, return_inverse=True)
        self.el_face = inv.reshape(4, self.nel).T            # columns: faces (012), (013), (023), (123)
        el = np.concatenate([T[:, [0, 1]], T[:, [0, 2]], T[:, [0, 3]], T[:, [1, 2]], T[:, [1, 3]], T[:, [2, 3]]])
        self.E, inv = np.unique(el, axis=0, return_inverse=True)
        self.el_edge = inv.reshape(6, self.nel).T            # columns: 01, 02, 03, 12, 13, 23
        self.nf, self.ne = len(self.F), len(self.E)
        ekey = {(int(a), int(b)): q for q, (a, b) in enumerate(self.E)}
        F = self.F
        self.face_edge = np.array([[ekey[(a, b)], ekey[(a, c)], ekey[(b, c)]] for a, b, c in F.tolist()], dtype=np.int64)
        fkey = {tuple(f): q for q, f in enumerate(F.tolist())}
        self.bdr_face = np.array([fkey[tuple(t)] for t in self.Btri.tolist()], dtype=np.int64)
        X = self.V[T]
        self.det = np.linalg.det(np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=1))
        self.vol = np.abs(self.det) / 6.0
        XF = self.V[F]
        N = 0.5 * np.cross(XF[:, 1] - XF[:, 0], XF[:, 2] - XF[:, 0])   # area vector of the orientation (a, b, c), a < b < c
        # Facet orientation: ascending vertex order, except that BOUNDARY facets point outward (as in an mfem::Mesh,
        # whose boundary faces inherit the orientation of their only element): the topology coarsening adds the unsigned
        # boundary-attribute coupling to the signed AE-facet products (Topology.cpp:735-748), which only groups the
        # facets of a boundary patch when their signs agree.
        s = np.sign(self.det)
        out012, out013, out023, out123 = -s, s, -s, s                  # outward sign of the element's faces, cf. topology()
        self.fsign = np.ones(self.nf)
        cnt = np.bincount(self.el_face.ravel(), minlength=self.nf)
        for col, o in enumerate((out012, out013, out023, out123)):
            f = self.el_face[:, col]
            b = cnt[f] == 1
            self.fsign[f[b]] = o[b]
        self.N = N * self.fsign[:, None]
        self.tvec = self.V[self.E[:, 1]] - self.V[self.E[:, 0]]

    def vertex_coords(self):
        return self.V

    def facet_area(self):
        return np.linalg.norm(self.N, axis=1)

    def ridge_length(self):
        return np.linalg.norm(self.tvec, axis=1)

    # ---- uniform refinement (children of e: 8e .. 8e+7; new vertex of edge q: nv + q)
    def refine(self):
        m = lambda a, b: self.nv + self.el_edge[:, {(0, 1): 0, (0, 2): 1, (0, 3): 2, (1, 2): 3, (1, 3): 4, (2, 3): 5}[(a, b)]]
        v = [self.T[:, q] for q in range(4)]
        m01, m02, m03, m12, m13, m23 = m(0, 1), m(0, 2), m(0, 3), m(1, 2), m(1, 3), m(2, 3)
        kids = [(v[0], m01, m02, m03), (m01, v[1], m12, m13), (m02, m12, v[2], m23), (m03, m13, m23, v[3]),
                (m01, m02, m03, m13), (m01, m02, m12, m13), (m02, m03, m13, m23), (m02, m12, m13, m23)]
        Tn = np.stack([np.stack(k, axis=1) for k in kids], axis=1).reshape(-1, 4)     # (nel, 8, 4) -> 8e + c
        Vn = np.concatenate([self.V, 0.5 * (self.V[self.E[:, 0]] + self.V[self.E[:, 1]])])
        fe = self.face_edge[self.bdr_face]                     # edges ab, ac, bc of every boundary triangle
        a, b, c = self.Btri[:, 0], self.Btri[:, 1], self.Btri[:, 2]
        mab, mac, mbc = self.nv + fe[:, 0], self.nv + fe[:, 1], self.nv + fe[:, 2]
        Bn = np.stack([np.stack(t, axis=1) for t in ((a, mab, mac), (mab, b, mbc), (mac, mbc, c), (mab, mac, mbc))], axis=1).reshape(-1, 3)
        return TetMesh(Vn, Tn, Bn, np.repeat(self.Battr, 4))

    # ---- topology
    def topology(self):
        nel, nf, ne, nv = self.nel, self.nf, self.ne, self.nv
        s = np.sign(self.det)
        r = np.repeat(np.arange(nel), 4)
        # boundary of [v0 v1 v2 v3]: +(123) -(023) +(013) -(012); columns of el_face: (012), (013), (023), (123)
        sg = np.stack([-s, s, -s, s], axis=1) * self.fsign[self.el_face]
        B0 = sp.csr_matrix((sg.ravel(), (r, self.el_face.ravel())), shape=(nel, nf))
        r = np.repeat(np.arange(nf), 3)
        B1 = sp.csr_matrix(((np.tile([1.0, -1.0, 1.0], nf).reshape(nf, 3) * self.fsign[:, None]).ravel(),
                            (r, self.face_edge.ravel())), shape=(nf, ne))                                   # +-(ab - ac + bc)
        r = np.repeat(np.arange(ne), 2)
        B2 = sp.csr_matrix((np.tile([-1.0, 1.0], ne), (r, self.E.ravel())), shape=(ne, nv))
        nattr = int(self.Battr.max()) if len(self.Battr) else 1
        fb = sp.csr_matrix((np.ones(len(self.bdr_face)), (self.bdr_face, self.Battr - 1)), shape=(nf, nattr))
        return amge.Topology([B0, B1, B2], fb, 3)

    # ---- local matrices (vectorised over the entities)
    @staticmethod
    def _I(measure, n, denom):
        """integral of lambda_a lambda_b over a simplex with n vertices: measure (1 + delta_ab) / denom"""
        return measure[:, None, None] * (np.ones((n, n)) + np.eye(n))[None] / denom

    def _grads(self):
        """gradients of the barycentric coordinates of every tetrahedron: (nel, 4, 3)"""
        X = self.V[self.T]
        A = np.concatenate([np.ones((self.nel, 4, 1)), X], axis=2)       # rows [1 x y z]
        return np.linalg.inv(A)[:, 1:, :].transpose(0, 2, 1)

    def h1_element_mass(self):
        return self._I(self.vol, 4, 20.0)

    def nd_element_mass(self):
        g, I = self._grads(), self._I(self.vol, 4, 20.0)
        G = np.einsum("eia,eja->eij", g, g)
        ed = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        M = np.empty((self.nel, 6, 6))
        for p, (i, j) in enumerate(ed):
            for q, (k, l) in enumerate(ed):
                M[:, p, q] = I[:, i, k] * G[:, j, l] - I[:, i, l] * G[:, j, k] - I[:, j, k] * G[:, i, l] + I[:, j, l] * G[:, i, k]
        return M

    def rt_element_mass(self):
        g, I = self._grads(), self._I(self.vol, 4, 20.0)
        fc = [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
        # w_f = 2 (l_a g_b x g_c + l_b g_c x g_a + l_c g_a x g_b): coefficient vector of l_v in w_f
        C = np.zeros((self.nel, 4, 4, 3))
        for p, (a, b, c) in enumerate(fc):
            C[:, p, a] = 2.0 * np.cross(g[:, b], g[:, c])
            C[:, p, b] = 2.0 * np.cross(g[:, c], g[:, a])
            C[:, p, c] = 2.0 * np.cross(g[:, a], g[:, b])
        o = self.fsign[self.el_face]                  # basis function of a re-oriented (boundary) facet changes sign
        return np.einsum("epax,eqbx,eab->epq", C, C, I) * o[:, :, None] * o[:, None, :]

    def _face_grads(self):
        """surface gradients of the barycentric coordinates of every triangle: (nf, 3, 3)"""
        X = self.V[self.F]
        e1, e2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
        g11, g12, g22 = np.einsum("fa,fa->f", e1, e1), np.einsum("fa,fa->f", e1, e2), np.einsum("fa,fa->f", e2, e2)
        dt = g11 * g22 - g12 * g12
        gb = (g22[:, None] * e1 - g12[:, None] * e2) / dt[:, None]
        gc = (g11[:, None] * e2 - g12[:, None] * e1) / dt[:, None]
        return np.stack([-gb - gc, gb, gc], axis=1)

    def h1_facet_mass(self):
        return self._I(self.facet_area(), 3, 12.0)

    def nd_facet_mass(self):
        g, I = self._face_grads(), self._I(self.facet_area(), 3, 12.0)
        G = np.einsum("fia,fja->fij", g, g)
        ed = [(0, 1), (0, 2), (1, 2)]
        M = np.empty((self.nf, 3, 3))
        for p, (i, j) in enumerate(ed):
            for q, (k, l) in enumerate(ed):
                M[:, p, q] = I[:, i, k] * G[:, j, l] - I[:, i, l] * G[:, j, k] - I[:, j, k] * G[:, i, l] + I[:, j, l] * G[:, i, k]
        return M

    def h1_ridge_mass(self):
        return self._I(self.ridge_length(), 2, 6.0)


def _blocks(M):
    return sp.block_diag(list(M), format="csr")


def fine_sequence_tet(mesh, topo=None, alpha=None, beta=None, jstart=0):
    """DeRhamSequence3D_FE at lowest order on tetrahedra (DeRhamSequenceFE.cpp:633-684) + order-0 upscaling targets
    (SetUpscalingTargets, :927-982).  alpha / beta: per-element weights of the L2 and H(div) mass matrices."""
    topo = topo or mesh.topology()
    seq = amge.Sequence(topo, 4)
    seq.mesh = mesh
    seq.jstart = jstart
    for j in range(4):
        dh = amge.DofHandler(3 - j, topo)
        dh.ndofs = topo.n[3 - j]
        for c in range(3 - j + 1):
            dh.entity_dof[c] = sp.identity(topo.n[c], format="csr") if c == 3 - j else topo.conn(c, 3 - j)
        seq.dof[j] = dh
    a_el = np.ones(mesh.nel) if alpha is None else np.asarray(alpha, dtype=np.float64)
    b_el = np.ones(mesh.nel) if beta is None else np.asarray(beta, dtype=np.float64)
    seq.D = [topo.B[2].copy(), topo.B[1].copy(), amge._canon(sp.diags(1.0 / mesh.vol) @ topo.B[0])]
    seq.M[(3, 0)] = sp.diags(mesh.vol * a_el).tocsr()
    seq.M[(2, 0)] = _blocks(mesh.rt_element_mass() * b_el[:, None, None])
    seq.M[(2, 1)] = sp.diags(1.0 / mesh.facet_area()).tocsr()
    seq.l2_const = np.ones(mesh.nel)
    seq.targets[3] = np.ones((mesh.nel, 1))
    seq.targets[2] = mesh.N.copy()                 # fluxes of e_x, e_y, e_z
    if jstart <= 1:
        seq.M[(1, 0)] = _blocks(mesh.nd_element_mass())
        seq.M[(1, 1)] = _blocks(mesh.nd_facet_mass())
        seq.M[(1, 2)] = sp.diags(1.0 / mesh.ridge_length()).tocsr()
        seq.targets[1] = mesh.tvec.copy()          # circulations of e_x, e_y, e_z
    if jstart <= 0:
        seq.M[(0, 0)] = _blocks(mesh.h1_element_mass())
        seq.M[(0, 1)] = _blocks(mesh.h1_facet_mass())
        seq.M[(0, 2)] = _blocks(mesh.h1_ridge_mass())
        seq.M[(0, 3)] = sp.identity(mesh.nv, format="csr")
        X = mesh.V
        seq.targets[0] = np.stack([np.ones(mesh.nv), X[:, 2], X[:, 1], X[:, 0]], axis=1)    # 1, z, y, x
    return seq


def build_hierarchy(mesh0, nref, nlevels, alpha=None, beta=None, jstart=0, svd_tol=1e-9):
    """MultigridTest0Form.cpp:147-375: refine mesh0 nref times, agglomerate back nlevels-1 times by derefinement
    (partition = element / 8), Coarsen() level by level.  Returns (finest mesh, [sequences])."""
    assert nlevels - 1 <= nref
    mesh = mesh0
    for _ in range(nref):
        mesh = mesh.refine()
    topos = [mesh.topology()]
    n = mesh.nel
    for _ in range(nlevels - 1):
        topos.append(topos[-1].coarsen(np.arange(n) // 8))
        n //= 8
    seqs = [fine_sequence_tet(mesh, topos[0], alpha=alpha, beta=beta, jstart=jstart)]
    for l in range(nlevels - 1):
        seqs[l].svd_tol = svd_tol
        seqs.append(seqs[l].coarsen())
    return mesh, seqs
