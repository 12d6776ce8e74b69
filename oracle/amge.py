"""Oracle (TEST INFRASTRUCTURE ONLY) for ParElag's AMGe *coarsen* path.

CPU restatement in numpy/scipy (+ LAPACK dsytrf/dsytrs/dgesvd/dgetrf/dgetrs through
scipy, the same routines the reference calls) of

  * AgglomeratedTopology::CoarsenLocalPartitioning   src/topology/Topology.cpp:685-828
    with findMinimalIntersectionSets               src/structures/minimalIntersectionSet.cpp:43-130
  * MFEMRefinedMeshPartitioner::Partition            src/partitioning/MFEMRefinedMeshPartitioner.cpp:48-90
  * GeometricBoxPartitioner::doPartition             src/partitioning/GeometricBoxPartitioner.cpp:20-79
  * connectedComponents                              src/structures/connectedComponents.cpp:23-87
  * AgglomeratedTopologyCheck, DeAgglomerateBad...   src/topology/AgglomeratedTopologyCheck.cpp:25-316, Topology.cpp:1151-1214
  * DofAgglomeration                                 src/amge/DOFAgglomeration.cpp:33-315,503-645
  * DofHandlerALG numbering / tables                 src/amge/DofHandler.cpp:694-1463
  * DeRhamSequence::Coarsen and helpers              src/amge/DeRhamSequence.cpp:572-692,1416-3048
  * FacetSaddlePoint / RidgePeakSaddlePoint          src/linalg/solver_core/ParELAG_SaddlePointSolver.cpp:49-189
  * CochainProjector                                 src/amge/CochainProjector.cpp:53-261,416-441
  * SVD_Calculator::ComputeON, Deflate               src/linalg/dense/ParELAG_SVDCalculator.cpp:192-284,
                                                     src/linalg/dense/ParELAG_InnerProduct.cpp:157-169
  * ComputeTrueP / ComputeTrueD / GetP(ess)          src/amge/DeRhamSequence.cpp:1082-1253
  * the fine level (DeRhamSequence3D_FE for lowest-order spaces on an axis-aligned
    structured hexahedral mesh: H1 / Nedelec / Raviart-Thomas / L2 with MFEM's dof
    conventions -- point values, circulations, fluxes, cell values)
                                                     src/amge/DeRhamSequenceFE.cpp:184-227,311-335,633-722,799-925

Numbering conventions are this oracle's own and documented here (MFEM's mesh
numbering is not reproducible offline): entities of the structured
mesh are numbered lexicographically with x fastest; all topology / dof tables are
kept in canonical CSR form (ascending column indices per row); agglomerates are
numbered lexicographically on the derefined grid.  Integer parity of the CUDA path is
checked against THESE tables.

PARITY PIN: `upscaling_errors()` reproduces the reference's numbering- and
sign-invariant golden values of testsuite/CMakeLists.txt:114-176 (form0, form1,
form2); see tests/test_oracle_goldens.py.  The topology checks (connectedComponents,
AgglomeratedTopologyCheck, de-agglomeration) reproduce the messages of the reference's eight
`twentyseven.exe` topology tests (testsuite/CMakeLists.txt:34-92) including the entity numbers,
and the geometric box partitioner its two goldens (:187-193, :254-258); see
tests/test_topology_check_cpu.py.

SVD sign convention (both oracle and CUDA path): every retained left singular vector
is scaled so that its largest-magnitude entry (first one on ties) is positive.

Single rank only: dof == true dof (SharingMap is the identity).
"""
import numpy as np
import scipy.sparse as sp
from scipy.linalg import lapack

RANGET, NULLSPACE = 1, 2
M1D = np.array([[1.0 / 3.0, 1.0 / 6.0], [1.0 / 6.0, 1.0 / 3.0]])


# ----------------------------------------------------------------------------
# small sparse helpers (canonical CSR)
# ----------------------------------------------------------------------------
def _canon(A):
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


def mult_orientation(A, B):
    """TopologyTable MultOrientation: product, drop |.|<1e-10, keep the sign
    (src/topology/TopologyTable.cpp:130-139)."""
    C = _canon(A @ B)
    C.data[np.abs(C.data) < 1e-10] = 0.0
    C.eliminate_zeros()
    C.data = np.sign(C.data)
    return C


def row(A, i):
    return A.indices[A.indptr[i]:A.indptr[i + 1]]


def rowvals(A, i):
    return A.data[A.indptr[i]:A.indptr[i + 1]]


def block_diag_csr(blocks):
    """ElementalMatricesContainer::GetAsSparseMatrix
    (src/amge/ElementalMatricesContainer.cpp:55-212): one dense block per entity."""
    sizes = np.array([b.shape[0] for b in blocks], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offs[-1])
    if n == 0:
        return sp.csr_matrix((0, 0))
    rows, cols, vals = [], [], []
    for b, o in zip(blocks, offs[:-1]):
        m = b.shape[0]
        if m == 0:
            continue
        r, c = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        rows.append((r + o).ravel())
        cols.append((c + o).ravel())
        vals.append(np.asarray(b, dtype=np.float64).ravel())
    M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return M


# ----------------------------------------------------------------------------
# topology
# ----------------------------------------------------------------------------
def find_minimal_intersection_sets(Z, skip_diag_less_than=0.5):
    """findMinimalIntersectionSets: entity x MIS table with +-1 orientation."""
    tol = 1e-10
    Z = _canon(Z)
    n = Z.shape[0]
    diag = Z.diagonal()
    member = diag - skip_diag_less_than > -tol
    mis_of = -np.ones(n, dtype=np.int64)
    sign_of = np.zeros(n)
    cur = 0
    for i in range(n):
        if member[i] and mis_of[i] == -1:
            Zii = diag[i]
            cols, vals = row(Z, i), rowvals(Z, i)
            for j, Zij in zip(cols, vals):
                if abs(diag[j] - Zii) < tol and (abs(Zij - Zii) < tol or abs(Zij + Zii) < tol):
                    mis_of[j] = cur
                    sign_of[j] = Zij / Zii
            cur += 1
    idx = np.nonzero(mis_of >= 0)[0]
    return _canon(sp.csr_matrix((sign_of[idx], (idx, mis_of[idx])), shape=(n, cur)))


class Topology:
    """AgglomeratedTopology: B[c] = signed incidence (entities of codim c) x (codim c+1)."""

    def __init__(self, B, facet_bdr=None, ndim=3):
        self.ndim = ndim
        self.B = [_canon(b) for b in B]
        self.n = [self.B[0].shape[0]] + [b.shape[1] for b in self.B]
        self.facet_bdr = None if facet_bdr is None else _canon(facet_bdr)
        self.AE_entity = None       # set on the FINE topology by coarsen()
        self.coarser = None
        self.partition = None
        self._conn = {}

    def conn(self, big, small):
        """GetConnectivity(big, small): boolean entity->sub-entity table."""
        if (big, small) not in self._conn:
            C = abs(self.B[big])
            for c in range(big + 1, small):
                C = C @ abs(self.B[c])
            C = _canon(C)
            C.data[:] = 1.0
            self._conn[(big, small)] = C
        return self._conn[(big, small)]

    def element_element(self):
        """LocalElementElementTable (Topology.cpp:281-293): pattern of B_0 B_0^T"""
        C = _canon(abs(self.B[0]) @ abs(self.B[0]).T)
        C.data[:] = 1.0
        return C

    def coarsen(self, partition, check_topology=False, preserve_material_interfaces=False):
        """CoarsenLocalPartitioning(partition, check_topology, preserve_material=0) (Topology.cpp:685-828): disconnected
        partitions are split and empty ones removed (connectedComponents), then, codimension by codimension, the
        agglomerated entities are the minimal intersection sets; with check_topology every stage is followed by
        ShowBadAgglomeratedEntities / MarkBadAgglomeratedEntities / DeAgglomerateBadAgglomeratedEntities
        (AgglomeratedTopologyCheck.cpp, Topology.cpp:421-434,1151-1214).  The messages the reference prints are
        collected in self.messages."""
        partition = np.array(partition, dtype=np.int64)
        self.messages = []
        if preserve_material_interfaces:        # connectedComponents.cpp:90-96: the material-aware form is a stub
            self.messages.append("WARNING: this form of connectedComponents not implemented yet.")
            nAE = int(partition.max()) + 1
        else:
            nAE = connected_components(partition, self.element_element())
        self.partition = partition
        AE_el = _canon(sp.csr_matrix((np.ones(len(partition)), (partition, np.arange(len(partition)))),
                                     shape=(nAE, len(partition))))
        AEe = [AE_el]
        if check_topology:
            AEe[0] = self._check_stage(AEe, 0)
        cB = []
        for icodim in range(self.ndim):
            AE_fc = mult_orientation(AEe[icodim], self.B[icodim])
            Z = _canon(AE_fc.T @ AE_fc)
            if icodim == 0 and self.facet_bdr is not None:
                Z = _canon(Z + self.facet_bdr @ self.facet_bdr.T)
            fc_AF = find_minimal_intersection_sets(Z, 0.5)
            AEe.append(_canon(fc_AF.T))
            if check_topology:
                AEe[icodim + 1] = self._check_stage(AEe, icodim + 1)
                fc_AF = _canon(AEe[icodim + 1].T)
            cB.append(mult_orientation(AE_fc, fc_AF))
        cbdr = None
        if self.facet_bdr is not None:
            cbdr = mult_orientation(AEe[1], self.facet_bdr)
        self.AE_entity = AEe
        self.coarser = Topology(cB, cbdr, self.ndim)
        return self.coarser

    def _check_stage(self, AEe, codim):
        """Topology.cpp:728-739 (codim 0) and CheckHFacetsTopology (:421-434)"""
        self.messages += show_bad_agglomerated_entities(self, AEe, codim)
        isbad = mark_bad_agglomerated_entities(self, AEe, codim)
        if isbad.sum() == 0:
            return AEe[codim]
        new = deagglomerate_bad_entities(AEe[codim], isbad)
        self.messages += ["Correcting agglomerated topology for icodim: %d" % codim,
                          "  original number of agglomerates: %d" % AEe[codim].shape[0],
                          "  number which were bad: %d" % int(isbad.sum()),
                          "  number of new agglomerates after de-agglomeration: %d" % (new.shape[0] - AEe[codim].shape[0])]
        return new

    def show_me(self):
        """AgglomeratedTopology::ShowMe (Topology.cpp:310-352), one rank: entity counts and the Euler characteristic"""
        names = ["N_elements", "N_facets  ", "N_ridges  ", "N_peaks   "]
        out = ["  %s = %10d%10d" % (names[c], self.n[c], self.n[c]) for c in range(self.ndim + 1)]
        chi = sum(self.n[c] for c in range(self.ndim, -1, -2)) - sum(self.n[c] for c in range(self.ndim - 1, -1, -2))
        out.append("Euler Characteristic = %10d%10d" % (chi, chi))
        return out


def connected_components(partition, conn):
    """connectedComponents (src/structures/connectedComponents.cpp:23-87): every partition is split into its connected
    components with respect to the element-element table; component c of partition p becomes offset[p] + c, components
    numbered in the order a scan of the elements meets them, so empty partitions disappear and connected partitions keep
    their relative order.  partition is rewritten in place; returns the number of agglomerates."""
    n = len(partition)
    if n == 0:
        return 0
    npart = int(partition.max()) + 1
    comp = -np.ones(n, dtype=np.int64)
    ncomp = np.zeros(npart, dtype=np.int64)
    I, J = conn.indptr, conn.indices
    for node in range(n):
        if partition[node] < 0 or comp[node] >= 0:
            continue
        comp[node] = ncomp[partition[node]]
        ncomp[partition[node]] += 1
        stack = [node]
        while stack:
            i = stack.pop()
            for k in J[I[i]:I[i + 1]]:
                if partition[k] == partition[i] and comp[k] < 0:
                    comp[k] = comp[i]
                    stack.append(k)
    offs = np.concatenate([[0], np.cumsum(ncomp)])
    partition[:] = offs[partition] + comp
    return int(offs[-1])


def betti_numbers(topo, AEe, codim):
    """AgglomeratedTopologyCheck::computeBettiNumbersAgglomeratedEntities (AgglomeratedTopologyCheck.cpp:242-316): for
    every agglomerated entity of codimension codim the ranks of the boundary operators restricted to its fine entities
    of every lower dimension; betti(l) = dim_k[i+1] - rank_k[i] - rank_k[i+1], l = nLowerDims - i - 1
    (0: connected components, ..., top: holes)."""
    nlow = topo.ndim - codim
    if nlow == 0:
        return np.zeros((0, 0), dtype=np.int64)
    tabs = [AEe[codim]]
    for i in range(nlow):
        T = _canon(abs(tabs[i]) @ abs(topo.B[codim + i]))
        T.data[:] = 1.0
        tabs.append(T)
    nAE = tabs[0].shape[0]
    betti = np.zeros((nAE, nlow), dtype=np.int64)
    for a in range(nAE):
        ents = [row(t, a) for t in tabs]
        dim_k = [len(e) for e in ents]
        rank_k = [0] * (nlow + 1)
        for i in range(nlow):
            if dim_k[i] and dim_k[i + 1]:
                d = topo.B[codim + i][ents[i]][:, ents[i + 1]].toarray()
                rank_k[i] = int(np.linalg.matrix_rank(d, tol=1e-9))
        for i in range(nlow):
            betti[a, nlow - i - 1] = dim_k[i + 1] - rank_k[i] - rank_k[i + 1]
    return betti


def additional_topology_check(topo, AEe, codim, isbad, messages=None):
    """AgglomeratedTopologyCheck::additionalTopologyCheck (AgglomeratedTopologyCheck.cpp:25-82): on the boundary of an
    agglomerated element (codim 0) / agglomerated facet (codim 1) every boundary ridge (peak) must be adjacent to exactly
    two boundary facets (ridges)"""
    bf = _canon(AEe[codim] @ topo.B[codim])
    bf.eliminate_zeros()
    bf = abs(bf)
    fe = abs(topo.B[codim + 1])
    be = _canon(bf @ fe)
    for a in range(bf.shape[0]):
        rows_, cols_ = row(bf, a), row(be, a)
        loc = fe[rows_][:, cols_]
        twos = np.asarray(loc.sum(axis=0)).ravel()
        if abs(twos.sum() - 2 * len(twos)) > 1e-10:
            if messages is not None:
                messages.append("    codim %d iAE %d has bad connectivity (eg boundary edge adjacent to >2 boundary faces)." % (codim, a))
            isbad[a] = 1


def mark_bad_agglomerated_entities(topo, AEe, codim):
    """AgglomeratedTopologyCheck::MarkBadAgglomeratedEntities (AgglomeratedTopologyCheck.cpp:84-142)"""
    betti = betti_numbers(topo, AEe, codim)
    isbad = np.zeros(betti.shape[0], dtype=np.int64)
    if codim <= 2 and betti.size:
        isbad[betti[:, 0] != 1] = 1
        if codim <= 1:
            isbad[np.any(betti[:, 1:] != 0, axis=1)] = 1
    if (topo.ndim == 2 and codim == 0) or (topo.ndim == 3 and codim in (0, 1)):
        additional_topology_check(topo, AEe, codim, isbad)
    return isbad


def show_bad_agglomerated_entities(topo, AEe, codim):
    """AgglomeratedTopologyCheck::ShowBadAgglomeratedEntities and showBadAgglomerated{Elements,Facets,Ridges}
    (AgglomeratedTopologyCheck.cpp:144-240): the reference's messages, one list entry per line"""
    betti = betti_numbers(topo, AEe, codim)
    out = []
    name = ["Element", "Facet", "Ridge"]
    if codim <= 2:
        for a in range(betti.shape[0]):
            if betti[a, 0] != 1:
                out.append("    %s %d is disconnected. The number of connected components is %d" % (name[codim], a, betti[a, 0]))
            if codim <= 1:
                for i in range(1, betti.shape[1]):
                    if betti[a, i] != 0:
                        what = "holes" if (codim == 1 or i == topo.ndim - 1) else "tunnels"
                        out.append("    %s %d has %d %s." % (name[codim], a, betti[a, i], what))
    if (topo.ndim == 2 and codim == 0) or (topo.ndim == 3 and codim in (0, 1)):
        additional_topology_check(topo, AEe, codim, np.zeros(betti.shape[0], dtype=np.int64), out)
    return out


def deagglomerate_bad_entities(AEE, isbad):
    """AgglomeratedTopology::DeAgglomerateBadAgglomeratedEntities (Topology.cpp:1151-1214): every fine entity of a bad
    agglomerated entity becomes an agglomerated entity of its own, in place (later ones are renumbered)"""
    I = [0]
    for a in range(AEE.shape[0]):
        lo, hi = AEE.indptr[a], AEE.indptr[a + 1]
        if isbad[a]:
            I.extend(range(lo + 1, hi + 1))
        else:
            I.append(hi)
    return sp.csr_matrix((AEE.data.copy(), AEE.indices.copy(), np.array(I, dtype=AEE.indptr.dtype)), shape=(len(I) - 1, AEE.shape[1]))


def refined_partition(dims_fine, ratio=(2, 2, 2)):
    """MFEMRefinedMeshPartitioner: agglomerate = parent element of one uniform refinement; on the lexicographic
    structured grid AE(i,j,k) = (i//2, j//2, k//2).  Grids that are not a multiple of the ratio take the logical
    Cartesian agglomeration (LogicalPartitioner.hpp:46-103 with CoarsenLogicalCartesianOperator,
    CartesianPartitioner.hpp:113-131): same coarse index = same agglomerate, ragged last blocks, parts numbered in the
    order the scan of the fine elements meets them (= lexicographic order of the coarse indices)."""
    nx, ny, nz = dims_fine
    rx, ry, rz = ratio
    cx, cy = -(-nx // rx), -(-ny // ry)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return (((k // rz) * cy + (j // ry)) * cx + (i // rx)).ravel()


def geometric_box_partition(centroids, bmin, bmax, num_partitions):
    """GeometricBoxPartitioner::doPartition (src/partitioning/GeometricBoxPartitioner.cpp:20-79): the bounding box is cut
    into round(extent / target_radius) boxes per direction, target_radius = (volume / num_partitions)^(1/dim); an element
    belongs to the box that holds the mean of its vertices; partition = ix + nx (iy + ny iz).  Empty boxes keep their
    number in the reference; here the partition is compacted (CoarsenLocalPartitioning needs contiguous ids)."""
    centroids = np.asarray(centroids, dtype=np.float64)
    bmin, bmax = np.asarray(bmin, dtype=np.float64), np.asarray(bmax, dtype=np.float64)
    dim = centroids.shape[1]
    ext = bmax - bmin
    radius = (np.prod(ext) / float(num_partitions)) ** (1.0 / dim)
    ndir = np.array([int(e / radius + 0.5) for e in ext], dtype=np.int64)
    pr = ext / ndir
    which = ((centroids - bmin) / pr).astype(np.int64)
    part = which[:, 0].copy()
    mult = 1
    for a in range(1, dim):
        mult *= ndir[a - 1]
        part += mult * which[:, a]
    _, compact = np.unique(part, return_inverse=True)
    return compact.astype(np.int64)


def coarse_dims(d, ratio=(2, 2, 2)):
    return tuple(-(-a // r) for a, r in zip(d, ratio))


# ----------------------------------------------------------------------------
# structured hexahedral mesh + lowest-order fine de Rham sequence
# ----------------------------------------------------------------------------
class HexMesh:
    """nx x ny x nz axis-aligned cells of size (hx,hy,hz) on [0,Lx]x[0,Ly]x[0,Lz].
    Numbering (x fastest): element (i,j,k) -> i + nx*(j + ny*k);
    facets: x-normal faces, then y-normal, then z-normal; ridges: x-, y-, z-edges;
    peaks: vertices.  Global orientation of every facet/ridge is the positive axis.
    Boundary attributes follow mfem::Mesh::Make3D: z=0:1, y=0:2, x=L:3, y=L:4, x=0:5, z=L:6."""

    def __init__(self, nx, ny, nz, L=(1.0, 1.0, 1.0)):
        self.dims = (nx, ny, nz)
        self.h = (L[0] / nx, L[1] / ny, L[2] / nz)
        self.nel = nx * ny * nz
        self.nf = ((nx + 1) * ny * nz, nx * (ny + 1) * nz, nx * ny * (nz + 1))
        self.ne = (nx * (ny + 1) * (nz + 1), (nx + 1) * ny * (nz + 1), (nx + 1) * (ny + 1) * nz)
        self.nv = (nx + 1) * (ny + 1) * (nz + 1)

    # -- index maps
    def el(self, i, j, k):
        nx, ny, nz = self.dims
        return i + nx * (j + ny * k)

    def fx(self, i, j, k):
        nx, ny, nz = self.dims
        return i + (nx + 1) * (j + ny * k)

    def fy(self, i, j, k):
        nx, ny, nz = self.dims
        return self.nf[0] + i + nx * (j + (ny + 1) * k)

    def fz(self, i, j, k):
        nx, ny, nz = self.dims
        return self.nf[0] + self.nf[1] + i + nx * (j + ny * k)

    def ex(self, i, j, k):
        nx, ny, nz = self.dims
        return i + nx * (j + (ny + 1) * k)

    def ey(self, i, j, k):
        nx, ny, nz = self.dims
        return self.ne[0] + i + (nx + 1) * (j + ny * k)

    def ez(self, i, j, k):
        nx, ny, nz = self.dims
        return self.ne[0] + self.ne[1] + i + (nx + 1) * (j + (ny + 1) * k)

    def vx(self, i, j, k):
        nx, ny, nz = self.dims
        return i + (nx + 1) * (j + (ny + 1) * k)

    def _grid(self, ni, nj, nk):
        k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
        return i.ravel(), j.ravel(), k.ravel()

    def topology(self):
        nx, ny, nz = self.dims
        nf, ne = sum(self.nf), sum(self.ne)
        # B0: element x facet, +1 if the facet's +axis normal points out of the element
        i, j, k = self._grid(nx, ny, nz)
        e = self.el(i, j, k)
        r = np.concatenate([e] * 6)
        c = np.concatenate([self.fx(i, j, k), self.fx(i + 1, j, k), self.fy(i, j, k), self.fy(i, j + 1, k),
                            self.fz(i, j, k), self.fz(i, j, k + 1)])
        v = np.concatenate([-np.ones(len(e)), np.ones(len(e))] * 3)
        B0 = sp.csr_matrix((v, (r, c)), shape=(self.nel, nf))
        # B1: facet x ridge (discrete curl): right-hand rule around the +axis normal
        rows, cols, vals = [], [], []

        def add(f, edges_signs):
            for ed, s in edges_signs:
                rows.append(f); cols.append(ed); vals.append(np.full(len(f), s))
        i, j, k = self._grid(nx + 1, ny, nz)      # x-faces: dEz/dy - dEy/dz
        add(self.fx(i, j, k), [(self.ez(i, j + 1, k), 1.0), (self.ez(i, j, k), -1.0),
                               (self.ey(i, j, k + 1), -1.0), (self.ey(i, j, k), 1.0)])
        i, j, k = self._grid(nx, ny + 1, nz)      # y-faces: dEx/dz - dEz/dx
        add(self.fy(i, j, k), [(self.ex(i, j, k + 1), 1.0), (self.ex(i, j, k), -1.0),
                               (self.ez(i + 1, j, k), -1.0), (self.ez(i, j, k), 1.0)])
        i, j, k = self._grid(nx, ny, nz + 1)      # z-faces: dEy/dx - dEx/dy
        add(self.fz(i, j, k), [(self.ey(i + 1, j, k), 1.0), (self.ey(i, j, k), -1.0),
                               (self.ex(i, j + 1, k), -1.0), (self.ex(i, j, k), 1.0)])
        B1 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nf, ne))
        # B2: ridge x peak (discrete gradient): -1 at the tail, +1 at the head
        rows, cols, vals = [], [], []
        i, j, k = self._grid(nx, ny + 1, nz + 1)
        rows += [self.ex(i, j, k)] * 2; cols += [self.vx(i, j, k), self.vx(i + 1, j, k)]
        vals += [-np.ones(len(i)), np.ones(len(i))]
        i, j, k = self._grid(nx + 1, ny, nz + 1)
        rows += [self.ey(i, j, k)] * 2; cols += [self.vx(i, j, k), self.vx(i, j + 1, k)]
        vals += [-np.ones(len(i)), np.ones(len(i))]
        i, j, k = self._grid(nx + 1, ny + 1, nz)
        rows += [self.ez(i, j, k)] * 2; cols += [self.vx(i, j, k), self.vx(i, j, k + 1)]
        vals += [-np.ones(len(i)), np.ones(len(i))]
        B2 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ne, self.nv))
        # facet -> boundary attribute (0-based column = attribute-1)
        rows, cols = [], []
        j, k = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
        rows += [self.fx(0, j.ravel(), k.ravel()), self.fx(nx, j.ravel(), k.ravel())]
        cols += [np.full(j.size, 4), np.full(j.size, 2)]
        i, k = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
        rows += [self.fy(i.ravel(), 0, k.ravel()), self.fy(i.ravel(), ny, k.ravel())]
        cols += [np.full(i.size, 1), np.full(i.size, 3)]
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        rows += [self.fz(i.ravel(), j.ravel(), 0), self.fz(i.ravel(), j.ravel(), nz)]
        cols += [np.full(i.size, 0), np.full(i.size, 5)]
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        fbdr = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(nf, 6))
        return Topology([B0, B1, B2], fbdr, 3)

    def facet_area(self):
        hx, hy, hz = self.h
        return np.concatenate([np.full(self.nf[0], hy * hz), np.full(self.nf[1], hx * hz), np.full(self.nf[2], hx * hy)])

    def ridge_length(self):
        hx, hy, hz = self.h
        return np.concatenate([np.full(self.ne[0], hx), np.full(self.ne[1], hy), np.full(self.ne[2], hz)])

    def vertex_coords(self):
        nx, ny, nz = self.dims
        i, j, k = self._grid(nx + 1, ny + 1, nz + 1)
        return np.stack([i * self.h[0], j * self.h[1], k * self.h[2]], axis=1)


def mfem_hex_numbering(nx, ny, nz):
    """Face and edge numbers mfem::Mesh gives the entities of its Cartesian hexahedral mesh (mfem::Mesh::Make3D without
    space-filling-curve ordering; vertices and elements are lexicographic, x fastest, like HexMesh): faces and edges are
    numbered in the order a scan over the elements meets them, local faces in the order z-, y-, x+, y+, x-, z+
    (mfem::Geometry::Constants<CUBE>::FaceVert), local edges in the order of mfem's hexahedron edge table
    (01 12 32 03 | 45 56 76 47 | 04 15 26 37 with vertices 0..3 counter-clockwise at z-, 4..7 above them).
    AgglomeratedTopology numbers facets / ridges as the RT0 / Nedelec dofs of the mesh, i.e. by these numbers
    (Topology.cpp:85-141), and the minimal intersection sets inherit the order (minimalIntersectionSet.cpp:96-127), so the
    entity numbers in the reference's topology messages (testsuite/CMakeLists.txt:57-75) can be reproduced.
    Returns (facet_perm, ridge_perm): lexicographic number -> mfem number."""
    m = HexMesh(nx, ny, nz)
    fperm = -np.ones(sum(m.nf), dtype=np.int64)
    eperm = -np.ones(sum(m.ne), dtype=np.int64)
    nfc = nec = 0
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                for f in (m.fz(i, j, k), m.fy(i, j, k), m.fx(i + 1, j, k), m.fy(i, j + 1, k), m.fx(i, j, k), m.fz(i, j, k + 1)):
                    if fperm[f] < 0:
                        fperm[f] = nfc
                        nfc += 1
                for e in (m.ex(i, j, k), m.ey(i + 1, j, k), m.ex(i, j + 1, k), m.ey(i, j, k),
                          m.ex(i, j, k + 1), m.ey(i + 1, j, k + 1), m.ex(i, j + 1, k + 1), m.ey(i, j, k + 1),
                          m.ez(i, j, k), m.ez(i + 1, j, k), m.ez(i + 1, j + 1, k), m.ez(i, j + 1, k)):
                    if eperm[e] < 0:
                        eperm[e] = nec
                        nec += 1
    return fperm, eperm


def renumbered_topology(topo, facet_perm=None, ridge_perm=None):
    """the same topology with facets / ridges renumbered (old number -> new number)"""
    B = [b.copy() for b in topo.B]
    fb = topo.facet_bdr
    if facet_perm is not None:
        Pf = sp.csr_matrix((np.ones(len(facet_perm)), (np.arange(len(facet_perm)), facet_perm)), shape=(len(facet_perm),) * 2)
        B[0] = B[0] @ Pf
        B[1] = Pf.T @ B[1]
        fb = None if fb is None else Pf.T @ fb
    if ridge_perm is not None:
        Pr = sp.csr_matrix((np.ones(len(ridge_perm)), (np.arange(len(ridge_perm)), ridge_perm)), shape=(len(ridge_perm),) * 2)
        B[1] = B[1] @ Pr
        B[2] = Pr.T @ B[2]
    return Topology(B, fb, topo.ndim)


class DeformedHexMesh(HexMesh):
    """HexMesh whose vertices have been moved: every cell is a trilinear hexahedron (mfem's isoparametric Q1 map).
    Numbering, orientation and boundary attributes are those of the index grid.  This is the geometry of
    examples/3DHdivWeakScaling.cpp:148-158 (uniformly refined unit cube, then y += exp(z)/2, x += sin(y)).
    Provides what the lowest-order H(div)-L2 part of DeRhamSequence3D_FE needs (DeRhamSequenceFE.cpp:633-684):
    cell volumes (MassIntegrator, P0), RT0 element mass matrices (VectorFEMassIntegrator: contravariant Piola map,
    Gauss rule of order OrderW + 2 = 4, i.e. 3 points per direction), the facet weight |N| at the facet centre
    (VolumetricFEMassIntegrator on the RT trace element, one-point rule; InterpolatePV_HdivTraces, :810-857)
    and the normal N itself (RT_HexahedronElement::Project of a constant field)."""

    def __init__(self, nx, ny, nz, deform, L=(1.0, 1.0, 1.0)):
        super().__init__(nx, ny, nz, L=L)
        self.X = np.asarray(deform(HexMesh.vertex_coords(self)), dtype=np.float64)

    def vertex_coords(self):
        return self.X

    @staticmethod
    def _gauss(n):
        x, w = np.polynomial.legendre.leggauss(n)
        return 0.5 * (x + 1.0), 0.5 * w

    def _corner(self, a, b, c):
        nx, ny, nz = self.dims
        i, j, k = self._grid(nx, ny, nz)
        return self.X[self.vx(i + a, j + b, k + c)]

    def jacobian_columns(self, xh, yh, zh):
        """dr/dxh, dr/dyh, dr/dzh of the trilinear map at one reference point, for all cells: three (nel, 3) arrays"""
        sx, sy, sz, d = (1 - xh, xh), (1 - yh, yh), (1 - zh, zh), (-1.0, 1.0)
        abc = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
        rx = sum(self._corner(a, b, c) * (d[a] * sy[b] * sz[c]) for a, b, c in abc)
        ry = sum(self._corner(a, b, c) * (sx[a] * d[b] * sz[c]) for a, b, c in abc)
        rz = sum(self._corner(a, b, c) * (sx[a] * sy[b] * d[c]) for a, b, c in abc)
        return rx, ry, rz

    def cell_volumes(self):
        g, w = self._gauss(2)                      # det J is quadratic per variable: exact
        v = np.zeros(self.nel)
        for a, wa in zip(g, w):
            for b, wb in zip(g, w):
                for c, wc in zip(g, w):
                    rx, ry, rz = self.jacobian_columns(a, b, c)
                    v += wa * wb * wc * np.einsum("ei,ei->e", rx, np.cross(ry, rz))
        return v

    def rt0_element_mass(self, npts=3):
        """(nel, 6, 6), local order x-,x+,y-,y+,z-,z+, every basis function with unit flux along the +index axis:
        v = J vhat / det J, vhat_{x-} = (1-xh, 0, 0), vhat_{x+} = (xh, 0, 0), ..."""
        g, w = self._gauss(npts)
        M = np.zeros((self.nel, 6, 6))
        for a, wa in zip(g, w):
            for b, wb in zip(g, w):
                for c, wc in zip(g, w):
                    rx, ry, rz = self.jacobian_columns(a, b, c)
                    det = np.einsum("ei,ei->e", rx, np.cross(ry, rz))
                    V = [rx * (1 - a), rx * a, ry * (1 - b), ry * b, rz * (1 - c), rz * c]
                    for p in range(6):
                        for q in range(6):
                            M[:, p, q] += (wa * wb * wc) * np.einsum("ei,ei->e", V[p], V[q]) / det
        return M

    def facet_normals(self):
        """dr/du x dr/dv at the centre of every facet, oriented along the +index axis (nf_total, 3)"""
        nx, ny, nz = self.dims
        X = self.X

        def mean_tangents(c):
            t1 = 0.5 * ((c(1, 0) - c(0, 0)) + (c(1, 1) - c(0, 1)))
            t2 = 0.5 * ((c(0, 1) - c(0, 0)) + (c(1, 1) - c(1, 0)))
            return t1, t2
        i, j, k = self._grid(nx + 1, ny, nz)
        ty, tz = mean_tangents(lambda b, c: X[self.vx(i, j + b, k + c)])
        Nx = np.cross(ty, tz)
        i, j, k = self._grid(nx, ny + 1, nz)
        tx, tz = mean_tangents(lambda a, c: X[self.vx(i + a, j, k + c)])
        Ny = np.cross(tz, tx)
        i, j, k = self._grid(nx, ny, nz + 1)
        tx, ty = mean_tangents(lambda a, b: X[self.vx(i + a, j + b, k)])
        Nz = np.cross(tx, ty)
        return np.concatenate([Nx, Ny, Nz], axis=0)

    def facet_area(self):
        return np.linalg.norm(self.facet_normals(), axis=1)

    # ---- H(curl) part (lowest-order Nedelec: unit circulation along the +index axis of every edge)
    def edge_vectors(self):
        """end point minus start point of every ridge (ne_total, 3): x-, y-, z-edges"""
        nx, ny, nz = self.dims
        X = self.X
        i, j, k = self._grid(nx, ny + 1, nz + 1)
        tx = X[self.vx(i + 1, j, k)] - X[self.vx(i, j, k)]
        i, j, k = self._grid(nx + 1, ny, nz + 1)
        ty = X[self.vx(i, j + 1, k)] - X[self.vx(i, j, k)]
        i, j, k = self._grid(nx + 1, ny + 1, nz)
        tz = X[self.vx(i, j, k + 1)] - X[self.vx(i, j, k)]
        return np.concatenate([tx, ty, tz], axis=0)

    def ridge_length(self):
        return np.linalg.norm(self.edge_vectors(), axis=1)

    def nd0_element_mass(self, npts=3):
        """(nel, 12, 12): VectorFEMassIntegrator with the covariant Piola map w = J^-T what, Gauss rule of order
        OrderW + 2 = 4.  Local order = ascending edge id: x-edges (j,k),(j+1,k),(j,k+1),(j+1,k+1), y-edges
        (i,k),(i+1,k),(i,k+1),(i+1,k+1), z-edges (i,j),(i+1,j),(i,j+1),(i+1,j+1)."""
        phi = lambda c, t: (1.0 - t) if c == 0 else t
        g, w = self._gauss(npts)
        M = np.zeros((self.nel, 12, 12))
        pairs = ((0, 0), (1, 0), (0, 1), (1, 1))
        for a, wa in zip(g, w):
            for b, wb in zip(g, w):
                for c, wc in zip(g, w):
                    rx, ry, rz = self.jacobian_columns(a, b, c)
                    J = np.stack([rx, ry, rz], axis=2)
                    det = np.linalg.det(J)
                    JinvT = np.transpose(np.linalg.inv(J), (0, 2, 1))
                    W = ([np.array([phi(p, b) * phi(q, c), 0.0, 0.0]) for p, q in pairs]
                         + [np.array([0.0, phi(p, a) * phi(q, c), 0.0]) for p, q in pairs]
                         + [np.array([0.0, 0.0, phi(p, a) * phi(q, b)]) for p, q in pairs])
                    V = [JinvT @ wv for wv in W]
                    for p in range(12):
                        for q in range(12):
                            M[:, p, q] += (wa * wb * wc) * det * np.einsum("ei,ei->e", V[p], V[q])
        return M

    def _facet_corner_fn(self, axis):
        """corner(u, v) of every facet of one axis; in-plane axes (u, v): x-faces (y, z), y-faces (x, z), z-faces (x, y)"""
        nx, ny, nz = self.dims
        X = self.X
        if axis == 0:
            i, j, k = self._grid(nx + 1, ny, nz)
            return lambda u, v: X[self.vx(i, j + u, k + v)]
        if axis == 1:
            i, j, k = self._grid(nx, ny + 1, nz)
            return lambda u, v: X[self.vx(i + u, j, k + v)]
        i, j, k = self._grid(nx, ny, nz + 1)
        return lambda u, v: X[self.vx(i + u, j + v, k)]

    @staticmethod
    def _facet_map(c, u, v):
        """tangent map (n, 3, 2), its pseudo-inverse transposed G = J (J^T J)^-1 (covariant map of 2D Nedelec shapes
        into the facet's tangent plane) and the surface weight sqrt(det J^T J) at the reference point (u, v)"""
        tu = (c(1, 0) - c(0, 0)) * (1 - v) + (c(1, 1) - c(0, 1)) * v
        tv = (c(0, 1) - c(0, 0)) * (1 - u) + (c(1, 1) - c(1, 0)) * u
        Jf = np.stack([tu, tv], axis=2)
        JtJ = np.einsum("eji,ejk->eik", Jf, Jf)
        G = np.einsum("eij,ejk->eik", Jf, np.linalg.inv(JtJ))
        return Jf, G, np.sqrt(np.linalg.det(JtJ))

    def nd0_facet_mass(self, npts=2):
        """three arrays (nf_axis, 4, 4): ND_3D_FacetMassIntegrator (bilinIntegrators.cpp:106-157), Gauss rule of order
        OrderW + 2 = 3.  Local order = ascending edge id: the two edges along u (at v = 0, 1), then the two along v."""
        phi = lambda c, t: (1.0 - t) if c == 0 else t
        g, w = self._gauss(npts)
        out = []
        for axis in range(3):
            c = self._facet_corner_fn(axis)
            M = np.zeros((c(0, 0).shape[0], 4, 4))
            for u, wu in zip(g, w):
                for v, wv in zip(g, w):
                    _, G, wt = self._facet_map(c, u, v)
                    W = [np.array([phi(0, v), 0.0]), np.array([phi(1, v), 0.0]), np.array([0.0, phi(0, u)]), np.array([0.0, phi(1, u)])]
                    V = [G @ x for x in W]
                    for p in range(4):
                        for q in range(4):
                            M[:, p, q] += (wu * wv) * wt * np.einsum("ei,ei->e", V[p], V[q])
            out.append(M)
        return out

    # ---- H1 part (trilinear nodal basis)
    def h1_element_mass(self, npts=3):
        """(nel, 8, 8): MassIntegrator, Gauss rule of order 2 + OrderW = 4; local order = ascending vertex id
        (i fastest, then j, then k)"""
        phi = lambda c, t: (1.0 - t) if c == 0 else t
        g, w = self._gauss(npts)
        M = np.zeros((self.nel, 8, 8))
        for a, wa in zip(g, w):
            for b, wb in zip(g, w):
                for c, wc in zip(g, w):
                    rx, ry, rz = self.jacobian_columns(a, b, c)
                    det = np.einsum("ei,ei->e", rx, np.cross(ry, rz))
                    sh = np.array([phi(p, a) * phi(q, b) * phi(r, c) for r in (0, 1) for q in (0, 1) for p in (0, 1)])
                    M += (wa * wb * wc) * det[:, None, None] * np.outer(sh, sh)[None, :, :]
        return M

    def h1_facet_mass(self, npts=2):
        """three arrays (nf_axis, 4, 4): MassIntegrator on the bilinear faces, Gauss rule of order 2 + OrderW = 3;
        local order = ascending vertex id (first in-plane axis fastest)"""
        phi = lambda c, t: (1.0 - t) if c == 0 else t
        g, w = self._gauss(npts)
        out = []
        for axis in range(3):
            c = self._facet_corner_fn(axis)
            M = np.zeros((c(0, 0).shape[0], 4, 4))
            for u, wu in zip(g, w):
                for v, wv in zip(g, w):
                    _, _, wt = self._facet_map(c, u, v)
                    sh = np.array([phi(p, u) * phi(q, v) for q in (0, 1) for p in (0, 1)])
                    M += (wu * wv) * wt[:, None, None] * np.outer(sh, sh)[None, :, :]
            out.append(M)
        return out

    def boundary_tangent_rhs_bottom(self, f):
        """VectorFEBoundaryTangentLFIntegrator on boundary attribute 1 (z-index 0): (n x f, w) with the outward normal,
        2x2 Gauss rule (order 2 * el.GetOrder() = 2); returns the Nedelec load vector"""
        phi = lambda c, t: (1.0 - t) if c == 0 else t
        nx, ny, nz = self.dims
        X = self.X
        g, w = self._gauss(2)
        b = np.zeros(sum(self.ne))
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        i, j = i.ravel(), j.ravel()
        c = lambda a, bb: X[self.vx(i + a, j + bb, 0)]
        edges = [self.ex(i, j, 0), self.ex(i, j + 1, 0), self.ey(i, j, 0), self.ey(i + 1, j, 0)]
        f = np.asarray(f, dtype=np.float64)
        for u, wu in zip(g, w):
            for v, wv in zip(g, w):
                Jf, G, wt = self._facet_map(c, u, v)
                n = np.cross(Jf[:, :, 0], Jf[:, :, 1])
                n = -n / np.linalg.norm(n, axis=1)[:, None]            # outward: -(t_x x t_y)
                nxf = np.cross(n, f[None, :])
                W = [np.array([phi(0, v), 0.0]), np.array([phi(1, v), 0.0]), np.array([0.0, phi(0, u)]), np.array([0.0, phi(1, u)])]
                for e, x in zip(edges, W):
                    np.add.at(b, e, (wu * wv) * wt * np.einsum("ei,ei->e", G @ x, nxf))
        return b


class DofHandler:
    """Common part of DofHandlerFE / DofHandlerALG: entity_dof[c] for c <= max_codim_base."""

    def __init__(self, max_codim_base, topo):
        self.mcb = max_codim_base
        self.topo = topo
        self.entity_dof = [None] * (max_codim_base + 1)
        self.ndofs = 0
        # ALG only
        self.dof_type = None
        self.n_rangeT = [None] * (max_codim_base + 1)
        self.n_null = [None] * (max_codim_base + 1)
        self.int_offsets = [None] * (max_codim_base + 1)
        self.type_ndofs = [0] * (max_codim_base + 2)

    # ---- ALG (src/amge/DofHandler.cpp:878-1460)
    def alg_init(self):
        for c in range(self.mcb + 1):
            self.n_rangeT[c] = np.zeros(self.topo.n[c], dtype=np.int64)
            self.n_null[c] = np.zeros(self.topo.n[c], dtype=np.int64)
        self.dof_type = {}

    def build_entity_dof_table(self, c):
        """computeOffset(c) + build{Peak,Ridge,Facet,Element}DofTable."""
        counts = self.n_rangeT[c] + self.n_null[c]
        start = self.type_ndofs[c + 1] if c < self.mcb else 0
        assert self.ndofs == start
        offs = start + np.concatenate([[0], np.cumsum(counts)])
        self.int_offsets[c] = offs
        self.ndofs = int(offs[-1])
        self.type_ndofs[c] = self.ndofs
        nent = self.topo.n[c]
        rows, cols = [], []
        for small in range(self.mcb, c, -1):     # PEAK first ... down to c+1
            C = self.topo.conn(c, small)
            so = self.int_offsets[small]
            for e in range(nent):
                for s in row(C, e):
                    cols.extend(range(so[s], so[s + 1]))
                    rows.extend([e] * int(so[s + 1] - so[s]))
        for e in range(nent):
            cols.extend(range(offs[e], offs[e + 1]))
            rows.extend([e] * int(offs[e + 1] - offs[e]))
        # rows keep the reference order [lower-dimensional carriers ..., own interior];
        # with canonical conn tables this order is ascending
        M = sp.csr_matrix((np.ones(len(rows)), (np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64))),
                          shape=(nent, self.ndofs))
        self.entity_dof[c] = _canon(M)
        for cc in range(c + 1, self.mcb + 1):   # widen earlier tables
            self.entity_dof[cc] = sp.csr_matrix((self.entity_dof[cc].data, self.entity_dof[cc].indices,
                                                 self.entity_dof[cc].indptr), shape=(self.topo.n[cc], self.ndofs))

    def interior_dofs(self, c, e):
        o = self.int_offsets[c]
        return np.arange(o[e], o[e + 1])

    def dofs_on_bdr(self, c, e):
        out = []
        for small in range(self.mcb, c, -1):
            for s in row(self.topo.conn(c, small), e):
                out.extend(self.interior_dofs(small, s))
        return np.array(out, dtype=np.int64)

    def rangeT_dofs(self, c, e):
        return [d for d in self.interior_dofs(c, e) if self.dof_type[d] == RANGET]

    def null_dofs(self, c, e):
        return [d for d in self.interior_dofs(c, e) if self.dof_type[d] == NULLSPACE]

    def mark_bdr_dofs(self, ess_attr):
        """MarkDofsOnSelectedBndr (DofHandler.cpp:812-853): dofs of facets whose boundary
        attribute is selected.  (For the FE level MFEM's GetEssentialVDofs marks the same
        set: all dofs in the closure of selected boundary faces.)"""
        marker = np.zeros(self.ndofs, dtype=bool)
        fb = self.topo.facet_bdr
        ess_attr = np.asarray(ess_attr)
        FD = self.entity_dof[1] if self.mcb >= 1 else None
        if FD is None:
            return marker
        for f in range(fb.shape[0]):
            a = row(fb, f)
            if len(a) == 1 and ess_attr[a[0]]:
                marker[row(FD, f)] = True
        return marker


class DofAgg:
    """DofAgglomeration (DOFAgglomeration.cpp:33-315): AE -> dof rows with interior dofs
    first (sorted by (separator type, dof id)); ADof = position in that CSR."""

    def __init__(self, topo, dof):
        self.topo, self.dof = topo, dof
        ncod = dof.mcb + 1
        AE_dof = [_canon(abs(topo.AE_entity[c]) @ abs(dof.entity_dof[c])) for c in range(ncod)]
        sep = np.zeros(dof.ndofs, dtype=np.int64)
        for c in range(1, ncod):
            sep[AE_dof[c].indices] = c
        self.sep = sep
        self.I, self.J, self.nint = [], [], []
        for c in range(ncod):
            A = AE_dof[c]
            J = A.indices.copy()
            nint = np.zeros(A.shape[0], dtype=np.int64)
            for a in range(A.shape[0]):
                s, e = A.indptr[a], A.indptr[a + 1]
                r = J[s:e]
                if dof.mcb > c:
                    order = np.lexsort((r, sep[r]))
                    J[s:e] = r[order]
                    nint[a] = int(np.sum(sep[r] == c))
                else:
                    nint[a] = e - s
            self.I.append(A.indptr.astype(np.int64))
            self.J.append(J.astype(np.int64))
            self.nint.append(nint)
        self._adof_rdof = {}

    def nAE(self, c):
        return len(self.I[c]) - 1

    def rng(self, c, a):
        """(start, start+nint, end) in ADof numbering."""
        return int(self.I[c][a]), int(self.I[c][a] + self.nint[c][a]), int(self.I[c][a + 1])

    def dofs(self, c, a):
        s, m, e = self.rng(c, a)
        return self.J[c][s:m], self.J[c][m:e]

    def adof_dof(self, c):
        n = len(self.J[c])
        return sp.csr_matrix((np.ones(n), (np.arange(n), self.J[c])), shape=(n, self.dof.ndofs))

    def adof_rdof(self, c):
        """ADof_rDof: couples each agglomerated dof with the repeated dofs (entity-local
        copies) of the fine entities that make up the agglomerate."""
        if c in self._adof_rdof:
            return self._adof_rdof[c]
        ED = self.dof.entity_dof[c]
        AEe = self.topo.AE_entity[c]
        nrd = ED.nnz
        ent_of_rdof = np.repeat(np.arange(ED.shape[0]), np.diff(ED.indptr))
        AE_of_ent = -np.ones(ED.shape[0], dtype=np.int64)
        ae_rows = np.repeat(np.arange(AEe.shape[0]), np.diff(AEe.indptr))
        AE_of_ent[AEe.indices] = ae_rows
        AE_of_rdof = AE_of_ent[ent_of_rdof]
        keep = AE_of_rdof >= 0
        nd = self.dof.ndofs
        key_adof = np.repeat(np.arange(self.nAE(c)), np.diff(self.I[c])) * nd + self.J[c]
        order = np.argsort(key_adof)
        key_r = AE_of_rdof[keep] * nd + ED.indices[keep]
        pos = np.searchsorted(key_adof[order], key_r)
        adof = order[pos]
        assert np.all(key_adof[adof] == key_r)
        M = sp.csr_matrix((ED.data[keep], (adof, np.nonzero(keep)[0])), shape=(len(self.J[c]), nrd))
        self._adof_rdof[c] = M
        return M

    def assemble_agg_matrix(self, c, M_e, other=None):
        """AssembleAgglomerateMatrix: ADof_rDof * M_e * ADof_rDof^T (block diagonal)."""
        R = self.adof_rdof(c)
        Pm = (other or self).adof_rdof(c)
        return _canon(R @ M_e @ Pm.T)

    @staticmethod
    def distribute(c, D_g, rng_agg, dom_agg):
        """DistributeAgglomerateMatrix(range!=null, domain!=null) -> Distribute():
        per-AE blocks D_g[AE range dofs, AE domain dofs] in ADof numbering."""
        R, Cm = rng_agg.adof_dof(c), dom_agg.adof_dof(c)
        T = (R @ D_g @ Cm.T).tocoo()
        ae_r = np.repeat(np.arange(rng_agg.nAE(c)), np.diff(rng_agg.I[c]))
        ae_c = np.repeat(np.arange(dom_agg.nAE(c)), np.diff(dom_agg.I[c]))
        keep = ae_r[T.row] == ae_c[T.col]
        return _canon(sp.csr_matrix((T.data[keep], (T.row[keep], T.col[keep])), shape=T.shape))


# ----------------------------------------------------------------------------
# dense helpers
# ----------------------------------------------------------------------------
SIGN_TIE_REL = 1e-6


def fix_sign(U):
    """Canonical sign: the first entry (in index order) whose magnitude is within a relative
    SIGN_TIE_REL of the column's largest magnitude is positive.  Equal to "largest-magnitude entry
    positive" when that entry is unique; independent of rounding when congruent fine entities give
    entries of equal magnitude (the reference's sign is LAPACK's, i.e. arbitrary: every quantity
    the reference's tests pin is invariant under it)."""
    U = np.array(U, dtype=np.float64, copy=True)
    for j in range(U.shape[1]):
        col = U[:, j]
        if col.size == 0:
            continue
        a = np.abs(col)
        p = int(np.argmax(a >= a.max() * (1.0 - SIGN_TIE_REL)))
        if col[p] < 0:
            U[:, j] = -col
    return U


def svd_on(A):
    """SVD_Calculator::ComputeON (dgesvd JOBU='O', JOBVT='N'): returns (U, s) with
    min(m,n) columns, canonical sign."""
    A = np.asarray(A, dtype=np.float64)
    if A.shape[0] == 0 or A.shape[1] == 0:
        return np.zeros((A.shape[0], 0)), np.zeros(0)
    u, s, vt, info = lapack.dgesvd(A, compute_uv=1, full_matrices=0)
    assert info == 0
    return fix_sign(u), s


def svd_on_weighted(M, A):
    """ComputeON(sqrt_w, A, s) for diagonal M, ComputeON(W, A, s) (symmetric square
    root through dsyev) otherwise (SVDCalculator.cpp:247-284)."""
    offd = M - np.diag(np.diag(M))
    if not np.any(offd):
        w = np.sqrt(np.diag(M))
        U, s = svd_on(A * w[:, None])
        return fix_sign(U / w[:, None]), s
    ev, V = np.linalg.eigh(M)
    sq = np.sqrt(ev)
    X = (V * sq) @ V.T
    U, s = svd_on(X @ A)
    Xi = (V / sq) @ V.T
    return fix_sign(Xi @ U), s


class LDL:
    """LDLCalculator: dsytrf('L') / dsytrs (src/linalg/dense/ParELAG_LDLCalculator.cpp:32-93)."""

    def __init__(self, A):
        self.n = A.shape[0]
        if self.n:
            self.ldu, self.ipiv, info = lapack.dsytrf(np.asfortranarray(A), lower=1)
            assert info == 0, "LDL factorization failed (singular local saddle point?)"

    def solve(self, rhs):
        if self.n == 0 or rhs.shape[1] == 0:
            return np.zeros_like(rhs)
        x, info = lapack.dsytrs(self.ldu, self.ipiv, np.asfortranarray(rhs), lower=1)
        assert info == 0
        return x


def dof_functional(Ploc, Mii):
    """CochainProjector::CreateDofFunctional: (P^T M P)^{-1} P^T M via dgetrf/dgetrs."""
    nc = Ploc.shape[1]
    if nc == 0:
        return np.zeros((0, Ploc.shape[0]))
    MlP = Mii @ Ploc
    cM = Ploc.T @ MlP
    lu, piv, info = lapack.dgetrf(np.asfortranarray(cM))
    assert info == 0
    x, info = lapack.dgetrs(lu, piv, np.asfortranarray(MlP.T))
    assert info == 0
    return x


# ----------------------------------------------------------------------------
# de Rham sequence
# ----------------------------------------------------------------------------
class Sequence:
    """One level of the de Rham sequence (DeRhamSequenceFE on level 0, DeRhamSequenceAlg
    below).  M[(j,c)] is the DG-like mass matrix of form j on entities of codim c in rDof
    numbering; D[j] maps form j -> j+1; P[j]/Pi[j] connect to the coarser level."""

    def __init__(self, topo, nforms=4):
        self.topo = topo
        self.nforms = nforms
        self.ndim = nforms - 1
        self.jstart = 0
        self.dof = [None] * nforms
        self.D = [None] * (nforms - 1)
        self.M = {}
        self.targets = [None] * nforms
        self.P = [None] * nforms
        self.Pi = [None] * nforms
        self.l2_const = None
        self.svd_tol = 1e-9
        self.smallest_entry = np.finfo(float).eps
        self.coarser = None
        self.finer = None
        self.mesh = None       # fine level only
        self.stats = {}

    # --- operators handed to the solver layer
    def mass_operator(self, j):
        """ComputeMassOperator(j): assemble the element mass matrices, rDof_dof^T M_e rDof_dof
        (DofHandler.cpp:270-281)."""
        ED = self.dof[j].entity_dof[0]
        n = ED.nnz
        R = sp.csr_matrix((ED.data, (np.arange(n), ED.indices)), shape=(n, self.dof[j].ndofs))
        return _canon(R.T @ self.M[(j, 0)] @ R)

    def get_P(self, j, ess_attr=None):
        """GetP(j, ess): copy of P with the columns of essential COARSE dofs zeroed; the
        zeros stay in the pattern (SparseMatrix::EliminateCols)."""
        P = self.P[j].copy()
        if ess_attr is not None:
            marker = self.coarser.dof[j].mark_bdr_dofs(ess_attr)
            P.data[marker[P.indices]] = 0.0
        return P

    def get_D(self, j, ess_attr=None):
        D = self.D[j].copy()
        if ess_attr is not None:
            marker = self.dof[j].mark_bdr_dofs(ess_attr)
            D.data[marker[D.indices]] = 0.0
        return D

    # --- PV traces
    def pv_traces(self, c):
        j = self.nforms - 1 - c
        nd = self.dof[j].ndofs
        AEe = self.topo.AE_entity[c]
        if self.mesh is not None:       # DeRhamSequence3D_FE::computePVTraces
            if c == 0:
                return np.ones(nd)
            pv = np.zeros(nd)
            if c == 1:
                pv[AEe.indices] = AEe.data * self.mesh.facet_area()[AEe.indices]
            elif c == 2:
                pv[AEe.indices] = AEe.data * self.mesh.ridge_length()[AEe.indices]
            else:
                pv[AEe.indices] = 1.0
            return pv
        # DeRhamSequenceAlg::computePVTraces: +-1 on the first dof of every member entity
        pv = np.zeros(nd)
        ED = self.dof[j].entity_dof[c]
        pv[ED.indices[ED.indptr[AEe.indices]]] = AEe.data
        return pv

    # --- Coarsen
    def coarsen(self):
        topo, ctopo = self.topo, self.topo.coarser
        assert ctopo is not None, "coarsen the topology first"
        cs = Sequence(ctopo, self.nforms)
        cs.jstart = self.jstart
        cs.svd_tol = self.svd_tol
        self.coarser, cs.finer = cs, self
        self.agg = [None] * self.nforms
        for j in range(self.jstart, self.nforms):
            self.agg[j] = DofAgg(topo, self.dof[j])
        self._Pcoo = [None] * self.nforms
        self._func = [None] * self.nforms
        for codim in range(self.nforms):
            j = self.nforms - codim - 1
            if j < self.jstart:
                break
            cs.dof[j] = DofHandler(codim, ctopo)
            cs.dof[j].alg_init()
            self._Pcoo[j] = ([], [], [])
            self._func[j] = [dict() for _ in range(codim + 1)]
            self._coarse_traces(j)
            if codim > 0:
                self._h_facet_extension(j)
                if codim > 1:
                    self._h_ridge_peak_extension(j, self.nforms - j - 3)
                    if codim > 2:
                        self._h_ridge_peak_extension(j, self.nforms - j - 4)
            self.P[j] = self._P_csr(j, cs.dof[j].ndofs)
            if codim > 0:
                r, c, v = self._Dcoo[j]
                cs.D[j] = _raw_coo(np.concatenate(r), np.concatenate(c), np.concatenate(v),
                                   (cs.dof[j + 1].ndofs, cs.dof[j].ndofs))
            self.Pi[j] = self._compute_projector(j)
        cs.targets = [None] * self.nforms
        for j in range(self.jstart, self.nforms):
            cs.targets[j] = self.Pi[j] @ self.targets[j]
        cs.l2_const = self.Pi[self.nforms - 1] @ self.l2_const
        return cs

    def _P_csr(self, j, ncols):
        r, c, v = self._Pcoo[j]
        if len(r) == 0:
            return sp.csr_matrix((self.dof[j].ndofs, ncols))
        return _raw_coo(np.concatenate(r), np.concatenate(c), np.concatenate(v), (self.dof[j].ndofs, ncols))

    def _P_add(self, j, rows, cols, block):
        rows, cols = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)
        if len(rows) == 0 or len(cols) == 0:
            return
        rr, cc = np.meshgrid(rows, cols, indexing="ij")
        r, c, v = self._Pcoo[j]
        r.append(rr.ravel()); c.append(cc.ravel()); v.append(np.asarray(block, dtype=np.float64).ravel())

    def _midx(self, j, c):
        return (j, c)

    # ---- traces (DeRhamSequence.cpp:1521-2085)
    def _coarse_traces(self, j):
        cs = self.coarser
        codim = self.ndim - j
        agg = self.agg[j]
        cdof = cs.dof[j]
        nAE = agg.nAE(codim)
        pv = self.pv_traces(codim)
        if j == 0:      # Compute0formCoarseTraces
            for a in range(nAE):
                ints, _ = agg.dofs(codim, a)
                assert len(ints) == 1, "topology error: disconnected coarse peak"
                self._P_add(0, ints, [a], np.ones((1, 1)))
                cdof.dof_type[a] = RANGET
                cdof.n_rangeT[codim][a] = 1
                self._func[0][codim][a] = np.ones((1, 1))
            cdof.build_entity_dof_table(codim)
            cs.M[(0, codim)] = sp.identity(nAE, format="csr")
            return
        M_d = agg.assemble_agg_matrix(codim, self.M[(j, codim)])
        T = self.targets[j]
        nT = 0 if T is None else T.shape[1]
        ploc, masses, ndofs = [None] * nAE, [None] * nAE, np.zeros(nAE, dtype=np.int64)
        for a in range(nAE):
            s, m, e = agg.rng(codim, a)
            dofs = agg.J[codim][s:e]
            loc_pv = pv[dofs]
            Mloc = M_d[s:e, s:e].toarray()
            pvMpv = float(loc_pv @ (Mloc @ loc_pv))
            if nT > 0:
                lt = T[dofs, :].copy()
                # Deflate(targets, pv, M-inner product)
                sc = -1.0 / pvMpv
                for t in range(nT):
                    lt[:, t] = lt[:, t] + (float(loc_pv @ (Mloc @ lt[:, t])) * sc) * loc_pv
                U, sv = svd_on_weighted(Mloc, lt)
            else:
                U, sv = np.zeros((len(dofs), 0)), np.zeros(0)
            s_max_tol = pvMpv * self.svd_tol
            k = 0
            while k < len(sv) and not (sv[k] < s_max_tol):
                k += 1
            ndofs[a] = k + 1
            p = np.empty((len(dofs), k + 1))
            p[:, 0] = loc_pv
            p[:, 1:] = U[:, :k] * np.sqrt(pvMpv)
            cm = p.T @ (Mloc @ p)
            masses[a] = 0.5 * (cm + cm.T)
            ploc[a] = p
            self._func[j][codim][a] = dof_functional(p, Mloc)
        cnt = 0
        for a in range(nAE):
            cdof.dof_type[cnt] = RANGET
            for q in range(1, ndofs[a]):
                cdof.dof_type[cnt + q] = NULLSPACE
            cnt += int(ndofs[a])
            cdof.n_rangeT[codim][a] = 1
            cdof.n_null[codim][a] = ndofs[a] - 1
        cdof.build_entity_dof_table(codim)
        ED = cdof.entity_dof[codim]
        for a in range(nAE):
            s, m, e = agg.rng(codim, a)
            self._P_add(j, agg.J[codim][s:e], row(ED, a), ploc[a])
        cs.M[(j, codim)] = block_diag_csr(masses)
        self.stats[("trace_null", j)] = int(ndofs.sum() - nAE)

    # ---- local matrices shared by the extension stages
    def _local_setup(self, j, cdom):
        agg, aggp = self.agg[j], self.agg[j + 1]
        M_d = agg.assemble_agg_matrix(cdom, self.M[(j, cdom)])
        D_d = DofAgg.distribute(cdom, self.D[j], aggp, agg)
        W_d = aggp.assemble_agg_matrix(cdom, self.M[(j + 1, cdom)])
        return M_d, D_d, W_d

    def _current_Rt(self, j):
        """TransposeAbstractSparseMatrix(P_[j]) at the start of a stage."""
        r, c, v = self._Pcoo[j]
        ncol = self.coarser.dof[j].ndofs
        if len(r) == 0:
            return sp.csr_matrix((ncol, self.dof[j].ndofs))
        P = _raw_coo(np.concatenate(r), np.concatenate(c), np.concatenate(v), (self.dof[j].ndofs, ncol))
        return sp.csr_matrix(P.T)

    # ---- hFacetExtension (DeRhamSequence.cpp:2214-2581)
    def _h_facet_extension(self, j):
        cs = self.coarser
        cbdr, cdom = self.nforms - j - 1, self.nforms - j - 2
        agg, aggp = self.agg[j], self.agg[j + 1]
        ucd, pcd = cs.dof[j], cs.dof[j + 1]
        M_d, D_d, W_d = self._local_setup(j, cdom)
        Rt = self._current_Rt(j)
        Pp = self.P[j + 1].tocsc()
        nAE = cs.topo.n[cdom]
        Bc = cs.topo.B[cdom]
        EDb = ucd.entity_dof[cbdr]
        self._Dcoo = getattr(self, "_Dcoo", [None] * self.nforms)
        Dr, Dc, Dv = [], [], []
        T = self.targets[j]
        nT = 0 if T is None else T.shape[1]
        counter = Rt.shape[0]
        masses = [None] * nAE
        n_rt = n_null = 0
        for a in range(nAE):
            us, um, ue = agg.rng(cdom, a)
            ps, pm, pe = aggp.rng(cdom, a)
            nu, npi = um - us, pm - ps
            u_int, u_bdr = agg.J[cdom][us:um], agg.J[cdom][um:ue]
            p_int = aggp.J[cdom][ps:pm]
            Mall = M_d[us:ue, us:ue].toarray()
            Mloc, Mib = Mall[:nu, :nu], Mall[:nu, nu:]
            Dall = D_d[ps:pe, us:ue].toarray()
            Wall = W_d[ps:pe, ps:pe].toarray()
            Ball = Wall @ Dall
            Bloc, Bib = Ball[:npi, :nu], Ball[:npi, nu:]
            Wloc = Wall[:npi, :npi]
            # PV dof of form j+1 on this AE and the T block
            pv_c = pcd.rangeT_dofs(cdom, a)
            assert len(pv_c) >= 1
            pvloc = Pp[:, [pv_c[0]]].toarray()[p_int, 0]
            tloc = Wloc @ pvloc
            n = nu + npi + 1
            A = np.zeros((n, n))
            A[:nu, :nu] = Mloc
            A[nu:nu + npi, :nu] = Bloc
            A[:nu, nu:nu + npi] = Bloc.T
            A[nu + npi, nu:nu + npi] = tloc
            A[nu:nu + npi, nu + npi] = tloc
            ldl = LDL(A)
            # (3) harmonic extension of the boundary traces
            cb = np.concatenate([row(EDb, f) for f in row(Bc, a)]).astype(np.int64) if len(row(Bc, a)) else np.zeros(0, dtype=np.int64)
            Rb = Rt[cb][:, u_bdr].toarray().T          # (bdr fine dofs) x (bdr coarse dofs)
            rhs = np.zeros((n, len(cb)))
            rhs[:nu] = -(Mib @ Rb)
            rhs[nu:nu + npi] = -(Bib @ Rb)
            sol = ldl.solve(rhs)
            ext = sol[:nu]
            self._P_add(j, u_int, cb, ext)
            lam = sol[nu + npi]
            drow = np.where(np.abs(lam) > self.smallest_entry, -lam, 0.0)
            Dr.append(np.full(len(cb), pv_c[0])); Dc.append(cb); Dv.append(drow)
            # (4) RangeT bubbles: one per NullSpace dof of form j+1 on this AE
            pnull = pcd.null_dofs(cdom, a)
            nrt = len(pnull)
            ucd.n_rangeT[cdom][a] = nrt
            n_rt += nrt
            c_rt = np.arange(counter, counter + nrt, dtype=np.int64)
            counter += nrt
            bub = np.zeros((nu, 0))
            if nrt:
                sub = Pp[:, pnull].toarray()[p_int, :]
                rhs = np.zeros((n, nrt))
                rhs[nu:nu + npi] = Wloc @ sub
                bub = ldl.solve(rhs)[:nu]
                self._P_add(j, u_int, c_rt, bub)
                for q in range(nrt):
                    Dr.append(np.array([pnull[q]])); Dc.append(np.array([c_rt[q]])); Dv.append(np.array([1.0]))
                    ucd.dof_type[int(c_rt[q])] = RANGET
            # (5) NullSpace dofs: targets minus their extension, SVD, truncate
            nul = np.zeros((nu, 0))
            if nu > nrt and nT > 0:
                tint, tbdr = T[u_int, :].copy(), T[u_bdr, :]
                rhs = np.zeros((n, nT))
                rhs[:nu] = -(Mib @ tbdr)
                rhs[nu:nu + npi] = Bloc @ tint
                sol = ldl.solve(rhs)
                tint = tint - sol[:nu]
                U, sv = svd_on(tint)
                k = 0
                while k < len(sv) and not (sv[k] < self.svd_tol):
                    k += 1
                nul = U[:, :k]
            k = nul.shape[1]
            c_nu = np.arange(counter, counter + k, dtype=np.int64)
            counter += k
            for d in c_nu:
                ucd.dof_type[int(d)] = NULLSPACE
            ucd.n_null[cdom][a] = k
            n_null += k
            if k:
                self._P_add(j, u_int, c_nu, nul)
            # (5') dof functionals of the interior coarse dofs
            self._func[j][cdom][a] = dof_functional(np.hstack([bub, nul]), Mloc)
            # (6) coarse element mass: [bdr | RangeT | Null]
            basis = np.zeros((ue - us, len(cb) + nrt + k))
            basis[:nu, :len(cb)] = ext
            basis[nu:, :len(cb)] = Rb
            basis[:nu, len(cb):len(cb) + nrt] = bub
            basis[:nu, len(cb) + nrt:] = nul
            cm = basis.T @ (Mall @ basis)
            masses[a] = 0.5 * (cm + cm.T)
        ucd.build_entity_dof_table(cdom)
        cs.M[(j, cdom)] = block_diag_csr(masses)
        self._Dcoo[j] = [Dr, Dc, Dv]
        self.stats[("facet_ext", j)] = (n_rt, n_null)

    # ---- hRidgePeakExtension (DeRhamSequence.cpp:2629-3048)
    def _h_ridge_peak_extension(self, j, cdom):
        cs = self.coarser
        ridge_stuff = (cdom == self.nforms - j - 3)
        agg, aggp, aggq = self.agg[j], self.agg[j + 1], self.agg[j + 2]
        ucd, pcd = cs.dof[j], cs.dof[j + 1]
        M_d, D_d, W_d = self._local_setup(j, cdom)
        # minusC = -(D2_d^T W2_d D2_d), GetMinusC (DeRhamSequence.cpp:2583-2607)
        D2_d = DofAgg.distribute(cdom, self.D[j + 1], aggq, aggp)
        W2_d = aggq.assemble_agg_matrix(cdom, self.M[(j + 2, cdom)])
        mC_d = _canon(-(D2_d.T @ W2_d @ D2_d))
        # PDc = P_{j+1} * coarse D_j (as built so far)
        Dr, Dc, Dv = self._Dcoo[j]
        Dcoarse = _raw_coo(np.concatenate(Dr), np.concatenate(Dc), np.concatenate(Dv),
                           (pcd.ndofs, ucd.ndofs)) if len(Dr) else sp.csr_matrix((pcd.ndofs, ucd.ndofs))
        PDc = sp.csr_matrix(self.P[j + 1] @ Dcoarse).tocsc()
        Rt = self._current_Rt(j)
        Pp = self.P[j + 1].tocsc()
        nAE = cs.topo.n[cdom]
        T = self.targets[j]
        nT = 0 if T is None else T.shape[1]
        counter = Rt.shape[0]
        masses = [None] * nAE
        n_rt = n_null = 0
        for a in range(nAE):
            us, um, ue = agg.rng(cdom, a)
            ps, pm, pe = aggp.rng(cdom, a)
            nu, npi = um - us, pm - ps
            u_int, u_bdr = agg.J[cdom][us:um], agg.J[cdom][um:ue]
            p_all = aggp.J[cdom][ps:pe]
            p_int = p_all[:npi]
            Mall = M_d[us:ue, us:ue].toarray()
            Mloc, Mib = Mall[:nu, :nu], Mall[:nu, nu:]
            Dall = D_d[ps:pe, us:ue].toarray()
            Wall = W_d[ps:pe, ps:pe].toarray()
            Ball = Wall @ Dall
            Bloc = Ball[:npi, :nu]
            W_iA = Wall[:npi, :]
            Wloc = Wall[:npi, :npi]
            mC = mC_d[ps:pm, ps:pm].toarray()
            n = nu + npi
            A = np.zeros((n, n))
            A[:nu, :nu] = Mloc
            A[nu:, :nu] = Bloc
            A[:nu, nu:] = Bloc.T
            A[nu:, nu:] = mC
            ldl = LDL(A) if nu > 0 else None
            cb = ucd.dofs_on_bdr(cdom, a)
            Rb = Rt[cb][:, u_bdr].toarray().T if len(cb) else np.zeros((len(u_bdr), 0))
            rhs = np.zeros((n, len(cb)))
            rhs[:nu] = -(Mib @ Rb)
            # -W_iA * (D_loc P_bdr) + W_iA * (P_{j+1} D_c)_loc
            DRb = Dall[:, nu:] @ Rb
            PDloc = PDc[:, cb].toarray()[p_all, :] if len(cb) else np.zeros((len(p_all), 0))
            rhs[nu:] = W_iA @ (PDloc - DRb)
            ext = ldl.solve(rhs)[:nu] if nu > 0 else np.zeros((0, len(cb)))
            self._P_add(j, u_int, cb, ext)
            # (4) RangeT bubbles
            pnull = pcd.null_dofs(cdom, a)
            nrt = len(pnull)
            ucd.n_rangeT[cdom][a] = nrt
            n_rt += nrt
            c_rt = np.arange(counter, counter + nrt, dtype=np.int64)
            counter += nrt
            bub = np.zeros((nu, 0))
            if nrt:
                sub = Pp[:, pnull].toarray()[p_int, :]
                rhs = np.zeros((n, nrt))
                rhs[nu:] = Wloc @ sub
                bub = ldl.solve(rhs)[:nu] if nu > 0 else np.zeros((0, nrt))
                self._P_add(j, u_int, c_rt, bub)
                for q in range(nrt):
                    Dr.append(np.array([pnull[q]])); Dc.append(np.array([c_rt[q]])); Dv.append(np.array([1.0]))
                    ucd.dof_type[int(c_rt[q])] = RANGET
            # (5) NullSpace dofs (ridge extension only)
            nul = np.zeros((nu, 0))
            if ridge_stuff and nu > nrt and nT > 0:
                tint, tbdr = T[u_int, :].copy(), T[u_bdr, :]
                rhs = np.zeros((n, nT))
                rhs[:nu] = -(Mib @ tbdr)
                rhs[nu:] = Bloc @ tint
                sol = ldl.solve(rhs)
                tint = tint - sol[:nu]
                U, sv = svd_on(tint)
                k = 0
                while k < len(sv) and not (sv[k] < self.svd_tol):
                    k += 1
                nul = U[:, :k]
            k = nul.shape[1]
            c_nu = np.arange(counter, counter + k, dtype=np.int64)
            counter += k
            for d in c_nu:
                ucd.dof_type[int(d)] = NULLSPACE
            ucd.n_null[cdom][a] = k
            n_null += k
            if k:
                self._P_add(j, u_int, c_nu, nul)
            self._func[j][cdom][a] = dof_functional(np.hstack([bub, nul]), Mloc)
            basis = np.zeros((ue - us, len(cb) + nrt + k))
            basis[:nu, :len(cb)] = ext
            basis[nu:, :len(cb)] = Rb
            basis[:nu, len(cb):len(cb) + nrt] = bub
            basis[:nu, len(cb) + nrt:] = nul
            cm = basis.T @ (Mall @ basis)
            masses[a] = 0.5 * (cm + cm.T)
        ucd.build_entity_dof_table(cdom)
        cs.M[(j, cdom)] = block_diag_csr(masses)
        self.stats[("ridgepeak_ext", j, cdom)] = (n_rt, n_null)

    # ---- CochainProjector::ComputeProjector (CochainProjector.cpp:219-261,416-441)
    def _compute_projector(self, j):
        cs = self.coarser
        cdof, agg = cs.dof[j], self.agg[j]
        nf, nc = self.dof[j].ndofs, cdof.ndofs
        P = self.P[j]

        def hat(c):
            rows, cols, vals = [], [], []
            for e in range(cs.topo.n[c]):
                F = self._func[j][c][e]
                fi, _ = agg.dofs(c, e)
                ci = cdof.interior_dofs(c, e)
                if F.size == 0:
                    continue
                rr, cc = np.meshgrid(ci, fi, indexing="ij")
                rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(F.ravel())
            if not rows:
                return sp.csr_matrix((nc, nf))
            return _raw_coo(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (nc, nf))

        base = cdof.mcb
        Pi = hat(base)
        for c in range(base - 1, -1, -1):
            h = hat(c)
            Pi = _canon(Pi + h - h @ (P @ Pi))
        return Pi


def _raw_coo(r, c, v, shape):
    """COO -> CSR keeping explicit zeros (no duplicate entries are ever produced by
    the callers: every (row, col) pair is written once, like SparseMatrix::Set)."""
    M = sp.coo_matrix((v, (r, c)), shape=shape).tocsr()
    M.sort_indices()
    return M


# ----------------------------------------------------------------------------
# fine level
# ----------------------------------------------------------------------------
def fine_sequence(mesh, topo=None, upscaling_order=0, alpha=None, beta=None, jstart=0):
    """DeRhamSequence3D_FE at lowest order on a structured hex mesh + upscaling targets
    (SetUpscalingTargets).  alpha / beta: optional per-element weights of the L2 and
    H(div) element mass matrices (ReplaceMassIntegrator in the drivers)."""
    topo = topo or mesh.topology()
    if isinstance(mesh, DeformedHexMesh):
        return _fine_sequence_deformed(mesh, topo, alpha, beta, jstart)
    seq = Sequence(topo, 4)
    seq.mesh = mesh
    seq.jstart = jstart
    hx, hy, hz = mesh.h
    vol = hx * hy * hz
    nel = mesh.nel
    ent_dim = {0: 3, 1: 2, 2: 1, 3: 0}
    # dof handlers: dof of form j == entity of codim 3-j; entity_dof[c] = closure table
    for j in range(4):
        dh = DofHandler(3 - j, topo)
        dh.ndofs = topo.n[3 - j]
        for c in range(3 - j + 1):
            if c == 3 - j:
                dh.entity_dof[c] = sp.identity(topo.n[c], format="csr")
            else:
                dh.entity_dof[c] = topo.conn(c, 3 - j)
        seq.dof[j] = dh
    seq.D = [topo.B[2].copy(), topo.B[1].copy(), _canon(topo.B[0] * (1.0 / vol))]
    a_el = np.ones(nel) if alpha is None else np.asarray(alpha, dtype=np.float64)
    b_el = np.ones(nel) if beta is None else np.asarray(beta, dtype=np.float64)

    def rep(block, n, w=None):
        m = block.shape[0]
        data = np.tile(block.ravel(), n)
        if w is not None:
            data = data * np.repeat(w, m * m)
        rr = (np.arange(n)[:, None] * m + np.repeat(np.arange(m), m)[None, :]).ravel()
        cc = (np.arange(n)[:, None] * m + np.tile(np.arange(m), m)[None, :]).ravel()
        return sp.csr_matrix((data, (rr, cc)), shape=(n * m, n * m))

    def bd(*blocks):
        return np.asarray(sp.block_diag(blocks).todense())

    # form 3 (L2, cell values)
    seq.M[(3, 0)] = rep(np.array([[vol]]), nel, a_el)
    # form 2 (RT0, fluxes): element (local order x-,x+,y-,y+,z-,z+), facet
    if b_el.ndim == 2:
        # diagonal tensor coefficient (beta_x, beta_y, beta_z) per element (VectorFunctionCoefficient in
        # VectorFEMassIntegrator, e.g. the SPE10 inverse permeability): the three axis blocks scale separately
        assert b_el.shape == (nel, 3)
        blocks = [sp.block_diag([hx / (hy * hz) * M1D * b[0], hy / (hx * hz) * M1D * b[1], hz / (hx * hy) * M1D * b[2]]) for b in b_el]
        seq.M[(2, 0)] = sp.block_diag(blocks, format="csr")
    else:
        seq.M[(2, 0)] = rep(bd(hx / (hy * hz) * M1D, hy / (hx * hz) * M1D, hz / (hx * hy) * M1D), nel, b_el)
    seq.M[(2, 1)] = sp.diags(1.0 / mesh.facet_area()).tocsr()
    # form 1 (Nedelec, circulations): element (4 x-edges, 4 y-edges, 4 z-edges), facet, ridge
    K = np.kron(M1D, M1D)
    seq.M[(1, 0)] = rep(bd(hy * hz / hx * K, hx * hz / hy * K, hx * hy / hz * K), nel)
    fm = []
    # x-face: its edges in ascending id are 2 y-edges (at k, k+1) then 2 z-edges (at j, j+1)
    fm.append(rep(bd(hz / hy * M1D, hy / hz * M1D), mesh.nf[0]))
    # y-face: 2 x-edges (k,k+1), 2 z-edges (i,i+1)
    fm.append(rep(bd(hz / hx * M1D, hx / hz * M1D), mesh.nf[1]))
    # z-face: 2 x-edges (j,j+1), 2 y-edges (i,i+1)
    fm.append(rep(bd(hy / hx * M1D, hx / hy * M1D), mesh.nf[2]))
    seq.M[(1, 1)] = sp.block_diag(fm).tocsr()
    seq.M[(1, 2)] = sp.diags(1.0 / mesh.ridge_length()).tocsr()
    # form 0 (H1, vertex values)
    seq.M[(0, 0)] = rep(vol * np.kron(M1D, K), nel)
    seq.M[(0, 1)] = sp.block_diag([rep(hy * hz * K, mesh.nf[0]), rep(hx * hz * K, mesh.nf[1]),
                                   rep(hx * hy * K, mesh.nf[2])]).tocsr()
    seq.M[(0, 2)] = sp.block_diag([rep(hx * M1D, mesh.ne[0]), rep(hy * M1D, mesh.ne[1]),
                                   rep(hz * M1D, mesh.ne[2])]).tocsr()
    seq.M[(0, 3)] = sp.identity(mesh.nv, format="csr")
    seq.l2_const = np.ones(nel)
    # targets (Coefficient.cpp:20-274): scalar monomials / vector monomials e_c x^a y^b z^c
    X = mesh.vertex_coords()

    def monos(order):
        out = []
        for om in range(order + 1):
            for ox in range(om + 1):
                for oy in range(om - ox + 1):
                    out.append((ox, oy, om - ox - oy))
        return out
    assert upscaling_order == 0, "oracle fine level implements upscaling order 0"
    seq.targets[3] = np.ones((nel, 1))
    area, length = mesh.facet_area(), mesh.ridge_length()
    T2 = np.zeros((sum(mesh.nf), 3))
    T2[:mesh.nf[0], 0] = area[:mesh.nf[0]]
    T2[mesh.nf[0]:mesh.nf[0] + mesh.nf[1], 1] = area[mesh.nf[0]:mesh.nf[0] + mesh.nf[1]]
    T2[mesh.nf[0] + mesh.nf[1]:, 2] = area[mesh.nf[0] + mesh.nf[1]:]
    seq.targets[2] = T2
    T1 = np.zeros((sum(mesh.ne), 3))
    T1[:mesh.ne[0], 0] = length[:mesh.ne[0]]
    T1[mesh.ne[0]:mesh.ne[0] + mesh.ne[1], 1] = length[mesh.ne[0]:mesh.ne[0] + mesh.ne[1]]
    T1[mesh.ne[0] + mesh.ne[1]:, 2] = length[mesh.ne[0] + mesh.ne[1]:]
    seq.targets[1] = T1
    seq.targets[0] = np.stack([X[:, 0] ** a * X[:, 1] ** b * X[:, 2] ** c for (a, b, c) in monos(1)], axis=1)
    return seq


def _fine_sequence_deformed(mesh, topo, alpha, beta, jstart):
    """DeRhamSequence3D_FE at lowest order on trilinear hexahedra (DeRhamSequenceFE.cpp:633-684): mass matrices by
    mfem's Gauss rules, topological D_0, D_1, D_2 = flux / volume, order-0 upscaling targets."""
    seq = Sequence(topo, 4)
    seq.mesh = mesh
    seq.jstart = jstart
    for j in range(4):
        dh = DofHandler(3 - j, topo)
        dh.ndofs = topo.n[3 - j]
        for c in range(3 - j + 1):
            dh.entity_dof[c] = sp.identity(topo.n[c], format="csr") if c == 3 - j else topo.conn(c, 3 - j)
        seq.dof[j] = dh
    vol = mesh.cell_volumes()
    a_el = np.ones(mesh.nel) if alpha is None else np.asarray(alpha, dtype=np.float64)
    b_el = np.ones(mesh.nel) if beta is None else np.asarray(beta, dtype=np.float64)
    # DivergenceInterpolator2 (bilinIntegrators.hpp:272-290): L2 projection of the divergence = net flux / volume
    seq.D = [topo.B[2].copy(), topo.B[1].copy(), _canon(sp.diags(1.0 / vol) @ topo.B[0])]
    seq.M[(3, 0)] = sp.diags(vol * a_el).tocsr()
    seq.M[(2, 0)] = sp.block_diag(list(mesh.rt0_element_mass() * b_el[:, None, None]), format="csr")
    N = mesh.facet_normals()
    seq.M[(2, 1)] = sp.diags(1.0 / np.linalg.norm(N, axis=1)).tocsr()
    seq.l2_const = np.ones(mesh.nel)
    seq.targets[3] = np.ones((mesh.nel, 1))
    seq.targets[2] = N.copy()            # fluxes of e_x, e_y, e_z
    if jstart <= 1:
        seq.M[(1, 0)] = sp.block_diag(list(mesh.nd0_element_mass()), format="csr")
        seq.M[(1, 1)] = sp.block_diag([sp.block_diag(list(m), format="csr") for m in mesh.nd0_facet_mass()], format="csr")
        t = mesh.edge_vectors()
        seq.M[(1, 2)] = sp.diags(1.0 / np.linalg.norm(t, axis=1)).tocsr()   # VolumetricFEMassIntegrator on the edges
        seq.targets[1] = t.copy()        # circulations of e_x, e_y, e_z
    if jstart <= 0:
        seq.M[(0, 0)] = sp.block_diag(list(mesh.h1_element_mass()), format="csr")
        seq.M[(0, 1)] = sp.block_diag([sp.block_diag(list(m), format="csr") for m in mesh.h1_facet_mass()], format="csr")
        L = mesh.ridge_length()
        M1D_ = np.array([[1.0 / 3.0, 1.0 / 6.0], [1.0 / 6.0, 1.0 / 3.0]])
        seq.M[(0, 2)] = sp.block_diag([l * M1D_ for l in L], format="csr")      # straight edges: exact
        seq.M[(0, 3)] = sp.identity(mesh.nv, format="csr")
        X = mesh.vertex_coords()
        seq.targets[0] = np.stack([np.ones(mesh.nv), X[:, 2], X[:, 1], X[:, 0]], axis=1)   # 1, z, y, x
    return seq


def hex_centroids(mesh):
    X = mesh.vertex_coords()
    nx, ny, nz = mesh.dims
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    c = np.zeros((len(i), 3))
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                c += X[mesh.vx(i + di, j + dj, k + dk)]
    return c / 8.0


def build_hierarchy(dims, nlevels, L=(1.0, 1.0, 1.0), alpha=None, beta=None, jstart=0, svd_tol=1e-9, deform=None, partitioner="derefine"):
    """Drivers' steps 3-4 (examples/MultigridTest2Form.cpp:248-375): agglomerate the
    topology nlevels-1 times by derefinement, then Coarsen() level by level.  deform: optional map of the
    (nv, 3) vertex coordinates (trilinear hexahedra; needs jstart >= 2)."""
    mesh = HexMesh(*dims, L=L) if deform is None else DeformedHexMesh(*dims, deform=deform, L=L)
    topos = [mesh.topology()]
    d = dims
    if partitioner == "geometric":
        # testsuite/UpscalingGeneralForm.cpp:249-256,367-385: two levels only, number of partitions = half the element
        # count of the once-coarser mesh ("coarsen more aggressively")
        assert nlevels <= 2
        if nlevels == 2:
            X = mesh.vertex_coords()
            nparts = max(1, (dims[0] * dims[1] * dims[2]) // 8 // 2)
            topos.append(topos[-1].coarsen(geometric_box_partition(hex_centroids(mesh), X.min(axis=0), X.max(axis=0), nparts)))
    else:
        for _ in range(nlevels - 1):
            topos.append(topos[-1].coarsen(refined_partition(d)))
            d = coarse_dims(d)
    seqs = [fine_sequence(mesh, topos[0], alpha=alpha, beta=beta, jstart=jstart)]
    for l in range(nlevels - 1):
        seqs[l].svd_tol = svd_tol
        seqs.append(seqs[l].coarsen())
    return mesh, seqs


# ----------------------------------------------------------------------------
# checks and golden reproduction
# ----------------------------------------------------------------------------
def check_invariants(seq, tol=1e-9):
    """DeRhamSequence::CheckInvariants (DeRhamSequence.cpp:694-970): M_c = P^T M_f P,
    D_{j+1} D_j = 0, D_f P_j = P_{j+1} D_c, Pi P = I, Pi reproduces targets."""
    cs = seq.coarser
    out = {}
    for j in range(seq.jstart, seq.nforms):
        Mf, Mc = seq.mass_operator(j), cs.mass_operator(j)
        P = seq.P[j]
        out[("M", j)] = abs(Mc - P.T @ Mf @ P).max() / max(abs(Mc).max(), 1e-300)
        PiP = (seq.Pi[j] @ P).toarray()
        out[("PiP", j)] = np.abs(PiP - np.eye(PiP.shape[0])).max()
        out[("target", j)] = np.abs(P @ cs.targets[j] - seq.targets[j]).max()
        if j < seq.nforms - 1:
            out[("DP", j)] = abs(seq.D[j] @ P - seq.P[j + 1] @ cs.D[j]).max()
        if j < seq.nforms - 2:
            out[("DD", j)] = abs(cs.D[j + 1] @ cs.D[j]).max()
    return out


def upscaling_errors(form, nref=1, base=(2, 2, 2), partitioner="derefine"):
    """testsuite/UpscalingGeneralForm.cpp (--form F --nref_parallel nref, default cube of
    2x2x2 hexes): A = M_F + D^T W D with essential (zero) data on attributes 2-5, a
    natural boundary term -1 on attribute 1; solve on every level; report
    ||u_h - P u_H||_M and ||D(u_h - P u_H)||_W on the finest level for the coarsest."""
    import scipy.sparse.linalg as spl
    dims = tuple(b * 2 ** nref for b in base)
    mesh, seqs = build_hierarchy(dims, nref + 1, partitioner=partitioner)
    ess = np.array([0, 1, 1, 1, 1, 0])
    nx, ny, nz = dims
    hx, hy, hz = mesh.h
    f = seqs[0]
    nd = f.dof[form].ndofs
    b = np.zeros(nd)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i, j = i.ravel(), j.ravel()
    if form == 0:      # BoundaryLFIntegrator(-1) on z=0
        for di in (0, 1):
            for dj in (0, 1):
                np.add.at(b, mesh.vx(i + di, j + dj, 0), -hx * hy / 4.0)
    elif form == 1:    # (n x f, v) on z=0, f=(1,1,1), n=(0,0,-1): n x f = (1,-1,0)
        for dj in (0, 1):
            np.add.at(b, mesh.ex(i, j + dj, 0), 1.0 * hy / 2.0)
        for di in (0, 1):
            np.add.at(b, mesh.ey(i + di, j, 0), -1.0 * hx / 2.0)
    else:              # (f, v.n) with f=-1 on z=0 (outward normal -z, dof oriented +z)
        b[mesh.fz(i, j, 0)] = 1.0
    sols, Ms, Ws, Ds = [], [], [], []
    rhs = b
    for k, s in enumerate(seqs):
        M, W, D = s.mass_operator(form), s.mass_operator(form + 1), s.D[form]
        A = _canon(M + D.T @ W @ D).tolil()
        marker = s.dof[form].mark_bdr_dofs(ess)
        r = rhs.copy()
        idx = np.nonzero(marker)[0]
        A = A.tocsr()
        keep = sp.diags((~marker).astype(float))
        A = keep @ A @ keep + sp.diags(marker.astype(float))
        r[idx] = 0.0
        sols.append(spl.spsolve(A.tocsc(), r))
        Ms.append(M); Ws.append(W); Ds.append(D)
        if k + 1 < len(seqs):
            rhs = s.P[form].T @ rhs
    uH = sols[-1]
    for k in range(len(seqs) - 2, -1, -1):
        uH = seqs[k].P[form] @ uH
    diff = uH - sols[0]
    e_l2 = float(np.sqrt(diff @ (Ms[0] @ diff)))
    dd = Ds[0] @ diff
    e_en = float(np.sqrt(dd @ (Ws[0] @ dd)))
    return e_l2, e_en, seqs


def upscaling_form2_amge(nref=2, base=(2, 2, 2)):
    """examples/Upscaling2FormAMGe.cpp:44-48,86-145,225-296,307-444 with --meshfile none (the structured
    2x2x2-hexahedra cube, nref parallel refinements, nref+1 levels): H(div) problem A = M_2 + D_2^T W D_2,
    all boundary attributes essential with zero data, right-hand side (f, v) with f = e_z
    (VectorFEDomainLFIntegrator) restricted level by level with P^T; every level is solved and the errors
    ||P..P u_H - u_h||_M and ||D(P..P u_H - u_h)||_W against the fine solution are reported for the
    coarsest level first (UpscalingPieces.cpp:143-166).  Returns [(u_err, du_err) coarsest, ..., level 1]."""
    import scipy.sparse.linalg as spl
    dims = tuple(b * 2 ** nref for b in base)
    mesh, seqs = build_hierarchy(dims, nref + 1)
    ess = np.ones(6, dtype=int)
    nx, ny, nz = dims
    hz = mesh.h[2]
    form = 2
    f = seqs[0]
    # RT0 basis of a z-face has unit normal flux: (e_z, phi) = hz/2 per adjacent element
    b = np.zeros(f.dof[form].ndofs)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz + 1), indexing="ij")
    adjacent = np.where((k == 0) | (k == nz), 1.0, 2.0)
    b[mesh.fz(i.ravel(), j.ravel(), k.ravel())] = (hz / 2.0) * adjacent.ravel()
    sols, rhs = [], b
    for lev, s in enumerate(seqs):
        M, W, D = s.mass_operator(form), s.mass_operator(form + 1), s.D[form]
        A = _canon(M + D.T @ W @ D)
        marker = s.dof[form].mark_bdr_dofs(ess)
        keep = sp.diags((~marker).astype(float))
        A = keep @ A @ keep + sp.diags(marker.astype(float))
        r = rhs.copy()
        r[marker] = 0.0
        sols.append(spl.spsolve(A.tocsc(), r))
        if lev + 1 < len(seqs):
            rhs = s.P[form].T @ rhs
    M0, W0, D0 = f.mass_operator(form), f.mass_operator(form + 1), f.D[form]
    out = []
    for lev in range(len(seqs) - 1, 0, -1):
        u = sols[lev]
        for q in range(lev - 1, -1, -1):
            u = seqs[q].P[form] @ u
        d = u - sols[0]
        dd = D0 @ d
        out.append((float(np.sqrt(d @ (M0 @ d))), float(np.sqrt(dd @ (W0 @ dd)))))
    return out


def weak_scaling_deformation(X):
    """examples/3DHdivWeakScaling.cpp:148-158: y += exp(z)/2, then x += sin(y)"""
    X = np.array(X, dtype=np.float64)
    X[:, 1] += 0.5 * np.exp(X[:, 2])
    X[:, 0] += np.sin(X[:, 1])
    return X


def hdiv_weak_scaling_errors(nref=2, base=(1, 1, 1)):
    """examples/3DHdivWeakScaling.cpp on one rank (--nref_parallel nref): the unit cube of base cells refined nref
    times, deformed (weak_scaling_deformation), nref+1 levels; H(div) problem A = M_2 + D_2^T W D_2 with
    essential (zero) data on attributes 2-5, natural data -1 on attribute 1 (z = 0), 0 on attribute 6
    (:53-66,175-180,240-245); every level is solved and the errors against the fine solution are reported,
    coarsest level first (UpscalingPieces.cpp:143-166).  Returns [(u_err, du_err), ...]."""
    import scipy.sparse.linalg as spl
    dims = tuple(b * 2 ** nref for b in base)
    mesh, seqs = build_hierarchy(dims, nref + 1, jstart=2, deform=weak_scaling_deformation)
    ess = np.array([0, 1, 1, 1, 1, 0])
    nx, ny, nz = dims
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    b = np.zeros(seqs[0].dof[2].ndofs)
    b[mesh.fz(i.ravel(), j.ravel(), 0)] = 1.0           # (f, v.n), f = -1, outward normal -z, dof oriented +z
    sols, rhs = [], b
    for lev, s in enumerate(seqs):
        M, W, D = s.mass_operator(2), s.mass_operator(3), s.D[2]
        A = _canon(M + D.T @ W @ D)
        marker = s.dof[2].mark_bdr_dofs(ess)
        keep = sp.diags((~marker).astype(float))
        A = keep @ A @ keep + sp.diags(marker.astype(float))
        r = rhs.copy()
        r[marker] = 0.0
        sols.append(spl.spsolve(A.tocsc(), r))
        if lev + 1 < len(seqs):
            rhs = s.P[2].T @ rhs
    f = seqs[0]
    M0, W0, D0 = f.mass_operator(2), f.mass_operator(3), f.D[2]
    out = []
    for lev in range(len(seqs) - 1, 0, -1):
        u = sols[lev]
        for q in range(lev - 1, -1, -1):
            u = seqs[q].P[2] @ u
        d = u - sols[0]
        dd = D0 @ d
        out.append((float(np.sqrt(d @ (M0 @ d))), float(np.sqrt(dd @ (W0 @ dd)))))
    return out


def hcurl_weak_scaling_errors(nref=2, base=(1, 1, 1)):
    """examples/3DHcurlWeakScaling.cpp on one rank (--nref_parallel nref): the deformed cube of
    hdiv_weak_scaling_errors, H(curl) problem A = M_1 + D_1^T M_2 D_1, essential (zero) data on attributes 2-5,
    natural data (n x f, v), f = (1, 1, 1), on attribute 1 (:181-184,247-256).  Errors against the fine solution,
    coarsest level first.  All levels are solved directly here; the reference solves iteratively (rtol 1e-6)."""
    import scipy.sparse.linalg as spl
    dims = tuple(b * 2 ** nref for b in base)
    mesh, seqs = build_hierarchy(dims, nref + 1, jstart=1, deform=weak_scaling_deformation)
    ess = np.array([0, 1, 1, 1, 1, 0])
    rhs = mesh.boundary_tangent_rhs_bottom((1.0, 1.0, 1.0))
    sols = []
    for lev, s in enumerate(seqs):
        M, W, D = s.mass_operator(1), s.mass_operator(2), s.D[1]
        A = _canon(M + D.T @ W @ D)
        marker = s.dof[1].mark_bdr_dofs(ess)
        keep = sp.diags((~marker).astype(float))
        A = keep @ A @ keep + sp.diags(marker.astype(float))
        r = rhs.copy()
        r[marker] = 0.0
        sols.append(spl.spsolve(A.tocsc(), r))
        if lev + 1 < len(seqs):
            rhs = s.P[1].T @ rhs
    f = seqs[0]
    M0, W0, D0 = f.mass_operator(1), f.mass_operator(2), f.D[1]
    out = []
    for lev in range(len(seqs) - 1, 0, -1):
        u = sols[lev]
        for q in range(lev - 1, -1, -1):
            u = seqs[q].P[1] @ u
        d = u - sols[0]
        dd = D0 @ d
        out.append((float(np.sqrt(d @ (M0 @ d))), float(np.sqrt(dd @ (W0 @ dd)))))
    return out


def logical_partition(elem_elem, logical, ratio):
    """LogicalPartitioner::Partition with CoarsenLogicalCartesianOperatorMaterialId (src/partitioning/LogicalPartitioner.hpp:46-103,
    CartesianPartitioner.hpp:133-150): flood fill over the element-element table; an element joins the partition of a
    neighbour when both have the same coarse logical index (i / rx, j / ry, k / rz, material id); partitions are numbered
    in the order a scan of the elements starts them.  logical: (n, 4) integer array (i, j, k, material id)."""
    logical = np.asarray(logical, dtype=np.int64)
    coarse = logical.copy()
    coarse[:, :3] //= np.asarray(ratio, dtype=np.int64)
    n = len(logical)
    part = -np.ones(n, dtype=np.int64)
    I, J = elem_elem.indptr, elem_elem.indices
    nparts = 0
    for e in range(n):
        if part[e] >= 0:
            continue
        part[e] = nparts
        queue = [e]
        head = 0
        while head < len(queue):
            i = queue[head]
            head += 1
            for k in J[I[i]:I[i + 1]]:
                if part[k] < 0 and np.array_equal(coarse[i], coarse[k]):
                    part[k] = nparts
                    queue.append(k)
        nparts += 1
    return part


def logical_demo_material_ids(N=(12, 12, 12)):
    """examples/LogicalPartitionerDemo.cpp:144-158: attribute 1 everywhere; in every layer k the four corner elements
    and the element k Nx Ny + int(Nx Ny / 2 - Nx / 2) get an attribute of their own, counting up from 2"""
    nx, ny, nz = N
    mat = np.ones(nx * ny * nz, dtype=np.int64)
    attr = 2
    for kk in range(nz):
        for e in (kk * nx * ny, kk * nx * ny + nx - 1, kk * nx * ny + nx * ny - 1, kk * nx * ny + nx * ny - nx,
                  kk * nx * ny + int(0.5 * nx * ny - 0.5 * nx)):
            mat[e] = attr
            attr += 1
    return mat


def logical_demo_topologies(N=(12, 12, 12), nlevels=4, ratio=(2, 2, 2), material=None):
    """examples/LogicalPartitionerDemo.cpp:203-226: LogicalPartitioner level after level,
    CoarsenLocalPartitioning(partitioning, check_topology = 1, preserve_material_interfaces = 1).
    Returns (mesh, [topology per level], [messages per coarsening step])."""
    nx, ny, nz = N
    mesh = HexMesh(nx, ny, nz)
    mat = logical_demo_material_ids(N) if material is None else np.asarray(material, dtype=np.int64)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    logical = np.stack([i.ravel(), j.ravel(), k.ravel(), mat], axis=1)
    topos, messages = [mesh.topology()], []
    for l in range(nlevels - 1):
        t = topos[l]
        part = logical_partition(t.element_element(), logical, ratio)
        topos.append(t.coarsen(part, check_topology=True, preserve_material_interfaces=True))
        messages.append(list(t.messages))
        first = t.AE_entity[0].indices[t.AE_entity[0].indptr[:-1]]          # ComputeCoarseLogical: the first element decides
        logical = logical[first].copy()
        logical[:, :3] //= np.asarray(ratio, dtype=np.int64)
    return mesh, topos, messages


def h1_upscaling_errors(levels):
    """examples/LogicalPartitionerDemo.cpp:330-470: A = M_0 + D_0^T M_1 D_0, zero essential data on the whole boundary,
    right-hand side (1, v) restricted with P^T, every level solved; errors of the coarsest level first against the fine
    solution.  levels: list of (M_0, M_1, D_0, boundary marker, P_0 or None) per level, fine first."""
    import scipy.sparse.linalg as spl
    M0, W0, D0 = levels[0][0], levels[0][1], levels[0][2]
    rhs = np.asarray(M0.sum(axis=1)).ravel()
    sols = []
    for lev, (M, W, D, marker, P) in enumerate(levels):
        keep = sp.diags((~marker).astype(float))
        A = keep @ _canon(M + D.T @ W @ D) @ keep + sp.diags(marker.astype(float))
        r = rhs.copy()
        r[marker] = 0.0
        sols.append(spl.spsolve(sp.csc_matrix(A), r))
        if P is not None:
            rhs = P.T @ rhs
    out = []
    for lev in range(len(levels) - 1, 0, -1):
        u = sols[lev]
        for q in range(lev - 1, -1, -1):
            u = levels[q][4] @ u
        d = u - sols[0]
        dd = D0 @ d
        out.append((float(np.sqrt(d @ (M0 @ d))), float(np.sqrt(dd @ (W0 @ dd)))))
    return out


def logical_partitioner_demo_errors(N=(12, 12, 12), nlevels=4, ratio=(2, 2, 2), return_all=False):
    """examples/LogicalPartitionerDemo.cpp (--Nx 12 --Ny 12 --Nz 12, one rank): Cartesian mesh whose four corner
    columns and one interior column carry a material id of their own per layer, logical Cartesian agglomeration by
    2 x 2 x 2 with the material ids kept apart and the topology check on (agglomerates pierced by the interior column
    have a tunnel and are de-agglomerated), all four forms coarsened (jFormStart = 0), then the H1 problem on every level.
    Returns [(u_err, du_err) for the coarsest level, ..., level 1] against the fine solution."""
    mesh, topos, messages = logical_demo_topologies(N, nlevels, ratio)
    seqs = [fine_sequence(mesh, topos[0], jstart=0)]
    for l in range(nlevels - 1):
        seqs[l].svd_tol = 1e-9
        seqs.append(seqs[l].coarsen())
    ess = np.ones(6, dtype=int)
    levels = [(s.mass_operator(0), s.mass_operator(1), s.D[0], s.dof[0].mark_bdr_dofs(ess), s.P[0] if l + 1 < nlevels else None)
              for l, s in enumerate(seqs)]
    out = h1_upscaling_errors(levels)
    if return_all:
        return out, topos, seqs, messages
    return out


def embedded_demo_errors(nref=2, return_all=False):
    """examples/EmbeddedMeshPartitionerDemo.cpp with --mesh none --par_ref_levels 2 (structured path): the 2 x 2 x 2 cube
    refined nref times, derefinement agglomeration (the material ids of the two coarse layers never cut a parent element)
    with the topology check on (CoarsenLocalPartitioning(partitioning, 1, 0)), all four forms coarsened, then the H1
    problem A = M_0 + D_0^T M_1 D_0 with ESSENTIAL DATA u = 1 on the whole boundary and no load: the boundary data of
    the coarse levels is the cochain projection Pi of the fine data (:436-446), eliminated with EliminateRowCol (:483-487).
    Returns [(u_err, du_err) for the coarsest level, ..., level 1] against the fine solution."""
    import scipy.sparse.linalg as spl
    n = 2 * 2 ** nref
    dims = (n, n, n)
    mesh = HexMesh(*dims)
    topos = [mesh.topology()]
    d = dims
    messages = []
    for _ in range(nref):
        topos.append(topos[-1].coarsen(refined_partition(d), check_topology=True))
        messages += topos[-2].messages
        d = coarse_dims(d)
    seqs = [fine_sequence(mesh, topos[0], jstart=0)]
    for l in range(nref):
        seqs[l].svd_tol = 1e-9
        seqs.append(seqs[l].coarsen())
    form = 0
    ess = np.ones(6, dtype=int)
    marker0 = seqs[0].dof[form].mark_bdr_dofs(ess)
    ess_data = [np.where(marker0, 1.0, 0.0)]          # ProjectBdrCoefficient(1): nodal values on the boundary vertices
    for l in range(nref):
        ess_data.append(seqs[l].Pi[form] @ ess_data[l])
    sols = []
    for lev, s in enumerate(seqs):
        M, W, D = s.mass_operator(form), s.mass_operator(form + 1), s.D[form]
        A = _canon(M + D.T @ W @ D).tocsc()
        marker = s.dof[form].mark_bdr_dofs(ess)
        g = np.where(marker, ess_data[lev], 0.0)
        r = -(A @ g)                                    # EliminateRowCol: rhs -= A[:, m] g_m, then rhs_m = g_m
        keep = sp.diags((~marker).astype(float))
        Ae = keep @ A @ keep + sp.diags(marker.astype(float))
        r[marker] = g[marker]
        sols.append(spl.spsolve(sp.csc_matrix(Ae), r))
    M0, W0, D0 = seqs[0].mass_operator(form), seqs[0].mass_operator(form + 1), seqs[0].D[form]
    out = []
    for lev in range(len(seqs) - 1, 0, -1):
        u = sols[lev]
        for q in range(lev - 1, -1, -1):
            u = seqs[q].P[form] @ u
        dlt = u - sols[0]
        dd = D0 @ dlt
        out.append((float(np.sqrt(dlt @ (M0 @ dlt))), float(np.sqrt(dd @ (W0 @ dd)))))
    if return_all:
        return out, seqs, messages
    return out
