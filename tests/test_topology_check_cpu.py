"""Topology checks of CoarsenLocalPartitioning(partitioning, check_topology = true) -- connectedComponents,
AgglomeratedTopologyCheck (Betti numbers of the agglomerated entities, boundary connectivity), de-agglomeration of bad
entities -- and the geometric box partitioner.  CPU only.

(1) The oracle (oracle/amge.py) is pinned to the reference's own goldens: the PASS_REGULAR_EXPRESSIONs of the eight
    `twentyseven.exe` topology tests (testsuite/CMakeLists.txt:34-92), of `geometric_form1` (:187-193) and of
    `geometric_partitioner` (:254-258), committed in tests/golden/reference_topology_messages.json.  The entity NUMBERS in
    "Facet 3 is disconnected." / "Ridge 16 is disconnected." are reproduced too: the oracle runs on mfem's face / edge
    numbering of the Cartesian mesh (amge.mfem_hex_numbering), which the minimal intersection sets inherit.
(2) The product's host code (parelag_b200/src/amge_topology.hpp) must give the oracle's tables bit for bit and the same
    messages, in the repository's own (lexicographic) numbering."""
import json
import os
import re

import numpy as np
import pytest

from oracle import amge, tets
from parelag_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_topology_messages.json")))


def partitioning(name):
    """the hand-made partitionings of testsuite/twentyseven.cpp:44-180 (3 x 3 x 3 hexahedra, element = i + 3 j + 9 k;
    `discedge`: 3 x 3 x 4; `tet`: the 48 tetrahedra of mfem's 2 x 2 x 2 cube)"""
    if name == "disconnected":
        p = np.ones(27, dtype=np.int64); p[[0, 26]] = 0
    elif name == "donut":
        p = np.ones(27, dtype=np.int64)
        for i in range(3):
            for j in range(3):
                p[9 * i + 3 * j + 1] = 0
        p[13] = 1
    elif name == "void":
        p = np.ones(27, dtype=np.int64); p[13] = 0
    elif name == "discface":
        p = np.zeros(27, dtype=np.int64); p[:9] = 1; p[12:15] = 2
    elif name == "facehole":
        p = np.full(27, 2, dtype=np.int64); p[:9] = 0; p[13] = 1
    elif name == "discedge":
        p = np.full(36, 4, dtype=np.int64)
        for i in range(4):
            p[9 + i] = 0; p[18 + i] = 1
        for i in range(5, 9):
            p[9 + i] = 2; p[18 + i] = 3
        p[:9] = 0; p[27:36] = 3
    elif name == "connectivity":
        p = np.zeros(48, dtype=np.int64); p[8] = 1
    elif name == "sharededge":
        p = np.zeros(27, dtype=np.int64); p[[0, 4, 5, 9, 14, 18, 21, 22, 23]] = 1
    else:
        raise KeyError(name)
    return p


def mfem_tet_cube(n):
    """mfem::Mesh(n, n, n, Element::TETRAHEDRON): every hexahedron (vertices 0..3 counter-clockwise at the bottom, 4..7
    above them) is cut into the six tetrahedra around its diagonal 0-6, in this order (mfem mesh.cpp, Make3D)"""
    g = np.arange(n + 1) / n
    k, j, i = np.meshgrid(g, g, g, indexing="ij")
    V = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    vid = lambda a, b, c: a + (n + 1) * (b + (n + 1) * c)
    hex_to_tet = [(0, 1, 2, 6), (0, 5, 1, 6), (0, 4, 5, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6)]
    T = []
    for c in range(n):
        for b in range(n):
            for a in range(n):
                ind = [vid(a, b, c), vid(a + 1, b, c), vid(a + 1, b + 1, c), vid(a, b + 1, c),
                       vid(a, b, c + 1), vid(a + 1, b, c + 1), vid(a + 1, b + 1, c + 1), vid(a, b + 1, c + 1)]
                T += [[ind[q] for q in t] for t in hex_to_tet]
    T = np.array(T, dtype=np.int64)
    faces = {}
    for t in T:
        for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
            key = tuple(sorted(int(t[q]) for q in f))
            faces[key] = faces.get(key, 0) + 1
    Bt, Ba = [], []
    for key, cnt in faces.items():
        if cnt == 1:
            X = V[list(key)]
            for ax, val, attr in ((2, 0.0, 1), (1, 0.0, 2), (0, 1.0, 3), (1, 1.0, 4), (0, 0.0, 5), (2, 1.0, 6)):
                if np.all(np.abs(X[:, ax] - val) < 1e-12):
                    Bt.append(key); Ba.append(attr)
                    break
    return V, T, np.array(Bt, dtype=np.int64), np.array(Ba, dtype=np.int64)


def oracle_case(name, mfem_numbering):
    if name == "connectivity":
        topo = tets.TetMesh(*mfem_tet_cube(2)).topology()
    else:
        dims = (3, 3, 4) if name == "discedge" else (3, 3, 3)
        topo = amge.HexMesh(*dims).topology()
        if mfem_numbering:
            topo = amge.renumbered_topology(topo, *amge.mfem_hex_numbering(*dims))
    coarse = topo.coarsen(partitioning(name), check_topology=True)
    return topo, coarse


@pytest.mark.parametrize("name", ["disconnected", "donut", "void", "discface", "facehole", "discedge", "connectivity", "sharededge"])
def test_oracle_reproduces_the_reference_topology_goldens(name):
    topo, coarse = oracle_case(name, mfem_numbering=True)
    out = "\n".join(coarse.show_me() + topo.messages)
    assert re.search(GOLD[name]["pass_regular_expression"], out), out
    for c in range(2):      # twentyseven.cpp:305-325: the coarse boundary operators still form a complex
        assert abs(coarse.B[c] @ coarse.B[c + 1]).max() == 0


def test_oracle_geometric_partitioner_goldens():
    """testsuite/CMakeLists.txt:254-258 (test_GeometricBoxPartitioner --x-elem 12 --y-elem 16 --partitions 9: 2-d unit square)
    and :187-193 (UpscalingGeneralForm --form 1 --geometric: form1's numbers)"""
    i, j = np.meshgrid(np.arange(12), np.arange(16), indexing="ij")
    cen = np.stack([(i.ravel() + 0.5) / 12, (j.ravel() + 0.5) / 16], axis=1)
    part = amge.geometric_box_partition(cen, [0.0, 0.0], [1.0, 1.0], 9)
    sizes = np.bincount(part)
    assert re.search(GOLD["geometric_partitioner"]["pass_regular_expression"], "  mean size: %g" % sizes.mean())
    assert (sizes.max(), sizes.min(), len(sizes)) == (24, 20, 9)
    e_l2, e_en, _ = amge.upscaling_errors(1, nref=1, partitioner="geometric")
    out = "u l2-like errors: %.4e \nu energy-like errors: %.4e" % (e_l2, e_en)
    assert re.search(GOLD["geometric_form1"]["pass_regular_expression"], out), out


def same(A, B):
    A = A.tocsr(); A.sort_indices(); B = B.tocsr(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data))


@pytest.mark.parametrize("name", ["disconnected", "donut", "void", "discface", "facehole", "discedge", "connectivity", "sharededge"])
def test_product_topology_check_matches_oracle(name):
    """same partitionings through the product (C ABI: pe_api_set_topology_options with the given partitioning and
    check_topology = 1): messages equal line by line, AE -> entity tables and coarse boundary operators bit-exact"""
    topo, coarse = oracle_case(name, mfem_numbering=False)
    api.set_topology_options("user", True, partitioning(name))
    try:
        if name == "connectivity":
            S = api.Sequence.tet(*mfem_tet_cube(2), 0, 2, svd_tol=-1.0)
        else:
            S = api.Sequence.hex((3, 3, 4) if name == "discedge" else (3, 3, 3), 2, svd_tol=-1.0)
        assert api.topology_log() == topo.messages
        assert S.show_topology(1) == coarse.show_me()
        for c in range(4):
            assert same(S.get_csr(0, "AE", c), topo.AE_entity[c]), c
        for c in range(3):
            assert same(S.get_csr(1, "B", c), coarse.B[c]), c
        assert same(S.get_csr(1, "FB"), coarse.facet_bdr)
        S.free()
    finally:
        api.set_topology_options()


def test_product_geometric_partitioner_matches_oracle():
    """GeometricBoxPartitioner on an anisotropic box (4 boxes of 3 x 2 x 2 hexahedra) and on the unit cube (= derefinement)"""
    for dims, L in (((6, 4, 2), (1.5, 1.0, 0.5)), ((4, 4, 4), (1.0, 1.0, 1.0))):
        mesh = amge.HexMesh(*dims, L=L)
        topo = mesh.topology()
        X = mesh.vertex_coords()
        nparts = max(1, dims[0] * dims[1] * dims[2] // 16)
        part = amge.geometric_box_partition(amge.hex_centroids(mesh), X.min(axis=0), X.max(axis=0), nparts)
        coarse = topo.coarsen(part)
        api.set_topology_options("geometric", True)
        try:
            S = api.Sequence.hex(dims, 2, L=L, svd_tol=-1.0)
            assert api.topology_log() == []
            for c in range(4):
                assert same(S.get_csr(0, "AE", c), topo.AE_entity[c]), c
            for c in range(3):
                assert same(S.get_csr(1, "B", c), coarse.B[c]), c
            S.free()
        finally:
            api.set_topology_options()
    assert np.array_equal(part, amge.refined_partition((4, 4, 4)))


def test_connected_components_renumbering():
    """connectedComponents.cpp:23-87: components of a partition are numbered in scan order after the partitions before it;
    empty partitions vanish"""
    topo = amge.HexMesh(4, 1, 1).topology()
    part = np.array([3, 1, 3, 1])
    n = amge.connected_components(part, topo.element_element())
    assert n == 4 and part.tolist() == [2, 0, 3, 1]


def cmake_regex_matches(pattern, out):
    """PASS_REGULAR_EXPRESSION is a ;-separated list of alternatives"""
    return any(re.search(alt, out) for alt in pattern.split(";"))


def test_oracle_reproduces_the_logical_partitioner_golden():
    """examples/CMakeLists.txt:104-110 (LogicalPartitionerDemo --Nx 12 --Ny 12 --Nz 12): four levels by the logical
    Cartesian partitioner with material ids and the topology check on (the 4 x 4 x 4 blocks pierced by the interior material
    column have a tunnel and are de-agglomerated), Coarsen() of all four forms over irregular agglomerates (single-element
    agglomerates next to 8- and 64-element ones), H1 problem on every level: all six published numbers"""
    out, topos, seqs, messages = amge.logical_partitioner_demo_errors(return_all=True)
    text = "u l2-like errors: %s \nu energy-like errors: %s" % (" ".join("%.4e" % a for a, _ in out), " ".join("%.4e" % b for _, b in out))
    assert cmake_regex_matches(GOLD["logical_partitioner"]["pass_regular_expression"], text), text
    assert any("has 1 tunnels." in m for m in messages[1]) and any("has 1 tunnels." in m for m in messages[2])
    for s in seqs[:-1]:
        for k, v in amge.check_invariants(s).items():
            assert v < 1e-9, (k, v)


def test_product_logical_partitioner_matches_oracle():
    """the product's LogicalPartitioner with material ids + topology check, four levels: tables and messages"""
    N, nlev = (12, 12, 12), 4
    mesh, topos, messages = amge.logical_demo_topologies(N, nlev)
    api.set_topology_options("logical", True, amge.logical_demo_material_ids(N))
    try:
        S = api.Sequence.hex(N, nlev, svd_tol=-1.0)
        assert api.topology_log() == [m for step in messages for m in step]
        for l in range(nlev):
            assert S.show_topology(l) == topos[l].show_me()
            for c in range(3):
                assert same(S.get_csr(l, "B", c), topos[l].B[c]), (l, c)
            assert same(S.get_csr(l, "FB"), topos[l].facet_bdr)
            if l + 1 < nlev:
                for c in range(4):
                    assert same(S.get_csr(l, "AE", c), topos[l].AE_entity[c]), (l, c)
        S.free()
    finally:
        api.set_topology_options()


def test_oracle_reproduces_the_embedded_mesh_partitioner_golden():
    """examples/CMakeLists.txt:122-128 (EmbeddedMeshPartitionerDemo --mesh none --par_ref_levels 2): three levels on the 8^3
    cube, H1 problem with essential data u = 1 on the whole boundary; the coarse boundary data is the cochain projection
    Pi of the fine data -- pins the projector matrix (CochainProjector::ComputeProjector) on non-zero data"""
    out, seqs, messages = amge.embedded_demo_errors(return_all=True)
    text = "u l2-like errors: %s \nu energy-like errors: %s" % (" ".join("%.4e" % a for a, _ in out), " ".join("%.4e" % b for _, b in out))
    assert cmake_regex_matches(GOLD["embedded_mesh_partitioner"]["pass_regular_expression"], text), text
    assert messages == []          # the derefinement agglomerates pass the topology check


@pytest.mark.parametrize("seed", range(24))
def test_random_partitionings_product_equals_oracle(seed):
    """fuzz: random (every third: perturbed blocky) partitionings of small Cartesian meshes with the topology check on --
    disconnected parts, tunnels, holes, pinched boundaries, disconnected facets and ridges in arbitrary combination.
    Product and oracle must report the same lines and produce the same tables bit for bit, and the repaired coarse topology
    must still be a complex (B_c B_{c+1} = 0)."""
    rng = np.random.default_rng(seed)
    dims = tuple(int(x) for x in rng.integers(2, 5, size=3))
    nel = dims[0] * dims[1] * dims[2]
    part = rng.integers(0, int(rng.integers(1, max(2, nel // 3))), size=nel)
    if seed % 3 == 0:
        i, j, k = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
        part = ((i // 2) + 3 * (j // 2) + 9 * (k // 2)).transpose(2, 1, 0).ravel().copy()
        part[rng.integers(0, nel, size=3)] = part[rng.integers(0, nel, size=3)]
    topo = amge.HexMesh(*dims).topology()
    coarse = topo.coarsen(part, check_topology=True)
    for c in range(2):
        assert abs(coarse.B[c] @ coarse.B[c + 1]).max() == 0
    api.set_topology_options("user", True, part)
    try:
        S = api.Sequence.hex(dims, 2, svd_tol=-1.0)
        assert api.topology_log() == topo.messages
        assert S.show_topology(1) == coarse.show_me()
        for c in range(4):
            assert same(S.get_csr(0, "AE", c), topo.AE_entity[c]), c
        for c in range(3):
            assert same(S.get_csr(1, "B", c), coarse.B[c]), c
        assert same(S.get_csr(1, "FB"), coarse.facet_bdr)
        S.free()
    finally:
        api.set_topology_options()


@pytest.mark.parametrize("seed", range(8))
def test_oracle_coarsen_keeps_the_invariants_on_repaired_random_topologies(seed):
    """Coarsen() of all four forms on the topology the check leaves behind for a random partitioning (single-element
    agglomerates, L-shaped agglomerated facets, agglomerates of very different sizes): the CheckInvariants identities
    (DeRhamSequence.cpp:694-970) must hold -- what the GPU path is compared with on such topologies"""
    rng = np.random.default_rng(seed)
    dims = tuple(int(x) for x in rng.integers(2, 5, size=3))
    nel = dims[0] * dims[1] * dims[2]
    part = rng.integers(0, int(rng.integers(1, max(2, nel // 3))), size=nel)
    mesh = amge.HexMesh(*dims)
    topo = mesh.topology()
    topo.coarsen(part, check_topology=True)
    seq = amge.fine_sequence(mesh, topo, jstart=0)
    seq.svd_tol = 1e-9
    seq.coarsen()
    for k, v in amge.check_invariants(seq).items():
        assert v < 1e-8, (k, v)
