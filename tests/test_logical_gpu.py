"""examples/LogicalPartitionerDemo.cpp through the product path: the logical Cartesian partitioner with material ids
and the topology check (host), then Coarsen() of all four forms on the GPU over the irregular agglomerates that result
(single-element agglomerates next to 8- and 64-element ones, de-agglomerated blocks), four levels.  The reference's own
golden (examples/CMakeLists.txt:104-110) is recomputed from the PRODUCT's operators, then every level is compared with
the oracle entry by entry."""
import json
import os
import re

import numpy as np
import pytest

from parelag_b200 import api
from oracle import amge
from tests.test_coarsen_gpu import compare_levels

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_topology_messages.json")))


def test_logical_partitioner_demo_golden_from_the_product_operators():
    api.session()
    N, nlev = (12, 12, 12), 4
    out, topos, seqs, messages = amge.logical_partitioner_demo_errors(return_all=True)
    api.set_topology_options("logical", True, amge.logical_demo_material_ids(N))
    try:
        S = api.Sequence.hex(N, nlev, jstart=0)
        assert api.topology_log() == [m for step in messages for m in step]
    finally:
        api.set_topology_options()
    levels = [(S.get_csr(l, "M", 0), S.get_csr(l, "M", 1), S.get_csr(l, "D", 0), S.get_bdr_mask(l, 0) != 0,
               S.get_csr(l, "P", 0) if l + 1 < nlev else None) for l in range(nlev)]
    errs = amge.h1_upscaling_errors(levels)
    text = "u l2-like errors: %s \nu energy-like errors: %s" % (" ".join("%.4e" % a for a, _ in errs), " ".join("%.4e" % b for _, b in errs))
    assert any(re.search(alt, text) for alt in GOLD["logical_partitioner"]["pass_regular_expression"].split(";")), text
    compare_levels(S, seqs, tol=1e-10, null_tol=1e-8)
    S.free()


def test_geometric_partitioner_coarsen_on_the_gpu():
    """UpscalingGeneralForm --geometric on an anisotropic box: agglomerates of 3 x 2 x 2 hexahedra"""
    api.session()
    dims, L = (6, 4, 2), (1.5, 1.0, 0.5)
    mesh = amge.HexMesh(*dims, L=L)
    topo = mesh.topology()
    X = mesh.vertex_coords()
    topo.coarsen(amge.geometric_box_partition(amge.hex_centroids(mesh), X.min(axis=0), X.max(axis=0), 48 // 16))
    seqs = [amge.fine_sequence(mesh, topo, jstart=0)]
    seqs[0].svd_tol = 1e-9
    seqs.append(seqs[0].coarsen())
    api.set_topology_options("geometric", True)
    try:
        S = api.Sequence.hex(dims, 2, L=L, jstart=0)
    finally:
        api.set_topology_options()
    compare_levels(S, seqs, tol=1e-10)
    S.free()
