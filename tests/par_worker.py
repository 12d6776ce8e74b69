"""Worker of the world_size-2 host-logic tests (launched by tests/test_par_cpu.py through
torch.distributed.run, gloo backend, CPU only): SharingMap numbering, box-decomposed topology
hierarchy, Assemble / IgnoreNonLocalRange into hypre's ParCSR layout, comm package -- all checked
against the oracle's single-domain objects on the undecomposed mesh."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parelag_b200 import api, par          # noqa: E402
from oracle import amge, drivers           # noqa: E402


def gather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def main():
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    comm = par.HostComm()
    api.set_host_comm(comm)

    # ---- T1: true numbering of shared items (two-rank layout)
    if size == 2:
        t1_numbering(comm, rank)
    run_boxes(comm, rank, size)


def t1_numbering(comm, rank):
    keys = [100, 101, 102, 200 + rank]                      # 100..102 held by both ranks, one private item each
    sharers = [[0, 1], [0, 1], [0, 1], [rank]]
    if rank == 1:
        keys = keys[::-1]; sharers = sharers[::-1]           # different local order
    gid, owner, start, cnt, tot = par.number_items(comm, keys, sharers)
    assert tot == 5 and cnt == (4 if rank == 0 else 1), (cnt, tot)
    allk = gather(dict(zip(keys, gid.tolist())))
    assert all(allk[0][k] == allk[1][k] for k in (100, 101, 102))
    assert sorted(set(allk[0].values()) | set(allk[1].values())) == list(range(5))
    assert all(o == 0 for k, o in zip(keys, owner) if k < 200)



def run_boxes(comm, rank, size):
    # ---- T2: box-decomposed hierarchy (topology + fine dof maps), 2x1x1 / 2x2x1 / 2x2x2 boxes of 4^3 hexahedra
    n, lev = 4, 3
    procs = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[size]
    S = api.Sequence.hex_par(procs, (n, n, n), lev, L=(1.0, 1.0, 1.0), jstart=1, svd_tol=-1.0)
    N = (procs[0] * n, procs[1] * n, procs[2] * n)
    expect = {1: N[0] * (N[1] + 1) * (N[2] + 1) + (N[0] + 1) * N[1] * (N[2] + 1) + (N[0] + 1) * (N[1] + 1) * N[2],
              2: (N[0] + 1) * N[1] * N[2] + N[0] * (N[1] + 1) * N[2] + N[0] * N[1] * (N[2] + 1),
              3: N[0] * N[1] * N[2]}
    maps = {}
    for form in (1, 2, 3):
        m = S.dofmap(0, form)
        maps[form] = m
        assert m["nglobal"] == expect[form], (form, m["nglobal"], expect[form])
        counts = gather(m["ntrue"])
        assert sum(counts) == expect[form]
        assert m["start"] == sum(counts[:rank])
        both = gather(dict(zip(m["key"].tolist(), m["gid"].tolist())))
        holders = {}
        for r in range(size):
            for k, g in both[r].items():
                if k in both[rank]:
                    assert both[rank][k] == g                # one true id per key on every holder
                holders.setdefault(k, []).append(r)
        allg = set()
        for r in range(size):
            allg |= set(both[r].values())
        assert allg == set(range(expect[form]))
        # a shared dof is owned by the smallest rank holding it
        for i, k in enumerate(m["key"].tolist()):
            assert m["owner"][i] == min(holders[k]), (form, k, m["owner"][i], holders[k])
        if size == 2:
            shared = set(both[0]) & set(both[1])
            assert len(shared) == {1: 2 * n * (n + 1), 2: n * n, 3: 0}[form]
    # pseudo boundary attribute on the interface: the coarse facets of both sides match one to one
    FB = S.get_csr(1, "FB")
    if size == 2:
        iface_attr = 7 if rank == 0 else 6                    # x+ face of rank 0, x- face of rank 1
        ncf = int((FB.indices == iface_attr).sum())
        assert ncf == (n // 2) ** 2 and FB.shape[1] == 12, (ncf, FB.shape)

    # ---- T3: Assemble(dofTrueDof, A_local, dofTrueDof) == the single-domain operator
    mesh_loc = amge.HexMesh(n, n, n, L=(1.0, 1.0, 1.0))
    seq_loc = amge.fine_sequence(mesh_loc, jstart=1)
    A_loc, _ = drivers.system_matrix(seq_loc, 2, np.zeros(6, dtype=np.int32))
    m = maps[2]
    rr = (m["start"], m["start"] + m["ntrue"])
    M = par.assemble(comm, 0, A_loc, m["gid"], m["owner"], m["gid"], m["owner"], rr, m["nglobal"], rr, m["nglobal"])
    d = M.arrays()
    assert np.all(np.diff(d["col_map_offd"]) > 0)
    _, blk = par.parcsr_rows_to_global(d)
    blocks = gather((d["first_row"], blk))
    A_true = sp.vstack([b for _, b in sorted(blocks, key=lambda t: t[0])]).tocsr()
    mesh_g = amge.HexMesh(*N, L=tuple(float(p) for p in procs))
    seq_g = amge.fine_sequence(mesh_g, jstart=1)
    A_g, _ = drivers.system_matrix(seq_g, 2, np.zeros(6, dtype=np.int32))
    # true id -> global face number of the undecomposed mesh (the dof key carries it)
    perm = np.empty(m["nglobal"], dtype=np.int64)
    for mm in gather((m["gid"], (m["key"] >> 8) & ((1 << 52) - 1))):
        perm[mm[0]] = mm[1]
    Pm = sp.csr_matrix((np.ones(len(perm)), (np.arange(len(perm)), perm)), shape=(len(perm), len(perm)))
    A_perm = (Pm.T @ A_true @ Pm).tocsr()
    A_perm.sort_indices(); A_g.sort_indices()
    assert np.array_equal(A_perm.indptr, A_g.indptr) and np.array_equal(A_perm.indices, A_g.indices)
    assert np.abs(A_perm.data - A_g.data).max() <= 1e-13 * np.abs(A_g.data).max()
    # comm package: what I receive is what the neighbour sends, ghost by ghost
    sends = gather({int(p): (d["send_map_elmts"][d["send_map_starts"][k]:d["send_map_starts"][k + 1]] + d["first_col"]).tolist()
                    for k, p in enumerate(d["send_procs"])})
    for k, p in enumerate(d["recv_procs"]):
        want = d["col_map_offd"][d["recv_vec_starts"][k]:d["recv_vec_starts"][k + 1]].tolist()
        assert sends[int(p)][rank] == want
    M.free()

    # ---- T3b: SharingMap::Assemble / Distribute on vectors
    ones = S.assemble_vector(0, 2, np.ones(len(m["gid"])))
    assert set(np.unique(ones).tolist()) <= {1.0, 2.0}                      # a face has at most two holders
    assert int(sum(gather(float(ones.sum())))) == int(sum(gather(len(m["gid"]))))
    tv = np.arange(m["start"], m["start"] + m["ntrue"], dtype=np.float64)   # true vector = its own global id
    loc = S.distribute_vector(0, 2, tv)
    assert np.array_equal(loc, m["gid"].astype(np.float64))

    # ---- T4: IgnoreNonLocalRange keeps the owner's rows only (D_2 : H(div) -> L2)
    D = sp.csr_matrix(seq_loc.D[2])
    m3 = maps[3]
    Dm = par.assemble(comm, 1, D, m3["gid"], m3["owner"], m["gid"], m["owner"], (m3["start"], m3["start"] + m3["ntrue"]),
                      m3["nglobal"], rr, m["nglobal"])
    dd = Dm.arrays()
    _, blk = par.parcsr_rows_to_global(dd)
    D_true = sp.vstack([b for _, b in sorted(gather((dd["first_row"], blk)), key=lambda t: t[0])]).tocsr()
    perm3 = np.empty(m3["nglobal"], dtype=np.int64)
    for mm in gather((m3["gid"], (m3["key"] >> 8) & ((1 << 52) - 1))):
        perm3[mm[0]] = mm[1]
    P3 = sp.csr_matrix((np.ones(len(perm3)), (np.arange(len(perm3)), perm3)), shape=(len(perm3), len(perm3)))
    D_perm = (P3.T @ D_true @ Pm).tocsr()
    D_g = sp.csr_matrix(seq_g.D[2])
    assert abs(D_perm - D_g).max() <= 1e-13 * abs(D_g).max()
    Dm.free()
    S.free()

    # ---- T5: the 3DHdivWeakScaling geometry (trilinear hexahedra, examples/3DHdivWeakScaling.cpp:148-158) on the box
    # decomposition: every rank builds the fine sequence of ITS box from its own vertices; the assembled H(div) operator
    # equals the single-domain operator of the undecomposed deformed mesh
    Lg = tuple(float(p) for p in procs)
    X = api.box_vertex_coords(procs, api.rank_box(procs, rank), (n, n, n), api.weak_scaling_deformation, domain=Lg)
    Sd = api.Sequence.hex_par(procs, (n, n, n), lev, jstart=1, svd_tol=-1.0, coords=X)
    Ml, Wl, Dl = Sd.get_csr(0, "M", 2), Sd.get_csr(0, "M", 3), Sd.get_csr(0, "D", 2)
    A_loc = sp.csr_matrix(Ml + Dl.T @ Wl @ Dl)
    md = Sd.dofmap(0, 2)
    assert np.array_equal(md["gid"], m["gid"])                                 # numbering does not depend on the geometry
    Md = par.assemble(comm, 0, A_loc, md["gid"], md["owner"], md["gid"], md["owner"], rr, md["nglobal"], rr, md["nglobal"])
    _, blk = par.parcsr_rows_to_global(Md.arrays())
    A_true = sp.vstack([b for _, b in sorted(gather((Md.arrays()["first_row"], blk)), key=lambda t: t[0])]).tocsr()
    mesh_d = amge.DeformedHexMesh(*N, deform=amge.weak_scaling_deformation, L=Lg)
    assert np.array_equal(mesh_d.vertex_coords()[mesh_d.vx(*[rank_off for rank_off in (api.rank_box(procs, rank)[0] * n,
                                                                                      api.rank_box(procs, rank)[1] * n,
                                                                                      api.rank_box(procs, rank)[2] * n)])], X[0])
    seq_d = amge.fine_sequence(mesh_d, jstart=1)
    A_d = sp.csr_matrix(seq_d.mass_operator(2) + seq_d.D[2].T @ seq_d.mass_operator(3) @ seq_d.D[2])
    A_dp = (Pm.T @ A_true @ Pm).tocsr()
    assert abs(A_dp - A_d).max() <= 1e-13 * abs(A_d).max()
    Md.free()
    Sd.free()
    dist.barrier()
    if rank == 0:
        print("PAR_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
