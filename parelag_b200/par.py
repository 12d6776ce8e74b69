"""Multi-rank harness glue: the host communicator callback table (include/parelag_b200_par.h)
backed by torch.distributed -- the role MPI_Comm plays in a ParElag build -- and ctypes bindings
of the host-side SharingMap / ParCSR assembly entry points.

Setup-time exchanges run over a gloo group (CPU tensors); the data path of the solver (halo
exchange, dot products) is NCCL inside the C library and never goes through this file."""
import ctypes as C

import numpy as np

from .capi import _chk, _ptr, _i32, _f64, lib, ParCSRHost

_AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
_A2A = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                   C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64))


class PeHostComm(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("user", C.c_void_p), ("allgather", _AG), ("alltoallv", _A2A)]


def _bytes_view(ptr, n):
    if n == 0 or not ptr:
        return np.empty(0, dtype=np.uint8)
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))


class HostComm:
    """pe_host_comm over a torch.distributed process group (gloo), or a trivial one-rank table."""

    def __init__(self, group=None, serial=False):
        self.group = group
        if serial:
            self.rank, self.size, self.dist = 0, 1, None
        else:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self._ag = _AG(self._allgather)
        self._a2a = _A2A(self._alltoallv)
        self.c = PeHostComm(self.rank, self.size, None, self._ag, self._a2a)

    def _allgather(self, user, send, nbytes, recv):
        try:
            s = _bytes_view(send, nbytes)
            r = _bytes_view(recv, nbytes * self.size)
            if self.size == 1:
                r[:] = s
                return 0
            import torch
            out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.size)]
            self.dist.all_gather(out, torch.from_numpy(s.copy()), group=self.group)
            for k, t in enumerate(out):
                r[k * nbytes:(k + 1) * nbytes] = t.numpy()
            return 0
        except Exception as e:      # never let an exception cross the C boundary
            print("HostComm.allgather failed:", repr(e), flush=True)
            return 1

    def _alltoallv(self, user, send, sb, sd, recv, rb, rd):
        try:
            n = self.size
            sb = [sb[i] for i in range(n)]; sd = [sd[i] for i in range(n)]
            rb = [rb[i] for i in range(n)]; rd = [rd[i] for i in range(n)]
            s = _bytes_view(send, sd[-1] + sb[-1])
            r = _bytes_view(recv, rd[-1] + rb[-1])
            if n == 1:
                r[rd[0]:rd[0] + rb[0]] = s[sd[0]:sd[0] + sb[0]]
                return 0
            import torch
            # the C side packs consecutively in rank order
            assert all(sd[i] == sum(sb[:i]) for i in range(n)) and all(rd[i] == sum(rb[:i]) for i in range(n))
            st = torch.from_numpy(s.copy()) if len(s) else torch.empty(0, dtype=torch.uint8)
            rt = torch.empty(sum(rb), dtype=torch.uint8)
            self.dist.all_to_all_single(rt, st, rb, sb, group=self.group)
            if len(r):
                r[:] = rt.numpy()
            return 0
        except Exception as e:
            print("HostComm.alltoallv failed:", repr(e), flush=True)
            return 1

    def ptr(self):
        return C.byref(self.c)


def number_items(comm, key, sharers):
    """pe_par_number_items: sharers = list of rank lists per item -> (gid, owner, start, count, total)"""
    n = len(key)
    key = np.ascontiguousarray(key, dtype=np.int64)
    sI = np.zeros(n + 1, dtype=np.int32)
    sI[1:] = np.cumsum([len(s) for s in sharers])
    sJ = _i32(np.concatenate([np.asarray(s, dtype=np.int32) for s in sharers]) if n else np.empty(0, dtype=np.int32))
    gid, owner = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int32)
    st, cnt, tot = C.c_int64(), C.c_int64(), C.c_int64()
    _chk(lib().pe_par_number_items(comm.ptr(), n, _ptr(key), _ptr(sI), _ptr(sJ), _ptr(gid), _ptr(owner),
                                   C.byref(st), C.byref(cnt), C.byref(tot)))
    return gid, owner, st.value, cnt.value, tot.value


class OwnedParCSR:
    """Host ParCSR (hypre layout) produced by pe_par_assemble."""

    def __init__(self, h):
        self.h = h
        lib().pe_parcsr_owned_view.restype = C.POINTER(ParCSRHost)
        self.view = lib().pe_parcsr_owned_view(h).contents

    def arrays(self):
        v = self.view
        n = v.num_rows

        def arr(p, m, dt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=(m,)).copy() if m else np.empty(0, dtype=dt)
        dI = arr(v.diag_i, n + 1, C.c_int32)
        oI = arr(v.offd_i, n + 1, C.c_int32)
        out = dict(first_row=v.first_row_index, first_col=v.first_col_diag, nrows=n, ncols_diag=v.num_cols_diag,
                   global_rows=v.global_num_rows, global_cols=v.global_num_cols,
                   diag_i=dI, diag_j=arr(v.diag_j, int(dI[-1]), C.c_int32), diag_a=arr(v.diag_data, int(dI[-1]), C.c_double),
                   offd_i=oI, offd_j=arr(v.offd_j, int(oI[-1]), C.c_int32), offd_a=arr(v.offd_data, int(oI[-1]), C.c_double),
                   col_map_offd=arr(v.col_map_offd, v.num_cols_offd, C.c_int64),
                   send_procs=arr(v.send_procs, v.num_sends, C.c_int32),
                   send_map_starts=arr(v.send_map_starts, v.num_sends + 1, C.c_int32),
                   recv_procs=arr(v.recv_procs, v.num_recvs, C.c_int32),
                   recv_vec_starts=arr(v.recv_vec_starts, v.num_recvs + 1, C.c_int32))
        out["send_map_elmts"] = arr(v.send_map_elmts, int(out["send_map_starts"][-1]) if v.num_sends else 0, C.c_int32)
        return out

    def free(self):
        if self.h:
            lib().pe_parcsr_owned_free(self.h)
            self.h = None


def assemble(comm, mode, A, row_gid, row_owner, col_gid, col_owner, row_range, global_rows, col_range, global_cols):
    """pe_par_assemble on a scipy CSR matrix in local dof numbering (mode 0 Assemble, 1 IgnoreNonLocalRange)."""
    A = A.tocsr()
    I, J, V = _i32(A.indptr), _i32(A.indices), _f64(A.data)
    rg, ro = np.ascontiguousarray(row_gid, dtype=np.int64), _i32(row_owner)
    cg, co = np.ascontiguousarray(col_gid, dtype=np.int64), _i32(col_owner)
    h = C.c_void_p()
    _chk(lib().pe_par_assemble(comm.ptr(), mode, A.shape[0], A.shape[1], _ptr(I), _ptr(J), _ptr(V), _ptr(rg), _ptr(ro),
                               _ptr(cg), _ptr(co), C.c_int64(row_range[0]), C.c_int64(row_range[1]), C.c_int64(global_rows),
                               C.c_int64(col_range[0]), C.c_int64(col_range[1]), C.c_int64(global_cols), C.byref(h)))
    return OwnedParCSR(h)


def parcsr_rows_to_global(d):
    """rows of a (host or downloaded) ParCSR piece as a scipy CSR block in GLOBAL column numbering"""
    import scipy.sparse as sp
    n = d["nrows"]
    rows_d = np.repeat(np.arange(n), np.diff(d["diag_i"]))
    rows_o = np.repeat(np.arange(n), np.diff(d["offd_i"]))
    cols = np.concatenate([d["diag_j"].astype(np.int64) + d["first_col"],
                           d["col_map_offd"][d["offd_j"]] if len(d["offd_j"]) else np.empty(0, dtype=np.int64)])
    vals = np.concatenate([d["diag_a"], d["offd_a"]])
    rows = np.concatenate([rows_d, rows_o])
    # keep explicit zeros: build the pattern and the values separately
    M = sp.coo_matrix((np.ones(len(vals)), (rows, cols)), shape=(n, d["global_cols"])).tocsr()
    Vv = sp.coo_matrix((vals, (rows, cols)), shape=(n, d["global_cols"])).tocsr()
    return M, Vv
