"""Worker of the multi-GPU parity test (one process per GPU, launched by tests/test_par_gpu.py through
torch.distributed.run): the box-decomposed AMGe hierarchy -- Coarsen() on every rank, SharingMaps,
ComputeTrueP / ComputeTrueD, distributed RAP, ParCSR SpMV / MatvecT with NCCL halo exchange, hybrid
Gauss-Seidel, PCG -- checked against the oracle's SINGLE-DOMAIN objects on the undecomposed mesh."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parelag_b200 import api, capi, par    # noqa: E402
from tests import parity_checks   # noqa: E402



def main():
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    ids = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    halo = os.environ.get("PE_TEST_HALO", "p2p")
    capi.set_tuning(capi.TUNE_P2P_HALO, 1 if halo == "p2p" else 0)
    ctx = api.session(rank, size, local, ids[0])
    comm = par.HostComm()
    api.set_host_comm(comm)
    # the path under test is the one that runs: NVLink peer-memory halo (CUDA IPC arenas mapped) or NCCL send/recv
    assert capi.lib().pe_ctx_p2p_enabled(ctx.h) == (1 if halo == "p2p" else 0), "halo path %s not active" % halo
    if os.environ.get("PE_TEST_CASE", "hdiv") == "darcy":
        rep = parity_checks.multi_rank_darcy(ctx, rank, size)
    else:
        rep = parity_checks.multi_rank(ctx, rank, size, deform=os.environ.get("PE_TEST_DEFORM", "0") == "1")
    dist.barrier()
    if rank == 0:
        print("PAR_GPU_WORKER_OK halo=%s %s" % (halo, rep))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
