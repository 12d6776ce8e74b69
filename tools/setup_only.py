#!/usr/bin/env python
"""Setup stages only (sequence, system assembly, BuildSolver) with the reference's timer names.  GPU box.
usage: setup_only.py [n] [levels]"""
import json, os, sys, time, ctypes, resource
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parelag_b200 import api, capi
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = api.session()
api.lib().pe_api_timer_clear()
t0 = time.perf_counter(); S = api.Sequence.hex((n, n, n), levels, jstart=0); ctx.sync(); t1 = time.perf_counter()
A = S.assemble_system(ctx, 0, 2, bench.ESS); ctx.sync(); t2 = time.perf_counter()
solver = api.Solver(api.library_xml(bench.library("multicolor")), "PCG with Auxiliary Space Preconditioner", A, S, 0, 2, bench.ESS)
ctx.sync(); t3 = time.perf_counter()
st6 = (ctypes.c_double * 6)(); capi.lib().pe_local_stage_seconds(st6, 1)
names = ["Host arena reserve (parallel first touch)", "Mesh Agglomeration -- Level 0", "Mesh Agglomeration -- Level 1",
         "DeRhamSequence Construction -- Level 0", "DeRhamSequence Construction -- Level 1", "DeRhamSequence Construction -- Level 2",
         "Coarsen: DofAgglomeration", "Coarsen: batched traces (H2D + kernels + D2H)", "Coarsen: extension prepare (host)",
         "Coarsen: batched extension (H2D + kernels + D2H)", "Coarsen: extension commit (host)", "Coarsen: finalize P and D",
         "Coarsen: project targets", "Build smoother: level 0", "Assemble linear system"]
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PE_")}, "sequence_s": t1 - t0, "assemble_s": t2 - t1,
                  "build_solver_s": t3 - t2, "total_s": t3 - t0, "ext_h2d_s": st6[0], "ext_kernel_s": st6[1], "ext_d2h_s": st6[2],
                  "ext_h2d_GB": st6[3] / 1e9, "rss_GB": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6,
                  "timers": {k: round(api.timer(k), 3) for k in names}}), flush=True)
