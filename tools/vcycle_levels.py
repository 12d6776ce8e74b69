#!/usr/bin/env python
"""Diagnostic: V-cycle time as a function of hierarchy depth (per-level cost), and the
fine-level smoother alone.  GPU only.  usage: vcycle_levels.py [n] [levels]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parelag_b200 import api, capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = api.session()
S = api.Sequence.hex((n, n, n), levels, jstart=1)
ESS = bench.ESS
rng = np.random.default_rng(0)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    t_host = time.perf_counter() - t0
    ms = ctx.timer_stop()
    return ms / reps, 1e3 * t_host / reps


for maxlev in range(1, levels + 1):
    lib = bench.library("multicolor")
    lib["AMGe-HIP-GS_2"][1]["Maximum levels"] = maxlev
    A = S.assemble_system(ctx, 0, 2, ESS)
    nd = A.info()[0]
    name = "AMGe-HIP-GS_2" if maxlev > 1 else "Hiptmair-GS-GS"
    solver = api.Solver(api.library_xml(lib), name, A, S, 0, 2, ESS)
    r, z = capi.Vec(ctx, data=rng.standard_normal(nd)), capi.Vec(ctx, nd)
    l0 = ctx.launch_count()
    solver.mult_device(r, z, False)
    nl = ctx.launch_count() - l0
    ms, host = timeit(lambda: solver.mult_device(r, z, False))
    print("max levels %d: %8.3f ms GPU  (host enqueue %8.3f ms)  %d launches" % (maxlev, ms, host, nl), flush=True)
    solver.free()
