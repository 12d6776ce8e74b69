#!/usr/bin/env python
"""A/B of scheduling knobs on the headline V-cycle (GPU box): one sequence, one system matrix, a new
solver per knob setting (group size of the SELL kernels, programmatic dependent launch, SELL/CSR
crossover).  usage: ab_vcycle.py [n] [levels] [reps]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parelag_b200 import api, capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = api.session()
t0 = time.perf_counter()
S = api.Sequence.hex((n, n, n), levels, jstart=1)
ESS = bench.ESS
A = S.assemble_system(ctx, 0, 2, ESS)
nd = A.info()[0]
print("setup %.1f s, %d dofs" % (time.perf_counter() - t0, nd), flush=True)
rng = np.random.default_rng(0)
r, z = capi.Vec(ctx, data=rng.standard_normal(nd)), capi.Vec(ctx, nd)
ref = None
variants = [("baseline: group 4, no PDL", dict(group=4, pdl=0, keep=0)),
            ("auto group, PDL", dict(group=0, pdl=1, keep=0)),
            ("auto group, PDL, 50% of gather lines evict_last", dict(group=0, pdl=1, keep=50)),
            ("auto group, PDL, 75% evict_last", dict(group=0, pdl=1, keep=75)),
            ("auto group, PDL, 100% evict_last", dict(group=0, pdl=1, keep=100)),
            ("auto group, no PDL, 75% evict_last", dict(group=0, pdl=0, keep=75))]
for name, kv in variants:
    capi.set_tuning(capi.TUNE_SELL_GROUP, kv["group"])
    capi.set_tuning(capi.TUNE_PDL, kv["pdl"])
    capi.set_tuning(capi.TUNE_GATHER_KEEP_PCT, kv["keep"])
    A = S.assemble_system(ctx, 0, 2, ESS)       # the solver shares ownership of its operator
    solver = api.Solver(api.library_xml(bench.library("multicolor")), "PCG with Auxiliary Space Preconditioner",
                        A, S, 0, 2, ESS)
    for _ in range(3):
        solver.prec_mult_device(r, z)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        solver.prec_mult_device(r, z)
    ms = ctx.timer_stop() / reps
    zz = z.download()
    same = True if ref is None else bool(np.array_equal(ref, zz))
    if ref is None:
        ref = zz
    ctx.profile(True)
    for _ in range(5):
        solver.prec_mult_device(r, z)
    ctx.profile(False)
    prof = {i: ctx.profile_get(i) for i in range(4)}
    gbs = {i: (v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0) for i, v in prof.items()}
    print(json.dumps({"variant": name, "ms_per_vcycle": ms, "bit_identical_to_first": same,
                      "GBs_sell_spmv": gbs[0], "GBs_fine_gs": gbs[1], "GBs_coarse_gs": gbs[3],
                      "us_fine_gs": 1e3 * prof[1][1] / max(prof[1][0], 1), "us_coarse_gs": 1e3 * prof[3][1] / max(prof[3][0], 1)}),
          flush=True)
    solver.free()

# per-level cost with the default knobs: V-cycle time as a function of hierarchy depth
capi.set_tuning(capi.TUNE_SELL_GROUP, 0)
capi.set_tuning(capi.TUNE_PDL, 1)
capi.set_tuning(capi.TUNE_GATHER_KEEP_PCT, 0)
for maxlev in range(1, levels + 1):
    lib = bench.library("multicolor")
    lib["AMGe-HIP-GS_2"][1]["Maximum levels"] = maxlev
    name = "AMGe-HIP-GS_2" if maxlev > 1 else "Hiptmair-GS-GS"
    A = S.assemble_system(ctx, 0, 2, ESS)
    solver = api.Solver(api.library_xml(lib), name, A, S, 0, 2, ESS)
    for _ in range(3):
        solver.mult_device(r, z, False)
    ctx.sync()
    ctx.timer_start()
    for _ in range(10):
        solver.mult_device(r, z, False)
    print(json.dumps({"max_levels": maxlev, "ms": ctx.timer_stop() / 10}), flush=True)
    solver.free()
