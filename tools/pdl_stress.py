#!/usr/bin/env python
"""Determinism stress of the scheduling knobs on the headline V-cycle (GPU box): every V-cycle output of every
(group, PDL) setting must be bit-identical to the first one.  usage: pdl_stress.py [n] [levels] [reps]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parelag_b200 import api, capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = api.session()
S = api.Sequence.hex((n, n, n), levels, jstart=1)
ESS = bench.ESS
ref = None
for group, pdl in [(0, 0), (0, 1), (8, 0), (8, 1), (4, 0), (4, 1), (12, 1)]:
    capi.set_tuning(capi.TUNE_SELL_GROUP, group)
    capi.set_tuning(capi.TUNE_PDL, pdl)
    A = S.assemble_system(ctx, 0, 2, ESS)
    nd = A.info()[0]
    solver = api.Solver(api.library_xml(bench.library("multicolor")), "PCG with Auxiliary Space Preconditioner",
                        A, S, 0, 2, ESS)
    r, z = capi.Vec(ctx, data=np.random.default_rng(0).standard_normal(nd)), capi.Vec(ctx, nd)
    bad, worst = 0, 0.0
    for it in range(reps):
        solver.prec_mult_device(r, z)
        zz = z.download()
        if ref is None:
            ref = zz
        if not np.array_equal(ref, zz):
            bad += 1
            worst = max(worst, float(np.abs(ref - zz).max() / np.abs(ref).max()))
    print(json.dumps({"group": group, "pdl": pdl, "vcycles": reps, "not_bit_identical": bad, "worst_rel_diff": worst}), flush=True)
    solver.free()
