#!/usr/bin/env python
"""Generates tests/golden/hdiv_amge_small.json from the CPU oracle (oracle/amge.py, oracle/solve.py), which is itself
pinned to the reference's published goldens (reference_upscaling_norms.json, tests/test_oracle_goldens.py).  The
reference cannot be built here, so these vectors are oracle outputs on fixed seeded inputs: they guard the
oracle against drift (tests/test_goldens_cpu.py) and give the GPU path a committed fixture to match
(tests/test_goldens_gpu.py).  Run from the repository root:  python tests/golden/make_goldens.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import amge, drivers, solve as orc   # noqa: E402


def case(dims, levels, form, seed):
    mesh, seqs = amge.build_hierarchy(dims, levels)
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], form, ess)
    b = np.random.default_rng(seed).standard_normal(A.shape[0])
    b[marker] = 0.0
    H = drivers.amge_pcg_solver(seqs, form, ess, A)
    x, it, conv, hist = orc.pcg(A, H.mult, b, rtol=1e-6, atol=1e-6, max_iter=100)
    ops = {}
    for l in range(levels - 1):
        for j in range(4):
            P = seqs[l].P[j].tocsr()
            ops["P_level%d_form%d" % (l, j)] = {"shape": list(P.shape), "nnz": int(P.nnz),
                                                "fro": float(np.sqrt((P.data ** 2).sum())), "abs_sum": float(np.abs(P.data).sum())}
    for l in range(1, levels):
        for j in range(3):
            D = seqs[l].D[j].tocsr()
            ops["D_level%d_form%d" % (l, j)] = {"shape": list(D.shape), "nnz": int(D.nnz),
                                                "fro": float(np.sqrt((D.data ** 2).sum())), "abs_sum": float(np.abs(D.data).sum())}
    return {"dims": list(dims), "levels": levels, "form": form, "rhs_seed": seed,
            "solver": "PCG(rtol=atol=1e-6) preconditioned by the AMGe V-cycle of oracle/drivers.py:library_entries(form)",
            "system": {"n": int(A.shape[0]), "nonzero_values": int(np.count_nonzero(A.data)), "fro": float(np.sqrt((A.data ** 2).sum()))},
            "pcg": {"iterations": int(it), "converged": bool(conv), "Br_r_history": [float(h) for h in hist],
                    "solution_l2": float(np.linalg.norm(x))},
            "operators": ops}


if __name__ == "__main__":
    out = {"_generator": "tests/golden/make_goldens.py", "_tolerances": "integers exact; norms 1e-12 relative; PCG history 1e-9 relative",
           "cases": [case((4, 4, 4), 2, 2, 0), case((4, 4, 4), 2, 1, 1), case((8, 4, 4), 3, 0, 2)]}
    with open(os.path.join(ROOT, "tests", "golden", "hdiv_amge_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written", len(out["cases"]), "cases")
