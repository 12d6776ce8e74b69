"""Diagnostic: cost per launch of the small-level GS sweeps, stream vs CUDA graph (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parelag_b200 import capi
from tests.util import laplace3d
ctx = capi.Ctx()
for m in (12, 26, 52, 104):
    A = laplace3d(m, m, m)
    n = A.shape[0]
    dA = capi.Mat.from_scipy(ctx, A)
    b = capi.Vec(ctx, data=np.ones(n)); x = capi.Vec(ctx, n)
    sg = capi.Smoother(ctx, dA, type=2, ordering=capi.GS_MULTICOLOR)
    l0 = ctx.launch_count(); sg.apply(b, x, True); nl = ctx.launch_count() - l0
    for _ in range(5): sg.apply(b, x, True)
    ctx.sync(); ctx.timer_start()
    for _ in range(50): sg.apply(b, x, True)
    ms = ctx.timer_stop() / 50
    ctx.graph_begin()
    for _ in range(10): sg.apply(b, x, True)
    g = ctx.graph_end()
    for _ in range(3): ctx.graph_launch(g)
    ctx.sync(); ctx.timer_start()
    for _ in range(5): ctx.graph_launch(g)
    msg = ctx.timer_stop() / 50
    print("n=%8d: %d launches/apply, stream %.1f us/launch, graph %.1f us/launch" % (n, nl, 1e3 * ms / nl, 1e3 * msg / nl), flush=True)
