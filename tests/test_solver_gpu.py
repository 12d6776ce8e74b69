"""End-to-end parity of the ParameterList-driven solver (C++ SolverLibrary -> AMGe
Hierarchy -> Hiptmair / hypre smoothers -> PCG, all on the GPU through the C ABI)
against the CPU oracle on the same mesh and coefficients.

Contract (BASELINE.json north_star): coarse-operator patterns bit-exact, values to
1e-12 relative; per-iteration PCG "(B r, r)" to 1e-9 relative with the same iteration
count +-1."""
import numpy as np
import pytest

from parelag_b200 import api
from oracle import amge, drivers, solve as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sess():
    ctx = api.session()
    yield ctx


@pytest.fixture(scope="module")
def hier():
    mesh, seqs = amge.build_hierarchy((8, 8, 8), 3)
    return mesh, seqs


def make_sequence(seqs):
    S = api.Sequence(4, len(seqs))
    for l, s in enumerate(seqs):
        for j in range(4):
            S.set_bdr_mask(l, j, drivers.bdr_mask(s.dof[j]))
            if j < 3:
                S.set_D(l, j, s.D[j])
            if l + 1 < len(seqs):
                S.set_P(l, j, s.P[j])
    return S


@pytest.fixture
def gs_kernel(request):
    from parelag_b200 import capi
    old = capi.get_tuning(capi.TUNE_SELL_MIN_ROWS)
    capi.set_tuning(capi.TUNE_SELL_MIN_ROWS, 0 if request.param == "sell" else 1 << 30)
    yield request.param
    capi.set_tuning(capi.TUNE_SELL_MIN_ROWS, old)


@pytest.mark.parametrize("form,ordering,gs_kernel", [(0, "natural", "csr"), (2, "natural", "csr"), (1, "natural", "csr"),
                                                     (2, "multicolor", "sell"), (0, "multicolor", "sell"),
                                                     (2, "multicolor", "csr"), (1, "multicolor", "csr")], indirect=["gs_kernel"])
def test_pcg_amge_residual_history(sess, hier, form, ordering, gs_kernel):
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], form, ess)
    rng = np.random.default_rng(11 + form)
    b = rng.standard_normal(A.shape[0])
    b[marker] = 0.0
    # oracle
    H = drivers.amge_pcg_solver(seqs, form, ess, A, ordering=ordering)
    xo, ito, convo, histo = orc.pcg(A, H.mult, b, rtol=1e-6, atol=1e-6, max_iter=300)
    # product
    xml = api.library_xml(drivers.library_entries(form, ordering=ordering))
    S = make_sequence(seqs)
    solver = api.Solver(xml, "PCG-AMGe", A, S, 0, form, ess)
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and convo
    assert abs(it - ito) <= 1
    m = min(len(hist), len(histo))
    assert m >= 3
    rel = np.abs(hist[:m] - np.array(histo[:m])) / np.abs(np.array(histo[:m]))
    assert rel.max() < 1e-9, rel
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)
    solver.free(); S.free()


@pytest.mark.parametrize("form", [0, 1, 2])
def test_galerkin_hierarchy_matches_oracle(sess, hier, form):
    """A_{l+1} = P^T A P + FixZeroRows: pattern bit-exact, values 1e-12 (relative to max)."""
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], form, ess)
    xml = api.library_xml(drivers.library_entries(form))
    S = make_sequence(seqs)
    solver = api.Solver(xml, "AMGe", A, S, 0, form, ess)
    assert solver.num_levels() == 3
    Ao = A
    for l in range(1, 3):
        Ao, _ = orc.fix_zero_rows(orc.rap(Ao, seqs[l - 1].get_P(form, ess)))
        Ag = solver.level_matrix(l)
        assert np.array_equal(Ag.indptr, Ao.indptr)
        assert np.array_equal(Ag.indices, Ao.indices)
        assert np.max(np.abs(Ag.data - Ao.data)) <= 1e-12 * np.max(np.abs(Ao.data))
    # one V-cycle applied to a vector
    rng = np.random.default_rng(5)
    r = rng.standard_normal(A.shape[0]); r[marker] = 0
    H = drivers.amge_pcg_solver(seqs, form, ess, A)
    zo = H.mult(r)
    z = solver.mult(r)
    assert np.linalg.norm(z - zo) <= 1e-11 * np.linalg.norm(zo)
    solver.free(); S.free()


def test_api_error_behaviour(sess, hier):
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    A, _ = drivers.system_matrix(seqs[0], 0, ess)
    xml = api.library_xml(drivers.library_entries(0))
    from parelag_b200.capi import PEError
    with pytest.raises(PEError, match="not in the library"):
        api.Solver(xml, "No such solver", A, None, 0, 0, ess)
    bad = xml.replace("L1 Gauss-Seidel", "Kaczmarz")
    S = make_sequence(seqs)
    with pytest.raises(PEError, match="not supported on the GPU path"):
        api.Solver(bad, "PCG-AMGe", A, S, 0, 0, ess)
    with pytest.raises(PEError, match="no DeRhamSequence"):
        api.Solver(xml, "AMGe", A, None, 0, 0, ess)


@pytest.mark.parametrize("gs_kernel", ["sell"], indirect=True)
def test_vcycle_persistent_program_matches_direct(sess, hier, gs_kernel):
    """Hierarchy::Mult recorded into the persistent program kernel (pe_prog.cu) must reproduce the
    V-cycle executed kernel by kernel: same arithmetic per row, only the dot-product reduction
    tree differs (1e-13), and it must actually be the path that runs."""
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    form = 2
    A, marker = drivers.system_matrix(seqs[0], form, ess)
    S = make_sequence(seqs)
    rng = np.random.default_rng(3)
    r = rng.standard_normal(A.shape[0]); r[marker] = 0
    outs = {}
    for use in (True, False):
        entries = drivers.library_entries(form, ordering="multicolor")
        entries["AMGe"][1]["Use persistent program"] = use
        entries["AMGe"][1]["Use CUDA graph"] = use
        solver = api.Solver(api.library_xml(entries), "PCG-AMGe", A, S, 0, form, ess)
        from parelag_b200 import capi
        b, z = capi.Vec(sess, data=r), capi.Vec(sess, len(r))
        for _ in range(4):           # 1: warm, 2: record + first launch, 3+: replay
            solver.prec_mult_device(b, z)
        outs[use] = z.download()
        prog = solver.program()
        if use:
            assert prog is not None and prog[0] > 50 and prog[1] > 0
            t, us, by = solver.program_profile(sess)
            assert len(t) == prog[0] and np.all(us >= 0) and abs(by.sum() - prog[1]) < 1e-6 * prog[1]
            z2 = z.download()
            assert np.array_equal(z2, outs[use])        # replays are bit-reproducible
        else:
            assert prog is None
        solver.free()
    H = drivers.amge_pcg_solver(seqs, form, ess, A, ordering="multicolor")
    zo = H.mult(r)
    assert np.linalg.norm(outs[True] - outs[False]) <= 1e-13 * np.linalg.norm(zo)
    assert np.linalg.norm(outs[True] - zo) <= 1e-11 * np.linalg.norm(zo)
    S.free()


@pytest.mark.parametrize("name", ["GMRES", "FGMRES", "BICGSTAB", "MINRES"])
def test_other_krylov_solvers(sess, hier, name):
    """KrylovSolver "Solver name" = GMRES / FGMRES / BICGSTAB / MINRES (mfem::*Solver::Mult restated) with an
    l1-Jacobi preconditioner (SPD, order independent) on the H1 system: monitored norm history vs the oracle."""
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], 0, ess)
    rng = np.random.default_rng(17)
    b = rng.standard_normal(A.shape[0]); b[marker] = 0.0
    So = orc.Smoother(A, type=1)
    prec = lambda r: So.apply(r, np.zeros_like(r), False)
    fn = {"GMRES": orc.gmres, "FGMRES": orc.fgmres, "BICGSTAB": orc.bicgstab, "MINRES": orc.minres}[name]
    kw = dict(rtol=1e-8, atol=1e-12, max_iter=200)
    if name in ("GMRES", "FGMRES"):
        kw["restart"] = 20
    xo, ito, convo, histo = fn(A, prec, b, **kw)
    lib = {"L1J": ("Hypre", {"Type": "L1 Jacobi", "Sweeps": 1}),
           "K": ("Krylov", {"Solver name": name, "Preconditioner": "L1J", "Print level": -1, "Maximum iterations": 200,
                            "Relative tolerance": 1e-8, "Absolute tolerance": 1e-12, "Restart size": 20})}
    solver = api.Solver(api.library_xml(lib), "K", A, None, 0, 0, ess)
    x = solver.mult(b)
    hist, it, conv = solver.history()
    assert conv and convo and abs(it - ito) <= 1, (name, it, ito, conv, convo)
    m = min(len(hist), len(histo))
    ho = np.array(histo[:m])
    sel = ho > 1e-6 * ho[0]
    assert (np.abs(hist[:m] - ho)[sel] / ho[sel]).max() < 1e-8
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    solver.free()


def test_stationary_solver_matches_reference_iteration(sess, hier):
    """"Stationary Iteration" factory (ParELAG_StationarySolverFactory.hpp:27-107) around an l1-Jacobi corrector:
    ||r_k|| history, iteration count and stopping rule of ParELAG_StationarySolver.cpp:41-147 (updated residual, at
    least one correction, accumulated-ratio test) against the oracle's restatement."""
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    A, marker = drivers.system_matrix(seqs[0], 0, ess)
    rng = np.random.default_rng(23)
    b = rng.standard_normal(A.shape[0]); b[marker] = 0.0
    So = orc.Smoother(A, type=1)
    corr = lambda r: So.apply(r, np.zeros_like(r), False)
    for rtol, maxit in ((0.0, 7), (0.3, 200)):
        xo, ito, convo, histo = orc.stationary(A, corr, b, rtol=rtol, atol=0.0, max_iter=maxit)
        lib = {"L1J": ("Hypre", {"Type": "L1 Jacobi", "Sweeps": 1}),
               "SI": ("Stationary Iteration", {"Solver": "L1J", "Maximum Iterations": maxit, "Relative Tolerance": rtol,
                                               "Absolute Tolerance": 0.0, "Print Iterations": False})}
        solver = api.Solver(api.library_xml(lib), "SI", A, None, 0, 0, ess)
        x = solver.mult(b)
        hist, it, conv = solver.history()
        assert it == ito and conv == convo and len(hist) == len(histo), (it, ito, conv, convo)
        assert np.max(np.abs(hist - np.array(histo)) / np.array(histo)) < 1e-12
        assert np.linalg.norm(x - xo) <= 1e-12 * np.linalg.norm(xo)
        # solver mode (initial guess given): the residual of the guess starts the iteration
        x1 = solver.mult(b, x0=xo)
        xo1 = orc.stationary(A, corr, b, rtol=rtol, atol=0.0, max_iter=maxit, x0=xo)[0]
        assert np.linalg.norm(x1 - xo1) <= 1e-12 * np.linalg.norm(xo1)
        solver.free()
    # the default is ONE iteration (factory default "Maximum Iterations" = 1)
    lib = {"L1J": ("Hypre", {"Type": "L1 Jacobi", "Sweeps": 1}), "SI": ("Stationary Iteration", {"Solver": "L1J"})}
    solver = api.Solver(api.library_xml(lib), "SI", A, None, 0, 0, ess)
    x = solver.mult(b)
    assert solver.history()[1] == 1 and np.linalg.norm(x - corr(b)) <= 1e-13 * np.linalg.norm(x)
    solver.free()


def test_hiptmair_mult_transpose(sess, hier):
    """HiptmairSmoother::MultTranspose (ParELAG_HiptmairSmoother.cpp:79-109): auxiliary-space correction first, primary
    sweep last; preconditioner mode restricts B itself, solver mode the residual."""
    mesh, seqs = hier
    ess = np.ones(6, dtype=np.int32)
    for form in (1, 2):
        A, marker = drivers.system_matrix(seqs[0], form, ess)
        D = seqs[0].get_D(form - 1, ess)
        kw = lambda M: dict(type=1)
        Ho = orc.Hiptmair(A, D, kw, kw)
        lib = {"L1J": ("Hypre", {"Type": "L1 Jacobi", "Sweeps": 1}),
               "HIP": ("Hiptmair", {"Primary Smoother": "L1J", "Auxiliary Smoother": "L1J"})}
        S = make_sequence(seqs[:1])
        solver = api.Solver(api.library_xml(lib), "HIP", A, S, 0, form, ess)
        rng = np.random.default_rng(31 + form)
        b = rng.standard_normal(A.shape[0]); b[marker] = 0.0
        x0 = rng.standard_normal(A.shape[0]); x0[marker] = 0.0
        xt = solver.mult_transpose(b)
        xo = Ho.apply_transpose(b, np.zeros_like(b), False)
        assert np.linalg.norm(xt - xo) <= 1e-12 * np.linalg.norm(xo), form
        xt = solver.mult_transpose(b, x0=x0)
        xo = Ho.apply_transpose(b, x0, True)
        assert np.linalg.norm(xt - xo) <= 1e-12 * np.linalg.norm(xo), form
        # and it is the transpose sequence, not Mult
        xm = solver.mult(b, x0=x0)
        assert np.linalg.norm(xm - Ho.apply(b, x0, True)) <= 1e-12 * np.linalg.norm(xm)
        assert np.linalg.norm(xm - xt) > 1e-6 * np.linalg.norm(xm)
        solver.free(); S.free()
