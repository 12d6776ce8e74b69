#!/usr/bin/env python
"""The flow of the reference's drivers examples/MultigridTest{0,1,2}Form.cpp on this library, driven by a parameter list in
the reference's XML layout ("Problem parameters" / "Output control" / "Preconditioner Library"):

    python examples/multigrid_test.py --form 2 [-f my_parameters.xml] [--dry-run] [--print-parameters]

mesh ("TestingMesh": the cube of 2 x 2 x 2 hexahedra, examples/testing_helpers/Build3DHexMesh.hpp) -> serial + parallel
refinements -> AgglomeratedTopology by derefinement -> DeRhamSequence, Coarsen() once per parallel refinement -> for every
level between "Start level" and "Stop level": assemble A = [M_form +] D^T M_{form+1} D with all boundary attributes essential,
and for every entry of "List of linear solvers": BuildSolver, Mult, "Initial / Final residual norm" as the drivers print them.

An example of use, not a parity test: the right-hand side is the load of the constant field (the mass matrix applied to the
interpolant of 1 resp. (1, 1, 1)) instead of the drivers' manufactured solution.  --dry-run parses and validates the list
(every solver of the list must resolve to a factory) without touching a GPU."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parelag_b200 import api  # noqa: E402


def build_test_parameters(form, solvers=("PCG-AMGe",)):
    """The default list, built in code like the reference's examples/testing_helpers/Create{0,1,2}FormParameterList.hpp
    ("BuildTestParameters"): the cube refined once in serial and twice in parallel, PCG preconditioned by the AMGe
    V-cycle with (Hiptmair) l1-Gauss-Seidel smoothers and the PCG-GS coarse solver of spe10_example_parameters.xml."""
    def plist(name, body):
        return ['<ParameterList name="%s">' % name] + ["  " + l for l in body] + ["</ParameterList>"]

    def par(name, value):
        if isinstance(value, bool):
            t, v = "bool", "true" if value else "false"
        elif isinstance(value, int):
            t, v = "int", str(value)
        elif isinstance(value, float):
            t, v = "double", repr(value)
        elif isinstance(value, (list, tuple)):
            t, v = ("vector(int)", " ".join(str(x) for x in value)) if value and isinstance(value[0], int) else ("list(string)", ", ".join(value))
        else:
            t, v = "string", value
        return '<Parameter name="%s" type="%s" value="%s"/>' % (name, t, v)

    def solver(name, typ, params):
        return plist(name, [par("Type", typ)] + plist("Solver Parameters", [par(k, v) for k, v in params.items()]))
    smoother = "Hiptmair-GS-GS" if form > 0 else "Gauss-Seidel"
    lib = solver("PCG-AMGe", "Krylov", {"Solver name": "PCG", "Preconditioner": "AMGe", "Print level": 1, "Maximum iterations": 300,
                                        "Relative tolerance": 1e-6, "Absolute tolerance": 1e-6})
    lib += solver("AMGe", "AMGe", {"Maximum levels": -1, "Forms": [form], "PreSmoother": smoother, "PostSmoother": smoother,
                                   "Coarse solver": "PCG-GS", "Cycle type": "V-cycle"})
    lib += solver("PCG-GS", "Krylov", {"Solver name": "PCG", "Preconditioner": "Gauss-Seidel", "Print level": -1, "Maximum iterations": 3,
                                       "Relative tolerance": 1e-4, "Absolute tolerance": 1e-4})
    if form > 0:
        lib += solver("Hiptmair-GS-GS", "Hiptmair", {"Primary Smoother": "Gauss-Seidel", "Auxiliary Smoother": "Gauss-Seidel"})
    lib += solver("Gauss-Seidel", "Hypre", {"Type": "L1 Gauss-Seidel", "Sweeps": 1, "GS ordering": "multicolor"})
    doc = plist("Default",
                plist("Problem parameters", [par("Mesh file", "TestingMesh"), par("Serial refinement levels", 1), par("Parallel refinement levels", 2),
                                             par("Finite element order", 0), par("Upscaling order", 0), par("Start level", 0), par("Stop level", 0),
                                             par("List of linear solvers", list(solvers))]) +
                plist("Output control", [par("Visualize solution", False), par("Print timings", True), par("Show progress", True)]) +
                plist("Preconditioner Library", lib))
    return "\n".join(doc) + "\n"


def read_parameters(xml):
    """{'Problem parameters/Mesh file': 'TestingMesh', ...} with int / float / bool / list values"""
    out = {}
    for line in api.parameterlist_dump(xml):
        key, typ, val = line.split("\t", 2)
        if typ == "bool":
            out[key] = val == "true"
        elif typ in ("int", "long", "long long", "unsigned long", "unsigned long long"):
            out[key] = int(val)
        elif typ in ("double", "float", "long double"):
            out[key] = float(val)
        elif typ == "list(string)":
            out[key] = [v for v in val.split(",") if v]
        elif typ == "vector(int)":
            out[key] = [int(v) for v in val.split()]
        else:
            out[key] = val
    return out


def library_document(xml):
    """the <ParameterList name="Preconditioner Library"> element of the master list as a document of its own"""
    start = xml.index('<ParameterList name="Preconditioner Library">')
    depth, pos = 0, start
    while True:
        nxt_open, nxt_close = xml.find("<ParameterList", pos + 1), xml.find("</ParameterList>", pos + 1)
        if nxt_open != -1 and nxt_open < nxt_close:
            if not xml[nxt_open:xml.index(">", nxt_open) + 1].endswith("/>"):
                depth += 1
            pos = nxt_open
        else:
            if depth == 0:
                return xml[start:nxt_close + len("</ParameterList>")]
            depth -= 1
            pos = nxt_close


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--form", type=int, required=True, choices=[0, 1, 2], help="0: H1 (MultigridTest0Form), 1: H(curl), 2: H(div)")
    ap.add_argument("-f", "--xml-file", default="BuildTestParameters", help="XML parameter list (default: the list build_test_parameters() builds)")
    ap.add_argument("--print-parameters", action="store_true", help="print the parameter list in use and exit")
    ap.add_argument("--dry-run", action="store_true", help="parse and validate the parameter list, print the plan, no GPU")
    args = ap.parse_args()
    xml = build_test_parameters(args.form) if args.xml_file == "BuildTestParameters" else open(args.xml_file).read()
    if args.print_parameters:
        print(xml, end="")
        return
    p = read_parameters(xml)
    prob = lambda k, d: p.get("Problem parameters/" + k, d)
    meshfile = prob("Mesh file", "TestingMesh")
    ser, par = prob("Serial refinement levels", -1), prob("Parallel refinement levels", 2)
    if ser < 0:
        ser = 0                    # "refine until there are 6 elements per rank": the 8 elements of the cube are enough for one rank
    start, stop = prob("Start level", 0), prob("Stop level", 0)
    if stop < 0:
        stop = par
    solvers = prob("List of linear solvers", [])
    show, timings = p.get("Output control/Show progress", True), p.get("Output control/Print timings", True)
    if prob("Finite element order", 0) != 0 or prob("Upscaling order", 0) != 0:
        raise SystemExit("only lowest-order elements and upscaling order 0 are built")
    if meshfile != "TestingMesh":
        raise SystemExit('this example builds the "TestingMesh" cube; tetrahedral mesh files go through api.Sequence.tet_from_file')
    lib_xml = library_document(xml)
    report = {n: (t, s) for n, t, s in api.library_factories(lib_xml)}
    for name in solvers:
        if name not in report:
            raise SystemExit('solver "%s" of "List of linear solvers" is not in the Preconditioner Library' % name)
        if report[name][1] != "ok":
            raise SystemExit('solver "%s" cannot be built on the GPU path: %s' % (name, report[name][1]))
    n = 2 * 2 ** (ser + par)
    levels = par + 1
    print("\n" + "*" * 50 + "\n*  Mesh: %s\n*\n*              FE order: 0\n*       Upscaling order: 0\n*\n"
          "*    Serial refinements: %d\n*  Parallel refinements: %d\n*       Fine mesh size: %d hexahedra, %d levels\n"
          % (meshfile, ser, par, n ** 3, levels) + "*" * 50)
    if args.dry_run:
        print("dry run: form %d, levels %d..%d, solvers %s: ok" % (args.form, start, stop, solvers))
        return
    ctx = api.session()
    if show:
        print("-- Building the topology, the fine DeRhamSequence and coarsening to all levels...")
    S = api.Sequence.hex((n, n, n), levels, jstart=0 if args.form < 2 else 1)
    ess = np.ones(6, dtype=np.int32)
    for level in range(start, stop + 1):
        if show:
            print("-- Assembling the linear system on level %d..." % level)
        A = S.assemble_system(ctx, level, args.form, ess)
        As = A.to_scipy()
        M = S.get_csr(level, "M", args.form)
        T = S.get_targets(level, args.form)
        field = T[:, 0] if args.form == 0 else T.sum(axis=1)
        b = M @ field
        b[S.get_bdr_mask(level, args.form) != 0] = 0.0
        for name in solvers:
            print("\n" + "*" * 50 + "\n*    Solving on level: %d\n*              A size: %dx%d\n*               A NNZ: %d\n*\n*              Solver: %s\n"
                  % (level, As.shape[0], As.shape[1], As.nnz, name) + "*" * 50 + "\n")
            solver = api.Solver(lib_xml, name, As, S, level, args.form, ess)
            print("Initial residual norm: %g" % np.linalg.norm(b))
            x = solver.mult(b)
            hist, its, conv = solver.history()
            print("Final residual norm: %g   (%d iterations, converged: %s)" % (np.linalg.norm(b - As @ x), its, conv))
            solver.free()
    if timings:
        for key in ("Mesh Agglomeration -- Level 1", "DeRhamSequence Construction -- Level 0", "DeRhamSequence Construction -- Level 1",
                    "Assemble linear system"):
            print("%-45s %.3f s" % (key, api.timer(key)))
    S.free()
    if show:
        print("-- Goodbye!\n")


if __name__ == "__main__":
    main()
