"""examples/multigrid_test.py (the flow of the reference's MultigridTest{0,1,2}Form drivers on a parameter list in the
reference's XML layout): the parts that need no GPU -- parsing, validation of the solver list, the plan."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "examples", "multigrid_test.py")


@pytest.mark.parametrize("form", [0, 1, 2])
def test_dry_run_of_the_shipped_parameter_lists(form):
    r = subprocess.run([sys.executable, DRIVER, "--form", str(form), "--dry-run"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dry run: form %d, levels 0..0, solvers ['PCG-AMGe']: ok" % form in r.stdout
    assert "Fine mesh size: 4096 hexahedra, 3 levels" in r.stdout          # 2 x 2 x 2, one serial + two parallel refinements


def test_a_list_that_needs_a_hypre_black_box_is_refused_with_the_reason(tmp_path):
    r = subprocess.run([sys.executable, DRIVER, "--form", "2", "--print-parameters"], capture_output=True, text=True, timeout=300)
    xml = r.stdout
    assert '<Parameter name="Coarse solver" type="string" value="PCG-GS"/>' in xml
    xml = xml.replace('<Parameter name="Coarse solver" type="string" value="PCG-GS"/>',
                      '<Parameter name="Coarse solver" type="string" value="ADS Solver"/>')
    xml = xml.replace('<ParameterList name="Preconditioner Library">',
                      '<ParameterList name="Preconditioner Library">\n    <ParameterList name="ADS Solver">'
                      '<Parameter name="Type" type="string" value="ADS"/></ParameterList>')
    path = tmp_path / "needs_ads.xml"
    path.write_text(xml)
    r = subprocess.run([sys.executable, DRIVER, "--form", "2", "-f", str(path), "--dry-run"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and 'unknown factory type "ADS"' in (r.stdout + r.stderr)
