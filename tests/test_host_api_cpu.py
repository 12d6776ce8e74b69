"""The C++ mirror of ParElag's solver-side classes used the way a C++ driver of the reference uses them (no device): builds
and runs tests/cpp/host_api_test.cpp against the in-tree headers and shared library."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_api(tmp_path):
    exe = str(tmp_path / "host_api_test")
    lib = os.path.join(ROOT, "parelag_b200", "lib")
    cmd = ["g++", "-O1", "-std=c++17", "-fopenmp", "-Wall", "-Wno-misleading-indentation", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "parelag_b200", "src"), os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-o", exe,
           "-L" + lib, "-lparelag_b200", "-Wl,-rpath," + lib]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_API_TEST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
