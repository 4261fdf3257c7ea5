"""C++ host API (nalu_b200/host): the reference's Kernel / ElemDataRequests / AssembleElemSolverAlgorithm /
LinearSystem interface on top of the C ABI.  tests/host_cpp/unit_tests.cpp is written like the reference's
own gtest cases; this wrapper builds it, hands it the golden tables as text files and runs it.
  * cpu mode: registration + error conventions, and that building a system without a device fails loudly
  * gpu mode: the reference's golden kernel tests and UnitTestTpetra through the CUDA path"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nalu_b200")
OUT = os.path.join(ROOT, "tests", "host_cpp", "_build")


def _build():
    from nalu_b200.host import build as hb
    hb.build()
    os.makedirs(os.path.join(OUT, "gold"), exist_ok=True)
    exe = os.path.join(OUT, "unit_tests")
    src = os.path.join(ROOT, "tests", "host_cpp", "unit_tests.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hb.OUT)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(PKG, "host", "include"), src, "-L", PKG,
                               "-lnalu_b200_host", "-lnalu_b200", "-Wl,-rpath," + PKG, "-o", exe])
    gdir = os.path.join(ROOT, "tests", "golden")
    for f in os.listdir(gdir):
        if not f.endswith(".json"):
            continue
        with open(os.path.join(gdir, f)) as fh:
            tables = json.load(fh)["tables"]
        for name, tab in tables.items():
            with open(os.path.join(OUT, "gold", f"{f[:-5]}__{name}.txt"), "w") as o:
                o.write(f"{len(tab['dims'])} " + " ".join(map(str, tab["dims"])) + "\n")
                o.write(" ".join(repr(float(x)) for x in tab["values"]) + "\n")
    return exe


def _run(mode):
    exe = _build()
    r = subprocess.run([exe, os.path.join(OUT, "gold"), mode], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout
    return r.stdout


def test_host_api_registration_and_errors_cpu():
    out = _run("cpu")
    assert "no-device error (expected)" in out


@pytest.mark.gpu
def test_host_api_reference_unit_tests_gpu():
    out = _run("gpu")
    assert "Tpetra.basic" in out
