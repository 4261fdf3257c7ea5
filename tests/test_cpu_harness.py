"""The CUDA kernels' per-thread phase functions are plain `__host__ __device__` code; these
harnesses compile them as ordinary C++ and execute the kernels thread by thread on the CPU
against the oracle, so the arithmetic and the tiling / flush logic are checked in the GPU-less
container (the `-m gpu` tests then check the real launches on the B200)."""
import os
import subprocess

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "tests", "cpu_harness")
B = os.path.join(H, "_build")


def _objs():
    os.makedirs(B, exist_ok=True)
    oracle.build()
    objs = []
    for f in ("hex8_me.c", "elem_kernels.c", "linsys.c"):
        o = os.path.join(B, f[:-2] + ".o")
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", f), "-o", o])
        objs.append(o)
    o = os.path.join(B, "graph.o")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-c",
                           os.path.join(ROOT, "nalu_b200", "csrc", "graph.cpp"), "-o", o])
    objs.append(o)
    return objs


@pytest.mark.parametrize("name", ["hex8_host_check", "general_host_check", "box2_host_check", "box5_host_check", "cont_host_check", "scalar_host_check",
                                  "slist_host_check"])
def test_kernel_phases_on_cpu(name):
    exe = os.path.join(B, name)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fopenmp", os.path.join(H, name + ".cpp")] + _objs() +
                          ["-o", exe, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
