"""Build libnalu_b200_host.so: the C++ mirror of the reference's Kernel / ElemDataRequests /
AssembleElemSolverAlgorithm / LinearSystem API on top of the C ABI (links libnalu_b200.so)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libnalu_b200_host.so")
SRC = [os.path.join(HERE, "src", "host.cpp")]
CXX = os.environ.get("CXX", "g++")
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-parameter"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root, _, files in os.walk(HERE):
        for f in files:
            if f.endswith((".h", ".cpp", ".py")) and os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return os.path.getmtime(os.path.join(PKG, "libnalu_b200.so")) > t if os.path.exists(os.path.join(PKG, "libnalu_b200.so")) else True


def build(force=False):
    from nalu_b200 import build as core
    core.build()
    if not (force or _stale()):
        return OUT
    cmd = [CXX] + FLAGS + ["-shared", "-o", OUT] + SRC + ["-L" + PKG, "-lnalu_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("host library build failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
