"""Build recipe for libnalu_b200.so: nvcc for sm_100a only, in-tree, no JIT cache.

    python -m nalu_b200.build [--force]

The library is compiled here (nvcc cross-compiles without a GPU) and travels with the
repo snapshot to the GPU box; nothing is installed into site-packages.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnalu_b200.so")
CU = ["capi.cu", "assemble_general.cu", "assemble_box.cu", "assemble_box2.cu", "assemble_box5.cu", "assemble_cont_box.cu", "assemble_scalar_box.cu", "assemble_slist_box.cu", "face_elem.cu", "assemble_urow.cu", "assemble_hex27.cu", "halo.cu"]
CPP = ["graph.cpp", "comm.cpp"]
HDR = ["internal.h", "graph.h", "general_elem.cuh", "hex8_dev.cuh", "box_tile.cuh", "box2_tile.cuh", "box5_tile.cuh", "cont_tile.cuh", "scalar_tile.cuh", "slist_tile.cuh", "../../include/nalu_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fopenmp,-O3",
         "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in CU + CPP + HDR + ["../build.py"]:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not (force or _stale()):
        return OUT
    objs = []
    log = []
    for f in CU + CPP:
        src = os.path.join(CSRC, f)
        obj = os.path.join(CSRC, os.path.splitext(f)[0] + ".o")
        cmd = [NVCC] + ARCH + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError("nvcc failed on " + f)
        objs.append(obj)
    cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xcompiler", "-fopenmp", "-lcudart", "-lgomp", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", OUT)
