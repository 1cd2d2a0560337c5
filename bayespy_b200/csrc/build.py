"""Build libbpk.so (sm_100a only) in-tree with nvcc.  No torch, no JIT cache.

Used by ``__graft_entry__.build()`` and ``python -m bayespy_b200.csrc.build``.
The shared object lands next to the package (``bayespy_b200/libbpk.so``) so
that it travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libbpk.so")
SOURCES = ["runtime.cu", "ewise.cu", "reduce.cu", "linalg.cu", "nodes.cu", "pca.cu", "pca_vb.cu", "pca_masked.cu", "gmm_vb.cu", "gmm.cu", "gmc.cu"]
HEADERS = ["common.cuh", "spd.cuh", "spd16.cuh", "warp_spd.cuh", "gmc_bcr3.cuh", "pca_common.cuh", "pca_vb_ops.cuh", os.path.join("..", "..", "include", "bpk.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default",
    "-prec-div=true", "-prec-sqrt=true", "-fmad=true",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libbpk cannot be built")
    return nvcc


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, ptxas_info=False):
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (s, out))
        elif verbose or ptxas_info:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libbpk build failed")
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
