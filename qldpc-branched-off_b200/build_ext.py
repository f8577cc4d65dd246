"""Build libqldpc_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build()).

Each .cu is compiled to an object under build/obj (re-used while the source, the headers and the flags are
unchanged; objects are compiled in parallel), then linked into the shared library.
Environment: NVCC, QB_EXTRA_NVCC_FLAGS (e.g. -DQB_EDGE_PROFILE for instrumented builds), QB_BUILD_OUT (output path).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "minsum.cu", "minsum_edge.cu", "minsum_edge_h2.cu", "minsum_edge_cluster.cu", "edge_layout.cu", "osd.cu", "osd_free.cu", "sampler.cu"]
OUT = os.environ.get("QB_BUILD_OUT", os.path.join(HERE, "libqldpc_b200.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
OBJDIR = os.path.join(HERE, "..", "build", "obj")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "qldpc_b200.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "qldpc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _object(src, extra, verbose):
    path = os.path.join(CSRC, src)
    h = hashlib.sha1()
    for p in [path] + _headers():
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS + extra).encode())
    obj = os.path.join(OBJDIR, f"{os.path.splitext(src)[0]}.{h.hexdigest()[:16]}.o")
    if verbose or not os.path.exists(obj):
        cmd = [NVCC] + FLAGS + extra + ["-c", path, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    extra = os.environ.get("QB_EXTRA_NVCC_FLAGS", "").split()
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _object(s, extra, verbose), SOURCES))
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs)
    keep = set(objs)
    for f in os.listdir(OBJDIR):     # drop objects of much older source revisions
        p = os.path.join(OBJDIR, f)
        if p not in keep and os.path.getmtime(p) < os.path.getmtime(OUT) - 900:
            os.remove(p)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
