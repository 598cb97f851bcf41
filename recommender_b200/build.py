"""Build librecsys_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m recommender_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librecsys_b200.so")
STAMP = os.path.join(LIB_DIR, "librecsys_b200.stamp")
SOURCES = ["api.cu", "gather.cu", "interaction.cu", "sparse_update.cu", "exchange.cu", "dense.cu", "mlp.cu", "p2p.cu", "criteo_input.cu", "din.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(PKG, "..", "include", "recsys_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())     # names, not paths: the GPU box mounts the repo elsewhere
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name: str, defines) -> str:
    """Tuning aid: lib/librecsys_b200_<name>.so compiled with extra -D flags (picked up through the
    RB_LIB_PATH environment variable by _lib.py).  Not used by the product path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, f"librecsys_b200_{name}.so")
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(LIB_DIR, f"{name}_" + os.path.basename(src).replace(".cu", ".o"))
        procs.append(subprocess.Popen([_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *[f"-D{d}" for d in defines],
                                       "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        objs.append(obj)
    for p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(o)
    subprocess.run([_nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                    "-lpthread", "-ldl", "-lrt"], check=True)
    for o in objs:
        os.remove(o)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if os.environ.get("RB_LIB_PATH"):
        return os.environ["RB_LIB_PATH"]
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp:
        return LIB_PATH
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src).replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {os.path.basename(src)} ====\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see recommender_b200/lib/build.log")
    link = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", help="name of a tuning build (lib/librecsys_b200_<name>.so)")
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    print(build_variant(a.variant, a.defines) if a.variant else build(a.force, a.verbose))
