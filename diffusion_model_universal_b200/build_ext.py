"""Build libdmu_b200.so (sm_100a) in-tree with nvcc.  No torch involved: the
library is a plain C-ABI shared object (include/dmu_b200.h)."""

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(OUT_DIR, "libdmu_b200.so")
SOURCES = ["process.cu", "norm.cu", "conv_simt.cu", "attention.cu", "attention_tc.cu", "conv_tc.cu", "conv_halo.cu", "conv_wgrad_halo.cu", "conv_edge.cu", "conv_stem.cu", "conv_edge_tc.cu", "pipeline.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest(paths, extra=""):
    """sha256 over the CONTENTS of a source, every header it may include and the compiler flags: what decides whether an
    object file is current (modification times do not survive a checkout or the copy to the GPU box)."""
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, digest):
    """True when `target` is missing or was built from other contents than `digest` says (recorded next to it in .sha)."""
    if not os.path.exists(target) or not os.path.exists(target + ".sha"):
        return True
    with open(target + ".sha") as f:
        return f.read().strip() != digest


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "dmu_b200.h"))
    jobs = []
    digests = {}
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        digests[obj] = _digest([src] + headers, " ".join(NVCC_FLAGS))
        if force or _stale(obj, digests[obj]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = obj + ".log"
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            if os.path.exists(obj + ".sha"):
                os.remove(obj + ".sha")
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".sha", "w") as f:
            f.write(digests[obj])
        return r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(compile_one, jobs):
            if verbose:
                print(out)
    objs = [os.path.join(OUT_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    lib_digest = hashlib.sha256("".join(digests[o] for o in objs).encode()).hexdigest()
    if force or jobs or _stale(LIB, lib_digest):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs      # (without the arch the link step adds an empty sm_52 stub cubin)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(LIB + ".sha", "w") as f:
            f.write(lib_digest)
    return LIB


def is_current() -> bool:
    """Does the shared library on disk correspond to the sources on disk?  (No compiler needed: used by the GPU-side checks.)"""
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "dmu_b200.h"))
    ds = [_digest([os.path.join(CSRC, s)] + headers, " ".join(NVCC_FLAGS)) for s in SOURCES]
    return not _stale(LIB, hashlib.sha256("".join(ds).encode()).hexdigest())


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
