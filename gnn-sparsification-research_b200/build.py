"""In-tree build of libgsp.so (nvcc, sm_100a only). No JIT cache: the .so sits next to this file so it
travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libgsp.so")
SOURCES = ["graph.cu", "intersect.cu", "intersect_owner.cu", "featcos.cu", "select.cu", "approx_er.cu", "backbone.cu", "gcn.cu", "topology.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: libgsp.so cannot be built")
    return path


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(PKG_DIR), "include", "gsp.h")]
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [src_path] + headers):
            jobs.append((src_path, obj))

    def compile_one(job):
        src_path, obj = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src_path, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src_path}:\n{res.stdout}\n{res.stderr}")

    with ThreadPoolExecutor(max_workers=min(4, max(1, len(jobs)))) as pool:
        list(pool.map(compile_one, jobs))
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(verbose=True))
