"""Build recipe for libhjb_b200.so (nvcc, sm_100a only).  Used by ``__graft_entry__.build()`` and runnable
as ``python -m q_learning_with_hjb_b200.build``.  The shared library is built IN-TREE next to this file so
that it travels to the GPU box with the repo snapshot; it is git-ignored."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build", "obj")
LIB_PATH = os.path.join(PKG_DIR, "libhjb_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # the fatbin (SASS + the line tables and PTX text -lineinfo attaches, which were more than half of every object) is stored
    # zstd-compressed and inflated by the driver at load (CUDA >= 12.8): libhjb_b200.so 120 MB -> ~15 MB, same SASS
    "--compress-mode=size",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libhjb_b200.so cannot be built")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "hjb_b200.h"))
    return sorted(hdrs)


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode())     # NOT the path: the GPU box builds from another directory
            h.update(fh.read())
    return h.hexdigest()


KERNEL_FAMILIES = {
    "rollout": ("hjb_common.cuh", "systems.cuh", "rollout_kernel.cuh"),
    "vhjb": ("hjb_common.cuh", "systems.cuh", "umma.cuh", "vhjb_epilogue.cuh", "vhjb_simt.cuh", "vhjb_tc.cuh", "vhjb_tc_res.cuh"),
}


def source_hash(family: str) -> str:
    """sha256 (16 hex digits) of the sources a kernel family is compiled from: stamps an ncu capture (profiles/traffic.json)
    so that bench.py only quotes DRAM traffic measured on the kernel it is timing."""
    return _digest([os.path.join(CSRC, f) for f in KERNEL_FAMILIES[family]])[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libhjb_b200.so.  Incremental: an object is rebuilt
    when its source, any header or the flags changed."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(_headers(), " ".join(NVCC_FLAGS))
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".stamp"
        want = _digest([src], hdr_digest)
        objs.append(obj)
        have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(obj) else ""
        if force or have != want:
            jobs.append((src, obj, stamp, want))

    def compile_one(job):
        src, obj, stamp, want = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as fh:
            fh.write(want)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(compile_one, jobs):
                if verbose and out:
                    print(out)
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--compress-mode=size", "-o", LIB_PATH, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
