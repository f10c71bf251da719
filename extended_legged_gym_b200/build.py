"""In-tree build of the C-ABI CUDA library (``libelg_b200.so``) for sm_100a.

``python -m extended_legged_gym_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels with the tree to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libelg_b200.so")
STAMP = LIB_PATH + ".stamp"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-fmad=false", "--use_fast_math=false"]


LINK_LIBS = ["-ldl"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(PKG_DIR), "include", "elg_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())      # (not the absolute path: the tree is copied to the GPU box)
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the B200 hot path cannot be built (there is no CPU fallback)")


def _compile_one(src: str, flags, common_fp: str, verbose: bool, tag: str = ""):
    """One translation unit -> object file under csrc/_obj (git-ignored), skipped when its fingerprint is unchanged."""
    obj_dir = os.path.join(CSRC, "_obj" + tag)
    os.makedirs(obj_dir, exist_ok=True)
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    h = hashlib.sha256(common_fp.encode())
    with open(src, "rb") as fh:
        h.update(fh.read())
    fp = h.hexdigest()
    if os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read().strip() == fp:
        return obj, 0, ""
    cmd = [nvcc_path()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode == 0:
        with open(obj + ".stamp", "w") as fh:
            fh.write(fp)
    return obj, res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """``defines`` / ``out``: diagnostic variants (e.g. ``-DELG_STEP_STAMPS`` -> a second library with the in-kernel
    clock stamps compiled in, scripts/step_stamps.py); the product library is the default call."""
    lib_path = out or LIB_PATH
    stamp = lib_path + ".stamp"
    tag = ("_" + hashlib.sha256(" ".join(defines).encode()).hexdigest()[:8]) if defines else ""
    fp = _fingerprint() + tag
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return lib_path
    from concurrent.futures import ThreadPoolExecutor
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math") and f != "-shared"] + ["-D" + d for d in defines]
    # headers + flags: a change there recompiles every translation unit; otherwise only the edited .cu files
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
            [os.path.join(os.path.dirname(PKG_DIR), "include", "elg_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    common_fp = h.hexdigest() + tag + ("force" + str(os.getpid()) if force else "")
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, flags, common_fp, verbose, tag), _sources()))
    for obj, rc, log in results:
        if verbose or rc != 0:
            sys.stderr.write(log)
        if rc != 0:
            raise RuntimeError("nvcc failed building " + obj)
    link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", lib_path] + \
        [r[0] for r in results] + LINK_LIBS
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libelg_b200.so")
    with open(stamp, "w") as fh:
        fh.write(fp)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
