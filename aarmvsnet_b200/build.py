"""Builds lib/libaarmvs_sm100.so from csrc/*.cu with nvcc for sm_100a (in-tree; the .so travels with the repo)."""
import concurrent.futures
import fcntl
import glob
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libaarmvs_sm100.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libaarmvs_sm100.so for sm_100a)")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("AARMVS_NVCC_EXTRA", "")).encode())
    return h.hexdigest()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu and link the C-ABI shared library. Returns its path."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    stamp = os.path.join(LIBDIR, "build.sha256")
    digest = _digest(deps)
    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest

    if not force and fresh():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    # several ranks of one job may get here at once (torch.distributed.run): one builds, the others wait on the lock and
    # then find the library fresh; the .so is linked under a temporary name and renamed into place, never seen half-written
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return LIB
        return _build_locked(srcs, stamp, digest, verbose)


def _build_locked(srcs, stamp, digest, verbose):
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("AARMVS_NVCC_EXTRA", "").split()
    objs = [os.path.join(LIBDIR, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc, *flags, "-I", INCLUDE, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as ex:
        for src, r in ex.map(compile_one, zip(srcs, objs)):
            logs.append(f"== {os.path.basename(src)}\n{r.stderr}")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    tmp = f"{LIB}.tmp.{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
