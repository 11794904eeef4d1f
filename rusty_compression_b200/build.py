"""Builds rusty_compression_b200/librc_b200.so (sm_100a only) with nvcc, in-tree.

    python -m rusty_compression_b200.build [--force]

The library has no torch dependency (pure CUDA runtime + dlopen'ed NCCL); it is the tested
artefact behind include/rc_api.h.  The built .so is git-ignored but travels to the GPU box."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librc_b200.so")
SOURCES = ["kernels_basic.cu", "gemm_generic.cu", "gemm_dmma.cu", "gemm_tf32.cu", "tsqr.cu", "pivqr.cu", "jacobi.cu",
           "trsm.cu", "chol.cu", "comm.cu", "host_api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/rc_api.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if verbose:
        print(f"[rc_b200] compiling {len(SOURCES)} CUDA sources for sm_100a ...", file=sys.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ldl", "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    peaks = os.path.join(OBJ, "rc_peaks")
    r = subprocess.run([NVCC] + FLAGS[:7] + [os.path.join(CSRC, "peaks.cu"), "-o", peaks], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"peaks build failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[rc_b200] built {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
