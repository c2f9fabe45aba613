"""Build liblvae_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every .cu of csrc/ is compiled to an object under build/ (in parallel, only when it or a header changed) and linked into
lib/liblvae_b200.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "liblvae_b200.so")
SOURCES = ["lvae_dense.cu", "lvae_gemm.cu", "lvae_blas.cu", "lvae_kld.cu", "lvae_kld64.cu", "lvae_kld_big.cu",
           "lvae_prep.cu", "lvae_prep3.cu", "lvae_subjects_fused3.cu", "lvae_subjects_big.cu",
           "lvae_kernel_grad.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "lvae_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hm = _headers_mtime()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in sources:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            jobs.append((s, [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]))

    def run(job):
        r = subprocess.run(job[1], capture_output=True, text=True)
        return job[0], r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, r in ex.map(run, jobs):
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {name}:\n" + r.stdout + r.stderr)
                if verbose:
                    print(f"==== {name}\n{r.stderr}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources]
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
