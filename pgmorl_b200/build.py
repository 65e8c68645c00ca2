"""In-tree build of the sm_100a shared library (C ABI in include/pgmorl_b200.h).

    python -m pgmorl_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting pgmorl_b200/libpgmorl_b200.so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpgmorl_b200.so")
OUT_DIAG = os.path.join(HERE, "libpgmorl_b200_diag.so")      # csrc/diag/*.cu + capi.cu: tcgen05 self-test / probes
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def diag_sources():
    d = os.path.join(CSRC, "diag")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))


def deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdr.append(os.path.join(HERE, "..", "include", "pgmorl_b200.h"))
    hdr.append(os.path.join(HERE, "..", "include", "pgmorl_b200_diag.h"))
    return hdr


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_trace():
    """Instrumented variant (per-phase clock marks in K3) -> libpgmorl_b200_trace.so; profiling aid only."""
    out = os.path.join(HERE, "libpgmorl_b200_trace.so")
    r = subprocess.run([NVCC] + FLAGS + ["-DPGM_K3_TRACE", "-shared", "-o", out] + sources() + ["-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return out


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = deps()
    jobs = []
    for src in sources() + diag_sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdr):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return src, r.returncode, r.stdout + r.stderr

    failed = False
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for src, rc, log in ex.map(compile_one, jobs):
            if rc != 0:
                failed = True
                sys.stderr.write(f"nvcc failed on {src}:\n{log}\n")
            elif verbose:
                sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources()]
    if force or jobs or _stale(OUT, objs):
        r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    # diagnostics library: the probes + the error helpers of capi.cu, nothing of the product kernels
    dobjs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in diag_sources()] + [os.path.join(OBJ, "capi.o")]
    if force or jobs or _stale(OUT_DIAG, dobjs):
        r = subprocess.run([NVCC, "-shared", "-o", OUT_DIAG] + dobjs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link (diag) failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
