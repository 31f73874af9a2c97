"""Build librapiddoc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librapiddoc_b200.so")
SOURCES = ["api.cu", "det.cu", "rec.cu", "ops.cu", "contours.cu", "table.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--use_fast_math" if os.environ.get("RDB_FAST_MATH") else "-DRDB_NO_FAST_MATH", "-cudart", "static"]
# aligned dynamic-smem base derived without leaving the shared address space: LDS/STS instead of generic LD/ST in the fused
# kernels (gemm_tc.cuh RDB_ALIGNED_SMEM).  A/B on a B200 (profiles/r02_ab_smem_base.txt): det 5908 -> 6109 pages/s, rec 97.6 ->
# 98.8 k crops/s, parity tests green; RDB_SMEM_BASE=generic builds the round-1 variant.
if os.environ.get("RDB_SMEM_BASE") != "generic":
    FLAGS.append("-DRDB_SMEM_SHARED_BASE")


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "rapiddoc_b200.h"))
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
