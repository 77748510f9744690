"""
Builds libfmgpu.so (the C-ABI product library) in-tree with nvcc for sm_100a.

    python -m findex_b200.build            # or findex_b200.build.build()

The .so is git-ignored but travels to the GPU box with gpurun snapshots.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfmgpu.so")
OBJ = os.path.join(HERE, "_build")
SOURCES = ["fmx_files.cpp", "fmx_regex.cpp", "fmx_kernels.cu", "fmx_regex_kernel.cu", "fmx_cub.cu", "fmx_build.cu", "fmx_api.cu"]
HEADERS = ["fmx_internal.h", "fmx_device.cuh", "fmx_kernels.cuh", "fmx_cub.cuh", "fmx_build.cuh", "../../include/fmgpu.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcudafe", "--diag_suppress=177", "-Wno-deprecated-declarations"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join([NVCC] + FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on " + s)
        if verbose and out:
            sys.stderr.write(out.decode())
    if force or procs or _newer(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
        subprocess.check_call(cmd)
        with open(stamp, "w") as f:
            f.write(flags_now)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
