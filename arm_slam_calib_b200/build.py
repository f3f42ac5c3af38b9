"""Builds libasc_b200.so (sm_100a only) in-tree with nvcc.  No torch extension machinery: the
library is a plain C-ABI shared object (include/asc.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libasc_b200.so")
HOST_LIB = os.path.join(HERE, "libasc_host.so")
SOURCES = ["asc_api.cu", "sim_generator.cpp"]
DEPS = SOURCES + ["kernels.cuh", "coarse_ml.cuh", "host_math.h", "default_params.inl", "host_only.cpp", os.path.join("..", "..", "include", "asc.h")]


STAMP = LIB + ".dofs"   # which DOF instantiations the existing library holds ("all" = the default list)


def needs_build(dof_list=None):
    if not os.path.exists(LIB):
        return True
    want = "all" if not dof_list else ",".join(map(str, dof_list))
    have = open(STAMP).read().strip() if os.path.exists(STAMP) else ""
    if want != have:
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_host(force=False):
    """libasc_host.so: generator + default parameters, plain g++ (no CUDA): what the reference arm and problem-only callers load."""
    srcs = [os.path.join(CSRC, "sim_generator.cpp"), os.path.join(CSRC, "host_only.cpp")]
    deps = srcs + [os.path.join(CSRC, "default_params.inl"), os.path.join(CSRC, "host_math.h"), os.path.join(HERE, "..", "include", "asc.h")]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    r = subprocess.run(["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-o", HOST_LIB] + srcs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return HOST_LIB


def build(force=False, verbose=False, dof_list=None):
    build_host(force)
    if not force and not needs_build(dof_list):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O3,-fopenmp", "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if dof_list:
        cmd += ["-DASC_DOF_LIST(X)=" + " ".join("X(%d)" % n for n in dof_list)]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lgomp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    with open(STAMP, "w") as f:
        f.write("all" if not dof_list else ",".join(map(str, dof_list)))
    return LIB


if __name__ == "__main__":
    dofs = [int(a) for a in sys.argv[1:] if a.isdigit()]
    build(force=True, verbose="-v" in sys.argv, dof_list=dofs or None)
    print(LIB)
