"""In-tree build of libparafem_b200.so (nvcc, sm_100a only) and of the host driver p121_b200."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libparafem_b200.so")
DRIVER = os.path.join(_HERE, "p121_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # separate IEEE multiply/add everywhere: bit-parity with the oracle (kernels are HBM-bound)
    "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-Wall",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("device.cu", "host.cpp")]
    deps = srcs + [os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "xx3_compat.cuh"),
                   os.path.join(ROOT, "include", "parafem_b200.h"), os.path.join(ROOT, "include", "parafem_xx3_compat.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-I", os.path.join(ROOT, "include"), "-I", "/usr/include",
                                        "-o", LIB] + srcs + ["-lgomp", "-ldl"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libparafem_b200.so")
        with open(os.path.join(_HERE, "ptxas_info.txt"), "w") as f:
            f.write(res.stdout)
    for name in ("p121_b200", "p12x_b200", "p1210_b200"):          # host drivers above the C-ABI (no Fortran compiler here)
        drv_src, exe = os.path.join(CSRC, name + ".cpp"), os.path.join(_HERE, name)
        if os.path.exists(drv_src) and (force or _stale(exe, [drv_src, LIB])):
            cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), drv_src,
                   "-o", exe, "-L", _HERE, "-lparafem_b200", "-Wl,-rpath,$ORIGIN"]
            subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
