"""In-tree build of libpnce.so (sm_100a only).  `python -m gan_variant_research_b200.build`."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PNCE_EXPERIMENTS=1: the build with the pnce_debug_* hooks the scripts under scratch/ use (never the shipped library)
EXPERIMENTS = os.environ.get("PNCE_EXPERIMENTS", "") not in ("", "0")
LIB = os.path.join(HERE, "libpnce_exp.so" if EXPERIMENTS else "libpnce.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-exceptions",      # plain C ABI: no libstdc++ at run time
    "-Xlinker", "--version-script=" + os.path.join(CSRC, "pnce.map"),
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".map")))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(HERE, "..", "include", "pnce.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr, os.path.abspath(__file__)])


def build(force=False, verbose=False):
    """Compile csrc/pnce_api.cu -> libpnce.so with nvcc (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("PNCE_NVCC_EXTRA", "").split() if EXPERIMENTS else []      # experiment builds only (e.g. -DPNCE_GP_PIPE=1)
    cmd = [nvcc] + NVCC_FLAGS + (["-DPNCE_EXPERIMENTS"] if EXPERIMENTS else []) + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB + ".tmp", os.path.join(CSRC, "pnce_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpnce.so")
    if verbose:
        sys.stderr.write(res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
