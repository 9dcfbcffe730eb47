"""Build libytb200.so in-tree with nvcc for sm_100a (the only target; no multi-arch dispatch)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "ytb200.cu"), os.path.join(HERE, "csrc", "kdtree.cpp"), os.path.join(HERE, "csrc", "zraster.cpp")]
OUT = os.path.join(HERE, "libytb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh", ".cpp"))] + [
        os.path.join(os.path.dirname(HERE), "include", "ytb200.h")]


def is_current():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if is_current() and not force:
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))


def build_variant(tag, defines, verbose=False):
    """Kernel experiments: libytb200_<tag>.so with extra -D flags (select with YTB200_LIB=<path>)."""
    out = os.path.join(HERE, "libytb200_%s.so" % tag)
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    subprocess.check_call(cmd)
    return out
