"""Build libytb200.so in-tree with nvcc for sm_100a (the only target; no multi-arch dispatch).

Every source under csrc/ is its own translation unit (objects in yt_b200/_build/, compiled in parallel, rebuilt when
the source or any header is newer), linked into one shared library."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
OUT = os.path.join(HERE, "libytb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]


def headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(os.path.dirname(HERE), "include", "ytb200.h")]


def sources():
    return units() + headers()


def is_current():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def _obj(src, tag=""):
    return os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + tag + ".o")


def _compile(src, obj, defines=(), verbose=False):
    cmd = [NVCC] + CFLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    subprocess.check_call(cmd)


def build(force=False, verbose=False, defines=(), only=None, tag="", out=None):
    """Compile what is out of date and link.  `defines` / `only` / `tag` / `out` are for kernel experiments
    (build_variant): extra -D flags applied to the units named in `only`, objects and library suffixed with `tag`."""
    out = out or OUT
    if not defines and is_current() and not force:
        return out
    os.makedirs(OBJ, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in headers())
    jobs, objs = [], []
    for src in units():
        special = bool(defines) and (only is None or os.path.basename(src) in only)
        obj = _obj(src, tag if special else "")
        objs.append(obj)
        stale = force or special or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((src, obj, tuple(defines) if special else (), verbose))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for f in [ex.submit(_compile, *j) for j in jobs]:
            f.result()
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
    return out


def build_variant(tag, defines, only=None, verbose=False):
    """Kernel experiments: libytb200_<tag>.so with extra -D flags (select with YTB200_LIB=<path>)."""
    return build(defines=defines, only=only, tag="_" + tag, out=os.path.join(HERE, "libytb200_%s.so" % tag), verbose=verbose)


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
