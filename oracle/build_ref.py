"""Build the UNMODIFIED reference ray caster into ``oracle/_ref`` (test infrastructure only).

This is *not* product code.  It compiles, straight from the read-only checkout
at ``/root/reference`` (no source is copied into this repository), the five
Cython extension modules that make up yt's software ray caster:

    yt/utilities/lib/image_samplers.pyx      (ImageSampler family)
    yt/utilities/lib/grid_traversal.pyx      (walk_volume)
    yt/utilities/lib/partitioned_grid.pyx    (PartitionedGrid / VolumeContainer)
    yt/utilities/lib/lenses.pyx              (extent + ray generation)
    yt/utilities/lib/_octree_raytracing.pyx  (cimported by image_samplers.pyx:34)
    yt/utilities/lib/amr_kdtools.pyx         (kd-tree Node, view-order traversal)
    yt/utilities/lib/interpolators.pyx       (ghost_zone_interpolate: no-ghost vertex centring)
  + yt/utilities/lib/fixed_interpolator.cpp  (offset_interpolate)

Outputs (``.so`` files, generated C++ and object files) go ONLY under
``oracle/_ref`` which is git-ignored but travels to the GPU box with gpurun.
The only Python-level dependency of those modules is ``from yt.funcs import
mylog`` (image_samplers.pyx:32); a two-line logging shim is written for it.

Usage:  python oracle/build_ref.py [--reference /root/reference]
"""
import argparse
import os
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

MODS = ["image_samplers", "grid_traversal", "partitioned_grid", "lenses",
        "_octree_raytracing", "amr_kdtools", "interpolators"]
NEED_FIXED_INTERP = {"image_samplers", "partitioned_grid"}


def is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    lib = os.path.join(OUT, "yt", "utilities", "lib")
    return all(os.path.exists(os.path.join(lib, m + suffix)) for m in MODS)


def build(reference="/root/reference", force=False):
    if is_built() and not force:
        return True
    libsrc = os.path.join(reference, "yt", "utilities", "lib")
    if not os.path.isdir(libsrc):
        return False
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution

    # /opt/gcc/bin/g++ (the image's default $CXX) cannot link -fopenmp; use the system one.
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    os.environ["LDSHARED"] = "/usr/bin/g++ -shared"

    exts = []
    for m in MODS:
        src = [os.path.join(libsrc, m + ".pyx")]
        if m in NEED_FIXED_INTERP:
            # the reference links this as a static library (setup.py:116-123)
            src.append(os.path.join(libsrc, "fixed_interpolator.cpp"))
        exts.append(Extension(
            "yt.utilities.lib." + m, src,
            include_dirs=[numpy.get_include(), libsrc],
            define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
            language="c++"))
    cwd = os.getcwd()
    os.makedirs(OUT, exist_ok=True)
    os.chdir(OUT)
    try:
        exts = cythonize(
            exts, build_dir=os.path.join(OUT, "build", "cy"),
            include_path=[reference],
            aliases={"LIB_DIR": [libsrc], "STD_LIBS": ["m"], "OMP_ARGS": ["-fopenmp"],
                     "FIXED_INTERP": [], "CPP14_FLAG": ["-std=c++14"]},
            compiler_directives={"language_level": 3}, quiet=True)
        dist = Distribution({"name": "yt_ref_oracle", "ext_modules": exts})
        cmd = dist.get_command_obj("build_ext")
        cmd.build_lib = OUT
        cmd.build_temp = os.path.join(OUT, "build", "obj")
        cmd.parallel = os.cpu_count()
        cmd.ensure_finalized()
        cmd.run()
    finally:
        os.chdir(cwd)
    # package skeleton + the yt.funcs shim (ours, two lines; not reference source)
    for d in ("yt", "yt/utilities", "yt/utilities/lib"):
        p = os.path.join(OUT, d, "__init__.py")
        if not os.path.exists(p):
            open(p, "w").close()
    with open(os.path.join(OUT, "yt", "funcs.py"), "w") as f:
        f.write('import logging\nmylog = logging.getLogger("yt")\n')
    return is_built()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
