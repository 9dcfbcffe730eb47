"""Generate tests/golden/ref_octree_images.npz: the UNMODIFIED reference's cast_through_octree (image_samplers.pyx:202-299
over _octree_raytracing.hpp), compiled from /root/reference into oracle/_ref, on the scenarios of
tests/common.py:octree_scenarios().  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_octree_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import common as T  # noqa: E402


def main():
    samplers, grid = O.load_ref()
    from yt.utilities.lib._octree_raytracing import _OctreeRayTracing
    out = {}
    for sc in T.octree_scenarios():
        out[sc.name] = T.run_octree_scenario(sc, samplers, grid, _OctreeRayTracing, num_threads=1)
        print("%-24s cells %6d  max %.6g" % (sc.name, len(sc.lvl), np.nanmax(out[sc.name])))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_octree_images.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
