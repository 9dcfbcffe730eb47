"""Golden images of the UNMODIFIED reference LightSourceRenderSampler (image_samplers.pyx:496-600, compiled into
oracle/_ref) on the volume-render scenarios of tests/common.py.  Run in the build container (needs /root/reference):
    python tests/golden/make_light_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

LIGHT_DIR = (0.3, -1.0, 0.5)
LIGHT_RGBA = (1.0, 0.7, 0.4, 1.0)
NAMES = ["vr_colour", "vr_colour_grey_zbuf", "vr_mask_nan", "vr_multivariate", "vr_8bricks_perspective_grey", "tiny"]


def run_light(sc, sampler_mod, grid_mod, num_threads=1, sampler_kwargs=None):
    """tests.common.run_scenario for the lit sampler."""
    from tests import common as T
    img = np.zeros((sc.cam.resolution[0], sc.cam.resolution[1], 4)) if sc.image0 is None else sc.image0.copy()
    args, kw = T.sampler_args(sc.cam, zbuffer=None if sc.zbuffer is None else sc.zbuffer.copy(), image=img)
    if sc.cam.lens.lens_type != "perspective":
        kw.pop("camera_data")
    kw.update(sampler_kwargs or {})
    s = sampler_mod.LightSourceRenderSampler(*args, sc.tf, sc.n_samples, light_dir=np.array(LIGHT_DIR, dtype="float64"),
                                             light_rgba=np.array(LIGHT_RGBA, dtype="float64"), **kw)
    grids = [grid_mod.PartitionedGrid(i, [T.zero_padded(d) for d in b[0]], b[1], np.array(b[2], "float64"), np.array(b[3], "float64"),
                                      np.array(b[4], "int64")) for i, b in enumerate(sc.bricks)]
    for i in sc.order:
        s(grids[i], num_threads=num_threads)
    return np.array(s.aimage, copy=True)


if __name__ == "__main__":
    from oracle import oracle as O
    from tests import common as T
    samplers, grid = O.load_ref()
    by_name = {s.name: s for s in T.scenarios()}
    out = {n: run_light(by_name[n], samplers, grid) for n in NAMES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_light_images.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
