"""The ray-lattice route of tests/test_fullsize_oracle_gpu.py checked against itself on the CPU: the reference fed a
sub-lattice of plane-parallel rays through per-pixel arrays, slab by slab, gives the pixels of its own full frame."""
import numpy as np
import pytest

from tests import common as T
from tests.test_fullsize_oracle_gpu import reference_on_lattice


@pytest.mark.parametrize("kind", ["vr", "proj"])
def test_lattice_route_reproduces_the_full_frame(kind):
    from oracle import oracle as O
    S, G = O.load_ref() if O.ref_available() else (O, O)
    N, R = 24, 96
    cam = T.make_camera(resolution=(R, R))
    tf = T.colour_tf(height=20.0)
    cells = np.random.default_rng(5).random((N, N, N))
    data = cells if kind == "proj" else O.vertex_centre(cells)
    img = np.zeros((R, R, 4))
    args, kw = T.sampler_args(cam, image=img)
    kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, tf, 10, **kw) if kind == "vr" else S.ProjectionSampler(*args, 10, **kw)
    s(G.PartitionedGrid(0, [data], np.ones((N, N, N), "uint8"), np.zeros(3), np.ones(3), np.array([N] * 3, "int64")), num_threads=1)
    full = np.array(s.aimage)
    ii = jj = np.arange(3, R, 8)
    c = 0 if kind == "proj" else 1
    want, _ = reference_on_lattice(kind, cam, args, ii, jj, N, lambda x0, x1: np.ascontiguousarray(data[x0:x1 + c]), c, tf=tf,
                                   n_slabs=3, threads=1)
    sub = full[np.ix_(ii, jj)]
    assert sub[:, :, :3].max() > 0
    assert np.abs(sub - want).max() <= 1e-11 * max(1.0, np.abs(sub).max())
