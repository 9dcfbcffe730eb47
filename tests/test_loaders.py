"""Host mirrors of the stream loaders against the reference (CPU): the processor grid of load_uniform_grid(nprocs>1)
(yt/utilities/decompose.py:69-148, golden values made by tests/golden/make_psize_golden.py from the reference module
itself) and the host no-ghost vertex centring (interpolators.pyx:177-227, bit-identical to the pinned oracle)."""
import json
import os

import numpy as np
import pytest

from yt_b200 import loaders

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psize_golden.json")


def test_get_psize_and_piece_edges_equal_the_reference():
    cases = json.load(open(GOLD))
    assert len(cases) >= 200
    for c in cases:
        ps = loaders.get_psize(c["shape"], c["pieces"])
        assert [int(x) for x in ps] == c["psize"], c
        e = loaders.piece_edges(c["shape"], ps)
        assert len(e) == int(np.prod(c["psize"]))      # (more pieces than cells along every axis: the reference keeps 1x1x1)
        assert [int(x) for x in e[0][1] - e[0][0]] == c["first_dims"]
        assert [int(x) for x in e[-1][1] - e[-1][0]] == c["last_dims"]


@pytest.mark.parametrize("nprocs", [2, 4, 6])
def test_uniform_grid_pieces_tile_the_domain_like_the_reference(nprocs):
    arr = np.random.default_rng(3).random((12, 10, 16))
    ds = loaders.load_uniform_grid({"density": arr}, arr.shape, nprocs=nprocs)
    gold = {(tuple(c["shape"]), c["pieces"]): c for c in json.load(open(GOLD))}
    assert len(ds.index.grids) == nprocs
    ps = loaders.get_psize(arr.shape, nprocs)
    assert int(np.prod(ps)) == nprocs
    total = np.zeros(arr.shape)
    for g in ds.index.grids:
        lo = g.get_global_startindex()
        hi = lo + g.ActiveDimensions
        total[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
        assert np.array_equal(g["density"], arr[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])
    assert np.all(total == 1)
    del gold


@pytest.mark.parametrize("shape", [(2, 2, 2), (5, 7, 3), (16, 9, 33)])
def test_host_vertex_centring_is_bit_identical_to_the_oracle(shape):
    from oracle import oracle as O
    c = np.random.default_rng(1).random(shape) * 10 - 3
    assert np.array_equal(loaders.vertex_centre_no_ghost(c), O.vertex_centre(c))


def test_host_vertex_centring_rejects_one_cell_axes():
    with pytest.raises(NotImplementedError):
        loaders.vertex_centre_no_ghost(np.ones((4, 1, 4)))
