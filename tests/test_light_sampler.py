"""LightSourceRenderSampler (image_samplers.pyx:496-600): the oracle's restatement against the
compiled reference and the golden images (CPU), the CUDA sampler against the oracle (GPU): FAST within the 1e-4 bar,
EXACT (fp64, reference evaluation order) to rounding."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from tests.golden.make_light_golden import NAMES, run_light

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_light_images.npz")
BY_NAME = {s.name: s for s in T.scenarios()}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden_bit_for_bit(name):
    got = run_light(BY_NAME[name], O, O, num_threads=1)
    assert np.array_equal(got, np.load(GOLD)[name], equal_nan=True)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_matches_compiled_reference(name, threads):
    samplers, grid = O.load_ref()
    want = run_light(BY_NAME[name], samplers, grid, num_threads=threads)
    got = run_light(BY_NAME[name], O, O, num_threads=threads)
    assert np.array_equal(got, want, equal_nan=True)


def test_lighting_changes_the_image_and_keeps_alpha():
    sc = BY_NAME["vr_colour"]
    lit = run_light(sc, O, O)
    plain = T.run_scenario(sc, O, O, num_threads=1)
    assert not np.allclose(lit[..., :3], plain[..., :3])
    assert np.array_equal(lit[..., 3], plain[..., 3])         # neither sampler touches alpha here


def _compare(got, want, tol, name):
    """NaN vertices make the gradient, hence the lit colour, NaN in the reference too (0 * NaN): same pixels must be NaN."""
    assert np.array_equal(np.isnan(got), np.isnan(want)), name
    ok = ~np.isnan(want[..., :3]).any(axis=2)
    cmax = np.abs(want[..., :3][ok]).sum(axis=1).max()
    assert np.abs(got[..., :3][ok] - want[..., :3][ok]).max() <= tol * cmax, name
    assert np.array_equal(got[..., 3], want[..., 3])          # alpha is never written by the lit sampler


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_fast_matches_oracle(name, gpu):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    sc = BY_NAME[name]
    want = run_light(sc, O, O, num_threads=0)
    got = run_light(sc, S, G)
    _compare(got, want, 1e-4, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_exact_matches_oracle_to_rounding(name, gpu):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    sc = BY_NAME[name]
    want = run_light(sc, O, O, num_threads=0)
    got = run_light(sc, S, G, sampler_kwargs={"storage": "f64", "precision": "exact"})
    _compare(got, want, 1e-10, name)
