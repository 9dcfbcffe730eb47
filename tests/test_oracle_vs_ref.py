"""The C oracle restates the reference bit-for-bit: compare it with the UNMODIFIED reference
samplers compiled from /root/reference into oracle/_ref (oracle/build_ref.py)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("sc", T.scenarios(), ids=lambda s: s.name)
@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_matches_reference_bitwise(sc, threads):
    ref_samplers, ref_grid = O.load_ref()
    ref = T.run_scenario(sc, ref_samplers, ref_grid, num_threads=threads)
    got = T.run_scenario(sc, O, O, num_threads=threads)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    assert np.array_equal(np.nan_to_num(ref), np.nan_to_num(got)), np.nanmax(np.abs(ref - got))
    assert np.isfinite(ref).any()
