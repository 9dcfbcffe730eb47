"""Input generators of the reference's test-suite that the ray-casting configs name (BASELINE.json configs[0]).

``fake_random_ds`` follows yt/testing.py:345-436 for grid fields: one ``RandomState(0x4D3D3D3)`` stream, one
``random_sample(ndims)`` draw per field in field order (density first, then the three velocities), minus 0.5 where
``negative``, times ``peak_value``; the arrays go to ``load_uniform_grid`` with ``nprocs`` pieces.  Particles, units
and non-Cartesian geometries are outside the path.
"""
import numpy as np

from .loaders import load_uniform_grid

_DEFAULT_FIELDS = ("density", "velocity_x", "velocity_y", "velocity_z")


def fake_random_ds(ndims, peak_value=1.0, fields=None, negative=False, nprocs=1, bbox=None):
    prng = np.random.RandomState(0x4D3D3D3)
    if np.ndim(ndims) == 0:
        ndims = [ndims, ndims, ndims]
    elif len(ndims) != 3:
        raise AssertionError("ndims must be a scalar or have three entries")
    fields = tuple(fields) if fields else _DEFAULT_FIELDS
    if np.ndim(negative) == 0:
        negative = [bool(negative)] * len(fields)
    if len(negative) != len(fields):
        raise RuntimeError("`negative` needs one entry per field")
    data = {}
    for field, neg in zip(fields, negative):
        data[field] = (prng.random_sample(ndims) - (0.5 if neg else 0.0)) * peak_value
    return load_uniform_grid(data, ndims, nprocs=nprocs, bbox=bbox)
