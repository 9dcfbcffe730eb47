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


def fake_vr_orientation_test_ds(N=96, scale=1):
    """The orientation / handedness toy volume of yt/testing.py:653-738 on [-scale, scale]^3: background 1e-4, a
    sphere (1.0) at the origin, the three lower boundary layers two cells thick (0.3), one cube on +x (1.0), two on
    +y (0.8) and three on +z (0.6) -- so a rendering tells the axes and their signs apart."""
    lo, hi = -1.0 * scale, 1.0 * scale
    d = (hi - lo) / N
    axis = (np.arange(N) + 0.5) * d + lo
    x, y, z = np.meshgrid(axis, axis, axis, indexing="ij")
    arr = np.full((N, N, N), 1.0e-4, dtype=np.float64)
    mid = 0.5 * (lo + hi)
    arr[np.sqrt((x - mid) ** 2 + (y - mid) ** 2 + (z - mid) ** 2) < 0.05] = 1.0
    for c in (x, y, z):
        arr[abs(c - lo) < 2 * d] = 0.3
    half = 0.05 * scale

    def cube(along, centre, others):
        inside = (along > centre - half) & (along < centre + half)
        for o in others:
            inside &= (o > -half) & (o < half)
        return inside

    arr[cube(x, 0.75 * scale, (y, z))] = 1.0
    for c in (0.65 * scale, 0.85 * scale):
        arr[cube(y, c, (x, z))] = 0.8
    for c in (0.5 * scale, 0.7 * scale, 0.9 * scale):
        arr[cube(z, c, (x, y))] = 0.6
    return load_uniform_grid({"density": arr}, arr.shape, bbox=np.array([[lo, hi]] * 3))
