"""The per-table decision between LINE-form rows and (value, increment) rows (``ytb200_tf_create``), through its host-only
probe: no device needed.  LINE-form rows write a bin's line in the table's own row coordinate, ``y = d * fs + c``, which
saves the sample loop two instructions per lookup but rounds ``c`` at a magnitude of ``|row - origin| * |d|``; the guard
admits a table only if every bin keeps a relative accuracy of 5e-6 of its own mean magnitude for the best origin."""
import ctypes as C

import numpy as np

from yt_b200 import _lib as L


def _probe(tables):
    lib = L.load()
    y = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
    origin, worst = C.c_int(0), C.c_double(0.0)
    rc = lib.ytb200_tf_line_form_probe(len(y[0]), *[t.ctypes.data for t in y], C.byref(origin), C.byref(worst))
    assert rc in (0, 1)
    return bool(rc), origin.value, worst.value


def _restate(tables, origin):
    """The guard, restated in numpy."""
    worst = 0.0
    nb = len(tables[0])
    r = np.arange(nb - 1) + 2
    for y in tables:
        y = np.asarray(y, dtype=np.float64)
        df = y[1:].astype(np.float32) - y[:-1].astype(np.float32)
        c64 = y[:-1] - (r - origin) * df.astype(np.float64)
        err = np.abs(c64.astype(np.float32).astype(np.float64) - c64) + np.abs(df.astype(np.float64) - np.diff(y))
        lim = 5e-6 * 0.5 * (np.abs(y[:-1]) + np.abs(y[1:])) + 1e-9 * np.abs(y).max()
        with np.errstate(divide="ignore", invalid="ignore"):
            worst = max(worst, float(np.nanmax(np.where(lim > 0, err / lim, np.where(err > 0, np.inf, 0.0)))))
    return worst


def test_benchmark_table_takes_the_line_form():
    from bench import bench_tf
    tf = bench_tf()
    tables = [tf.funcs[c].y for c in range(4)]
    ok, origin, worst = _probe(tables)
    assert ok and 2 <= origin < tf.nbins + 2 and worst <= 1.0
    assert abs(worst - _restate(tables, origin)) <= 1e-9 * max(worst, 1.0)
    # the chosen origin is the best one
    assert worst <= min(_restate(tables, o) for o in range(2, tf.nbins + 2, 7)) * (1 + 1e-12)


def test_layered_colormap_tables_take_the_line_form():
    """What users build: ColorTransferFunction(bounds) (256 bins) + add_layers -- narrow gaussians all over the range, and a
    table that does NOT vanish at its bounds."""
    from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction
    tf = ColorTransferFunction((-3.0, 1.5))
    tf.add_layers(8)
    ok, origin, worst = _probe([tf.funcs[c].y for c in range(4)])
    assert ok and worst <= 1.0, worst


def test_a_noisy_table_keeps_value_and_increment_rows():
    """Large increments far from any origin: c = y - (row - origin) * d is rounded at a magnitude of ~40, more than a bin of
    mean 0.5 tolerates.  (A cliff between exactly representable values, 0 -> 1, is exact in either form and passes.)"""
    rng = np.random.default_rng(3)
    ok, origin, worst = _probe([rng.random(256) for _ in range(4)])
    assert not ok and worst > 1.0
    step = np.zeros(256)
    step[200:] = 1.0
    assert _probe([step, step, step, step])[0]


def test_degenerate_tables():
    n = 64
    z = np.zeros(n)
    assert _probe([z, z, z, z])[0]                       # all zero: trivially exact
    c = np.full(n, 0.25)
    assert _probe([c, c, c, c])[0]                       # constant: d = 0, c = y
    bad = z.copy()
    bad[7] = np.nan
    assert not _probe([z, z, z, bad])[0]                 # NaN never passes the guard
    big = np.linspace(0, 1, 300)
    assert not _probe([big, big, big, big])[0]           # more than 256 bins: value layout, no LINE form
