"""Colour maps for ``ColorTransferFunction.sample_colormap`` / ``map_to_colormap``.

The reference asks matplotlib for the map (transfer_functions.py:742-752).  matplotlib is not a
dependency here; a map is anything callable as ``cmap(x) -> (..., 4)`` RGBA for x in [0, 1].  A few
analytic maps are registered; ``matplotlib`` names are honoured when matplotlib is importable.  The
table *contents* are an input to both the reference and the CUDA path, so parity does not depend on
which map is used.
"""
import numpy as np

_registry = {}


def register(name, fn):
    _registry[name] = fn


def _from_nodes(nodes):
    xs = np.array([n[0] for n in nodes], dtype="float64")
    cs = np.array([n[1] for n in nodes], dtype="float64")

    def cmap(x):
        x = np.clip(np.asarray(x, dtype="float64"), 0.0, 1.0)
        rgb = np.stack([np.interp(x, xs, cs[:, k]) for k in range(3)], axis=-1)
        return np.concatenate([rgb, np.ones(rgb.shape[:-1] + (1,))], axis=-1)
    return cmap


# piecewise-linear stand-ins (same role as the maps yt defaults to)
register("spectral", _from_nodes([
    (0.00, (0.00, 0.00, 0.00)), (0.10, (0.47, 0.00, 0.53)), (0.20, (0.00, 0.00, 0.73)),
    (0.30, (0.00, 0.53, 0.87)), (0.40, (0.00, 0.67, 0.60)), (0.50, (0.00, 0.60, 0.00)),
    (0.60, (0.00, 0.87, 0.00)), (0.70, (0.80, 0.97, 0.00)), (0.80, (1.00, 0.67, 0.00)),
    (0.90, (0.93, 0.00, 0.00)), (1.00, (0.80, 0.80, 0.80))]))
register("arbre", _from_nodes([
    (0.00, (0.35, 0.20, 0.10)), (0.25, (0.80, 0.55, 0.15)), (0.50, (0.95, 0.95, 0.75)),
    (0.75, (0.30, 0.70, 0.45)), (1.00, (0.05, 0.30, 0.35))]))
register("gist_stern", _from_nodes([
    (0.00, (0.00, 0.00, 0.00)), (0.05, (1.00, 0.05, 0.10)), (0.25, (0.00, 0.25, 0.50)),
    (0.50, (0.50, 0.50, 1.00)), (0.735, (0.73, 0.73, 0.00)), (1.00, (1.00, 1.00, 1.00))]))
register("gray", _from_nodes([(0.0, (0.0, 0.0, 0.0)), (1.0, (1.0, 1.0, 1.0))]))
_registry["nipy_spectral"] = _registry["spectral"]
_registry["cmyt.arbre"] = _registry["arbre"]


def get(name):
    """Resolve a colormap: a callable is returned as is; a name is looked up in matplotlib, then here."""
    if callable(name):
        return name
    try:  # pragma: no cover - matplotlib is optional
        import matplotlib as mpl
        return mpl.colormaps[name]
    except Exception:
        pass
    if name in _registry:
        return _registry[name]
    raise KeyError("unknown colormap %r (registered: %s)" % (name, sorted(_registry)))
