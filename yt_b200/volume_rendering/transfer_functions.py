"""Transfer functions: the tables the VolumeRenderSampler turns into FieldInterpolationTables.

Mirrors the classes of yt/visualization/volume_rendering/transfer_functions.py that the hot path
consumes through the duck type of image_samplers.pyx:429-446:
``TransferFunction`` (8-71), ``MultiVariateTransferFunction`` (250-361), ``ColorTransferFunction``
(364-406 + add_gaussian/add_step/sample_colormap/map_to_colormap/add_layers, 408-893) and
``ProjectionTransferFunction`` (921-961).  Table values for the same calls are identical to the
reference's (same numpy expressions).
"""
import numpy as np

from . import colormaps


class TransferFunction:
    """One tabulated channel: ``y`` over ``nbins`` points spanning ``x_bounds``."""

    def __init__(self, x_bounds, nbins=256):
        self.pass_through = 0
        self.nbins = nbins
        self.x_bounds = [np.float64(xb) for xb in x_bounds]
        self.x = np.linspace(self.x_bounds[0], self.x_bounds[1], nbins, dtype="float64")
        self.y = np.zeros(nbins, dtype="float64")
        self.grad_field = -1
        self.light_source_v = self.light_source_c = np.zeros(3, "float64")
        self.features = []

    def _merge(self, vals):
        # a feature never lowers the table: y <- clip(max(new, y), 0, inf)
        self.y = np.clip(np.maximum(vals, self.y), 0.0, np.inf)

    def add_gaussian(self, location, width, height):
        """y <- max(y, height * exp(-(x - location)^2 / width))"""
        self._merge(height * np.exp(-((self.x - location) ** 2.0) / width))
        self.features.append(("gaussian", location, width, height))

    def add_line(self, start, stop):
        (x0, y0), (x1, y1) = start, stop
        vals = (y1 - y0) / (x1 - x0) * (self.x - x0) + y0
        vals[~((self.x >= x0) & (self.x <= x1))] = 0.0
        self._merge(vals)
        self.features.append(("line", start, stop))

    def add_step(self, start, stop, value):
        vals = np.zeros(self.x.shape, "float64")
        vals[(self.x >= start) & (self.x <= stop)] = value
        self._merge(vals)
        self.features.append(("step", start, stop, value))

    def add_filtered_planck(self, wavelength, trans):
        raise NotImplementedError("Planck tables are outside the ray-casting hot path")

    def clear(self):
        self.y[:] = 0.0
        self.features = []

    def __repr__(self):
        return "<Transfer Function Object>: x_bounds:(%3.2g, %3.2g) nbins:%3.2g features:%s" % (
            self.x_bounds[0], self.x_bounds[1], self.nbins, self.features)


class MultiVariateTransferFunction:
    """Up to six tables, each bound to a field, optionally weighted by a field and/or another table,
    and linked to the six integration channels."""

    def __init__(self, grey_opacity=False):
        self.n_field_tables = 0
        self.tables = []
        self.field_ids = [0] * 6
        self.weight_field_ids = [-1] * 6
        self.field_table_ids = [0] * 6
        self.weight_table_ids = [-1] * 6
        self.grad_field = -1
        self.light_source_v = self.light_source_c = np.zeros(3, "float64")
        self.grey_opacity = grey_opacity

    def add_field_table(self, table, field_id, weight_field_id=-1, weight_table_id=-1):
        k = self.n_field_tables
        self.tables.append(table)
        self.field_ids[k] = field_id
        self.weight_field_ids[k] = weight_field_id
        self.weight_table_ids[k] = weight_table_id
        self.n_field_tables = k + 1

    def link_channels(self, table_id, channels=0):
        for c in np.atleast_1d(channels):
            self.field_table_ids[int(c)] = table_id


class ColorTransferFunction(MultiVariateTransferFunction):
    """R, G, B tables weighted by the alpha table, all on field 0 (transfer_functions.py:385-406)."""

    def __init__(self, x_bounds, nbins=256, grey_opacity=False):
        super().__init__()
        self.x_bounds = [np.float64(xb) for xb in x_bounds]
        self.nbins = nbins
        self.red, self.green, self.blue, self.alpha = (TransferFunction(self.x_bounds, nbins) for _ in range(4))
        self.funcs = (self.red, self.green, self.blue, self.alpha)
        self.grey_opacity = grey_opacity
        self.features = []
        for i in range(3):
            self.add_field_table(self.funcs[i], 0, weight_table_id=3)
            self.link_channels(i, i)
        self.add_field_table(self.alpha, 0)
        self.link_channels(3, 3)

    def add_gaussian(self, location, width, height):
        if len(height) != 4:
            raise ValueError("height must be (red, green, blue, alpha)")
        for tf, v in zip(self.funcs, height):
            tf.add_gaussian(location, width, v)
        self.features.append(("gaussian", location, width, tuple(height)))

    def add_step(self, start, stop, value):
        if len(value) != 4:
            raise ValueError("value must be (red, green, blue, alpha)")
        for tf, v in zip(self.funcs, value):
            tf.add_step(start, stop, v)
        self.features.append(("step", start, stop, tuple(value)))

    def sample_colormap(self, v, w, alpha=None, colormap="gist_stern", col_bounds=None):
        v = np.float64(v)
        lo, hi = (self.x_bounds if col_bounds is None else col_bounds)
        r, g, b, a = colormaps.get(colormap)((v - lo) / (hi - lo))
        self.add_gaussian(v, w, [r, g, b, a if alpha is None else alpha])

    def map_to_colormap(self, mi, ma, scale=1.0, colormap="gist_stern", scale_func=None):
        mi, ma = np.float64(mi), np.float64(ma)
        span = self.x_bounds[1] - self.x_bounds[0]
        rel0 = max(int(self.nbins * (mi - self.x_bounds[0]) / span), 0)
        rel1 = min(int(self.nbins * (ma - self.x_bounds[0]) / span), self.nbins - 1) + 1
        tomap = np.linspace(0.0, 1.0, num=rel1 - rel0)
        cc = colormaps.get(colormap)(tomap)
        mult = 1.0 if scale_func is None else scale_func(tomap, 0.0, 1.0)
        self.red.y[rel0:rel1] = cc[:, 0] * mult
        self.green.y[rel0:rel1] = cc[:, 1] * mult
        self.blue.y[rel0:rel1] = cc[:, 2] * mult
        self.alpha.y[rel0:rel1] = scale * cc[:, 3] * mult
        self.features.append(("map_to_colormap", mi, ma, scale))

    def add_layers(self, N, w=None, mi=None, ma=None, alpha=None, colormap="gist_stern", col_bounds=None):
        lo, hi = (self.x_bounds if col_bounds is None else col_bounds)
        dist = hi - lo
        if mi is None:
            mi = lo + dist / (10.0 * N)
        if ma is None:
            ma = hi - dist / (10.0 * N)
        if w is None:
            w = max(0.001 * (ma - mi) / N, 1.0 / self.nbins)
        if alpha is None:
            alpha = np.ones(N, dtype="float64") if self.grey_opacity else np.logspace(-3, 0, N)
        for v, a in zip(np.mgrid[mi:ma:N * 1j], alpha):
            self.sample_colormap(v, w, a, colormap=colormap, col_bounds=col_bounds)

    def clear(self):
        for f in self.funcs:
            f.clear()
        self.features = []

    def __repr__(self):
        return "<Color Transfer Function Object>: x_bounds:[%3.2g, %3.2g] nbins:%d features:%s" % (
            self.x_bounds[0], self.x_bounds[1], self.nbins, self.features)


class ProjectionTransferFunction(MultiVariateTransferFunction):
    """Marks a source as a plain line integral (transfer_functions.py:921-961); the samplers it
    selects (ProjectionSampler / InterpolatedProjectionSampler) never evaluate these tables."""

    def __init__(self, x_bounds=(-1e60, 1e60), n_fields=1):
        if n_fields > 3:
            raise NotImplementedError("supplied %d but n_fields > 3 not implemented." % n_fields)
        super().__init__()
        self.x_bounds = [np.float64(xb) for xb in x_bounds]
        self.nbins = 2
        self.linear_mapping = TransferFunction(self.x_bounds, 2)
        self.linear_mapping.pass_through = 1
        self.link_channels(0, [0, 1, 2])
        for i in range(n_fields):
            self.add_field_table(self.linear_mapping, i)
            self.link_channels(i, i)
        self.link_channels(n_fields, [3, 4, 5])
