"""ZBuffer -- an RGBA image with a depth per pixel (interface of
yt/visualization/volume_rendering/zbuffer_array.py:4-79).

Opaque sources paint into it; the volume cast then starts from the painted colours (utils.py:79-84) and stops each
ray at ``clip(z, 0, 1)`` (image_samplers.pyx:185).
"""
import numpy as np


class ZBuffer:
    def __init__(self, rgba, z):
        if rgba.shape[: z.ndim] != z.shape:
            raise AssertionError("rgba %s and z %s do not describe the same pixels" % (rgba.shape, z.shape))
        self.rgba, self.z = rgba, z
        self.shape = z.shape

    def _merge(self, other):
        """Per pixel the nearer of the two buffers; where the depths tie the colour is left black, as in the
        reference (neither ``<`` nor ``>`` selects it, zbuffer_array.py:52-58)."""
        if self.shape != other.shape:
            raise AssertionError("buffers of different shape")
        mine = self.z < other.z
        if self.z.shape[1] == 1:      # flattened (non-rectangular) buffers: blend by the mask
            rgba = self.rgba * mine[:, None, :] + other.rgba * (1.0 - mine)[:, None, :]
        else:
            theirs = other.z < self.z
            rgba = np.zeros(self.rgba.shape)
            rgba[mine] = self.rgba[mine]
            rgba[theirs] = other.rgba[theirs]
        return rgba, np.minimum(self.z, other.z)

    def __add__(self, other):
        return ZBuffer(*self._merge(other))

    def __iadd__(self, other):
        self.rgba, self.z = self._merge(other)
        return self

    def __eq__(self, other):
        return bool((self.rgba == other.rgba).all() and (self.z == other.z).all())

    def paint(self, ind, value, z):
        """Set pixel ``ind`` to ``value`` if ``z`` is in front of what is there."""
        if z < self.z[ind]:
            self.rgba[ind], self.z[ind] = value, z
