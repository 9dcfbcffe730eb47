"""ZBuffer -- an RGBA image with a per-pixel depth (yt/visualization/volume_rendering/zbuffer_array.py:4-79).

Opaque sources paint into it; the volume sampler then stops each ray at ``clip(z, 0, 1)``
(image_samplers.pyx:185) and starts from the painted colours (utils.py:79-84).
"""
import numpy as np


class ZBuffer:
    def __init__(self, rgba, z):
        assert rgba.shape[: len(z.shape)] == z.shape
        self.rgba = rgba
        self.z = z
        self.shape = z.shape

    def __add__(self, other):
        """Per pixel, keep whichever buffer is nearer."""
        assert self.shape == other.shape
        nearer = self.z < other.z
        if self.z.shape[1] == 1:  # non-rectangular buffers
            rgba = self.rgba * nearer[:, None, :] + other.rgba * (1.0 - nearer)[:, None, :]
        else:
            farther = self.z > other.z
            rgba = np.zeros(self.rgba.shape)
            rgba[nearer] = self.rgba[nearer]
            rgba[farther] = other.rgba[farther]
        return ZBuffer(rgba, np.min([self.z, other.z], axis=0))

    def __iadd__(self, other):
        merged = self + other
        self.rgba, self.z = merged.rgba, merged.z
        return self

    def __eq__(self, other):
        return bool(np.all(self.rgba == other.rgba) and np.all(self.z == other.z))

    def paint(self, ind, value, z):
        if z < self.z[ind]:
            self.rgba[ind] = value
            self.z[ind] = z
