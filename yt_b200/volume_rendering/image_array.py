"""ImageArray -- ndarray subclass returned by Scene.render (yt/data_objects/image_array.py).

Only the image operations are restated: ``rescale`` (image_array.py:191-250, the normalisation the
1e-4 parity bar is stated in), ``write_png``/``save`` with sigma clipping (252-330) through a
self-contained PNG encoder (the reference goes through matplotlib)."""
import struct
import zlib

import numpy as np


def _png_bytes(rgba8):
    """Minimal RGBA8 PNG encoder; rgba8 is (rows, cols, 4) uint8."""
    h, w = rgba8.shape[:2]
    raw = b"".join(b"\x00" + rgba8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


class ImageArray(np.ndarray):
    def __new__(cls, input_array, info=None):
        obj = np.asarray(input_array).view(cls)
        obj.info = {} if info is None else info
        return obj

    def __array_finalize__(self, obj):
        self.info = getattr(obj, "info", None) or {}

    def rescale(self, cmax=None, amax=None, inline=True):
        """Divide RGB by max(R+G+B) and alpha by max(alpha), clip to [0, 1]."""
        assert len(self.shape) == 3 and self.shape[2] >= 3
        out = self if inline else self.copy()
        if cmax is None:
            cmax = np.asarray(self[:, :, :3]).sum(axis=2).max()
        if cmax > 0.0:
            np.multiply(self[:, :, :3], 1.0 / cmax, out[:, :, :3])
        if self.shape[2] == 4:
            if amax is None:
                amax = np.asarray(self[:, :, 3]).max()
            if amax > 0.0:
                np.multiply(self[:, :, 3], 1.0 / amax, out[:, :, 3])
        np.clip(out, 0.0, 1.0, out)
        return out

    def write_png(self, filename, sigma_clip=None, background="black", rescale=True):
        out = self.rescale(inline=False) if rescale else np.array(self, copy=True).view(ImageArray)
        if sigma_clip is not None:
            nz = np.asarray(out[:, :, :3])[np.asarray(out[:, :, :3]) != 0]
            if nz.size:
                clip_value = nz.mean() + sigma_clip * nz.std()
                np.clip(out[:, :, :3] / clip_value, 0.0, 1.0, out[:, :, :3])
        arr = np.asarray(out, dtype="float64")
        if arr.shape[2] == 3:
            arr = np.concatenate([arr, np.ones(arr.shape[:2] + (1,))], axis=2)
        if background == "black":
            arr = arr.copy()
            arr[:, :, 3] = 1.0
        rgba8 = (np.clip(arr, 0.0, 1.0) * 255).astype("uint8")
        data = _png_bytes(np.ascontiguousarray(rgba8.swapaxes(0, 1)[::-1]))
        if filename is None:
            return data
        if not filename.endswith(".png"):
            filename += ".png"
        with open(filename, "wb") as f:
            f.write(data)
        return filename

    save = write_png
