"""The two acceptance metrics of the path (BASELINE.json north_star), as plain numpy."""
import numpy as np


def normalised_diff(img, ref, grey):
    """Volume-rendered images: rescale both with the REFERENCE image's maxima (RGB: the largest r+g+b of a pixel;
    alpha: the largest alpha, grey-opacity renders only), clip to [0, 1] like the reference's image writer, and
    return the largest absolute difference.  The bar is 1e-4."""
    cmax = ref[:, :, :3].sum(axis=2).max()
    out = 0.0
    if cmax > 0:
        out = np.abs(np.clip(img[:, :, :3] / cmax, 0, 1) - np.clip(ref[:, :, :3] / cmax, 0, 1)).max()
    if grey:
        amax = ref[:, :, 3].max()
        if amax > 0:
            out = max(out, np.abs(np.clip(img[:, :, 3] / amax, 0, 1) - np.clip(ref[:, :, 3] / amax, 0, 1)).max())
    return float(out)


def projection_diff(img, ref):
    """Projection integrals: largest |difference| over the largest |reference| value.  The bar is 1e-6."""
    m = np.abs(ref).max()
    return float(np.abs(img - ref).max() / m) if m > 0 else float(np.abs(img).max())
