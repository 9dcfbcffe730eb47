"""TransferFunctionHelper -- chooses the default transfer function of a VolumeSource.

Interface of yt/visualization/volume_rendering/transfer_function_helper.py:10-127 (``set_field``, ``set_log``,
``set_bounds``, ``build_transfer_function``, ``setup_default`` and the ``field / log / bounds / tf / grey_opacity``
attributes the volume source reads); the profile plotting of the reference is not part of the ray-casting path.
"""
import numpy as np

from .transfer_functions import ColorTransferFunction

DEFAULT_BINS = 512          # transfer_function_helper.py:112-114
DEFAULT_LAYERS = 10         # :124
DEFAULT_COLORMAP = "nipy_spectral"


class TransferFunctionHelper:
    profiles = None

    def __init__(self, ds):
        self.ds = ds
        self.profiles = {}
        self.tf = None
        self.field, self.log = None, False
        self.bounds = None
        self.grey_opacity = False

    # -- what is rendered -----------------------------------------------------------------------------------
    def set_field(self, field):
        """Switching to another field adopts that field's default log/linear choice."""
        changed = field != self.field
        self.field = field
        if changed:
            self.log = self.ds._get_field_info(field).take_log

    def set_log(self, log):
        self.log = log

    def set_bounds(self, bounds=None):
        """``bounds`` are raw field values (not logarithms); by default the non-zero extrema of the whole dataset."""
        if bounds is None:
            lo, hi = self.ds.all_data().quantities["Extrema"](self.field, non_zero=True)
            bounds = [np.float64(lo), np.float64(hi)]
        if len(bounds) != 2:
            raise AssertionError("bounds must be [min, max]")
        if self.log and not (bounds[0] > 0.0 and bounds[1] > 0.0):
            raise AssertionError("a logarithmic transfer function needs positive bounds")
        self.bounds = bounds

    # -- the table ------------------------------------------------------------------------------------------
    def _table_range(self):
        if self.bounds is None:
            self.set_bounds()
        lo, hi = self.bounds
        return (np.log10(lo), np.log10(hi)) if self.log else (lo, hi)

    def build_transfer_function(self):
        self.tf = ColorTransferFunction(self._table_range(), grey_opacity=self.grey_opacity, nbins=DEFAULT_BINS)
        return self.tf

    def setup_default(self):
        """Ten Gaussian layers along the colour map, then the alpha table rescaled so that its mean is 2 -- the
        reference's "natural contrast" normalisation (transfer_function_helper.py:124-127)."""
        self.tf.add_layers(DEFAULT_LAYERS, colormap=DEFAULT_COLORMAP)
        alpha = self.tf.funcs[-1].y
        alpha *= 2.0 * alpha.size / alpha.sum()
