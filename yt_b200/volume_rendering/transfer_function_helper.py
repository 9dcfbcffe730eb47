"""TransferFunctionHelper -- default transfer function of a VolumeSource
(yt/visualization/volume_rendering/transfer_function_helper.py:10-127; plotting omitted)."""
import numpy as np

from .transfer_functions import ColorTransferFunction


class TransferFunctionHelper:
    profiles = None

    def __init__(self, ds):
        self.ds = ds
        self.field = None
        self.log = False
        self.tf = None
        self.bounds = None
        self.grey_opacity = False
        self.profiles = {}

    def set_bounds(self, bounds=None):
        if bounds is None:
            bounds = self.ds.all_data().quantities["Extrema"](self.field, non_zero=True)
            bounds = [np.float64(b) for b in bounds]
        self.bounds = bounds
        assert len(self.bounds) == 2
        if self.log:
            assert self.bounds[0] > 0.0
            assert self.bounds[1] > 0.0

    def set_field(self, field):
        if field != self.field:
            self.log = self.ds._get_field_info(field).take_log
        self.field = field

    def set_log(self, log):
        self.log = log

    def build_transfer_function(self):
        if self.bounds is None:
            self.set_bounds()
        if self.log:
            mi, ma = np.log10(self.bounds[0]), np.log10(self.bounds[1])
        else:
            mi, ma = self.bounds
        self.tf = ColorTransferFunction((mi, ma), grey_opacity=self.grey_opacity, nbins=512)
        return self.tf

    def setup_default(self):
        """10 Gaussian layers sampling 'nipy_spectral', alpha rescaled to a natural contrast."""
        self.tf.add_layers(10, colormap="nipy_spectral")
        factor = self.tf.funcs[-1].y.size / self.tf.funcs[-1].y.sum()
        self.tf.funcs[-1].y *= 2 * factor
