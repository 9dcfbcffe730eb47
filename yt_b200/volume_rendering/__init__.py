"""Scene-graph mirror of yt/visualization/volume_rendering (numpy only, code_length only).

Only what is needed to build sampler arguments and loop bricks is restated; the compute goes
through :mod:`yt_b200.image_samplers` into libytb200.so.
"""
from .camera import Camera  # noqa: F401
from .lens import PerspectiveLens, PlaneParallelLens, lenses  # noqa: F401
from .transfer_functions import (  # noqa: F401
    ColorTransferFunction,
    MultiVariateTransferFunction,
    ProjectionTransferFunction,
    TransferFunction,
)
