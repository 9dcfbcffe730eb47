"""Scene-graph mirror of yt/visualization/volume_rendering (numpy only, code_length only).

Only what is needed to build sampler arguments and loop bricks is restated; the compute goes
through :mod:`yt_b200.image_samplers` into libytb200.so.
"""
from .camera import Camera  # noqa: F401
from .image_array import ImageArray  # noqa: F401
from .lens import PerspectiveLens, PlaneParallelLens, lenses  # noqa: F401
from .off_axis_projection import off_axis_projection  # noqa: F401
from .render_source import (  # noqa: F401
    BoxSource,
    CoordinateVectorSource,
    GridSource,
    KDTreeVolumeSource,
    LineSource,
    OpaqueSource,
    PointSource,
    VolumeSource,
    create_volume_source,
)
from .scene import Scene  # noqa: F401
from .transfer_function_helper import TransferFunctionHelper  # noqa: F401
from .transfer_functions import (  # noqa: F401
    ColorTransferFunction,
    MultiVariateTransferFunction,
    ProjectionTransferFunction,
    TransferFunction,
)
from .volume_rendering import create_scene, volume_render  # noqa: F401
from .zbuffer_array import ZBuffer  # noqa: F401
