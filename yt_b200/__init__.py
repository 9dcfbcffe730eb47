"""yt_b200 -- B200-native (sm_100a) ray caster behind yt's volume-rendering API.

    import yt_b200 as yt
    ds = yt.load_uniform_grid({"density": arr}, arr.shape)
    sc = yt.create_scene(ds, lens_type="perspective")
    im = sc.render()

Importing this package never touches CUDA; the first sampler / context construction loads
``libytb200.so`` (built by ``__graft_entry__.build()``) and raises if it or a GPU is missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .loaders import load_amr_grids, load_octree_cells, load_uniform_grid  # noqa: F401
from .parallel import enable_parallelism  # noqa: F401
from .volume_rendering import (  # noqa: F401
    ColorTransferFunction,
    Scene,
    create_scene,
    create_volume_source,
    off_axis_projection,
    volume_render,
)
