"""Scene -- the container a user drives: sources + camera -> image.

Interface of yt/visualization/volume_rendering/scene.py:29-880 restricted to what reaches the ray caster:
``add_source / get_source / [] / opaque_sources / transparent_sources``, ``add_camera / camera``,
``render / composite / save`` and the ``annotate_*`` helpers.  Compositing order and the final 180-degree rotation
are the reference's (scene.py:611-625); all rendering work happens inside the sources.
"""
import numpy as np

from .camera import Camera
from .image_array import ImageArray
from .render_source import OpaqueSource
from .zbuffer_array import ZBuffer


class Scene:
    _current = None
    _camera = None

    def __init__(self):
        self.sources = {}              # insertion-ordered: the order sources are rendered in
        self._last_render = None
        self._sigma_clip = None

    # -- sources ----------------------------------------------------------------------------------------
    def add_source(self, render_source, keyname=None):
        key = "source_%02d" % len(self.sources) if keyname is None else keyname
        self.sources[key] = render_source
        return self

    __setitem__ = lambda self, key, value: self.add_source(value, key)   # noqa: E731

    def get_source(self, source_num=0):
        return tuple(self.sources.values())[source_num]

    def __getitem__(self, item):
        return self.sources[item] if item in self.sources else self.get_source(item)

    def _sources_where(self, opaque):
        return ((k, s) for k, s in self.sources.items() if isinstance(s, OpaqueSource) == opaque)

    @property
    def opaque_sources(self):
        return self._sources_where(True)

    @property
    def transparent_sources(self):
        return self._sources_where(False)

    # -- camera -----------------------------------------------------------------------------------------
    def add_camera(self, data_source=None, lens_type="plane-parallel", auto=False):
        self._camera = Camera(self, data_source, lens_type, auto)
        return self._camera

    @property
    def camera(self):
        return self._camera

    @camera.setter
    def camera(self, value):
        self._camera = value

    @camera.deleter
    def camera(self):
        self._camera = None

    def set_camera(self, camera):
        self.camera = camera

    def get_camera(self):
        return self.camera

    # -- rendering --------------------------------------------------------------------------------------
    def _validate(self):
        for source in self.sources.values():
            source._validate()

    def render(self, camera=None):
        camera = self.camera if camera is None else camera
        assert camera is not None
        self._validate()
        self._last_render = self.composite(camera=camera)
        return self._last_render

    def composite(self, camera=None):
        """Opaque sources paint one shared ZBuffer, in order; every transparent source is then cast from that
        buffer (initial colours + per-pixel stopping depth) and its result becomes the buffer's colours for the
        next one (scene.py:611-622).  The image is rotated by 180 degrees to agree with a PlotWindow plot."""
        camera = self.camera if camera is None else camera
        canvas = camera.lens.new_image(camera)
        from .utils import _filled
        depth = ZBuffer(canvas, _filled(canvas.shape[:2], np.inf))
        image = canvas
        for _, source in self.opaque_sources:
            source.render(camera, zbuffer=depth)
            image = source.zbuffer.rgba
        for _, source in self.transparent_sources:
            image = source.render(camera, zbuffer=depth)
            depth.rgba = image
        return ImageArray(np.rot90(image, k=2), info={"imtype": "rendering"})

    def save(self, fname=None, sigma_clip=None, render=True):
        fname = "Render_opaque.png" if fname is None else fname
        if not fname.endswith(".png"):
            fname += ".png"
        if render or self._last_render is None:
            self.render()
        return ImageArray(self._last_render).write_png(fname, sigma_clip=sigma_clip, background="black")

    # -- annotations: opaque sources with the reference's defaults (scene.py:755-880) --------------------------
    def annotate_domain(self, ds, color=None):
        from .render_source import BoxSource
        return self.add_source(BoxSource(ds.domain_left_edge, ds.domain_right_edge, color=color))

    def annotate_grids(self, data_source, alpha=0.3, cmap=None, min_level=None, max_level=None):
        from .render_source import GridSource
        return self.add_source(GridSource(data_source, alpha=alpha, cmap=cmap, min_level=min_level, max_level=max_level))

    def annotate_axes(self, colors=None, alpha=1.0, *, thickness=1):
        from .render_source import CoordinateVectorSource
        return self.add_source(CoordinateVectorSource(colors, alpha, thickness=thickness))

    # -- misc -------------------------------------------------------------------------------------------
    def arr(self, x, units=None):
        return np.asarray(x, dtype="float64")

    def __repr__(self):
        listing = "".join("    %s: %s\n" % kv for kv in self.sources.items())
        return "<Scene Object>:\nSources: \n%sCamera: \n    %s" % (listing, self.camera)
