"""Scene -- yt/visualization/volume_rendering/scene.py:29-760 (sources, camera, render, composite, save)."""
from collections import OrderedDict

import numpy as np

from .camera import Camera
from .image_array import ImageArray
from .render_source import OpaqueSource
from .zbuffer_array import ZBuffer


class Scene:
    _current = None
    _camera = None

    def __init__(self):
        self.sources = OrderedDict()
        self._last_render = None
        self._sigma_clip = None

    def get_source(self, source_num=0):
        return list(self.sources.values())[source_num]

    def __getitem__(self, item):
        if item in self.sources:
            return self.sources[item]
        return self.get_source(item)

    @property
    def opaque_sources(self):
        for k, source in self.sources.items():
            if isinstance(source, OpaqueSource):
                yield k, source

    @property
    def transparent_sources(self):
        for k, source in self.sources.items():
            if not isinstance(source, OpaqueSource):
                yield k, source

    def add_source(self, render_source, keyname=None):
        if keyname is None:
            keyname = "source_%02d" % len(self.sources)
        self.sources[keyname] = render_source
        return self

    def __setitem__(self, key, value):
        return self.add_source(value, key)

    def render(self, camera=None):
        if camera is None:
            camera = self.camera
        assert camera is not None
        self._validate()
        bmp = self.composite(camera=camera)
        self._last_render = bmp
        return bmp

    def _validate(self):
        for source in self.sources.values():
            source._validate()

    def composite(self, camera=None):
        """Opaque sources fill the z-buffer; transparent sources render from it (scene.py:581-625)."""
        if camera is None:
            camera = self.camera
        empty = camera.lens.new_image(camera)
        opaque = ZBuffer(empty, np.full(empty.shape[:2], np.inf))
        im = empty
        for _, source in self.opaque_sources:
            source.render(camera, zbuffer=opaque)
            im = source.zbuffer.rgba
        for _, source in self.transparent_sources:
            im = source.render(camera, zbuffer=opaque)
            opaque.rgba = im
        # rotate 180 degrees so the orientation agrees with a PlotWindow plot
        return ImageArray(np.rot90(im, k=2), info={"imtype": "rendering"})

    # -- annotations (scene.py:755-880): opaque sources added under fixed keys -------------------------------
    def annotate_domain(self, ds, color=None):
        from .render_source import BoxSource
        self.add_source(BoxSource(ds.domain_left_edge, ds.domain_right_edge, color=color))
        return self

    def annotate_grids(self, data_source, alpha=0.3, cmap=None, min_level=None, max_level=None):
        from .render_source import GridSource
        self.add_source(GridSource(data_source, alpha=alpha, cmap=cmap, min_level=min_level, max_level=max_level))
        return self

    def annotate_axes(self, colors=None, alpha=1.0, *, thickness=1):
        from .render_source import CoordinateVectorSource
        self.add_source(CoordinateVectorSource(colors, alpha, thickness=thickness))
        return self

    def save(self, fname=None, sigma_clip=None, render=True):
        if fname is None:
            fname = "Render_opaque.png"
        if not fname.endswith(".png"):
            fname += ".png"
        if render or self._last_render is None:
            self.render()
        return ImageArray(self._last_render).write_png(fname, sigma_clip=sigma_clip, background="black")

    def add_camera(self, data_source=None, lens_type="plane-parallel", auto=False):
        self._camera = Camera(self, data_source, lens_type, auto)
        return self.camera

    @property
    def camera(self):
        return self._camera

    @camera.setter
    def camera(self, value):
        self._camera = value

    @camera.deleter
    def camera(self):
        self._camera = None

    def arr(self, x, units=None):
        return np.asarray(x, dtype="float64")

    def __repr__(self):
        disp = "<Scene Object>:\nSources: \n"
        for k, v in self.sources.items():
            disp += "    %s: %s\n" % (k, v)
        disp += "Camera: \n    %s" % self.camera
        return disp
