"""Camera -- mirror of yt/visualization/volume_rendering/camera.py:47-740 without units.

Defaults (camera.py:127-137): position = domain right edge, focus = domain centre,
width = 1.5 * max(domain width) on all three axes, resolution 512 x 512.
"""
import numpy as np

from ..orientation import Orientation
from .lens import Lens, lenses


def _vec3(v):
    a = np.array(v, dtype="float64")
    if a.ndim == 0:
        a = np.array([float(a)] * 3)
    if a.shape != (3,):
        raise RuntimeError("camera properties need a scalar or 3 components")
    return a


def _rotation_matrix(theta, u):
    """Rodrigues rotation about unit vector u (yt/utilities/math_utils.py get_rotation_matrix)."""
    ux, uy, uz = u
    c, s = np.cos(theta), np.sin(theta)
    return np.array([
        [c + ux * ux * (1 - c), ux * uy * (1 - c) - uz * s, ux * uz * (1 - c) + uy * s],
        [uy * ux * (1 - c) + uz * s, c + uy * uy * (1 - c), uy * uz * (1 - c) - ux * s],
        [uz * ux * (1 - c) - uy * s, uz * uy * (1 - c) + ux * s, c + uz * uz * (1 - c)],
    ])


class Camera(Orientation):
    _moved = True

    def __init__(self, scene, data_source=None, lens_type="plane-parallel", auto=False):
        self.scene = scene
        self.lens = None
        self.north_vector = None
        self.normal_vector = None
        self.light = None
        self.data_source = data_source.all_data() if hasattr(data_source, "all_data") else data_source
        self._resolution = (512, 512)
        if self.data_source is not None:
            ds = self.data_source.ds
            self._focus = np.array(ds.domain_center, dtype="float64")
            self._position = np.array(ds.domain_right_edge, dtype="float64")
            self._width = np.array([1.5 * ds.domain_width.max()] * 3, dtype="float64")
            self._domain_center = np.array(ds.domain_center, dtype="float64")
            self._domain_width = np.array(ds.domain_width, dtype="float64")
        else:
            self._focus = np.zeros(3)
            self._width = np.ones(3)
            self._position = np.ones(3)
            self._domain_center = np.zeros(3)
            self._domain_width = np.ones(3)
        super().__init__(self._focus - self._position, self.north_vector, steady_north=False)
        self.set_lens(lens_type)
        if auto:
            if self.data_source is None:
                raise RuntimeError("auto=True needs a data source")
            self.set_defaults_from_data_source(self.data_source)

    # -- properties with the reference's side effects -------------------------------------------------
    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, value):
        p = _vec3(value)
        if np.array_equal(p, self._focus):
            raise RuntimeError("Cannot set the camera focus and position to the same value")
        self._position = p
        self.switch_orientation(normal_vector=self._focus - self._position, north_vector=self.north_vector)

    @property
    def focus(self):
        return self._focus

    @focus.setter
    def focus(self, value):
        f = _vec3(value)
        if np.array_equal(f, self._position):
            raise RuntimeError("Cannot set the camera focus and position to the same value")
        self._focus = f
        self.switch_orientation(normal_vector=self._focus - self._position, north_vector=None)

    @property
    def width(self):
        return self._width

    @width.setter
    def width(self, value):
        self._width = _vec3(value)
        self.switch_orientation()

    @property
    def resolution(self):
        return self._resolution

    @resolution.setter
    def resolution(self, value):
        if np.ndim(value) == 0:
            value = (int(value), int(value))
        elif len(value) != 2:
            raise RuntimeError
        self._resolution = (int(value[0]), int(value[1]))

    def set_resolution(self, resolution):
        self.resolution = resolution

    def get_resolution(self):
        return self.resolution

    def set_width(self, width):
        self.width = width

    def get_width(self):
        return self.width

    def set_position(self, position, north_vector=None):
        if north_vector is not None:
            self.north_vector = np.array(north_vector, dtype="float64")
        self.position = position

    def get_position(self):
        return self.position

    def set_focus(self, new_focus):
        self.focus = new_focus

    def get_focus(self):
        return self.focus

    # -- sampler arguments (camera.py:275-281) ------------------------------------------------------------
    def _get_sampler_params(self, render_source):
        params = self.lens._get_sampler_params(self, render_source)
        params.update(width=np.array(self.width, dtype="float64"))
        params.update(camera_data=np.vstack((self.position, self.width, self.unit_vectors)))
        return params

    def set_lens(self, lens_type):
        if isinstance(lens_type, Lens):
            self.lens = lens_type
        elif lens_type not in lenses:
            raise RuntimeError(
                "Lens type %s not in available list of available lens types (%s)"
                % (lens_type, ", ".join(repr(k) for k in lenses)))
        else:
            self.lens = lenses[lens_type]()
        self.lens.set_camera(self)

    def switch_orientation(self, normal_vector=None, north_vector=None):
        if north_vector is None:
            north_vector = self.north_vector
        if normal_vector is None:
            normal_vector = self.normal_vector
        self._setup_normalized_vectors(normal_vector, north_vector)
        if self.lens is not None:
            self.lens.setup_box_properties(self)

    def switch_view(self, normal_vector=None, north_vector=None):
        self.switch_orientation(normal_vector=normal_vector, north_vector=north_vector)
        self._moved = True

    def rotate(self, theta, rot_vector=None, rot_center=None):
        """camera.py:451-512"""
        rotate_all = rot_vector is not None
        if rot_vector is None:
            rot_vector = self.north_vector
        if rot_center is None:
            rot_center = self._position
        rot_vector = np.array(rot_vector, dtype="float64")
        rot_vector = rot_vector / np.linalg.norm(rot_vector)
        rot_center = np.array(rot_center, dtype="float64")
        R = _rotation_matrix(theta, rot_vector)
        new_position = np.dot(R, self._position - rot_center) + rot_center
        if (new_position == self._position).all():
            normal_vector = self.unit_vectors[2]
        else:
            normal_vector = rot_center - new_position
        normal_vector = normal_vector / np.sqrt((normal_vector ** 2).sum())
        if rotate_all:
            self.switch_view(normal_vector=np.dot(R, normal_vector), north_vector=np.dot(R, self.unit_vectors[1]))
        else:
            self.switch_view(normal_vector=np.dot(R, normal_vector))
        if (new_position != self._position).any():
            self.set_position(new_position)

    def pitch(self, theta, rot_center=None):
        self.rotate(theta, rot_vector=self.unit_vectors[0], rot_center=rot_center)

    def yaw(self, theta, rot_center=None):
        self.rotate(theta, rot_vector=self.unit_vectors[1], rot_center=rot_center)

    def roll(self, theta, rot_center=None):
        self.rotate(theta, rot_vector=self.unit_vectors[2], rot_center=rot_center)

    def zoom(self, factor):
        self._width[:2] = self._width[:2] / factor
        self.switch_orientation()

    # -- sequences of views (camera.py:598-760): generators that move the camera one step per iteration ---
    def iter_rotate(self, theta, n_steps, rot_vector=None, rot_center=None):
        dtheta = (1.0 * theta) / n_steps
        for i in range(n_steps):
            self.rotate(dtheta, rot_vector=rot_vector, rot_center=rot_center)
            yield i

    def iter_move(self, final, n_steps, exponential=False):
        """Towards ``final`` in ``n_steps`` equal displacements, or equal ratios per component if ``exponential``."""
        final = _vec3(final)
        if exponential:
            dx = ((final / self._position) * 1.0) ** (1.0 / n_steps)
        else:
            dx = (final - self._position) * 1.0 / n_steps
        for i in range(n_steps):
            self.set_position(self._position * dx if exponential else self._position + dx)
            yield i

    def iter_zoom(self, final, n_steps):
        f = final ** (1.0 / n_steps)
        for i in range(n_steps):
            self.zoom(f)
            yield i

    def set_defaults_from_data_source(self, data_source):
        """camera.py:310-346 (what ``auto=True`` does): look at the data source's centre from the domain's right edge,
        with a field of view as wide as the diagonal of the box of its cell centres."""
        ds = data_source.ds
        grids = [g for g, _ in data_source.blocks]
        lo = np.min([g.LeftEdge + 0.5 * g.dds for g in grids], axis=0)
        hi = np.max([g.RightEdge - 0.5 * g.dds for g in grids], axis=0)
        width = np.sqrt(((hi - lo) ** 2).sum())
        self._width = np.array([width] * 3, dtype="float64")
        self._focus = np.array(ds.domain_center, dtype="float64")
        self._position = np.array(ds.domain_right_edge, dtype="float64")
        self._domain_center = np.array(ds.domain_center, dtype="float64")
        self._domain_width = np.array(ds.domain_width, dtype="float64")
        Orientation.__init__(self, self._focus - self._position, self.north_vector, steady_north=False)
        if self.lens is not None:
            self.lens.setup_box_properties(self)
        self._moved = True

    def __repr__(self):
        return ("<Camera Object>:\n\tposition:%s\n\tfocus:%s\n\tnorth_vector:%s\n\twidth:%s\n\tresolution:%s\n"
                % (self.position, self.focus, self.north_vector, self.width, self.resolution))
