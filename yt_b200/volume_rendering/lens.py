"""Lenses: the ray sets handed to the samplers.

Argument construction follows yt/visualization/volume_rendering/lens.py (PlaneParallelLens
83-121, PerspectiveLens 166-240) operation by operation, so the fp64 ``vp_pos`` / ``vp_dir``
arrays equal the reference's for the same camera.  Only these two lenses are named by the
north star; the other four (stereo-perspective, fisheye, spherical, stereo-spherical, lens.py:295-844) only
generate per-pixel ray arrays, which the samplers' GENERIC lens mode casts like any others.
"""
import numpy as np


def _image_zeros(shape):
    """A zeroed fp64 image.  Page-locked when the CUDA library is there: the cast then reads the initial colours and
    writes the result in place over PCIe (``ytb200_render``) instead of through two staged copies -- this is an
    allocation choice, not a compute path; pure host uses of the lens mirror get an ordinary array."""
    try:
        from ..device import pinned_full
        return pinned_full(shape, 0.0)
    except Exception:   # noqa: BLE001  (library not built / no device: host-only use of the mirror)
        return np.zeros(shape, dtype="float64", order="C")


class Lens:
    lens_type = None

    def __init__(self):
        self.viewpoint = None
        self.sub_samples = 5
        self.num_threads = 0
        self.box_vectors = None
        self.origin = None
        self.back_center = None
        self.front_center = None
        self.sampler = None
        self.current_image = None

    def set_camera(self, camera):
        self.setup_box_properties(camera)

    def new_image(self, camera):
        """Fresh fp64 C-order (nx, ny, 4) image (lens.py:33-43)."""
        self.current_image = _image_zeros((camera.resolution[0], camera.resolution[1], 4))
        return self.current_image

    def setup_box_properties(self, camera):
        """back/front planes of the view box (lens.py:45-61)."""
        uv, w, c = camera.unit_vectors, camera.width, camera.focus
        self.box_vectors = np.array([uv[0] * w[0], uv[1] * w[1], uv[2] * w[2]])
        self.origin = c - 0.5 * w.dot(uv)
        self.back_center = c - 0.5 * w[2] * uv[2]
        self.front_center = c + 0.5 * w[2] * uv[2]
        self.set_viewpoint(camera)

    def set_viewpoint(self, camera):
        raise NotImplementedError("Need to choose viewpoint for this class")


class PlaneParallelLens(Lens):
    lens_type = "plane-parallel"

    def _get_sampler_params(self, camera, render_source):
        image = render_source.zbuffer.rgba if render_source.zbuffer is not None else self.new_image(camera)
        # packed (1,12,1) ray description: inv_mat column-major + back_center (lens.py:89-94)
        vp_pos = np.concatenate([camera.inv_mat.ravel("F"), self.back_center.ravel()])
        half0, half1 = camera.width[0] / 2.0, camera.width[1] / 2.0
        return {
            "vp_pos": vp_pos,
            "vp_dir": self.box_vectors[2],
            "center": self.back_center,
            "bounds": (-half0, half0, -half1, half1),
            "x_vec": camera.unit_vectors[0],
            "y_vec": camera.unit_vectors[1],
            "width": np.array(camera.width, dtype="float64"),
            "image": image,
            "lens_type": "plane-parallel",
        }

    def set_viewpoint(self, camera):
        # "infinitely far" behind the far plane: kd order == increasing t (lens.py:114-121)
        self.viewpoint = self.front_center + camera.unit_vectors[2] * 1.0e6 * camera.width[2]

    def project_to_plane(self, camera, pos, res=None):
        res = camera.resolution if res is None else res
        dx = np.array(np.dot(pos - self.origin, camera.unit_vectors[1]))
        dy = np.array(np.dot(pos - self.origin, camera.unit_vectors[0]))
        dz = np.array(np.dot(pos - self.front_center, -camera.unit_vectors[2]))
        px = (res[0] * (dy / camera.width[0])).astype("int64")
        py = (res[1] * (dx / camera.width[1])).astype("int64")
        return px, py, dz


class PerspectiveLens(Lens):
    lens_type = "perspective"

    def _get_sampler_params(self, camera, render_source):
        nx, ny = camera.resolution
        camera.width[1] = camera.width[0] * (ny / nx)  # lens.py:168-170
        east, north, normal = camera.unit_vectors
        px = np.linspace(-0.5, 0.5, nx)[np.newaxis, :]
        py = np.linspace(-0.5, 0.5, ny)[np.newaxis, :]
        sample_x = (camera.width[0] * np.array(east.reshape(3, 1) * px)).transpose()
        sample_y = (camera.width[1] * np.array(north.reshape(3, 1) * py)).transpose()
        sample_x = np.repeat(sample_x.reshape(nx, 1, 3), ny, axis=1)
        sample_y = np.repeat(sample_y.reshape(1, ny, 3), nx, axis=0)
        normal_vecs = np.tile(normal, nx * ny).reshape(nx, ny, 3)
        max_length = np.linalg.norm(camera.position - camera._domain_center) + 0.5 * np.linalg.norm(camera._domain_width)
        vectors = (sample_x + sample_y + normal_vecs * camera.width[2]) * (max_length / camera.width[2])
        positions = np.tile(camera.position, nx * ny).reshape(nx, ny, 3)
        uv = np.ones(3, dtype="float64")
        return {
            "vp_pos": positions,
            "vp_dir": np.ascontiguousarray(vectors),
            "center": self.back_center,
            "bounds": (0.0, 1.0, 0.0, 1.0),
            "x_vec": uv,
            "y_vec": uv,
            "width": np.zeros(3, dtype="float64"),
            "image": self.new_image(camera),
            "lens_type": "perspective",
        }

    def set_viewpoint(self, camera):
        self.viewpoint = self.front_center  # lens.py:236-240 -- NOT the camera position

    def project_to_plane(self, camera, pos, res=None):
        """Pinhole projection of points onto the image plane (lens.py:242-288), operation for operation;
        points behind the camera are pushed outside the image instead of being dropped."""
        res = camera.resolution if res is None else res
        width = np.array(camera.width, dtype="float64")
        position = np.array(camera.position, dtype="float64")
        width[1] = width[0] * res[1] / res[0]
        sight_vector = np.array(pos, dtype="float64") - position
        pos1 = sight_vector          # the reference aliases these two arrays (lens.py:251-253)
        for i in range(sight_vector.shape[0]):
            norm = np.sqrt(np.dot(sight_vector[i], sight_vector[i]))
            if norm != 0:
                sight_vector[i] = sight_vector[i] / norm
        sight_center = camera.position + camera.width[2] * camera.unit_vectors[2]
        for i in range(sight_vector.shape[0]):
            sight_angle_cos = np.clip(np.dot(sight_vector[i], camera.unit_vectors[2]), -1.0, 1.0)
            if np.arccos(sight_angle_cos) < 0.5 * np.pi:
                sight_length = width[2] / sight_angle_cos
            else:
                sight_length = np.sqrt(width[0] ** 2 + width[1] ** 2)
                sight_length = sight_length / np.sqrt(1 - sight_angle_cos ** 2)
            pos1[i] = position + sight_length * sight_vector[i]
        dx = np.dot(pos1 - sight_center, camera.unit_vectors[0])
        dy = np.dot(pos1 - sight_center, camera.unit_vectors[1])
        dz = np.dot(np.asarray(pos, dtype="float64") - position, camera.unit_vectors[2])
        px = (res[0] * 0.5 + res[0] / width[0] * dx).astype("int64")
        py = (res[1] * 0.5 + res[1] / width[1] * dy).astype("int64")
        return px, py, dz


def get_rotation_matrix(theta, rot_vector):
    """Rotation by ``theta`` about ``rot_vector`` (yt/utilities/math_utils.py:1181-1207)."""
    ux, uy, uz = rot_vector[0], rot_vector[1], rot_vector[2]
    cost, sint = np.cos(theta), np.sin(theta)
    return np.array([
        [cost + ux ** 2 * (1 - cost), ux * uy * (1 - cost) - uz * sint, ux * uz * (1 - cost) + uy * sint],
        [uy * ux * (1 - cost) + uz * sint, cost + uy ** 2 * (1 - cost), uy * uz * (1 - cost) - ux * sint],
        [uz * ux * (1 - cost) - uy * sint, uz * uy * (1 - cost) + ux * sint, cost + uz ** 2 * (1 - cost)]])


def arr_fisheye_vectors(resolution, fov):
    """Unit vectors of a dome fisheye image (grid_traversal.pyx:322-350 with its default tiling arguments)."""
    fov_rad = fov * np.pi / 180.0
    px = ((2.0 * np.arange(resolution)) / resolution - 1.0)[:, None] * np.ones((1, resolution))
    py = np.ones((resolution, 1)) * ((2.0 * np.arange(resolution)) / resolution - 1.0)[None, :]
    r = np.sqrt(px * px + py * py)
    phi = np.arctan2(py, px)
    theta = r * fov_rad / 2.0
    vp = np.empty((resolution, resolution, 3), dtype="float64")
    vp[:, :, 0] = np.sin(theta) * np.cos(phi)
    vp[:, :, 1] = np.sin(theta) * np.sin(phi)
    vp[:, :, 2] = np.cos(theta)
    vp[r > 1.01] = 0.0
    return vp


def _max_length(camera, extra=0.0):
    return np.linalg.norm(camera.position - camera._domain_center) + 0.5 * np.linalg.norm(camera._domain_width) + extra


class StereoPerspectiveLens(Lens):
    """Two perspective ray sets side by side, one per eye (lens.py:295-518).  Per-pixel rays: GENERIC lens mode."""
    lens_type = "stereo-perspective"

    def __init__(self):
        super().__init__()
        self.disparity = None

    def _get_sampler_params(self, camera, render_source):
        camera.width[1] = camera.width[0] * (2.0 * camera.resolution[1] / camera.resolution[0])
        if self.disparity is None:
            self.disparity = 3.0 * camera.width[0] / camera.resolution[0]
        vl, pl = self._get_positions_vectors(camera, -self.disparity)
        vr, pr = self._get_positions_vectors(camera, self.disparity)
        uv = np.ones(3, dtype="float64")
        res = tuple(camera.resolution)
        image = self.new_image(camera).reshape(*res, 4)
        return {
            "vp_pos": np.ascontiguousarray(np.vstack([pl, pr]).reshape(*res, 3)),
            "vp_dir": np.ascontiguousarray(np.vstack([vl, vr]).reshape(*res, 3)),
            "center": self.back_center, "bounds": (0.0, 1.0, 0.0, 1.0), "x_vec": uv, "y_vec": uv,
            "width": np.zeros(3, dtype="float64"), "image": image, "lens_type": "stereo-perspective",
        }

    def _get_positions_vectors(self, camera, disparity):
        nxs = int(np.floor(camera.resolution[0]) / 2)
        ny = camera.resolution[1]
        east, north, normal = camera.unit_vectors
        R = get_rotation_matrix(-np.arctan2(disparity, camera.width[2]), north)
        east_rot, normal_rot = np.dot(R, east), np.dot(R, normal)
        px = np.linspace(-0.5, 0.5, nxs)[np.newaxis, :]
        py = np.linspace(-0.5, 0.5, ny)[np.newaxis, :]
        sample_x = (camera.width[0] * np.array(east_rot.reshape(3, 1) * px)).transpose()
        sample_y = (camera.width[1] * np.array(north.reshape(3, 1) * py)).transpose()
        sample_x = np.repeat(sample_x.reshape(nxs, 1, 3), ny, axis=1)
        sample_y = np.repeat(sample_y.reshape(1, ny, 3), nxs, axis=0)
        normal_vecs = np.tile(normal_rot, nxs * ny).reshape(nxs, ny, 3)
        east_vecs = np.tile(east_rot, nxs * ny).reshape(nxs, ny, 3)
        max_length = _max_length(camera, np.abs(self.disparity))
        vectors = (sample_x + sample_y + normal_vecs * camera.width[2]) * (max_length / camera.width[2])
        positions = np.tile(camera.position, nxs * ny).reshape(nxs, ny, 3) + east_vecs * disparity
        return vectors, positions

    def project_to_plane(self, camera, pos, res=None):
        res = camera.resolution if res is None else res
        camera.width[1] = camera.width[0] * (2.0 * res[1] / res[0])
        if self.disparity is None:
            self.disparity = 3.0 * camera.width[0] / camera.resolution[0]
        a = self._get_px_py_dz(camera, pos, res, -self.disparity)
        b = self._get_px_py_dz(camera, pos, res, self.disparity)
        return np.vstack([a[0], b[0]]), np.vstack([a[1], b[1]]), np.vstack([a[2], b[2]])

    def _get_px_py_dz(self, camera, pos, res, disparity):
        res0_h = np.floor(res[0]) / 2
        east, north, normal = camera.unit_vectors
        R = get_rotation_matrix(-np.arctan2(disparity, camera.width[2]), north)
        east_rot, normal_rot = np.dot(R, east), np.dot(R, normal)
        shift = np.asarray(camera.position, dtype="float64") + east * disparity
        width = np.array(camera.width, dtype="float64")
        sight_vector = np.array(pos, dtype="float64") - shift
        pos1 = sight_vector
        for i in range(sight_vector.shape[0]):
            sight_vector[i] = sight_vector[i] / np.sqrt(np.dot(sight_vector[i], sight_vector[i]))
        sight_center = shift + camera.width[2] * normal_rot
        for i in range(sight_vector.shape[0]):
            c = np.clip(np.dot(sight_vector[i], normal_rot), -1.0, 1.0)
            if np.arccos(c) < 0.5 * np.pi:
                sight_length = width[2] / c
            else:
                sight_length = np.sqrt(width[0] ** 2 + width[1] ** 2) / np.sqrt(1 - c ** 2)
            pos1[i] = shift + sight_length * sight_vector[i]
        dx = np.dot(pos1 - sight_center, east_rot)
        dy = np.dot(pos1 - sight_center, north)
        dz = np.dot(np.asarray(pos, dtype="float64") - shift, normal_rot)
        if disparity > 0:
            px = (res0_h * 0.5 + res0_h / camera.width[0] * dx).astype("int64")
        else:
            px = (res0_h * 1.5 + res0_h / camera.width[0] * dx).astype("int64")
        py = (res[1] * 0.5 + res[1] / camera.width[1] * dy).astype("int64")
        return px, py, dz

    def set_viewpoint(self, camera):
        self.viewpoint = self.front_center


class FisheyeLens(Lens):
    """Dome projection with a field of view ``fov`` degrees (lens.py:520-633); square image."""
    lens_type = "fisheye"

    def __init__(self):
        super().__init__()
        self.fov = 180.0
        self.radius = 1.0
        self.center = None
        self.rotation_matrix = np.eye(3)

    def setup_box_properties(self, camera):
        self.radius = camera.width.max()
        super().setup_box_properties(camera)
        self.set_viewpoint(camera)

    def new_image(self, camera):
        self.current_image = _image_zeros((camera.resolution[0], camera.resolution[0], 4))
        return self.current_image

    def _get_sampler_params(self, camera, render_source):
        shape = (camera.resolution[0], camera.resolution[0], 3)
        vp = -arr_fisheye_vectors(camera.resolution[0], self.fov).reshape(shape)
        vp = vp.dot(np.linalg.inv(self.rotation_matrix))
        vp *= self.radius
        uv = np.ones(3, dtype="float64")
        positions = np.full(shape, camera.position, dtype="float64")
        image = render_source.zbuffer.rgba if render_source.zbuffer is not None else self.new_image(camera)
        return {"vp_pos": positions, "vp_dir": np.ascontiguousarray(vp), "center": self.center, "bounds": (0.0, 1.0, 0.0, 1.0),
                "x_vec": uv, "y_vec": uv, "width": np.zeros(3, dtype="float64"), "image": image, "lens_type": "fisheye"}

    def set_viewpoint(self, camera):
        self.viewpoint = camera.position

    def project_to_plane(self, camera, pos, res=None):
        res = camera.resolution if res is None else res
        lpos = np.asarray(camera.position, dtype="float64") - np.asarray(pos, dtype="float64")
        lpos = lpos.dot(self.rotation_matrix)
        mag = (lpos * lpos).sum(axis=1) ** 0.5
        mag[mag == 0] = 1
        lpos /= mag[:, None]
        dz = mag / self.radius
        theta = np.arccos(lpos[:, 2])
        r = 2.0 * theta / (self.fov * np.pi / 180.0)
        phi = np.arctan2(lpos[:, 1], lpos[:, 0])
        px = (r * np.cos(phi) + 1.0) * res[0] / 2.0
        py = (r * np.sin(phi) + 1.0) * res[1] / 2.0
        return np.rint(px).astype("int64"), np.rint(py).astype("int64"), dz


def _spherical_frame(camera):
    R1 = get_rotation_matrix(0.5 * np.pi, [1, 0, 0])
    R2 = get_rotation_matrix(0.5 * np.pi, [0, 0, 1])
    return np.dot(R2, np.dot(R1, camera.unit_vectors))


class SphericalLens(Lens):
    """Equirectangular 360 x 180 degree projection (lens.py:635-729)."""
    lens_type = "spherical"

    def __init__(self):
        super().__init__()
        self.radius = 1.0
        self.center = None
        self.rotation_matrix = np.eye(3)

    def setup_box_properties(self, camera):
        self.radius = camera.width.max()
        super().setup_box_properties(camera)
        self.set_viewpoint(camera)

    def _get_sampler_params(self, camera, render_source):
        res = tuple(camera.resolution)
        px = np.linspace(-np.pi, np.pi, res[0], endpoint=True)[:, None]
        py = np.linspace(-np.pi / 2.0, np.pi / 2.0, res[1], endpoint=True)[None, :]
        positions = np.tile(camera.position, int(np.prod(res))).reshape(*res, 3)
        uv = _spherical_frame(camera)
        vectors = np.empty((*res, 3), dtype="float64", order="C")
        vectors[:, :, 0] = np.cos(px) * np.cos(py)
        vectors[:, :, 1] = np.sin(px) * np.cos(py)
        vectors[:, :, 2] = np.sin(py)
        vectors = np.dot(vectors * _max_length(camera), uv).reshape(*res, 3)
        image = render_source.zbuffer.rgba if render_source.zbuffer is not None else self.new_image(camera)
        dummy = np.ones(3, dtype="float64")
        return {"vp_pos": positions, "vp_dir": np.ascontiguousarray(vectors), "center": self.back_center,
                "bounds": (0.0, 1.0, 0.0, 1.0), "x_vec": dummy, "y_vec": dummy, "width": np.zeros(3, dtype="float64"),
                "image": image.reshape(*res, 4), "lens_type": "spherical"}

    def set_viewpoint(self, camera):
        self.viewpoint = camera.position

    def project_to_plane(self, camera, pos, res=None):
        res = camera.resolution if res is None else res
        lpos = np.asarray(camera.position, dtype="float64") - np.asarray(pos, dtype="float64")
        mag = (lpos * lpos).sum(axis=1) ** 0.5
        mag[mag == 0] = 1
        lpos /= mag[:, None]
        px = np.arctan2(lpos[:, 1], lpos[:, 0])
        py = np.arcsin(lpos[:, 2])
        dz = mag / self.radius
        px = ((-px + np.pi) / (2.0 * np.pi)) * res[0]
        py = ((-py + np.pi / 2.0) / np.pi) * res[1]
        return np.rint(px).astype("int64"), np.rint(py).astype("int64"), dz


class StereoSphericalLens(Lens):
    """Two equirectangular ray sets stacked along the image's second axis, one per eye (lens.py:731-844)."""
    lens_type = "stereo-spherical"

    def __init__(self):
        super().__init__()
        self.radius = 1.0
        self.center = None
        self.disparity = None
        self.rotation_matrix = np.eye(3)

    def setup_box_properties(self, camera):
        self.radius = camera.width.max()
        super().setup_box_properties(camera)
        self.set_viewpoint(camera)

    def _get_sampler_params(self, camera, render_source):
        if self.disparity is None:
            self.disparity = camera.width[0] / 1000.0
        res = tuple(camera.resolution)
        nys = int(np.floor(res[1]) / 2)
        px = np.linspace(-np.pi, np.pi, res[0], endpoint=True)[:, None]
        py = np.linspace(-np.pi / 2.0, np.pi / 2.0, nys, endpoint=True)[None, :]
        vectors = np.zeros((res[0], nys, 3), dtype="float64", order="C")
        vectors[:, :, 0] = np.cos(px) * np.cos(py)
        vectors[:, :, 1] = np.sin(px) * np.cos(py)
        vectors[:, :, 2] = np.sin(py)
        vectors = vectors * _max_length(camera, np.abs(self.disparity))
        uv = _spherical_frame(camera)
        vectors = np.dot(vectors, uv)
        vectors2 = np.zeros((res[0], nys, 3), dtype="float64", order="C")
        vectors2[:, :, 0] = -np.sin(px) * np.ones((1, nys))
        vectors2[:, :, 1] = np.cos(px) * np.ones((1, nys))
        vectors2 = np.dot(vectors2, uv)
        positions = np.tile(camera.position, res[0] * nys).reshape(res[0], nys, 3)
        positions_left = positions + vectors2 * self.disparity
        positions_right = positions + vectors2 * (-self.disparity)
        image = render_source.zbuffer.rgba if render_source.zbuffer is not None else self.new_image(camera)
        dummy = np.ones(3, dtype="float64")
        return {"vp_pos": np.ascontiguousarray(np.hstack([positions_left, positions_right]).reshape(*res, 3)),
                "vp_dir": np.ascontiguousarray(np.hstack([vectors, vectors]).reshape(*res, 3)),
                "center": self.back_center, "bounds": (0.0, 1.0, 0.0, 1.0), "x_vec": dummy, "y_vec": dummy,
                "width": np.zeros(3, dtype="float64"), "image": image.reshape(*res, 4), "lens_type": "stereo-spherical"}

    def set_viewpoint(self, camera):
        self.viewpoint = camera.position


lenses = {"plane-parallel": PlaneParallelLens, "perspective": PerspectiveLens, "stereo-perspective": StereoPerspectiveLens,
          "fisheye": FisheyeLens, "spherical": SphericalLens, "stereo-spherical": StereoSphericalLens}
