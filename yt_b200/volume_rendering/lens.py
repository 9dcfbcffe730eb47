"""Lenses: the ray sets handed to the samplers.

Argument construction follows yt/visualization/volume_rendering/lens.py (PlaneParallelLens
83-121, PerspectiveLens 166-240) operation by operation, so the fp64 ``vp_pos`` / ``vp_dir``
arrays equal the reference's for the same camera.  Only these two lenses are named by the
north star; every other lens of the reference produces per-pixel ray arrays too and would run
through the same GENERIC sampler mode.
"""
import numpy as np


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
        self.current_image = np.zeros((camera.resolution[0], camera.resolution[1], 4), dtype="float64", order="C")
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


lenses = {"plane-parallel": PlaneParallelLens, "perspective": PerspectiveLens}
