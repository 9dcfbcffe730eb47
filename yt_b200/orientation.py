"""Camera basis (east, north, normal) from a view direction.

Same arithmetic as the reference's ``Orientation`` (yt/utilities/orientation.py:27-85), so that
the sampler arguments built from it are bit-identical for the same inputs; unit handling is
dropped (everything is ``code_length`` by the time it reaches a sampler, utils.py:158-166).
"""
import numpy as np


def _unit(v):
    return v / np.sqrt(np.dot(v, v))


def camera_basis(normal, north=None, steady_north=False):
    """Return the 3x3 array whose rows are east, north, normal.

    Without ``north`` the east axis is ``e_k x normal`` where ``k`` maximises the component sum of
    ``normal x e_k`` (orientation.py:69-72); with ``north`` its component along ``normal`` is removed
    first when ``steady_north`` or when it is not already perpendicular (orientation.py:78-82).
    """
    n = np.array(normal, dtype="float64")
    if not np.dot(n, n) > 0:
        raise ValueError("normal_vector cannot be the zero vector.")
    n = _unit(n)
    if north is None:
        axes = np.identity(3)
        k = np.cross(n, axes).sum(axis=1).argmax()
        east = np.cross(axes[k, :], n).ravel()
        up = np.cross(n, east).ravel()
    else:
        up = np.array(north, dtype="float64")
        cosang = np.abs(np.dot(up, n) / np.linalg.norm(up) / np.linalg.norm(n))
        if np.isclose(cosang, 1.0, 1.0e-13):
            raise ValueError("normal_vector and north_vector cannot be aligned.")
        if steady_north or np.dot(up, n) != 0.0:
            up = up - np.dot(up, n) * n
        east = np.cross(up, n).ravel()
    return np.array([_unit(east), _unit(up), n])


class Orientation:
    """Holds ``unit_vectors`` (rows east/north/normal), ``normal_vector``, ``north_vector``, ``inv_mat``."""

    def __init__(self, normal_vector, north_vector=None, steady_north=False):
        self.steady_north = bool(steady_north) or north_vector is not None
        self.north_vector = None if north_vector is None else np.array(north_vector, dtype="float64")
        self._setup_normalized_vectors(normal_vector, self.north_vector)
        if self.north_vector is None:
            self.north_vector = self.unit_vectors[1]

    def _setup_normalized_vectors(self, normal_vector, north_vector):
        U = camera_basis(normal_vector, north_vector, self.steady_north)
        self.unit_vectors = U
        self.normal_vector = U[2]
        self.north_vector = U[1]
        self.inv_mat = np.linalg.pinv(U)  # orientation.py:85

    def switch_orientation(self, normal_vector=None, north_vector=None):
        if north_vector is None:
            north_vector = self.north_vector
        if normal_vector is None:
            normal_vector = self.normal_vector
        self._setup_normalized_vectors(normal_vector, north_vector)
