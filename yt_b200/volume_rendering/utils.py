"""Sampler factories -- yt/visualization/volume_rendering/utils.py:55-166, bound to the CUDA samplers.

This is the one place a yt maintainer would touch to switch the reference onto libytb200
(see INTEGRATION.md): the three ``new_*_sampler`` functions build the very same argument
tuples and hand them to :mod:`yt_b200.image_samplers` instead of ``yt.utilities.lib.image_samplers``.
"""
import numpy as np

from ..image_samplers import InterpolatedProjectionSampler, ProjectionSampler, VolumeRenderSampler


def data_source_or_all(data_source):
    if hasattr(data_source, "all_data") and not hasattr(data_source, "blocks"):
        data_source = data_source.all_data()
    if data_source is not None and not hasattr(data_source, "blocks"):
        raise RuntimeError("The data_source is not a valid 3D data container.\n"
                           "Expected an object of type YTSelectionContainer3D but received "
                           "an object of type %s." % type(data_source))
    return data_source


def ensure_code_unit_params(params):
    for name in ["center", "vp_pos", "vp_dir", "width"]:
        p = params[name]
        if hasattr(p, "in_units"):
            params[name] = p.in_units("code_length")
    bounds = params["bounds"]
    if hasattr(bounds[0], "units"):
        params["bounds"] = tuple(b.in_units("code_length").d for b in bounds)
    return params


def _common_args(params, render_source):
    return [np.atleast_3d(np.array(params["vp_pos"], dtype="float64")), np.atleast_3d(np.array(params["vp_dir"], dtype="float64")),
            None if params["center"] is None else np.array(params["center"], dtype="float64"), params["bounds"], params["image"],
            np.array(params["x_vec"], dtype="float64"), np.array(params["y_vec"], dtype="float64"),
            np.array(params["width"], dtype="float64"), render_source.volume_method]


def _filled(shape, value):
    """A constant fp64 array, page-locked when the CUDA library is there (see lens._image_zeros)."""
    try:
        from ..device import pinned_full
        return pinned_full(shape, value)
    except Exception:   # noqa: BLE001
        return np.full(shape, value, dtype="float64")


def _ones(shape):
    """The default z-buffer (image_samplers.pyx:78-80)."""
    return _filled(shape, 1.0)


def _zbuffer_kwarg(params, render_source):
    if render_source.zbuffer is not None:
        return render_source.zbuffer.z
    return _ones(params["image"].shape[:2])


def new_volume_render_sampler(camera, render_source):
    params = ensure_code_unit_params(camera._get_sampler_params(render_source))
    args = _common_args(params, render_source) + [render_source.transfer_function, render_source.num_samples]
    kwargs = {"lens_type": params["lens_type"]}
    if "camera_data" in params:
        kwargs["camera_data"] = params["camera_data"]
    if render_source.zbuffer is not None:
        kwargs["zbuffer"] = render_source.zbuffer.z
        args[4][:] = np.reshape(render_source.zbuffer.rgba[:], (camera.resolution[0], camera.resolution[1], 4))
    else:
        kwargs["zbuffer"] = _ones(params["image"].shape[:2])
    return VolumeRenderSampler(*args, **kwargs)


def new_interpolated_projection_sampler(camera, render_source):
    params = ensure_code_unit_params(camera._get_sampler_params(render_source))
    args = _common_args(params, render_source) + [render_source.num_samples]
    return InterpolatedProjectionSampler(*args, lens_type=params["lens_type"], zbuffer=_zbuffer_kwarg(params, render_source))


def new_projection_sampler(camera, render_source):
    params = ensure_code_unit_params(camera._get_sampler_params(render_source))
    args = _common_args(params, render_source) + [render_source.num_samples]
    return ProjectionSampler(*args, lens_type=params["lens_type"], zbuffer=_zbuffer_kwarg(params, render_source))


def sampler_arguments(camera, zbuffer=None, image=None, volume_method="KDTree"):
    """(args, kwargs) of ``ImageSampler.__init__`` for `camera` exactly as the factories above assemble them
    (utils.py:55-139), without a render source: ``args`` stops at ``volume_method`` (append the transfer function /
    ``num_samples`` the sampler class takes), ``kwargs`` holds ``lens_type``, ``camera_data`` (only
    ``new_volume_render_sampler`` forwards it -- pop it for the projection samplers) and ``zbuffer``.
    Used by the benchmark and the parity tests to drive the CUDA samplers and the CPU implementations they are checked against
    with one argument tuple."""
    import types
    p = ensure_code_unit_params(camera._get_sampler_params(types.SimpleNamespace(zbuffer=None)))
    nx, ny = camera.resolution
    img = np.zeros((nx, ny, 4), dtype="float64") if image is None else image
    p["image"] = img
    args = _common_args(p, types.SimpleNamespace(volume_method=volume_method))
    kwargs = {"lens_type": p["lens_type"], "camera_data": np.array(p["camera_data"], dtype="float64"),
              "zbuffer": np.ones((nx, ny), "float64") if zbuffer is None else zbuffer}
    return args, kwargs


def get_corners(le, re):
    return np.array([[le[0], le[1], le[2]], [re[0], le[1], le[2]], [re[0], re[1], le[2]], [le[0], re[1], le[2]],
                     [le[0], le[1], re[2]], [re[0], le[1], re[2]], [re[0], re[1], re[2]], [le[0], re[1], re[2]]], dtype="float64")
