"""``create_scene`` / ``volume_render`` -- yt/visualization/volume_rendering/volume_rendering.py:8-126."""
from .render_source import create_volume_source
from .scene import Scene
from .utils import data_source_or_all


def create_scene(data_source, field=None, lens_type="plane-parallel"):
    """A Scene with one VolumeSource over ``data_source`` (default field, default transfer function)
    and a default Camera with the requested lens."""
    data_source = data_source_or_all(data_source)
    sc = Scene()
    if field is None:
        field = data_source.ds.default_field
        if field not in data_source.ds.derived_field_list:
            raise RuntimeError("Could not find field '%s' in %s. Please specify a field in create_scene()" % (field, data_source.ds))
    sc.add_source(create_volume_source(data_source, field=field))
    sc.add_camera(data_source=data_source, lens_type=lens_type)
    return sc


def volume_render(data_source, field=None, fname=None, sigma_clip=None, lens_type="plane-parallel"):
    sc = create_scene(data_source, field=field, lens_type=lens_type)
    im = sc.render()
    if fname is not None:
        sc.save(fname=fname, sigma_clip=sigma_clip, render=False)
    return im, sc
