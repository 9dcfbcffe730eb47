"""``create_scene`` / ``volume_render`` -- the two entry points of
yt/visualization/volume_rendering/volume_rendering.py:8-126, for grid datasets."""
from .render_source import create_volume_source
from .scene import Scene
from .utils import data_source_or_all


def _default_field(ds):
    """The dataset's default field, provided it can actually be evaluated (volume_rendering.py:56-63)."""
    field = ds.default_field
    if field in ds.derived_field_list:
        return field
    raise RuntimeError("Could not find field '%s' in %s. Please specify a field in create_scene()" % (field, ds))


def create_scene(data_source, field=None, lens_type="plane-parallel"):
    """One VolumeSource over ``data_source`` (with the default transfer function of its field) seen by a default
    Camera through the requested lens."""
    source = data_source_or_all(data_source)
    scene = Scene()
    scene.add_source(create_volume_source(source, field=_default_field(source.ds) if field is None else field))
    scene.add_camera(data_source=source, lens_type=lens_type)
    return scene


def volume_render(data_source, field=None, fname=None, sigma_clip=None, lens_type="plane-parallel"):
    """(image, scene) of a default rendering; the image is written to ``fname`` if one is given."""
    scene = create_scene(data_source, field=field, lens_type=lens_type)
    image = scene.render()
    if fname is not None:
        scene.save(fname=fname, sigma_clip=sigma_clip, render=False)
    return image, scene
