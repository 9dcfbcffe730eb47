"""yt_b200 -- B200-native (sm_100a) ray caster behind yt's volume-rendering sampler API.

Importing this package never touches CUDA; the first sampler / context construction loads
``libytb200.so`` (built by ``__graft_entry__.build()``) and raises if it or a GPU is missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
