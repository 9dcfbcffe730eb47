"""Generate tests/golden/ref_images.npz by running the UNMODIFIED reference samplers.

The reference's own test-suite holds no numeric vectors for the ray caster, so the
golden vectors are outputs of the reference itself: its Cython samplers compiled from
/root/reference into oracle/_ref (oracle/build_ref.py), driven on the scenarios of
tests/common.py exactly as render_source.py:594-599 drives them.  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import common as T  # noqa: E402


def main():
    samplers, grid = O.load_ref()
    out = {}
    for sc in T.scenarios():
        out[sc.name] = T.run_scenario(sc, samplers, grid, num_threads=1)
        print("%-32s max %.6g" % (sc.name, np.nanmax(out[sc.name])))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_images.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
