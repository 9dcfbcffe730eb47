"""Generate tests/golden/kdtree_golden.npz from the reference's own kd-tree Node class
(yt/utilities/lib/amr_kdtools.pyx compiled unmodified into oracle/_ref).  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_kdtree_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import test_kdtree as T  # noqa: E402


def main():
    out = {}
    for name, ds in T.cases().items():
        nodes, orders = T.reference_dump(ds)
        ids = np.array(sorted(nodes), dtype="int64")
        out[name + "/node_id"] = ids
        out[name + "/grid"] = np.array([nodes[k][0] for k in ids], dtype="int64")
        out[name + "/split_dim"] = np.array([nodes[k][1] for k in ids], dtype="int64")
        for j, o in enumerate(orders):
            out[name + "/order%d" % j] = o
        print(name, len(ids), "nodes", len(orders[0]), "leaves")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kdtree_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
