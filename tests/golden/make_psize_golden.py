"""Golden processor grids from the reference's own decompose.get_psize (yt/utilities/decompose.py:69-105).
Run in the build container (needs /root/reference):  python tests/golden/make_psize_golden.py"""
import importlib.util
import json
import os

spec = importlib.util.spec_from_file_location("ref_decompose", "/root/reference/yt/utilities/decompose.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
import numpy as np  # noqa: E402

SHAPES = [(32, 32, 32), (64, 32, 16), (16, 64, 32), (10, 12, 14), (128, 128, 1), (7, 50, 50), (1024, 1024, 1024), (33, 17, 65), (2, 2, 2), (3, 100, 5)]
PIECES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 18, 24, 27, 30, 32, 36, 48, 60, 64, 100, 128, 210, 512]
out = []
for sh in SHAPES:
    for p in PIECES:
        ps = ref.get_psize(np.array(sh), p)
        le, re, shapes, slices, _ = ref.split_array(np.zeros(3), np.ones(3), np.array(sh), ps)
        out.append({"shape": list(sh), "pieces": p, "psize": [int(x) for x in ps],
                    "first_dims": [int(x) for x in shapes[0]], "last_dims": [int(x) for x in shapes[-1]]})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "psize_golden.json"), "w"), indent=0)
print(len(out), "cases")
