#!/usr/bin/env python
"""Oracle-vs-shipped-fixture residual under each documented restatement choice, toggled one at a time.

For bags 1, 2, 3, 5, 6 (click = fixture centroid, the author's key presses from tests/golden/fixtures.json) prints the
max corner distance (mm, file order) between the oracle's corners and process_data/pointgrey_lidar_<i>.txt for:
the shipped restatement, each single switch of ilcc2o_set_variant, extra 'd' presses, and +-k ulp perturbations of
the subspace-dogleg boundary solution (sensitivity probe).  Second number: distance to the shipped restatement's own
corners, i.e. how far that choice moves the end point.  CPU only; writes profiles/deviation_table_r2.md with --write.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

VARIANTS = [
    ("shipped restatement", 0, 0, 0),
    ("cost-only pass: x * (1/g) instead of x / g", 1, 0, 0),
    ("rotation axis exactly +-1 (not theta * (1/|theta|))", 2, 0, 0),
    ("PCL sequential fp32 centroid / covariance (PCA)", 4, 0, 0),
    ("PCL fp32 raw moments + eigen33 (plane refit)", 8, 0, 0),
    ("both PCL fp32 pieces", 12, 0, 0),
    ("Eigen 3.3 deflation threshold", 16, 0, 0),
    ("T=double pass: abs() resolving to ::abs(int) (GCC 5 hypothesis)", 32, 0, 0),
    ("'d' pressed 2 more times", 0, 1, 0),
    ("'d' pressed 4 more times", 0, 2, 0),
    ("boundary solution +1 ulp", 0, 0, 1),
    ("boundary solution -1 ulp", 0, 0, -1),
    ("boundary solution +4 ulp", 0, 0, 4),
    ("boundary solution +64 ulp", 0, 0, 64),
    ("boundary solution x (1 + 1.5e-8)", 0, 0, 1 << 26),
    ("boundary solution x (1 - 2.4e-7)", 0, 0, -(1 << 30)),
]


def main():
    import json
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures.json")))
    clouds = np.load(os.path.join(ROOT, "tests", "golden", "bag_clouds.npz"))
    bags = sorted(fx["key_flags"])
    L = O.lib()
    rows, base = [], {}
    for name, mask, dpairs, ulps in VARIANTS:
        L.ilcc2o_set_variant(mask, dpairs, ulps)
        cells = []
        for b in bags:
            f = np.asarray(fx["lidar_corners"][b])
            st, c, d = O.extract_frame(np.ascontiguousarray(clouds[f"xyzi_{b}"]), f.mean(0), flags=fx["key_flags"][b])
            assert st == 0
            if mask == 0 and dpairs == 0 and ulps == 0:
                base[b] = c.copy()
            cells.append((1e3 * np.linalg.norm(c - f, axis=1).max(), 1e3 * np.linalg.norm(c - base[b], axis=1).max(),
                          d["n_board"]))
        rows.append((name, cells))
    L.ilcc2o_set_variant(0, 0, 0)
    out = ["| restatement choice | " + " | ".join(f"bag {b}" for b in bags) + " |", "|---|" + "---|" * len(bags)]
    for name, cells in rows:
        out.append(f"| {name} | " + " | ".join(f"{a:.3f} / {m:.4f} ({n})" for a, m, n in cells) + " |")
    text = ("Max corner distance in mm: to the shipped fixture / to the shipped restatement's corners (board points).\n\n"
            + "\n".join(out) + "\n")
    print(text)
    if "--write" in sys.argv:
        open(os.path.join(ROOT, "profiles", "deviation_table_r2.md"), "w").write(text)


if __name__ == "__main__":
    main()
