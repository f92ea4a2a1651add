"""Golden vectors for the collision verdict, produced by the REFERENCE's own functions.

Imports /root/reference/scripts/collision_detection.py (with two shims: a stub matplotlib.pyplot and
np.Inf for NumPy 2) and scripts/collision_geometry.py in THIS container and records their answers on
seeded rectangle/rectangle and circle/rectangle cases.  The reference tree does not exist on the GPU
box, so only the resulting fixture (tests/golden/verdict_golden.npz) travels.
    python tests/golden/make_verdict_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts"


def main():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    sys.path.insert(0, REF)
    import collision_detection as cd          # the reference
    from collision_geometry import Circle, Rectangle
    rng = np.random.default_rng(20241017)
    rr, rc = [], []
    for _ in range(400):
        p1 = np.array([rng.uniform(0, 8), rng.uniform(0, 8), rng.uniform(-4, 4)])
        p2 = p1 + np.array([rng.uniform(-4.5, 4.5), rng.uniform(-4.5, 4.5), rng.uniform(-4, 4)])
        r1 = Rectangle(p1[0], p1[1], p1[2], length=3.0, width=2.0, pos="rear_axle_center", lb=1.0)
        r2 = Rectangle(p2[0], p2[1], p2[2], length=3.0, width=2.0, pos="rear_axle_center", lb=1.0)
        rr.append(list(p1) + list(p2) + [float(cd.collision_rect_and_rect(r1, r2))])
    for _ in range(400):
        p = np.array([rng.uniform(0, 8), rng.uniform(0, 8), rng.uniform(-4, 4)])
        c = np.array([p[0] + rng.uniform(-3.5, 3.5), p[1] + rng.uniform(-3.5, 3.5), rng.choice([0.5, 0.8, 1.0])])
        r = Rectangle(p[0], p[1], p[2], length=3.0, width=2.0, pos="rear_axle_center", lb=1.0)
        rc.append(list(p) + list(c) + [float(cd.collision_circle_and_rect(Circle(c[0], c[1], c[2]), r))])
    np.savez_compressed(os.path.join(HERE, "verdict_golden.npz"), rect_rect=np.asarray(rr), circle_rect=np.asarray(rc))
    print("rect/rect collisions:", int(np.asarray(rr)[:, -1].sum()), "circle/rect collisions:", int(np.asarray(rc)[:, -1].sum()))


if __name__ == "__main__":
    main()
