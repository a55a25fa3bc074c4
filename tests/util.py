"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENE_KEYS = ("spheres", "capsules", "z_capsules", "cuboids", "z_cuboids", "heightfield", "heightfield_centre", "heightfield_scale")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    scene = {k: z["scene_" + k] for k in SCENE_KEYS if "scene_" + k in z.files}
    return z, scene, str(z["robot"]), bool(z["sorted"])


def assert_validity_parity(got, want, clearance_fn, what, band=1e-5):
    """Bit-exact except entries whose minimum clearance is inside the +-band (metres); those are counted."""
    got = np.asarray(got).astype(np.uint8)
    want = np.asarray(want).astype(np.uint8)
    assert got.shape == want.shape
    bad = np.nonzero(got != want)[0]
    if bad.size == 0:
        return 0
    clr = clearance_fn(bad)
    outside = bad[np.abs(clr) >= band]
    assert outside.size == 0, (f"{what}: {outside.size} of {len(got)} decisions differ outside the {band} m clearance band "
                               f"(first indices {outside[:8]}, clearances {clr[np.abs(clr) >= band][:8]})")
    return int(bad.size)
