#!/usr/bin/env python
"""Writes tests/golden/*.npz: outputs of the CPU oracle (oracle/, the restatement of the reference shaders) on small
seeded inputs, committed so that (a) an accidental change of the oracle shows up as a diff against history and (b) the
GPU parity tests have fixed vectors next to the live oracle comparison. The reference itself cannot produce vectors
(Zig + DXC + Vulkan RT, no CPU path), so these are oracle outputs, not reference outputs.

    python tools/make_golden.py        # regenerates every file; run from the repo root
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from moonshine_b200 import scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402
from gpu_common import random_rays  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (scene factory, launches, ray probe centre, radius)
    "cfg1_single_triangle": (lambda: scenes.single_triangle(), 1, (0.0, 0.0, 0.0), 1.0),
    "cornell_32": (lambda: scenes.cornell_box((32, 32)), 2, (0.4, 0.5, 0.5), 0.3),
    "material_zoo": (lambda: scenes.material_zoo(), 1, (0.0, 0.0, 1.0), 1.2),
}


def build(name):
    maker, launches, center, radius = CASES[name]
    sc = maker()
    o = Oracle()
    scenes.load_scene(o, sc)
    img, counts = o.render(sc.lens, *sc.extent, n_launches=launches)
    rays = random_rays(17, 3000, np.array(center), radius)
    closest = o.trace(rays)
    short = rays.copy()
    short[:, 7] = 0.8
    occluded = o.trace(short, any_hit=True)["instance"] != 0xFFFFFFFF
    prim_rays, rng_state = o.primary_rays(sc.lens, *sc.extent, 3)
    return dict(image=img, ray_counts=np.asarray(counts, np.uint64), launches=np.uint32(launches), rays=rays,
                hit_instance=closest["instance"], hit_geometry=closest["geometry"], hit_primitive=closest["primitive"],
                hit_u_bits=closest["u"].view(np.uint32), hit_v_bits=closest["v"].view(np.uint32), occluded_within_0p8=occluded,
                rng_state_sample3=rng_state, alias_table=o.alias_table().view(np.uint8))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **build(name))
        print("wrote", name)


if __name__ == "__main__":
    main()
