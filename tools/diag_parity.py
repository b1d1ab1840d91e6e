#!/usr/bin/env python
"""GPU-vs-oracle image agreement, scene by scene: share of bit-identical pixels, share within 1e-4, ray counts. Run on a GPU box:
    python tools/diag_parity.py [--full]      (--full adds the 1920x1080 cfg3 frame and cfg2 at 512x512)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from moonshine_b200 import scenes  # noqa: E402
from gpu_common import gpu_render, oracle_render, identical_fraction, within_tol_fraction  # noqa: E402

cases = [("cfg1", scenes.single_triangle, 1), ("cornell96", lambda: scenes.cornell_box((96, 96)), 2), ("zoo", scenes.material_zoo, 2),
         ("cfg5-24", lambda: scenes.mesh_light_stress(24, (128, 128)), 1), ("cfg3-small", lambda: scenes.bathroom((160, 90), detail=0.03), 1)]
if "--full" in sys.argv:
    cases += [("cfg2-512", lambda: scenes.cornell_box((512, 512)), 2), ("cfg3-full", scenes.bathroom, 1)]
for name, maker, launches in cases:
    sc = maker()
    g, st, _ = gpu_render(sc, launches)
    r, counts, _ = oracle_render(sc, launches)
    d = np.abs(g[..., :3].astype(np.float64) - r[..., :3]) / np.maximum(1.0, np.abs(r[..., :3]))
    if name == "cfg3-full":
        bad = np.argwhere(~(g.view(np.uint32) == r.view(np.uint32)).all(axis=-1))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", "cfg3_full_diff.npz"), yx=bad, gpu=g[bad[:, 0], bad[:, 1]], oracle=r[bad[:, 0], bad[:, 1]])
    print(f"{name:12s} identical {identical_fraction(g, r):.6f}  within1e-4 {within_tol_fraction(g, r):.6f}  max rel diff {d.max():.3g}  "
          f"closest {st.rays_closest}/{int(counts[0])}  shadow {st.rays_shadow}/{int(counts[1])}", flush=True)
