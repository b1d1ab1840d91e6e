#!/usr/bin/env python
"""One bench step of the metric config (cfg3, 1920x1080, 64 spp) or a few Hydra-style 1-spp calls, without the bench harness,
for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py [step|interactive] [n]
Prints wall/device time and ray counts of what it ran (never a bench value when run under ncu)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moonshine_b200 import hydra, scenes  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "step"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = scenes.bathroom()
m = hydra.HdMoonshine()
scenes.load_scene(m, sc)
sensor, lens = m.CreateSensor(*sc.extent), m.CreateLens(sc.lens)
if mode == "step":
    spp = int(os.environ.get("SPP", 64))
    for _ in range(n):
        m.ResetStats()
        t = time.perf_counter()
        m.RenderLaunches(sensor, lens, spp)
        dt = time.perf_counter() - t
        st = m.GetStats()
        print(f"step {spp} spp: wall {dt * 1e3:.1f} ms device {st.device_ms:.1f} ms rays {st.rays_closest + st.rays_shadow} "
              f"-> {(st.rays_closest + st.rays_shadow) / st.device_ms / 1e3:.1f} Mrays/s launches {st.kernel_launches} waves {st.waves} tail paths {st.tail_paths}", flush=True)
else:
    for i in range(n):
        m.ResetStats()
        t = time.perf_counter()
        m.Render(sensor, lens)
        dt = time.perf_counter() - t
        st = m.GetStats()
        print(f"call {i}: wall {dt * 1e3:.2f} ms device {st.device_ms:.2f} ms rays {st.rays_closest + st.rays_shadow} launches {st.kernel_launches} tail paths {st.tail_paths}", flush=True)
