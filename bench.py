#!/usr/bin/env python
"""bench.py — Mrays/s of the B200 wavefront path tracer on BASELINE.json's metric config.

Workload (config.workload): cfg3, the procedural "Salle-de-bain-class" bathroom (moonshine_b200/scenes.py
`bathroom`, ~1.4 M triangles, 34 materials, sun+sky env map) at 1920x1080, offline constants
(offline/main.zig:106-111). One STEP = one wave batch of `batch_spp` samples per pixel over every pixel this
rank owns (default 64 spp = one 128 Mi-path wave at 1080p; the default 16 timed steps are the full 1024 spp of the metric config). With N
GPUs the image is sharded into 32x32 tiles round-robin over ranks and each step renders batch_spp*N samples
of the rank's tiles (weak scaling: per-GPU work fixed), followed by ONE NCCL reduce of the RGBA32F radiance
buffer to rank 0 inside the timed region - issued by the library itself (HdMoonshineExtCommInit / CommReduceSensor).

  value    Mrays/s = (closest-hit + shadow rays traced by all ranks) / max-over-ranks device time (CUDA events
           on the library's stream around each step, first k_generate .. last k_accumulate).
  e2e      same metric through the reference-facing call HdMoonshineRender (hydra/moonshine.h:74): per step
           one launch of samples_per_run = batch_spp samples + the D2H copy of the sensor into the host buffer,
           wall-clock, host<->device copies inside the timed region.
  roofline dominant kernel k_trace: algorithmic bytes (48 B per traced ray) / CUDA-event duration of the k_trace
           launches alone; the ray-reordering passes of sorted bounces (keys, radix sort, binning) are timed separately
           (roofline.reorder_share_of_step).
  cpu_baseline / --impl reference: the CPU oracle (line-by-line restatement of the reference shaders; the
           reference itself has no CPU path and cannot be built here) on the host cores, bounded pixel sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line. Native libraries write there too (NCCL prints its version banner with printf when
# NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes
# to the saved descriptor (emit()).
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

METRIC = "Mrays/sec at 1920x1080x1024spp"
UNIT = "Mrays/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def make_scene(args):
    from moonshine_b200 import scenes
    if args.workload == "cfg3":
        return scenes.bathroom((args.width, args.height), detail=args.detail)
    if args.workload == "cfg2":
        return scenes.cornell_box((512, 512))
    if args.workload == "cfg5":
        return scenes.mesh_light_stress(args.n_side, (2048, 2048))
    raise SystemExit(f"unknown workload {args.workload}")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------
# CPU legs (oracle) — the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------------
def cpu_oracle_run(scene, steps, warmup, target_seconds):
    """Times the oracle on all host cores; each step = 1 spp over a stride-subsampled pixel set."""
    from moonshine_b200 import scenes
    from oracle.binding import Oracle
    o = Oracle()
    scenes.load_scene(o, scene)
    W, H = scene.extent
    cores = os.cpu_count() or 1
    # probe throughput on ~2000 pixels (also builds the oracle's BVH outside the timed part)
    probe_stride = max(1, (W * H) // 2000)
    o.render(scene.lens, W, H, n_launches=1, pixel_stride=max(1, (W * H) // 64))
    t = time.perf_counter()
    _, c = o.render(scene.lens, W, H, n_launches=1, pixel_stride=probe_stride)
    dt = max(time.perf_counter() - t, 1e-6)
    rays_per_pixel = float(c.sum()) / max(1, (W * H) // probe_stride)
    speed = float(c.sum()) / dt
    per_step = target_seconds / max(1, steps + warmup)
    n_pix = max(64, min(W * H, int(speed * per_step / max(rays_per_pixel, 1.0))))
    stride = max(1, (W * H) // n_pix)
    for i in range(warmup):
        o.render(scene.lens, W, H, n_launches=1, sample_count=i, pixel_stride=stride)
    rays = 0
    t0 = time.perf_counter()
    for i in range(steps):
        _, c = o.render(scene.lens, W, H, n_launches=1, sample_count=warmup + i, pixel_stride=stride)
        rays += int(c.sum())
    dt = time.perf_counter() - t0
    return {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps x 1 spp over every {stride}th pixel of {W}x{H} ({(W * H) // stride} pixels/step), {dt:.1f} s, oracle/liboracle.so on {cores} threads",
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    scene = make_scene(args)
    steps = min(args.steps, 64)
    r = cpu_oracle_run(scene, steps, min(args.warmup, 3), target_seconds=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args, scene, 1), step=r["sample"]),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference has no CPU path and cannot be built here (Zig/DXC/Vulkan-RT absent): this is the CPU oracle, a line-by-line restatement of its shaders"}
    emit(line)
    return 0


def config_dict(args, scene, world):
    return {"workload": f"{args.workload}: {scene.name}, {scene.triangle_count()} triangles, {len(scene.materials)} materials, "
                        f"{scene.extent[0]}x{scene.extent[1]}, offline constants (max_bounces 1024, env 1, mesh 1)",
            "step": f"{args.batch_spp * world} spp over the rank's tiles ({args.batch_spp} spp x {world} GPUs, weak scaling)",
            "target_spp": 1024, "batch_spp": args.batch_spp, "parallelism": f"tile-dp{world}",
            "l2": "per-step working set (path queues, >4 GB) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--detail", type=float, default=1.0)
    ap.add_argument("--n-side", type=int, default=1000, help="cfg5: instances per side (1000 = 10^6 instances)")
    ap.add_argument("--batch-spp", type=int, default=64)
    ap.add_argument("--paths-in-flight", type=int, default=0, help="wavefront pool size in paths (0 = library default, 128 Mi)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped at 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ahead-tail-paths", type=int, default=0, help="measurement: k_tail threshold for the two e2e_interactive legs (0 = library default)")
    ap.add_argument("--render-ahead", type=int, default=3, help="waves kept in flight beyond the one a call waits for, in the e2e_interactive.render_ahead leg (1..3)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libmoonshine has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from moonshine_b200 import hydra, scenes
    t_scene = time.perf_counter()
    scene = make_scene(args)
    t_generated = time.perf_counter()
    W, H = scene.extent
    m = hydra.HdMoonshine(local_rank)
    if args.paths_in_flight:
        m.SetPathsInFlight(args.paths_in_flight)
    scenes.load_scene(m, scene)
    t_loaded = time.perf_counter()
    if world > 1:
        # the product's own multi-GPU path: the library holds the NCCL communicator (HdMoonshineExtCommInit shards the pixel tiles by
        # rank; HdMoonshineRender / CommReduceSensor end with one ncclReduce of the frame to rank 0). torch.distributed is only the
        # harness plumbing here (barrier, max/sum of the timings).
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(hydra.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)   # the 128-byte NCCL id reaches the other ranks over the harness's own process group
        m.CommInit(bytes(uid.cpu().numpy().tobytes()), rank, world)
    sensor, lens = m.CreateSensor(W, H), m.CreateLens(scene.lens)
    batch = args.batch_spp * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_first = time.perf_counter()
    m.RenderLaunches(sensor, lens, batch)          # first render: builds the acceleration structure (host, all cores) and uploads it
    accel_ms = m.GetStats().accel_sync_ms
    first_render_s = time.perf_counter() - t_first
    setup = {"generate_scene_s": t_generated - t_scene, "load_through_abi_s": t_loaded - t_generated,
             "accel_build_ms": accel_ms, "first_render_s": first_render_s,
             "note": "outside the timed region, as in offline/main.zig:99-114 (which prints these phases): procedural scene generation in "
                     "Python, replay through the C ABI, host BVH build + upload inside the first render"}
    for _ in range(max(3, args.warmup) - 1):
        m.RenderLaunches(sensor, lens, batch)
    if world > 1:
        m.CommReduceSensor(sensor, 0)

    # ---- timed region: K steps, inputs resident in HBM ------------------------------------------------
    barrier()
    m.ResetStats()
    clocks = ClockSampler(local_rank)
    clocks.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.RenderLaunches(sensor, lens, batch)
    reduce_ms = 0.0
    if world > 1:
        tr = time.perf_counter()
        m.CommReduceSensor(sensor, 0)   # ncclReduce on the library's stream + the root's copy to its host buffer; includes the wait for the slowest rank
        reduce_ms = (time.perf_counter() - tr) * 1e3
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop()
    st = m.GetStats()
    dev_ms = st.device_ms + reduce_ms

    def allreduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    rays_local = float(st.rays_closest + st.rays_shadow)
    rays_total = allreduce(rays_local, dist.ReduceOp.SUM if world > 1 else None)
    dev_ms_max = allreduce(dev_ms, dist.ReduceOp.MAX if world > 1 else None)
    wall_ms_max = allreduce(wall_ms, dist.ReduceOp.MAX if world > 1 else None)
    value = rays_total / (dev_ms_max / 1e3) / 1e6

    # ---- roofline of the dominant kernel (k_trace): CUDA events on the launching stream around every k_trace launch -------------
    # The timed loop above keeps two waves in flight on two streams, so an event interval around one stream's kernel also spans
    # the other stream's work. The per-kernel figures therefore come from a second pass over the same workload with ONE wave in
    # flight (same kernels, same data, same launch population), measured right here, live.
    prof_steps = max(1, min(args.steps, 4))
    m.SetSchedule(wave_contexts=1)
    m.RenderLaunches(sensor, lens, batch)          # untimed: re-learns the survival profile for this wave size
    m.SetProfiling(True)
    barrier()
    m.ResetStats()
    for _ in range(prof_steps):
        m.RenderLaunches(sensor, lens, batch)
    stp = m.GetStats()
    m.SetProfiling(False)
    m.SetSchedule()
    peak, peak_src = hbm_peak()
    # rays the k_trace launches themselves traced (k_tail traces its stragglers inline): 48 B each = 32 B ray read + 16 B hit
    # write / conditional NEE clear
    trace_bytes = 48.0 * float(stp.rays_closest + stp.rays_shadow)
    trace_s = stp.trace_ms / 1e3 if stp.trace_ms > 0 else float("nan")
    achieved = trace_bytes / trace_s / 1e9
    traffic = ncu_traffic()
    step_bytes = 224.0 * st.rays_closest + 112.0 * st.rays_shadow + 32.0 * st.samples
    roofline = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "traffic_source": (traffic or {}).get("source"),
                "peak_source": peak_src, "launches": int(stp.trace_launches), "avg_launch_ms": stp.trace_ms / max(1, stp.trace_launches),
                "kernel_share_of_step": stp.trace_ms / max(stp.device_ms, 1e-9),
                "reorder_share_of_step": stp.sort_ms / max(stp.device_ms, 1e-9),
                "algorithmic_bytes_per_launch": trace_bytes / max(1, stp.trace_launches),
                "measured_in": f"{prof_steps} extra steps of the same workload with one wave in flight ({stp.device_ms / prof_steps:.1f} ms per step); "
                               "the timed steps overlap two waves on two streams, which makes per-kernel event intervals overlap",
                "l2": (traffic or {}).get("l2"), "simt": (traffic or {}).get("simt"),
                "note": "software BVH traversal is latency / SIMT-efficiency bound, not streaming bound: see DESIGN.md and profiles/"}
    roofline_step = {"bytes_formula": "224*N_ext + 112*N_shadow + 32*N_samples (SURVEY.md 8d)", "achieved": step_bytes / (st.device_ms / 1e3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": step_bytes / (st.device_ms / 1e3) / 1e9 / peak}
    timed_launches = int(st.kernel_launches)
    max_bounce = int(st.max_bounce_reached)

    # ---- e2e: the reference-facing call, host buffers, copies inside the timed region ------------------
    e2e_steps = args.e2e_steps or min(args.steps, 32)
    c = scenes.Constants(samples_per_run=batch)
    m.SetSpecConstants(c)  # = pipeline re-creation in the reference; clears the sensor
    host = m.GetSensorData(sensor)
    m.Render(sensor, lens)
    barrier()
    m.ResetStats()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        if not m.Render(sensor, lens):
            raise SystemExit("HdMoonshineRender failed: " + m.last_error())
        checksum += float(host[H // 2, W // 2, 0])  # the host buffer is really read
    barrier()
    e2e_wall = time.perf_counter() - t0
    st2 = m.GetStats()
    e2e_rays = allreduce(float(st2.rays_closest + st2.rays_shadow), dist.ReduceOp.SUM if world > 1 else None)
    e2e_wall = allreduce(e2e_wall, dist.ReduceOp.MAX if world > 1 else None)
    e2e = {"value": e2e_rays / e2e_wall / 1e6, "unit": UNIT, "h2d_bytes_per_step": 52, "d2h_bytes_per_step": W * H * 16,
           "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
           "call": "HdMoonshineRender(sensor, lens) with samples_per_run = batch spp; every call ends with the assembled frame in rank 0's pinned host "
                   "buffer (N > 1: ncclReduce of the tiles to rank 0 inside the call, then the copy to the host)"}

    # ---- e2e_interactive: the call pattern of the Hydra delegate (hydra/hydra.zig:316-329: ONE sample per HdMoonshineRender,
    # sensor copied to the host after every call), same scene and constants as above -------------------------------------------
    if args.ahead_tail_paths:
        m.SetSchedule(tail_paths=args.ahead_tail_paths)
    m.SetSpecConstants(scenes.Constants(samples_per_run=world))   # weak scaling: every rank still renders 1 spp of the full frame's worth of paths
    m.Render(sensor, lens)
    m.Render(sensor, lens)
    barrier()
    m.ResetStats()
    n_int = 48
    t0 = time.perf_counter()
    for _ in range(n_int):
        if not m.Render(sensor, lens):
            raise SystemExit("HdMoonshineRender failed: " + m.last_error())
    barrier()
    int_wall = time.perf_counter() - t0
    st3 = m.GetStats()
    int_rays = allreduce(float(st3.rays_closest + st3.rays_shadow), dist.ReduceOp.SUM if world > 1 else None)
    int_wall = allreduce(int_wall, dist.ReduceOp.MAX if world > 1 else None)
    e2e_interactive = {"value": int_rays / int_wall / 1e6, "unit": UNIT, "calls": n_int, "ms_per_call": int_wall / n_int * 1e3,
                       "device_ms_per_call": st3.device_ms / n_int, "launches_per_call": st3.kernel_launches / n_int,
                       "call": "HdMoonshineRender(sensor, lens), samples_per_run = 1 per GPU-frame, sensor read back every call"}

    # the same call pattern with render-ahead on (HdMoonshineExtSetRenderAhead): every call also starts the next sample's wave, whose
    # bandwidth-bound head then runs under the latency-bound tail of the wave being waited for; frames are the same bits
    m.SetRenderAhead(args.render_ahead)
    m.Render(sensor, lens)
    m.Render(sensor, lens)
    barrier()
    m.ResetStats()
    t0 = time.perf_counter()
    for _ in range(n_int):
        if not m.Render(sensor, lens):
            raise SystemExit("HdMoonshineRender failed: " + m.last_error())
    barrier()
    ahead_wall = time.perf_counter() - t0
    st4 = m.GetStats()
    m.SetRenderAhead(0)
    if args.ahead_tail_paths:
        m.SetSchedule()
    ahead_rays = allreduce(float(st4.rays_closest + st4.rays_shadow), dist.ReduceOp.SUM if world > 1 else None)
    ahead_wall = allreduce(ahead_wall, dist.ReduceOp.MAX if world > 1 else None)
    e2e_interactive["render_ahead"] = {"value": ahead_rays / ahead_wall / 1e6, "unit": UNIT, "calls": n_int, "ms_per_call": ahead_wall / n_int * 1e3,
                                       "launches_per_call": st4.kernel_launches / n_int, "waves_ahead": args.render_ahead, "waves_dropped": int(st4.ahead_discarded),
                                       "call": "the same calls after HdMoonshineExtSetRenderAhead(waves_ahead); rays of claimed waves only"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_run(scene, 4, 1, target_seconds=20.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": dev_ms_max / args.steps, "wall_ms_per_step": wall_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, scene, world),
                "clocks": clk, "e2e": e2e, "e2e_interactive": e2e_interactive, "gpu_launches": timed_launches, "roofline": roofline, "roofline_step": roofline_step,
                "cpu_baseline": cpu, "rays": {"closest": int(st.rays_closest), "shadow": int(st.rays_shadow), "samples": int(st.samples), "max_bounce": max_bounce,
                                              "total_all_ranks": rays_total},
                "reduce_ms": reduce_ms, "setup": setup}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    m.Destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
