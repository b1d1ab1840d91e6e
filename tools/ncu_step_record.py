#!/usr/bin/env python
"""Regenerates profiles/trace_kernel_traffic.json - the per-launch DRAM / L2 / SIMT record bench.py attaches to `roofline` - from
an ncu launch list of ONE bench step (every launch of the step, not a selection, so `traffic` and `achieved` average over the same
launch population):

    MSNE_WAVE_CONTEXTS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,\\
smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \\
        --clock-control none --csv --log-file gpurun_out/step_metrics.csv python tools/profile_step.py step 2
    python tools/ncu_step_record.py gpurun_out/step_metrics.csv profiles/trace_kernel_traffic.json profiles/r02_step_kernels.txt profiles/r02_step_launches.csv

The first step in the list is the warm-up (it learns the survival profile); the record is taken from the LAST k_generate onwards."""
import collections
import csv
import json
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[idi]), {"kernel": r[ki].split("(")[0].replace("void ", "").replace("msne::", "")})
        d[r[mi]] = float(r[vi].replace(",", ""))
    return list(per.values())


def main(path, out_json, out_txt=None, out_csv=None):
    launches = load(path)
    # one step = from the generate of its first wave; the step of interest is the last complete one: find the k_generate launches and
    # keep everything from the (n_waves_per_step)-th last onwards - with one context a 64-spp step is a single wave
    gens = [i for i, d in enumerate(launches) if d["kernel"].startswith("k_generate")]
    step = launches[gens[-1]:]
    if out_csv:   # compact per-launch list of the step (the raw ncu csv is ~1 MB): what every figure below is computed from
        with open(out_csv, "w") as f:
            f.write("launch,kernel,time_ns,dram_read_bytes,dram_write_bytes,l2_bytes,warp_inst,lanes_per_inst,issue_active_pct\n")
            for i, d in enumerate(step):
                f.write(f"{i},{d['kernel'].replace(',', ';')},{d.get('gpu__time_duration.sum', 0):.0f},{d.get('dram__bytes_read.sum', 0):.0f},"
                        f"{d.get('dram__bytes_write.sum', 0):.0f},{d.get('lts__t_bytes.sum', 0):.0f},{d.get('smsp__inst_executed.sum', 0):.0f},"
                        f"{d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.3f},{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.2f}\n")
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(d["kernel"], collections.Counter())
        a["launches"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0.0) / 1e6
        a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["l2"] += d.get("lts__t_bytes.sum", 0.0)
        inst = d.get("smsp__inst_executed.sum", 0.0)
        a["inst"] += inst
        a["lane_inst"] += inst * d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
        a["issue_w"] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    total_ms = sum(a["ms"] for a in agg.values())
    lines = [f"{len(step)} launches of one step, {total_ms:.1f} ms of kernel time under ncu (serialised, cold caches: shares count, absolutes do not)",
             f"{'kernel':<26s} {'n':>4s} {'ms':>9s} {'share':>6s} {'DRAM GB':>9s} {'L2 GB':>9s} {'DRAM GB/s':>10s} {'L2 GB/s':>9s} {'lanes/inst':>10s} {'issue %':>8s}"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        s = a["ms"] / 1e3
        lines.append(f"{k:<26s} {a['launches']:4d} {a['ms']:9.2f} {100 * a['ms'] / total_ms:5.1f}% {a['dram'] / 1e9:9.2f} {a['l2'] / 1e9:9.2f} "
                     f"{a['dram'] / 1e9 / max(s, 1e-12):10.0f} {a['l2'] / 1e9 / max(s, 1e-12):9.0f} {a['lane_inst'] / max(a['inst'], 1):10.2f} {a['issue_w'] / max(a['ns'], 1):8.1f}")
    print("\n".join(lines))
    if out_txt:
        open(out_txt, "w").write("\n".join(lines) + "\n")
    tr = collections.Counter()
    for k, a in agg.items():
        if k.startswith("k_trace"):
            tr.update(a)
    if tr["launches"]:
        rec = {"kernel": "k_trace (all instantiations)", "launches": int(tr["launches"]),
               "dram_bytes_per_launch": tr["dram"] / tr["launches"], "dram_bytes_per_step": tr["dram"],
               "avg_launch_ms_under_ncu": tr["ms"] / tr["launches"], "share_of_step_under_ncu": tr["ms"] / total_ms,
               "l2": {"bytes_per_launch": tr["l2"] / tr["launches"], "gb_per_s_under_ncu": tr["l2"] / 1e9 / (tr["ms"] / 1e3)},
               "simt": {"lanes_per_inst": tr["lane_inst"] / max(tr["inst"], 1), "issue_active_pct": tr["issue_w"] / max(tr["ns"], 1)},
               "step_dram_bytes": sum(a["dram"] for a in agg.values()),
               "source": out_csv or path, "regenerate": "tools/ncu_step_record.py (see its docstring for the ncu command)",
               "note": "every k_trace launch of one 64-spp step (one 128 Mi-path wave, one wave in flight): the same launch population bench.py averages `achieved` over"}
        json.dump(rec, open(out_json, "w"), indent=1)
        print("wrote", out_json)


if __name__ == "__main__":
    main(*sys.argv[1:5])
