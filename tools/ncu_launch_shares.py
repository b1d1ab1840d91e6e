"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`):
per-kernel launch counts, time shares and average DRAM traffic per launch. Run here, no GPU needed.
  python tools/ncu_launch_shares.py gpurun_out/launches.csv [traffic.json]
"""
import collections
import csv
import json
import sys


def main(path, out_json=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[idi], {"kernel": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        name = d["kernel"].split("(")[0].replace("void ", "").replace("msne::", "")
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e6
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    print(f"{len(per)} launches, {total:.1f} ms (ncu gpu__time_duration per launch: serialised, cold cache)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<28s} {a[0]:5d} launches {a[1]:9.2f} ms {100 * a[1] / total:5.1f}%   DRAM {a[2] / 1e9:8.2f} GB total, {a[2] / max(1, a[0]) / 1e6:9.1f} MB/launch")
    if out_json:
        t = next((a for k, a in agg.items() if k.startswith("k_trace")), None)
        if t:
            json.dump({"kernel": "k_trace", "launches": t[0], "dram_bytes_per_launch": t[2] / t[0], "avg_launch_ms_under_ncu": t[1] / t[0],
                       "source": path, "note": "average over every k_trace launch of the captured waves (ncu, dram__bytes_read.sum + dram__bytes_write.sum)"},
                      open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
