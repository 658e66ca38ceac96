"""Builds profiles/traffic_r02.json from one per-leg ncu pass of bench.py:

    QD_BENCH_NCU_LEGS=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/legs_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu \
        --no-checks > gpurun_out/legs_ncu.log

In that mode bench.py puts ONE more call of every timed leg inside its own profiler range and records how many kernels
each such call launched (extra._ncu_calls) and which call belongs to which leg (ncu_call_index), so the ncu launch list
is cut without knowing kernel names.  The headline launch (translate_frames_kernel<6, 1024, 0> on the 10 M-read
workload) is taken from the first profiled frames launch of the strong-scaling leg, which runs the same workload.
usage: python profiles/legs_traffic.py gpurun_out/legs_traffic.csv gpurun_out/legs_ncu.log"""
import csv
import json
import os
import sys

csv_path, log_path = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
hdr = rows[0]
iid, ik, im, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value"))
kern = {}
for r in rows[1:]:
    kern.setdefault(int(r[iid]), {"name": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
launches = [kern[i] for i in sorted(kern) if "qd::" in kern[i]["name"] or "qd::" in kern[i]["name"].replace(" ", "")]
line = json.loads([l for l in open(log_path) if l.startswith('{"metric"')][-1])
calls = line["extra"]["_ncu_calls"]
assert sum(calls) == len(launches), (sum(calls), len(launches))
starts = [0]
for n in calls:
    starts.append(starts[-1] + n)


def legs_of(node, prefix=""):
    for k, v in node.items():
        if isinstance(v, dict):
            if "ncu_call_index" in v:
                yield prefix + k, v
            yield from legs_of(v, prefix + k + "/")


out = {}
for name, node in legs_of(line["extra"]):
    i = node["ncu_call_index"]
    got = launches[starts[i]:starts[i + 1]]
    r = node.get("roofline", {})
    names = sorted({g["name"].split("(")[0] for g in got})
    out[name] = {
        "dram_read": int(sum(g["dram__bytes_read.sum"] for g in got)), "dram_write": int(sum(g["dram__bytes_write.sum"] for g in got)),
        "kernels": len(got), "kernel_names": names, "ncu_us": round(sum(g["gpu__time_duration.sum"] for g in got) / 1e3, 2),
        "algorithmic_bytes_per_launch": r.get("algorithmic_bytes_per_launch"),
    }
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic_r02.json")
d = {"_comment": "dram__bytes_read/write.sum of ONE call of each leg of bench.py (all its kernels) from a per-leg ncu pass at bench size "
                 "(profiles/r02_legs_traffic.csv; ncu's default cache control: caches flushed before every kernel)"}
head = out.get("config3_strong_scaling")
if head and head["kernels"] == 1:
    d["translate_frames_kernel<6>"] = {"dram_read": head["dram_read"], "dram_write": head["dram_write"],
                                                "reads": line["config"]["reads_per_gpu"], "read_len": line["config"]["read_len"]}
d["legs"] = out
json.dump(d, open(path, "w"), indent=1)
for k, v in out.items():
    a = v["algorithmic_bytes_per_launch"]
    t = v["dram_read"] + v["dram_write"]
    print(f"{k:40s} kernels {v['kernels']:5d}  dram {t / 1e9:9.3f} GB  algorithmic {a / 1e9 if a else float('nan'):9.3f} GB  ratio {t / a if a else float('nan'):6.3f}  ncu {v['ncu_us']:10.2f} us")
