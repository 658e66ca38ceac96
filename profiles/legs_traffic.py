"""Builds the "legs" section of profiles/traffic_r01.json from one per-leg ncu pass of bench.py:

    QD_BENCH_NCU_LEGS=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/legs_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu \
        > gpurun_out/legs_ncu.log

bench.py then puts ONE call of every `extra` leg inside its own profiler range, so the CSV lists the kernels of the legs in
the order below.  usage: python profiles/legs_traffic.py gpurun_out/legs_traffic.csv gpurun_out/legs_ncu.log"""
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
    kern.setdefault(int(r[iid]), {"name": r[ik]})[r[im]] = float(r[iv])
launches = [kern[i] for i in sorted(kern) if "qd::" in kern[i]["name"]]

# (leg, kernels of one call, in launch order)
SCAN = ["scan_block_sums_kernel", "scan_spine_kernel", "scan_finish_kernel"]
LEGS = [
    ("parse", ["parse_count_kernel"] + SCAN + ["parse_write_kernel"]),
    ("translate_1_frame", ["translate_frames_kernel<1"]),
    ("translate_self_frames", ["translate_frames_kernel<3"]),
    ("all_frames_strict", ["translate_frames_kernel<6"]),
    ("all_frames_variable_length", SCAN + ["tile_info_kernel", "warp_map_kernel", "translate_frames_kernel<6"]),
    ("all_frames_iupac_t1", ["translate_frames_kernel<6"]),
    ("all_frames_iupac_t11", ["translate_frames_kernel<6"]),
    ("all_frames_iupac_t22", ["translate_frames_kernel<6"]),
    ("canonical_windows", ["canonical_tile_kernel<3, 42, 2>"]),
    ("canonical_window_keys/fingerprint_64", ["canonical_tile_kernel<3, 42, 1>"]),
    ("canonical_window_keys/packed_2bit", ["canonical_tile_kernel<3, 42, 0>"]),
    ("windows_all", ["windows_batch_kernel<42>"] * 2 + ["translate_frames_kernel<6"] + ["windows_batch_kernel<20>"] * 6),
    ("expansion_windows", ["expansion_count_tiles_kernel"] + SCAN + ["expansion_write_tiles_kernel"]),
]
line = json.loads([l for l in open(log_path) if l.startswith('{"metric"')][-1])


def leg_node(path):
    node = line["extra"]
    for k in path.split("/"):
        node = node[k]
    return node


out, at = {}, 0
for leg, names in LEGS:
    got = launches[at:at + len(names)]
    assert len(got) == len(names) and all(n in g["name"] for n, g in zip(names, got)), (leg, [g["name"] for g in got])
    at += len(names)
    node = leg_node(leg)
    out[leg] = {
        "dram_read": int(sum(g["dram__bytes_read.sum"] for g in got)), "dram_write": int(sum(g["dram__bytes_write.sum"] for g in got)),
        "kernels": len(got), "ncu_us": round(sum(g["gpu__time_duration.sum"] for g in got) / 1e3, 2),
        "algorithmic_bytes_per_launch": node.get("roofline", {}).get("algorithmic_bytes_per_launch") or
                                        (round(node["GB/s_algorithmic"] * node["ms"] * 1e6) if leg == "parse" else None),
    }
assert at == len(launches), (at, len(launches))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic_r01.json")
d = json.load(open(path))
d["legs"] = {"_comment": "dram bytes of ONE call of each `extra` leg of bench.py (all its kernels), from the per-leg ncu pass "
                         "(profiles/r01_legs_traffic.csv; default cache control: caches flushed before every kernel)", **out}
json.dump(d, open(path, "w"), indent=1)
for k, v in out.items():
    a = v["algorithmic_bytes_per_launch"]
    print(f"{k:40s} read {v['dram_read'] / 1e9:8.3f} GB  write {v['dram_write'] / 1e9:8.3f} GB  algorithmic {a / 1e9 if a else float('nan'):8.3f} GB  ncu {v['ncu_us']:9.2f} us")
