#!/usr/bin/env python
"""bench.py -- six-frame translation throughput of the quickdna hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A "step" is one pass of translate_all_frames over one batch of synthetic reads
(BASELINE.json config 3: 10 M x 1 kb reads per GPU, weak scaling, no data-path collective).

  value        nt/s, whole job, inputs resident in HBM (CUDA events around the K launches, max over ranks)
  e2e          nt/s through the host-pointer C-ABI call (qd_translate_frames_batch_uniform) with pinned host buffers:
               H2D + kernels + D2H inside the timed region; beside it the bare-copy ceiling of the box
               (tools/pcie_copy_bench) and, at N > 1, the same batch through ONE call that drives all N GPUs
  roofline     algorithmic bytes (n + 2*sum floor((n-k)/3) = 2996 B per 1 kb read) / launch time vs the measured HBM copy
               peak (MEASURED_PEAKS.json); roofline.legs = the same accounting for every other kernel of the path
  config.checks   every output byte of the headline launch, of the 250 M-nt reverse complement and of one full chunk of each
               window enumeration compared with the CPU oracle through a position-weighted linear checksum
  cpu_baseline the oracle's scalar restatement of the reference loops on a bounded sample (rank 0, N=1)
  extra        full detail of every leg (BASELINE.json configs 2, 4, 5 at their stated sizes, strong scaling of config 3)

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) instead.
"""

from __future__ import annotations

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

READ_LEN = 1000
ALGO_BYTES_PER_READ = READ_LEN + 2 * sum((READ_LEN - k) // 3 for k in range(3))  # 2996
METRIC = "translate_all_frames_nt_per_s"
UNIT = "nt/s"
CONFIG3_READS = 10_000_000  # BASELINE.json config 3
M64 = (1 << 64) - 1
TRAFFIC_FILES = ("traffic_r02.json", "traffic_r01.json")  # committed per-leg ncu captures, newest first


def _traffic_table():
    for name in TRAFFIC_FILES:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f), "profiles/" + name
        except Exception:
            continue
    return {}, None


def ncu_traffic(kernel: str, **match):
    """dram bytes per launch of `kernel` from the committed ncu capture, if it was taken on this workload."""
    table, _src = _traffic_table()
    e = table.get(kernel)
    if e and all(e.get(k) == v for k, v in match.items()):
        return e["dram_read"] + e["dram_write"]
    return None


def fill_leg_traffic(extra: dict):
    """roofline.traffic of the `extra` legs: dram bytes of ONE call of the leg (all its kernels) from the committed per-leg
    ncu capture (profiles/traffic_rNN.json "legs", taken with QD_BENCH_NCU_LEGS=1), when the capture was taken on the same
    workload (same algorithmic bytes per call).  A replayed constant, labelled as such (`traffic_source`)."""
    table, src = _traffic_table()
    legs = table.get("legs", {})

    def visit(name, node):
        if not isinstance(node, dict):
            return
        r = node.get("roofline")
        e = legs.get(name)
        if isinstance(r, dict) and r.get("traffic") is None and e and e.get("algorithmic_bytes_per_launch") == r.get("algorithmic_bytes_per_launch"):
            r["traffic"] = e["dram_read"] + e["dram_write"]
            r["traffic_source"] = f"{src} (ncu capture of this workload, committed; not measured in this run)"
        for k, v in node.items():
            if k != "roofline":
                visit(f"{name}/{k}", v)

    for k, v in extra.items():
        visit(k, v)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_CPU_INPUT = {}


def _cpu_input(seed: int, count: int):
    """The counter-based input stream of synth_np, generated once per (seed, size) outside any timed region: chunks on a
    thread pool (numpy releases the GIL), and a prefix of a larger cached stream is reused as is."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from quickdna_b200.synth import synth_np

    for key, buf in _CPU_INPUT.items():
        if isinstance(key, tuple) and key[0] == seed and key[1] >= count:
            return buf[:count]
    chunk = 1 << 22
    out = np.empty(count, dtype=np.uint8)

    def fill(c0):
        c1 = min(count, c0 + chunk)
        out[c0:c1] = synth_np(seed, c0, c1 - c0)

    with ThreadPoolExecutor(max(1, min(host_threads(), 32))) as ex:
        list(ex.map(fill, range(0, count, chunk)))
    _CPU_INPUT[(seed, count)] = out
    return out


def cpu_all_frames(n_reads: int, threads: int, reps: int, seed: int = 3):
    """Times the oracle's restatement of the Rust path (parse + translate_all_frames) on n_reads reads."""
    import oracle

    import numpy as np

    dna = _cpu_input(seed, n_reads * READ_LEN)
    # the output buffer is allocated and touched once, outside the timed call (a Rust caller's allocator reuses warm pages;
    # first-touch page faults of a fresh 3 GB buffer would otherwise be charged to the CPU arm)
    need = n_reads * 2 * (READ_LEN - 2)
    out = _CPU_INPUT.get("out")
    if out is None or out.size < need:
        out = _CPU_INPUT["out"] = np.zeros(need, dtype=np.uint8)
    oracle.bench_all_frames_batch(0, dna[: 2000 * READ_LEN], 2000, READ_LEN, threads, 1, out=out)  # warm the tables
    secs, _ = oracle.bench_all_frames_batch(0, dna, n_reads, READ_LEN, threads, reps, out=out[:need])
    return n_reads * READ_LEN * reps / secs, secs


def run_reference(args):
    """--impl reference: the CPU restatement of the reference on all host threads (rank 0 only).  Loads oracle/ alone:
    nothing of the product (no CUDA library, no torch) is mapped into this process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    threads = host_threads()
    n_reads = args.cpu_reads or 100_000 * max(1, min(threads, 64))
    # size one step to roughly a second of CPU work, K+W steps well within minutes
    t0 = time.perf_counter()
    for _ in range(args.warmup):
        cpu_all_frames(min(n_reads, 20_000), threads, 1)
    vals = []
    for _ in range(args.steps):
        v, _s = cpu_all_frames(n_reads, threads, 1)
        vals.append(v)
    total_s = time.perf_counter() - t0
    value = len(vals) * n_reads * READ_LEN / sum(n_reads * READ_LEN / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_reads * READ_LEN / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"translate_all_frames, NCBI table 1, {n_reads} x {READ_LEN}-nt ACGT reads per step (sample of config 3)",
                   "sampled_from": CONFIG3_READS, "reads_per_step": n_reads, "read_len": READ_LEN,
                   "impl_detail": "C port of src/rust_api.rs:263-270 + :310-322 (oracle/qd_oracle.c, gcc -O3 -march=native); no Rust toolchain here",
                   "reference_is": "a C port of the Rust crate, not the crate", "wall_s": total_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_reads} of the {CONFIG3_READS} reads x {READ_LEN} nt per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def copy_ceiling(n_gpus: int, mb_in: int):
    """The bare pinned-copy ceiling of the box with the traffic pattern of the e2e leg (tools/pcie_copy_bench: per GPU three
    streams, 64 MiB H2D + 2.016 x that D2H per chunk): one process with a thread per GPU, and a process per GPU."""
    exe = os.path.join(ROOT, "tools", "pcie_copy_bench")
    res = {}
    if not os.path.exists(exe):
        return {"error": "tools/pcie_copy_bench not built"}
    for mode in ("threads", "procs") if n_gpus > 1 else ("threads",):
        try:
            out = subprocess.run([exe, "--gpus", str(n_gpus), "--mode", mode, "--mb-in", str(mb_in), "--reps", "2"], capture_output=True, text=True,
                                 timeout=120)
            j = json.loads(out.stdout.strip().splitlines()[-1])
            res[mode] = {"h2d_GBps": j["aggregate_h2d_GBps"], "d2h_GBps": j["aggregate_d2h_GBps"], "per_gpu_GBps": j["per_gpu_GBps"]}
        except Exception as ex:  # the ceiling is context for the e2e number, never fatal
            res[mode] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=CONFIG3_READS, help="reads per GPU (weak scaling)")
    ap.add_argument("--e2e-reads", type=int, default=2_000_000, help="reads per GPU in the host-buffer (e2e) leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=0)
    ap.add_argument("--revcomp-nt", type=int, default=250_000_000)
    ap.add_argument("--window-seqs", type=int, default=80_000, help="99 999-nt sequences per GPU in the config-5 legs (64 GB / 8 GPUs)")
    ap.add_argument("--window-chunk-seqs", type=int, default=2048, help="sequences per ring chunk of the canonical leg (two 8.6 GB slots; all_windows: at most 512, two 6.3 GB slots)")
    ap.add_argument("--expansion-chunk", type=int, default=12_000_000, help="windows per ring chunk of the expansion leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip the whole-output checksum comparisons with the CPU oracle")
    ap.add_argument("--table", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0:
        g.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # host-side waits that must not spin on the GPUs
        dist.barrier()
    from quickdna_b200 import _native as N
    from quickdna_b200 import batch as B
    from quickdna_b200.synth import synth_torch

    ctx = N.default_context(local_rank)
    L = N.lib()
    n_reads = args.reads
    total_nt = n_reads * READ_LEN
    threads_all = host_threads()
    cpu_threads = max(1, threads_all // world)  # every rank checks its own shard on its share of the host cores
    do_checks = not args.no_checks
    oracle = None
    if do_checks or not args.no_cpu:
        import oracle  # the checker (test infrastructure): never on a timed GPU path

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_true(ok: bool) -> bool:
        if world == 1:
            return ok
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    checks = {}

    # ---- workload: this rank's shard of the batch, generated on the device (counter-based, seed 3)
    x = synth_torch(3, rank * total_nt, total_nt, dev)
    out = torch.empty(n_reads * 16 * B.runs(READ_LEN) * 6, dtype=torch.uint8, device=dev)

    def step():
        B.translate_frames_uniform_dev(x, n_reads, READ_LEN, out=out, table=args.table, frames=6, ctx=ctx)

    B.dev_error_reset(ctx)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu_legs = os.environ.get("QD_BENCH_NCU_LEGS") == "1"  # profile one launch of every `extra` leg instead of the timed region
    if not ncu_legs:
        torch.cuda.profiler.start()  # `ncu --profile-from-start off` lists exactly the timed region
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    if not ncu_legs:
        torch.cuda.profiler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    assert B.dev_error_fetch(x, uniform_len=READ_LEN, ctx=ctx) is None
    ms_per_step = ms / args.steps
    value = world * total_nt / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak()
    achieved = n_reads * ALGO_BYTES_PER_READ / (ms_per_step * 1e-3) / 1e9
    _tt, traffic_src = _traffic_table()
    traffic = ncu_traffic("translate_frames_kernel<6>", reads=n_reads, read_len=READ_LEN)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": (f"{traffic_src} (ncu --set full capture of this workload, committed; not measured in this run)" if traffic else None),
                "kernel": "translate_frames_kernel<6, 1024, 0>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n_reads * ALGO_BYTES_PER_READ}

    # ---- check: EVERY byte of the headline launch's output against the CPU oracle (position-weighted linear checksum)
    if do_checks and args.table == 1:
        t0 = time.perf_counter()
        got = B.u64(B.checksum64_dev(out, ctx=ctx))
        want = oracle.check_frames_uniform(3, rank * n_reads, n_reads, READ_LEN, 0, 6, cpu_threads)
        ok = all_ranks_true(got == want)
        checks["config3_all_output_bytes_equal_oracle"] = ok
        checks["config3_checked"] = f"{n_reads} reads = {out.numel()} output bytes per GPU, CPU oracle on {cpu_threads} threads, {time.perf_counter() - t0:.1f} s"
        assert ok, f"six-frame output differs from the oracle: checksum {got:#x} != {want:#x}"

    # ---- extra legs: every other kernel of the path on resident data
    extra, legs = {}, {}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        barrier()
        t_ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
        if ncu_legs:  # one more call, alone inside a profiler range: `ncu --profile-from-start off` sees each leg's kernels in order
            before = ctx.launch_count
            torch.cuda.profiler.start()
            fn()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            ncu_calls.append(ctx.launch_count - before)
        return t_ms

    ncu_calls = []  # kernel launches of every call profiled under QD_BENCH_NCU_LEGS, in order (profiles/legs_traffic.py cuts the ncu list with it)

    def leg(name, workload, t_ms, algo_bytes, kernel, units=None, **more):
        """One leg: full record in `extra`, compact (ms, GB/s, frac of the measured HBM peak) in roofline.legs."""
        gbs = algo_bytes / (t_ms * 1e-3) / 1e9
        rec = {"workload": workload, "ms": t_ms,
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None, "kernel": kernel,
                            "algorithmic_bytes_per_launch": algo_bytes}}
        for k, v in (units or {}).items():
            rec[k + "_per_s"] = world * v / (t_ms * 1e-3)
        rec.update(more)
        if ncu_legs and ncu_calls:
            rec["ncu_call_index"] = len(ncu_calls) - 1  # the profiled call that belongs to this leg
        extra[name] = rec
        legs[name] = {"ms": round(t_ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
        return rec

    if not args.no_extra:
        # ---- config 2: reverse complement of one 250 Mb sequence, four buffer sets rotated (> L2)
        n_rc = args.revcomp_nt
        bufs = [(x[i * n_rc:(i + 1) * n_rc], out[i * n_rc:(i + 1) * n_rc]) for i in range(4)]
        turn = [0]

        def rc_step(**kw):
            i = turn[0] = (turn[0] + 1) % 4
            B.revcomp_dev(bufs[i][0], out=bufs[i][1], ctx=ctx, **kw)

        rc_ms = timed(rc_step, reps=20, warm=4)
        leg("revcomp", f"reverse_complement, one {n_rc}-nt ACGT sequence per GPU, ASCII non-strict (config 2)", rc_ms, 2 * n_rc, "revcomp_kernel<1024>",
            units={"nt": n_rc})
        extra["revcomp"]["roofline"]["traffic"] = ncu_traffic("revcomp_kernel", nt=n_rc)
        if do_checks:
            # buffer set 0 holds stream bytes [rank * total_nt, + n_rc) of seed 3 and their reverse complement
            got = B.u64(B.checksum64_dev(bufs[0][1], ctx=ctx))
            want = oracle.check_revcomp(3, rank * total_nt, n_rc, False, cpu_threads)
            ok = all_ranks_true(got == want)
            checks["config2_all_output_bytes_equal_oracle"] = ok
            assert ok, f"reverse complement differs from the oracle: checksum {got:#x} != {want:#x}"
        variants = {}
        for name, n_v, kw in (("n_plus_1", n_rc + 1, {}), ("n_minus_1", n_rc - 1, {}), ("strict", n_rc, {"strict": True})):
            vb = [(x[i * (n_rc + 16): i * (n_rc + 16) + n_v], out[i * (n_rc + 16): i * (n_rc + 16) + n_v]) for i in range(4)]

            def v_step(vb=vb, kw=kw):
                i = turn[0] = (turn[0] + 1) % 4
                B.revcomp_dev(vb[i][0], out=vb[i][1], ctx=ctx, **kw)

            v_ms = timed(v_step, reps=20, warm=4)
            variants[name] = {"nt": n_v, "ms": v_ms, "GB/s": 2 * n_v / (v_ms * 1e-3) / 1e9}
            legs["revcomp_" + name] = {"ms": round(v_ms, 4), "GBps": round(2 * n_v / (v_ms * 1e-3) / 1e9, 1), "frac": round(2 * n_v / (v_ms * 1e-3) / 1e9 / peak, 4)}
        assert B.dev_error_fetch(x, ctx=ctx) is None
        # mask encoding (the typed Rust API: Vec<Nucleotide>): parse the ASCII once on the device, then reverse-complement masks
        masks, cnt = B.parse_dev(x[:n_rc], ctx=ctx)
        assert int(cnt.item()) == n_rc
        mo = out[:n_rc]
        v_ms = timed(lambda: B.revcomp_dev(masks[:n_rc], out=mo, enc=N.ENC_MASK, ctx=ctx), reps=10, warm=3)
        variants["mask_encoding"] = {"nt": n_rc, "ms": v_ms, "GB/s": 2 * n_rc / (v_ms * 1e-3) / 1e9, "note": "one buffer set reused (partly L2-resident)"}
        extra["revcomp_variants"] = variants
        # ASCII -> mask parse itself (src/rust_api.rs:310-322), 2 B per input byte
        p_ms = timed(lambda: B.parse_dev(x[:n_rc], out=masks, ctx=ctx), reps=20, warm=3)
        leg("parse", f"ASCII -> mask parse of {n_rc} nt", p_ms, 2 * n_rc, "parse kernels", units={"nt": n_rc})
        del masks, cnt

        # ---- the other forms of the frame kernel on the same resident data
        n_v = min(n_reads, 4_000_000)
        nt_v = n_v * READ_LEN
        f1 = READ_LEN // 3
        f3 = sum((READ_LEN - k) // 3 for k in range(3))
        ms1 = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=1, ctx=ctx))
        leg("translate_1_frame", f"translate (1 frame) on {n_v} x {READ_LEN}-nt reads per GPU", ms1, n_v * (READ_LEN + f1),
            "translate_frames_kernel<1, 1024, direct>", units={"nt": nt_v})
        ms3 = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=3, ctx=ctx))
        leg("translate_self_frames", f"translate_self_frames (3 frames) on {n_v} x {READ_LEN}-nt reads per GPU", ms3, n_v * (READ_LEN + f3),
            "translate_frames_kernel<3, 1024>", units={"nt": nt_v})
        mss = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=6, strict=True, ctx=ctx))
        leg("all_frames_strict", f"translate_all_frames, strict alphabet check, {n_v} x {READ_LEN}-nt ACGT reads per GPU", mss, n_v * ALGO_BYTES_PER_READ,
            "translate_frames_kernel<6, 1024>", units={"nt": nt_v})
        # config 3 as BASELINE.json words it (STRONG scaling): the 10 M reads split over the GPUs, this rank's share in one launch
        n_s = CONFIG3_READS // world
        if n_s <= n_reads:
            ms_s = timed(lambda: B.translate_frames_uniform_dev(x[: n_s * READ_LEN], n_s, READ_LEN, out=out, table=args.table, frames=6, ctx=ctx), reps=10, warm=3)
            r = leg("config3_strong_scaling", f"translate_all_frames, {CONFIG3_READS} x {READ_LEN}-nt reads split over {world} GPU(s): {n_s} per GPU, one launch",
                    ms_s, n_s * ALGO_BYTES_PER_READ, "translate_frames_kernel<6, 1024>", units={"nt": n_s * READ_LEN})
            r["scaling"] = "strong"
        # variable-length batch (lengths uniform in 500..1500, seed 4): CSR offsets on the device, scan + tile map included
        gen = torch.Generator(device="cpu").manual_seed(4 + rank)
        lens = torch.randint(500, 1501, (n_v,), generator=gen, dtype=torch.int64)
        offs = torch.zeros(n_v + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(lens, 0)
        keep = int(torch.searchsorted(offs, torch.tensor([min(nt_v, total_nt)]), right=True).item()) - 1
        offs_d = offs[: keep + 1].to(dev)
        tot_v = int(offs[keep].item())
        runs_v = int(((lens[:keep] + 47) // 48).sum().item())
        algo_v = int((lens[:keep] + 2 * ((lens[:keep]) // 3 + (lens[:keep] - 1) // 3 + (lens[:keep] - 2) // 3)).sum().item())
        launches_v0 = ctx.launch_count
        msv = timed(lambda: B.translate_frames_csr_dev(x[:tot_v], offs_d, out=out[: runs_v * 96], table=args.table, frames=6, ctx=ctx))
        leg("all_frames_variable_length", f"translate_all_frames on {keep} reads of 500..1500 nt (CSR offsets; the scan and map kernels timed too)", msv, algo_v,
            "translate_frames_kernel<6, 1024> + scan / csr_finish_kernel", units={"nt": tot_v}, launches_per_call=(ctx.launch_count - launches_v0) // 7)
        # IUPAC-ambiguous, mixed-case input (config 4): 5 % ambiguity codes, tables 1 / 11 / 22
        from quickdna_b200.synth import synth_iupac_np

        base = torch.from_numpy(synth_iupac_np(5, rank * 50_000_000, 50_000 * READ_LEN)).to(dev)
        n_a = 1_000_000
        xa = base.repeat(n_a // 50_000)
        for tbl in (1, 11, 22):
            msa = timed(lambda: B.translate_frames_uniform_dev(xa, n_a, READ_LEN, out=out, table=tbl, frames=6, ctx=ctx), reps=3)
            leg(f"all_frames_iupac_t{tbl}", f"translate_all_frames, NCBI table {tbl}, {n_a} x {READ_LEN}-nt reads, 5 % IUPAC ambiguity codes, 50 % lower case (config 4)",
                msa, n_a * ALGO_BYTES_PER_READ, "translate_frames_kernel<6, 1024>", units={"nt": n_a * READ_LEN})
        assert B.dev_error_fetch(xa, uniform_len=READ_LEN, ctx=ctx) is None
        if do_checks:
            # the IUPAC output of table 22 (the last launch): 50 000 distinct reads, every output byte against the oracle
            host_reads = base.cpu().numpy().reshape(50_000, READ_LEN)
            per = 16 * B.runs(READ_LEN) * 6
            got_img = out[: 2000 * per].cpu().numpy()
            ok = True
            for i in range(0, 2000, 97):
                fr = oracle.py_all_frames(22, host_reads[i].tobytes())
                for f in range(6):
                    o0 = i * per + B.slot_frame_offset(READ_LEN, 6, f)
                    ok &= got_img[o0:o0 + len(fr[f])].tobytes() == fr[f]
            checks["config4_iupac_sampled_reads_equal_oracle"] = all_ranks_true(bool(ok))
            assert ok
        B.dev_error_reset(ctx)
        B.translate_frames_uniform_dev(xa, n_a, READ_LEN, out=out, table=1, frames=6, strict=True, ctx=ctx)
        e = B.dev_error_fetch(xa, uniform_len=READ_LEN, strict=True, ctx=ctx)
        extra["all_frames_iupac_strict_error"] = None if e is None else {"kind": e.kind, "byte": chr(e.byte), "seq_index": e.seq_index,
                                                                       "pos": e.pos, "message": str(e)}
        if do_checks and e is not None:
            try:
                oracle.revcomp_bytes(base[:READ_LEN * (e.seq_index + 1)].cpu().numpy().tobytes(), strict=True)  # validates every byte in order
                checks["config4_strict_first_error_equals_oracle"] = False
            except oracle.OracleError as o:
                checks["config4_strict_first_error_equals_oracle"] = (o.kind, o.byte, o.pos) == (e.kind, e.byte, e.seq_index * READ_LEN + e.pos)
            assert checks["config4_strict_first_error_equals_oracle"]
        B.dev_error_reset(ctx)
        del xa, base

        # ---- config 5 at its stated size: 64 GB / 8 GPUs = 80 000 x 99 999-nt sequences per GPU, window output through a ring
        seq_len, w = 99_999, 42
        n_w = min(args.window_seqs, total_nt // seq_len)
        if n_w > 0:
            nwin = seq_len - w + 1
            M = lambda t: sum(int(v) & M64 for v in t.cpu().tolist()) & M64  # noqa: E731
            # canonical rows: 43 B per window, 336 GB of rows per GPU through a two-slot ring
            chunk_c = min(args.window_chunk_seqs, n_w)
            slot_c = (L.qd_canonical_chunk_bytes(chunk_c, seq_len, w) + 255) & ~255
            ring_c = out[: 2 * slot_c]
            B.dev_error_reset(ctx)
            B.canonical_windows_chunked_dev(x, min(n_w, 2 * chunk_c), seq_len, w, ring_c, chunk_c, checksums=False, ctx=ctx)
            cw_ms = timed(lambda: B.canonical_windows_chunked_dev(x, n_w, seq_len, w, ring_c, chunk_c, checksums=False, ctx=ctx), reps=1, warm=0)
            assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True, ctx=ctx) is None
            r = leg("canonical_windows", f"canonical 42-nt windows of {n_w} x {seq_len}-nt sequences per GPU (config 5, benches/canonical.rs shape): "
                    f"{n_w * nwin * w / 1e9:.0f} GB of rows through a ring of two {slot_c / 1e9:.2f} GB slots, {L.qd_chunk_count(n_w, chunk_c)} chunks",
                    cw_ms, n_w * nwin * (1 + w), "canonical_tile_kernel<3, 42, rows>", units={"windows": n_w * nwin, "nt": n_w * seq_len},
                    sequences_per_gpu=n_w, chunk_sequences=chunk_c)
            if do_checks:
                t0 = time.perf_counter()
                sums = B.canonical_windows_chunked_dev(x, n_w, seq_len, w, ring_c, chunk_c, checksums=True, ctx=ctx)
                want0 = oracle.check_canonical_windows(x[: chunk_c * seq_len].cpu().numpy(), chunk_c, seq_len, w, cpu_threads)
                ok = all_ranks_true((int(sums[0].item()) & M64) == want0)
                checks["config5_canonical_chunk0_all_bytes_equal_oracle"] = ok
                r["checksum_all_chunks"] = f"{M(sums):#018x}"
                r["check_seconds"] = time.perf_counter() - t0
                assert ok
            # one key per canonical window instead of its 42 bytes, same full shard, no ring needed for 8 B keys
            n_k = min(n_w, out.numel() // (nwin * 16))
            keys = {}
            for kind, label, kb in ((N.CANON_KEY_FP64, "fingerprint_64", 8), (N.CANON_KEY_PACKED, "packed_2bit", 16)):
                kout = out.view(torch.int32)
                k_ms = timed(lambda: B.canonical_window_keys_dev(x, n_k, seq_len, w, kind=kind, out=kout, ctx=ctx), reps=2, warm=1)
                k_gbs = n_k * nwin * (1 + kb) / (k_ms * 1e-3) / 1e9
                keys[label] = {"ms": k_ms, "windows_per_s": world * n_k * nwin / (k_ms * 1e-3), "key_bytes": kb, "sequences_per_gpu": n_k,
                               **({"ncu_call_index": len(ncu_calls) - 1} if ncu_legs and ncu_calls else {}),
                               "roofline": {"bound": "hbm", "achieved": k_gbs, "peak": peak, "unit": "GB/s", "frac": k_gbs / peak, "traffic": None,
                                            "kernel": f"canonical_tile_kernel<3, 42, {kind}>", "algorithmic_bytes_per_launch": n_k * nwin * (1 + kb),
                                            "note": "ALU-bound, not HBM-bound"}}
                legs["canonical_keys_" + label] = {"ms": round(k_ms, 4), "GBps": round(k_gbs, 1), "frac": round(k_gbs / peak, 4)}
            extra["canonical_window_keys"] = {"workload": f"one key per canonical 42-nt window of {n_k} x {seq_len}-nt sequences per GPU (SURVEY 8(f) rank 3)",
                                              "speedup_vs_materialised_rows": {k: (cw_ms * n_k / n_w) / v["ms"] for k, v in keys.items()}, **keys}
            # the all_windows shape: 124 output bytes per nucleotide, ~1 TB per GPU through a two-slot ring
            chunk_a = min(args.window_chunk_seqs, 512, n_w)  # two 6.3 GB slots inside the 20 GB output buffer
            aa_off = C.c_size_t()
            slot_a = (L.qd_windows_all_chunk_bytes(chunk_a, seq_len, 42, 20, C.byref(aa_off)) + 255) & ~255
            ring_a = out[: 2 * slot_a]
            counts = (C.c_size_t * 8)()
            L.qd_windows_all_batch_counts(seq_len, 42, 20, counts)
            cs = list(counts)
            wa_out = n_w * (2 * cs[0] * 42 + sum(cs[2:]) * 20)
            B.dev_error_reset(ctx)
            B.windows_all_chunked_dev(x, min(n_w, 2 * chunk_a), seq_len, ring_a, chunk_a, checksums=False, ctx=ctx)
            wa_ms = timed(lambda: B.windows_all_chunked_dev(x, n_w, seq_len, ring_a, chunk_a, checksums=False, ctx=ctx), reps=1, warm=0)
            assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True, ctx=ctx) is None
            r = leg("windows_all", f"all_windows shape (benches/all_windows.rs) for {n_w} x {seq_len}-nt strict sequences per GPU (config 5): {n_w * sum(cs)} windows, "
                    f"{wa_out / 1e9:.0f} GB through a ring of two {slot_a / 1e9:.2f} GB slots (per chunk: six-frame kernel + 8 window launches)",
                    wa_ms, n_w * seq_len + wa_out, "windows_batch_kernel (+ translate_frames_kernel<6>)", units={"windows": n_w * sum(cs), "nt": n_w * seq_len},
                    sequences_per_gpu=n_w, chunk_sequences=chunk_a)
            if do_checks:
                t0 = time.perf_counter()
                sums = B.windows_all_chunked_dev(x, n_w, seq_len, ring_a, chunk_a, checksums=True, ctx=ctx)
                want0 = oracle.check_windows_all(x[: chunk_a * seq_len].cpu().numpy(), chunk_a, seq_len, 42, 20, cpu_threads)
                ok = all_ranks_true((int(sums[0].item()) & M64) == want0)
                checks["config5_all_windows_chunk0_all_bytes_equal_oracle"] = ok
                r["checksum_all_chunks"] = f"{M(sums):#018x}"
                r["check_seconds"] = time.perf_counter() - t0
                assert ok
            # expansions of 42-nt windows with exactly two ambiguity codes, cap 16 (benches/expansions.rs:68-102 shape), the whole shard as windows
            n_win = (n_w * seq_len) // w
            win = x[: n_win * w].clone().view(n_win, w)
            g2 = torch.Generator(device=dev).manual_seed(11 + rank)
            codes = torch.tensor(list(b"WMYHRKDSVBN"), dtype=torch.uint8, device=dev)
            rows_idx = torch.arange(n_win, device=dev)
            p1 = torch.randint(0, w, (n_win,), generator=g2, device=dev)
            win[rows_idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
            p1 = (p1 + 1 + torch.randint(0, w - 1, (n_win,), generator=g2, device=dev)) % w
            win[rows_idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
            del p1, rows_idx
            chunk_e = min(args.expansion_chunk, n_win)
            slot_e = (L.qd_expansion_chunk_bytes(chunk_e, w, 16) + 255) & ~255
            ring_e = out[: 2 * slot_e]
            B.dev_error_reset(ctx)
            rows_t, _ = B.expansion_windows_chunked_dev(win[: min(n_win, 2 * chunk_e)], ring_e, chunk_e, 16, checksums=False, ctx=ctx)
            holder = {}

            def ex_step():
                holder["rows"], _s = B.expansion_windows_chunked_dev(win, ring_e, chunk_e, 16, checksums=False, ctx=ctx)

            ex_ms = timed(ex_step, reps=1, warm=0)
            assert B.dev_error_fetch(win.view(-1), ctx=ctx) is None
            n_rows = int(holder["rows"].sum().item())
            r = leg("expansion_windows", f"expansions of {n_win} 42-nt windows with exactly 2 ambiguity codes, cap 16, per GPU (config 5): {n_rows} rows = "
                    f"{n_rows * w / 1e9:.0f} GB through a ring of two {slot_e / 1e9:.2f} GB slots (per chunk: ONE kernel -- size_hint, row index and rows from one read of the windows)",
                    ex_ms, n_win * w + n_rows * w, "expansion_fused_tiles_kernel<ascii, 42>",
                    units={"rows": n_rows, "windows": n_win}, windows_per_gpu=n_win, chunk_windows=chunk_e)
            if do_checks:
                t0 = time.perf_counter()
                rows_c, sums = B.expansion_windows_chunked_dev(win, ring_e, chunk_e, 16, checksums=True, ctx=ctx)
                want0, want_rows0 = oracle.check_expansions(win[:chunk_e].cpu().numpy(), 16, cpu_threads)
                ok = all_ranks_true((int(sums[0].item()) & M64) == want0 and int(rows_c[0].item()) == want_rows0)
                checks["config5_expansions_chunk0_all_bytes_equal_oracle"] = ok
                r["checksum_all_chunks"] = f"{M(sums):#018x}"
                r["check_seconds"] = time.perf_counter() - t0
                assert ok
            del win

        # ---- BASELINE config 1 shape (one 29 901-nt genome, translate table 1) through the Python drop-in API:
        # latency of one call, host bytes in, host bytes out (launch + two copies; the reference's CPU path wins here)
        if rank == 0:
            import quickdna_b200 as qd
            from quickdna_b200.synth import synth_np

            genome = qd.DnaSequence(synth_np(1, 0, 29_901, b"acgt").tobytes())
            for _ in range(20):
                genome.translate(1)
                genome.reverse_complement()
            t0 = time.perf_counter()
            for _ in range(200):
                genome.translate(1)
            lat_tr = (time.perf_counter() - t0) / 200
            t0 = time.perf_counter()
            for _ in range(200):
                genome.reverse_complement()
            lat_rc = (time.perf_counter() - t0) / 200
            extra["covid_sized_calls"] = {"workload": "DnaSequence(29 901 nt).translate(1) / .reverse_complement(), Python API, host in / host out",
                                          "translate_ms_per_call": lat_tr * 1e3, "reverse_complement_ms_per_call": lat_rc * 1e3,
                                          "reference_readme_ms": {"translate": 0.02959, "reverse_complement": 0.02409,
                                                                  "note": "README.md:85,89, hardware unstated"}}
            if world == 1 and not args.no_cpu:
                raw = genome.seq
                t0 = time.perf_counter()
                for _ in range(200):
                    oracle.translate_bytes(1, raw)
                extra["covid_sized_calls"]["cpu_port_translate_ms_per_call"] = (time.perf_counter() - t0) / 200 * 1e3

        # ---- FASTA record scanner (src/fasta.rs) through the host-pointer call: text in, masks + records out
        if rank == 0 and world == 1:
            from quickdna_b200.synth import synth_np

            n_rec, lines_per, width = 20_000, 143, 70  # ~10 kb records, 70-column lines, ~200 MB of text
            body = synth_np(7, 0, n_rec * lines_per * width).reshape(n_rec, lines_per, width)
            lines = np.concatenate([body, np.full((n_rec, lines_per, 1), 10, np.uint8)], axis=2).reshape(n_rec, -1)
            hdr = np.frombuffer(b">synthetic record, 143 lines of 70 nt\n", dtype=np.uint8)
            text = np.concatenate([np.broadcast_to(hdr, (n_rec, hdr.size)), lines], axis=1).reshape(-1).tobytes()
            del body, lines
            t0 = time.perf_counter()
            scan = B.fasta_scan(text, ctx=ctx)  # first call: the context's device buffers grow to this size
            fa_first = time.perf_counter() - t0
            t0 = time.perf_counter()
            scan = B.fasta_scan(text, ctx=ctx)
            fa_s = time.perf_counter() - t0
            assert len(scan) == n_rec and scan.masks.size == n_rec * lines_per * width
            extra["fasta_scan"] = {"workload": f"{n_rec} records, {len(text)} bytes of FASTA text (70-column lines), host in / host out "
                                               "(pageable memory), one qd_fasta_scan call + Python record objects",
                                   "seconds": fa_s, "text_GB_per_s": len(text) / fa_s / 1e9, "first_call_seconds": fa_first}
            # FASTA text in, six frames of every record out: one fused call (masks stay on the device) against scan + batch translate
            B.fasta_translate_frames(text, ctx=ctx)
            t0 = time.perf_counter()
            fscan, fframes = B.fasta_translate_frames(text, ctx=ctx)
            ff_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            two = B.translate_frames_batch((scan.masks, scan.offsets), frames=6, enc=N.ENC_MASK, ctx=ctx)
            tb_s = time.perf_counter() - t0
            same = bool(np.array_equal(two.data, fframes.data))
            checks["fasta_translate_frames_equals_scan_then_translate"] = same
            assert same and len(fscan) == n_rec
            extra["fasta_translate_frames"] = {"workload": "the same text -> six frames of every record, host in / host out (pageable), one qd_fasta_translate_frames call",
                                               "seconds": ff_s, "text_GB_per_s": len(text) / ff_s / 1e9, "nt_per_s": scan.masks.size / ff_s,
                                               "scan_then_translate_seconds": fa_s + tb_s}
            del fscan, fframes, two
            if not args.no_cpu:
                t0 = time.perf_counter()
                ref = oracle.fasta_scan(text[: len(text) // 10])
                extra["fasta_scan"]["cpu_port_text_GB_per_s"] = (len(text) // 10) / (time.perf_counter() - t0) / 1e9
                assert len(ref) >= n_rec // 10
            del text, scan

    # ---- e2e: the host-pointer C-ABI call on pinned host buffers (H2D + kernels + D2H timed), every rank at once
    e2e = None
    if not args.no_e2e:
        er = min(args.e2e_reads, n_reads)
        per = 16 * B.runs(READ_LEN) * 6
        h_in = torch.empty(er * READ_LEN, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(er * per, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(x[: er * READ_LEN])
        torch.cuda.synchronize()
        err = N.QdErr()

        def e2e_step():
            rc = L.qd_translate_frames_batch_uniform(ctx.handle, args.table, N.ENC_ASCII, 0, 6, C.c_void_p(h_in.data_ptr()), er,
                                                     READ_LEN, C.c_void_p(h_out.data_ptr()), C.byref(err))
            ctx.check(rc, err)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": world * er * READ_LEN / e2e_s, "unit": UNIT, "h2d_bytes_per_step": er * READ_LEN,
               "d2h_bytes_per_step": er * per, "reads_per_step": er, "ms_per_step": e2e_s * 1e3,
               "api": "qd_translate_frames_batch_uniform (host pointers, pinned), one process per GPU",
               "pcie_GBps_per_gpu": {"h2d": er * READ_LEN / e2e_s / 1e9, "d2h": er * per / e2e_s / 1e9}}
        if do_checks and args.table == 1:
            # the result really is on the host, every byte of it: checksum of the host buffer against the oracle's
            got = oracle.checksum64(h_out.numpy(), 0) if er <= 2_000_000 else None
            if got is not None:
                want = oracle.check_frames_uniform(3, rank * n_reads, er, READ_LEN, 0, 6, cpu_threads)
                ok = all_ranks_true(got == want)
                checks["e2e_all_host_output_bytes_equal_oracle"] = ok
                assert ok
        # the same call on PAGEABLE host memory (a Rust Vec<u8> / numpy array): staged through pinned buffers by host threads
        # (default) and with the driver's own staged copies (staging off)
        if rank == 0 and world == 1 and not args.no_extra:
            p_in = h_in.numpy().copy()
            p_out = np.empty(er * per, dtype=np.uint8)
            pg = {}
            for label, threads in (("staged_by_host_threads", -1), ("driver_staged", 0)):
                assert L.qd_ctx_set_stage_threads(ctx.handle, threads) == 0

                def p_step():
                    rc = L.qd_translate_frames_batch_uniform(ctx.handle, args.table, N.ENC_ASCII, 0, 6, C.c_void_p(p_in.ctypes.data), er,
                                                             READ_LEN, C.c_void_p(p_out.ctypes.data), C.byref(err))
                    ctx.check(rc, err)

                p_step()
                t0 = time.perf_counter()
                for _ in range(2):
                    p_step()
                dt = (time.perf_counter() - t0) / 2
                pg[label] = {"value": er * READ_LEN / dt, "ms_per_step": dt * 1e3}
                if label == "staged_by_host_threads":
                    same = bool(np.array_equal(p_out, h_out.numpy()))
                    checks["e2e_pageable_output_equals_pinned_output"] = same
                    assert same
            assert L.qd_ctx_set_stage_threads(ctx.handle, -1) == 0
            e2e["pageable_host_memory"] = pg
            del p_in, p_out
        # beside it: what the box itself delivers for this copy pattern, and ONE call driving all the GPUs
        host_barrier()
        if rank == 0:
            e2e["copy_ceiling"] = copy_ceiling(world, er * READ_LEN >> 20)
            ceil = e2e["copy_ceiling"].get("procs" if world > 1 else "threads", {})
            if "h2d_GBps" in ceil:
                e2e["frac_of_copy_ceiling"] = (world * er * per / e2e_s / 1e9) / ceil["d2h_GBps"]
                e2e["limiter"] = ("device->host copies (2.016 B per nt) at the rate the box delivers while both directions run: the bare-copy "
                                  "microbench with the same chunking reaches the d2h_GBps above")
        host_barrier()
        if world > 1 and rank == 0:
            # one process, one call: qd_translate_frames_batch_multi partitions the batch over all N GPUs (other ranks idle)
            ctxs = [N.default_context(d) for d in range(world)]
            hm_in = torch.empty(world * er * READ_LEN, dtype=torch.uint8, pin_memory=True)
            hm_out = torch.empty(world * er * per, dtype=torch.uint8, pin_memory=True)
            for k in range(world):
                hm_in[k * er * READ_LEN:(k + 1) * er * READ_LEN].copy_(h_in)
            arr = N.context_array(ctxs)

            def multi_step():
                rc = L.qd_translate_frames_batch_multi(arr, world, args.table, N.ENC_ASCII, 0, 6, C.c_void_p(hm_in.data_ptr()), None, world * er, READ_LEN,
                                                       C.c_void_p(hm_out.data_ptr()), None, C.byref(err))
                ctx.check(rc, err)

            multi_step()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                multi_step()
            m_s = (time.perf_counter() - t0) / args.e2e_steps
            same = all(torch.equal(hm_out[k * er * per:(k + 1) * er * per], h_out) for k in range(world))
            checks["multi_gpu_call_equals_single_gpu_call"] = bool(same)
            e2e["single_process_multi_gpu_call"] = {"api": "qd_translate_frames_batch_multi (one call, one host thread per GPU)", "value": world * er * READ_LEN / m_s,
                                                    "ms_per_step": m_s * 1e3, "reads_per_step": world * er}
            del hm_in, hm_out
            assert same
        host_barrier()
        # config 2 end to end: one 250 Mb chromosome, host in / host out, through the drop-in call (chunked, copies overlapped)
        if rank == 0 and not args.no_extra:
            n_rc = min(args.revcomp_nt, h_in.numel())
            rc_out = h_out[:n_rc]
            res = {}
            pageable_src = torch.from_numpy(h_in[:n_rc].numpy().copy())
            for label, src in (("pinned", h_in[:n_rc]), ("pageable", pageable_src), ("pageable_driver_staged", pageable_src)):
                dst = rc_out if label == "pinned" else torch.empty(n_rc, dtype=torch.uint8)
                assert L.qd_ctx_set_stage_threads(ctx.handle, 0 if label == "pageable_driver_staged" else -1) == 0
                call = lambda: ctx.check(L.qd_revcomp(ctx.handle, N.ENC_ASCII, 0, C.c_void_p(src.data_ptr()), n_rc, C.c_void_p(dst.data_ptr()), C.byref(err)), err)  # noqa: E731
                call()
                t0 = time.perf_counter()
                for _ in range(3):
                    call()
                dt = (time.perf_counter() - t0) / 3
                res[label] = {"ms": dt * 1e3, "nt_per_s": n_rc / dt, "GBps_each_way": n_rc / dt / 1e9}
                if label == "pageable":
                    assert torch.equal(dst, rc_out)
            assert L.qd_ctx_set_stage_threads(ctx.handle, -1) == 0
            if do_checks:
                checks["config2_e2e_all_host_output_bytes_equal_oracle"] = oracle.checksum64(rc_out.numpy(), 0) == oracle.check_revcomp(3, rank * total_nt, n_rc, False, cpu_threads)
                assert checks["config2_e2e_all_host_output_bytes_equal_oracle"]
            extra["revcomp_e2e"] = {"workload": f"qd_revcomp of one {n_rc}-nt sequence, host in / host out (config 2 through the drop-in call; chunks of n/16 on three streams)",
                                    **res}
        del h_in, h_out

    # ---- CPU baseline beside it (rank 0, N=1 only): scalar port, one thread, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_reads or 200_000
        v1, s1 = cpu_all_frames(n_cpu, 1, 1)
        reps = max(1, min(20, int(8.0 / max(s1, 1e-3))))
        v1, s1 = cpu_all_frames(n_cpu, 1, reps)
        vN, sN = cpu_all_frames(n_cpu * min(threads_all, 16), threads_all, 1)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n_cpu} of the {CONFIG3_READS} reads x {READ_LEN} nt x {reps} reps ({s1:.1f} s), oracle/qd_oracle.c (C port of the Rust loops), one thread",
               "all_threads": {"value": vN, "cores": threads_all, "seconds": sN}}

    if ncu_legs:
        extra["_ncu_calls"] = ncu_calls
    fill_leg_traffic(extra)
    roofline["legs"] = legs
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"translate_all_frames (6 frames, NCBI table {args.table}), {n_reads} x {READ_LEN}-nt ACGT reads per GPU (config 3), ASCII in",
                       "reads_per_gpu": n_reads, "read_len": READ_LEN, "parallelism": f"{world} independent shards, no collective",
                       "l2": "inputs (10 GB) and outputs (20 GB) per step far exceed the 126 MB L2",
                       "checks": checks},
            # the long per-leg records first: a reader who keeps only the tail of the line still gets the roofline with
            # its compact per-leg table (revcomp, parse, windows ...), the CPU baseline and the end-to-end numbers
            "extra": extra, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
