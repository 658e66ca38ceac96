#!/usr/bin/env python
"""bench.py -- six-frame translation throughput of the quickdna hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A "step" is one pass of translate_all_frames over one batch of synthetic reads
(BASELINE.json config 3: 10 M x 1 kb reads per GPU, weak scaling, no data-path collective).

  value      nt/s, whole job, inputs resident in HBM (CUDA events around the K launches, max over ranks)
  e2e        nt/s through the host-pointer C-ABI call (qd_translate_frames_batch_uniform) with pinned
             host buffers: H2D + kernels + D2H inside the timed region
  roofline   algorithmic bytes (n + 2*sum floor((n-k)/3) = 2996 B per 1 kb read) / launch time vs the
             measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle's scalar restatement of the reference loops on a bounded sample (rank 0, N=1)
  extra      reverse complement of a 250 Mb sequence (BASELINE.json config 2), same accounting

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) instead.
"""

from __future__ import annotations

import argparse
import json
import os
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


def ncu_traffic(kernel: str, **match):
    """dram bytes per launch of `kernel` from the committed ncu capture, if it was taken on this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            e = json.load(f)[kernel]
        if all(e.get(k) == v for k, v in match.items()):
            return e["dram_read"] + e["dram_write"]
    except Exception:
        pass
    return None


def fill_leg_traffic(extra: dict):
    """roofline.traffic of the `extra` legs: dram bytes of ONE call of the leg (all its kernels) from the committed per-leg
    ncu capture (profiles/traffic_r01.json "legs", taken with QD_BENCH_NCU_LEGS=1), when the capture was taken on the same
    workload (same algorithmic bytes per call)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            legs = json.load(f).get("legs", {})
    except Exception:
        return

    def visit(name, node):
        if not isinstance(node, dict):
            return
        r = node.get("roofline")
        e = legs.get(name)
        if isinstance(r, dict) and r.get("traffic") is None and e and e.get("algorithmic_bytes_per_launch") == r.get("algorithmic_bytes_per_launch"):
            r["traffic"] = e["dram_read"] + e["dram_write"]
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

    for (s, c), buf in _CPU_INPUT.items():
        if s == seed and c >= count:
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

    dna = _cpu_input(seed, n_reads * READ_LEN)
    oracle.bench_all_frames_batch(0, dna[: 2000 * READ_LEN], 2000, READ_LEN, threads, 1)  # warm the tables / pages
    secs, _ = oracle.bench_all_frames_batch(0, dna, n_reads, READ_LEN, threads, reps)
    return n_reads * READ_LEN * reps / secs, secs


def run_reference(args):
    """--impl reference: the CPU restatement of the reference on all host threads (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g

    g.build()
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
        "config": {"workload": f"translate_all_frames, NCBI table 1, {n_reads} x {READ_LEN}-nt uniform ACGT reads per step "
                               "(bounded sample of BASELINE.json config 3)", "impl_detail":
                   "oracle/qd_oracle.c: scalar C restatement of src/rust_api.rs:263-270 after the parse of :310-322 "
                   "(Rust toolchain absent, reference not compilable here)", "wall_s": total_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_reads} reads x {READ_LEN} nt per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (weak scaling)")
    ap.add_argument("--e2e-reads", type=int, default=2_000_000, help="reads per GPU in the host-buffer (e2e) leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=0)
    ap.add_argument("--revcomp-nt", type=int, default=250_000_000)
    ap.add_argument("--expansion-windows", type=int, default=20_000_000, help="42-nt windows per GPU in the expansion leg")
    ap.add_argument("--windows-all-seqs", type=int, default=1000, help="99 999-nt sequences per GPU in the all_windows leg")
    ap.add_argument("--canon-seqs", type=int, default=4000, help="99 999-nt sequences per GPU in the canonical-window leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--table", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

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
    numa = None
    try:  # run (and allocate pinned memory) on the CPUs next to this rank's GPU: the e2e leg is PCIe / host-memory bound
        import pynvml

        pynvml.nvmlInit()
        before_aff = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = {"cpus_before": before_aff, "cpus_after": len(os.sched_getaffinity(0))}
    except Exception as ex:  # not fatal: the numbers are simply taken without the binding
        numa = {"error": type(ex).__name__}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    from quickdna_b200 import _native as N
    from quickdna_b200 import batch as B
    from quickdna_b200.synth import synth_torch

    ctx = N.default_context(local_rank)
    L = N.lib()
    n_reads = args.reads
    total_nt = n_reads * READ_LEN

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("translate_frames_kernel<6>", reads=n_reads, read_len=READ_LEN),
                "kernel": "translate_frames_kernel<6, 1024, 0>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n_reads * ALGO_BYTES_PER_READ}

    # ---- extra: reverse complement of one 250 Mb sequence (config 2), device-resident
    extra = {}
    if not args.no_extra:
        n_rc = args.revcomp_nt
        bufs = [(x[i * n_rc:(i + 1) * n_rc], out[i * n_rc:(i + 1) * n_rc]) for i in range(4)]  # rotate 4 sets (> L2)
        for i in range(3):
            B.revcomp_dev(bufs[i % 4][0], out=bufs[i % 4][1], ctx=ctx)
        barrier()
        reps = 20
        ev0.record()
        for i in range(reps):
            B.revcomp_dev(bufs[i % 4][0], out=bufs[i % 4][1], ctx=ctx)
        ev1.record()
        barrier()
        rc_ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
        rc_gbs = 2 * n_rc / (rc_ms * 1e-3) / 1e9
        extra["revcomp"] = {"workload": f"reverse_complement, one {n_rc}-nt ACGT sequence per GPU, ASCII non-strict",
                            "nt_per_s": world * n_rc / (rc_ms * 1e-3), "ms": rc_ms,
                            "roofline": {"bound": "hbm", "achieved": rc_gbs, "peak": peak, "unit": "GB/s", "frac": rc_gbs / peak,
                                         "traffic": ncu_traffic("revcomp_kernel", nt=n_rc), "kernel": "revcomp_kernel<1024>",
                                         "algorithmic_bytes_per_launch": 2 * n_rc}}

    # ---- extra: reverse-complement variants: misaligned length, strict alphabet, mask encoding
    if not args.no_extra:
        from quickdna_b200 import _native as _N

        n_rc = args.revcomp_nt
        variants = {}
        for name, n_v, kw in (("n_plus_1", n_rc + 1, {}), ("n_minus_1", n_rc - 1, {}), ("strict", n_rc, {"strict": True})):
            bufs = [(x[i * (n_rc + 16): i * (n_rc + 16) + n_v], out[i * (n_rc + 16): i * (n_rc + 16) + n_v]) for i in range(4)]
            for i in range(4):
                B.revcomp_dev(bufs[i][0], out=bufs[i][1], ctx=ctx, **kw)
            barrier()
            ev0.record()
            for i in range(20):
                B.revcomp_dev(bufs[i % 4][0], out=bufs[i % 4][1], ctx=ctx, **kw)
            ev1.record()
            barrier()
            v_ms = max_over_ranks(ev0.elapsed_time(ev1)) / 20
            variants[name] = {"nt": n_v, "ms": v_ms, "GB/s": 2 * n_v / (v_ms * 1e-3) / 1e9}
        assert B.dev_error_fetch(x, ctx=ctx) is None
        # mask encoding (the typed Rust API: Vec<Nucleotide>): parse the ASCII once on the device, then reverse-complement masks
        masks, cnt = B.parse_dev(x[:n_rc], ctx=ctx)
        assert int(cnt.item()) == n_rc
        mo = out[:n_rc]
        for _ in range(3):
            B.revcomp_dev(masks[:n_rc], out=mo, enc=_N.ENC_MASK, ctx=ctx)
        barrier()
        ev0.record()
        for _ in range(10):
            B.revcomp_dev(masks[:n_rc], out=mo, enc=_N.ENC_MASK, ctx=ctx)
        ev1.record()
        barrier()
        v_ms = max_over_ranks(ev0.elapsed_time(ev1)) / 10
        variants["mask_encoding"] = {"nt": n_rc, "ms": v_ms, "GB/s": 2 * n_rc / (v_ms * 1e-3) / 1e9, "note": "one buffer set reused (partly L2-resident)"}
        # ASCII -> mask parse itself (src/rust_api.rs:310-322), 2 B per input byte
        for _ in range(3):
            B.parse_dev(x[:n_rc], out=masks, ctx=ctx)
        barrier()
        ev0.record()
        for _ in range(20):
            B.parse_dev(x[:n_rc], out=masks, ctx=ctx)
        ev1.record()
        barrier()
        v_ms = max_over_ranks(ev0.elapsed_time(ev1)) / 20
        if ncu_legs:
            torch.cuda.profiler.start()
            B.parse_dev(x[:n_rc], out=masks, ctx=ctx)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        p_gbs = 2 * n_rc / (v_ms * 1e-3) / 1e9
        extra["parse"] = {"workload": f"ASCII -> mask parse of {n_rc} nt (count pass + scan + compacting write pass)", "ms": v_ms,
                          "nt_per_s": world * n_rc / (v_ms * 1e-3), "GB/s_algorithmic": p_gbs,
                          "roofline": {"bound": "hbm", "achieved": p_gbs, "peak": peak, "unit": "GB/s", "frac": p_gbs / peak, "traffic": None,
                                       "kernel": "parse_count_kernel + scan + parse_write_kernel", "algorithmic_bytes_per_launch": 2 * n_rc}}
        extra["revcomp_variants"] = variants
        del masks, cnt

    # ---- extra: the other forms of the frame kernel on the same resident data (a few launches each)
    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        barrier()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        barrier()
        t_ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
        if ncu_legs:  # one more call, alone inside a profiler range: `ncu --profile-from-start off` sees each leg's kernels in order
            torch.cuda.profiler.start()
            fn()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return t_ms

    if not args.no_extra:
        def leg(name, workload, ms, nt, algo_bytes, kernel, traffic=None):
            gbs = algo_bytes / (ms * 1e-3) / 1e9
            extra[name] = {"workload": workload, "nt_per_s": world * nt / (ms * 1e-3), "ms": ms,
                           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
                                        "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes}}

        n_v = min(n_reads, 4_000_000)
        nt_v = n_v * READ_LEN
        f1 = sum([(READ_LEN - 0) // 3])
        f3 = sum((READ_LEN - k) // 3 for k in range(3))
        ms1 = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=1, ctx=ctx))
        leg("translate_1_frame", f"translate (1 frame) on {n_v} x {READ_LEN}-nt reads per GPU", ms1, nt_v, n_v * (READ_LEN + f1),
            "translate_frames_kernel<1, 1024, direct>")
        ms3 = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=3, ctx=ctx))
        leg("translate_self_frames", f"translate_self_frames (3 frames) on {n_v} x {READ_LEN}-nt reads per GPU", ms3, nt_v,
            n_v * (READ_LEN + f3), "translate_frames_kernel<3, 1024>")
        mss = timed(lambda: B.translate_frames_uniform_dev(x[:nt_v], n_v, READ_LEN, out=out, table=args.table, frames=6, strict=True, ctx=ctx))
        leg("all_frames_strict", f"translate_all_frames, strict alphabet check, {n_v} x {READ_LEN}-nt ACGT reads per GPU", mss, nt_v,
            n_v * ALGO_BYTES_PER_READ, "translate_frames_kernel<6, 1024>")
        # variable-length batch (lengths uniform in 500..1500, seed 4): CSR offsets on the device, scan + tile map included
        g = torch.Generator(device="cpu").manual_seed(4 + rank)
        lens = torch.randint(500, 1501, (n_v,), generator=g, dtype=torch.int64)
        offs = torch.zeros(n_v + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(lens, 0)
        keep = int(torch.searchsorted(offs, torch.tensor([min(nt_v, total_nt)]), right=True).item()) - 1
        offs_d = offs[: keep + 1].to(dev)
        tot_v = int(offs[keep].item())
        runs_v = int(((lens[:keep] + 47) // 48).sum().item())
        algo_v = int((lens[:keep] + 2 * ((lens[:keep]) // 3 + (lens[:keep] - 1) // 3 + (lens[:keep] - 2) // 3)).sum().item())
        launches_v0 = ctx.launch_count
        msv = timed(lambda: B.translate_frames_csr_dev(x[:tot_v], offs_d, out=out[: runs_v * 96], table=args.table, frames=6, ctx=ctx))
        leg("all_frames_variable_length", f"translate_all_frames on {keep} reads of 500..1500 nt (CSR offsets, prefix scan and tile map "
            "kernels inside the timed region)", msv, tot_v, algo_v, "translate_frames_kernel<6, 1024> + scan/tile-map")
        extra["all_frames_variable_length"]["launches_per_call"] = (ctx.launch_count - launches_v0) // 7
        # IUPAC-ambiguous, mixed-case input (config 4): 5 % ambiguity codes, tables 1 / 11 / 22
        from quickdna_b200.synth import synth_iupac_np

        base = torch.from_numpy(synth_iupac_np(5, rank * 50_000_000, 50_000 * READ_LEN)).to(dev)
        n_a = 1_000_000
        xa = base.repeat(n_a // 50_000)
        for tbl in (1, 11, 22):
            msa = timed(lambda: B.translate_frames_uniform_dev(xa, n_a, READ_LEN, out=out, table=tbl, frames=6, ctx=ctx), reps=3)
            leg(f"all_frames_iupac_t{tbl}", f"translate_all_frames, NCBI table {tbl}, {n_a} x {READ_LEN}-nt reads with 5 % IUPAC "
                "ambiguity codes, 50 % lower case, non-strict", msa, n_a * READ_LEN, n_a * ALGO_BYTES_PER_READ,
                "translate_frames_kernel<6, 1024>")
        assert B.dev_error_fetch(xa, uniform_len=READ_LEN, ctx=ctx) is None
        B.dev_error_reset(ctx)
        B.translate_frames_uniform_dev(xa, n_a, READ_LEN, out=out, table=1, frames=6, strict=True, ctx=ctx)
        e = B.dev_error_fetch(xa, uniform_len=READ_LEN, strict=True, ctx=ctx)
        extra["all_frames_iupac_strict_error"] = None if e is None else {"kind": e.kind, "byte": chr(e.byte), "seq_index": e.seq_index,
                                                                       "pos": e.pos, "message": str(e)}
        B.dev_error_reset(ctx)
        del xa, base

    # ---- extra: canonical 42-nt windows (config 5 shape: 99 999-nt strict sequences), device-resident
    if not args.no_extra:
        seq_len, w = 99_999, 42
        n_seq = min(args.canon_seqs, total_nt // seq_len, out.numel() // ((seq_len - w + 1) * w))
        if n_seq > 0:
            nwin = n_seq * (seq_len - w + 1)
            cw_ms = timed(lambda: B.canonical_windows_batch_dev(x, n_seq, seq_len, w, out=out, ctx=ctx))
            cw_gbs = nwin * (1 + w) / (cw_ms * 1e-3) / 1e9
            extra["canonical_windows"] = {
                "workload": f"canonical 42-nt windows of {n_seq} x {seq_len}-nt ACGT sequences per GPU (benches/canonical.rs shape)",
                "windows_per_s": world * nwin / (cw_ms * 1e-3), "ms": cw_ms,
                "roofline": {"bound": "hbm", "achieved": cw_gbs, "peak": peak, "unit": "GB/s", "frac": cw_gbs / peak, "traffic": None,
                             "kernel": "canonical_tile_kernel<3, 42, rows>", "algorithmic_bytes_per_launch": nwin * (1 + w),
                             "note": "issue-bound (about 280 thread-instructions per 42-byte window), not HBM-bound"}}
            # the same windows as fixed-width keys: 64-bit fingerprints, then lossless 16-byte packed windows (two bit planes)
            keys = {}
            for kind, label, kb in ((N.CANON_KEY_FP64, "fingerprint_64", 8), (N.CANON_KEY_PACKED, "packed_2bit", 16)):
                kout = out.view(torch.int32)
                k_ms = timed(lambda: B.canonical_window_keys_dev(x, n_seq, seq_len, w, kind=kind, out=kout, ctx=ctx))
                k_gbs = nwin * (1 + kb) / (k_ms * 1e-3) / 1e9
                keys[label] = {"ms": k_ms, "windows_per_s": world * nwin / (k_ms * 1e-3), "key_bytes": kb,
                               "roofline": {"bound": "hbm", "achieved": k_gbs, "peak": peak, "unit": "GB/s", "frac": k_gbs / peak, "traffic": None,
                                            "kernel": f"canonical_tile_kernel<3, 42, {kind}>", "algorithmic_bytes_per_launch": nwin * (1 + kb),
                                            "note": "ALU-bound (about 150 thread-instructions per window), not HBM-bound"}}
            extra["canonical_window_keys"] = {
                "workload": f"one key per canonical 42-nt window of the same {n_seq} x {seq_len}-nt sequences (SURVEY 8(f) rank 3)",
                "speedup_vs_materialised_rows": {k: cw_ms / v["ms"] for k, v in keys.items()}, **keys}
            if rank == 0 and world == 1 and not args.no_cpu:
                import oracle

                m = oracle.masks_of(x[: seq_len].cpu().numpy())
                t0 = time.perf_counter()
                n_rep = 0
                while time.perf_counter() - t0 < 2.0:
                    oracle.canonical_windows(m, w)
                    n_rep += 1
                extra["canonical_windows"]["cpu_port"] = {"windows_per_s": n_rep * (seq_len - w + 1) / (time.perf_counter() - t0), "cores": 1,
                                                          "sample": f"{n_rep} x one {seq_len}-nt sequence, oracle/qd_oracle.c"}

    # ---- extra: the all_windows shape (benches/all_windows.rs: 42-nt DNA windows of both strands + 20-aa windows of the
    # six frames) materialised for a batch of 99 999-nt sequences: 124 output bytes per nucleotide
    if not args.no_extra:
        seq_len, n_sw = 99_999, min(args.windows_all_seqs, total_nt // 99_999)
        if n_sw > 0:
            counts = (__import__("ctypes").c_size_t * 8)()
            L.qd_windows_all_batch_counts(seq_len, 42, 20, counts)
            cs = list(counts)
            dna_bytes, aa_bytes = 2 * n_sw * cs[0] * 42, n_sw * sum(cs[2:]) * 20
            if dna_bytes + aa_bytes + 64 <= out.numel():
                d_out, a_out = out[:dna_bytes], out[(dna_bytes + 15) // 16 * 16:(dna_bytes + 15) // 16 * 16 + aa_bytes]
                B.dev_error_reset(ctx)
                wa_ms = timed(lambda: B.windows_all_batch_dev(x, n_sw, seq_len, 42, 20, dna_out=d_out, aa_out=a_out, ctx=ctx), reps=3)
                assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True, ctx=ctx) is None
                wa_bytes = n_sw * seq_len + dna_bytes + aa_bytes
                gbs = wa_bytes / (wa_ms * 1e-3) / 1e9
                extra["windows_all"] = {
                    "workload": f"all_windows shape for {n_sw} x {seq_len}-nt strict sequences per GPU: {n_sw * sum(cs)} windows "
                                "(six-frame kernel + 2 DNA-window launches + 6 amino-acid-window launches)",
                    "nt_per_s": world * n_sw * seq_len / (wa_ms * 1e-3), "windows_per_s": world * n_sw * sum(cs) / (wa_ms * 1e-3), "ms": wa_ms,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                                 "kernel": "windows_batch_kernel (+ translate_frames_kernel<6>)", "algorithmic_bytes_per_launch": wa_bytes}}

    # ---- extra: BASELINE config 1 shape (one 29 901-nt genome, translate table 1) through the Python drop-in API:
    # latency of one call, host bytes in, host bytes out (launch + two copies; the reference's CPU path wins here)
    if not args.no_extra and rank == 0:
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
            import oracle

            raw = genome.seq
            t0 = time.perf_counter()
            for _ in range(200):
                oracle.translate_bytes(1, raw)
            extra["covid_sized_calls"]["cpu_port_translate_ms_per_call"] = (time.perf_counter() - t0) / 200 * 1e3

    # ---- extra: FASTA record scanner (src/fasta.rs) through the host-pointer call: text in, masks + records out
    if not args.no_extra and rank == 0 and world == 1:
        import numpy as np

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
        if not args.no_cpu:
            import oracle

            t0 = time.perf_counter()
            ref = oracle.fasta_scan(text[: len(text) // 10])
            extra["fasta_scan"]["cpu_port_text_GB_per_s"] = (len(text) // 10) / (time.perf_counter() - t0) / 1e9
            assert len(ref) >= n_rec // 10
        del text, scan

    # ---- extra: expansions of 42-nt windows with exactly two ambiguity codes, cap 16 (benches/expansions.rs:68-102 shape)
    if not args.no_extra:
        n_win, w = args.expansion_windows, 42
        if n_win > 0 and n_win * w <= total_nt:
            win = x[: n_win * w].clone().view(n_win, w)
            g2 = torch.Generator(device=dev).manual_seed(11 + rank)
            codes = torch.tensor(list(b"WMYHRKDSVBN"), dtype=torch.uint8, device=dev)
            p1 = torch.randint(0, w, (n_win,), generator=g2, device=dev)
            p2 = (p1 + 1 + torch.randint(0, w - 1, (n_win,), generator=g2, device=dev)) % w
            rows_idx = torch.arange(n_win, device=dev)
            win[rows_idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
            win[rows_idx, p2] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
            del p1, p2, rows_idx
            B.dev_error_reset(ctx)
            c_d, oi_d, _ = B.expansion_windows_dev(win, 16, out=out, ctx=ctx)
            n_rows = int(oi_d[n_win].item())
            assert B.dev_error_fetch(win.view(-1), ctx=ctx) is None and n_rows * w <= out.numel()
            ex_ms = timed(lambda: B.expansion_windows_dev(win, 16, out=out, ctx=ctx))
            ex_bytes = n_win * (w + 16) + n_rows * w
            ex_gbs = ex_bytes / (ex_ms * 1e-3) / 1e9
            extra["expansion_windows"] = {
                "workload": f"expansions of {n_win} 42-nt windows with exactly 2 ambiguity codes, cap 16, per GPU: {n_rows} rows "
                            "(size_hint kernel + scan + write kernel per call)",
                "rows_per_s": world * n_rows / (ex_ms * 1e-3), "windows_per_s": world * n_win / (ex_ms * 1e-3), "ms": ex_ms,
                "roofline": {"bound": "hbm", "achieved": ex_gbs, "peak": peak, "unit": "GB/s", "frac": ex_gbs / peak, "traffic": None,
                             "kernel": "expansion_count_tiles_kernel<ascii, 42> + scan + expansion_write_tiles_kernel<42>", "algorithmic_bytes_per_launch": ex_bytes}}
            del win, c_d, oi_d

    # ---- e2e: the host-pointer C-ABI call on pinned host buffers (H2D + kernels + D2H timed)
    e2e = None
    if not args.no_e2e:
        import ctypes as C

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
        # the result really is on the host: compare some reads against a device-resident run on the same bytes
        chk = B.translate_frames_uniform_dev(x[: 64 * READ_LEN], 64, READ_LEN, table=args.table, frames=6, ctx=ctx)
        assert bytes(h_out[: 64 * per].numpy()) == bytes(chk.cpu().numpy())
        tail = B.translate_frames_uniform_dev(x[(er - 64) * READ_LEN: er * READ_LEN], 64, READ_LEN, table=args.table, frames=6, ctx=ctx)
        assert bytes(h_out[(er - 64) * per: er * per].numpy()) == bytes(tail.cpu().numpy())
        e2e = {"value": world * er * READ_LEN / e2e_s, "unit": UNIT, "h2d_bytes_per_step": er * READ_LEN,
               "d2h_bytes_per_step": er * per, "reads_per_step": er, "ms_per_step": e2e_s * 1e3,
               "api": "qd_translate_frames_batch_uniform (host pointers, pinned)", "cpu_affinity": numa}
        del h_in, h_out

    # ---- CPU baseline beside it (rank 0, N=1 only): scalar port, one thread, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_reads or 200_000
        v1, s1 = cpu_all_frames(n_cpu, 1, 1)
        reps = max(1, min(20, int(10.0 / max(s1, 1e-3))))
        v1, s1 = cpu_all_frames(n_cpu, 1, reps)
        threads = host_threads()
        vN, sN = cpu_all_frames(n_cpu * min(threads, 16), threads, 1)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n_cpu} reads x {READ_LEN} nt x {reps} reps ({s1:.1f} s), oracle/qd_oracle.c single thread",
               "all_threads": {"value": vN, "cores": threads, "seconds": sN}}

    fill_leg_traffic(extra)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"translate_all_frames (6 frames, NCBI table {args.table}) on {n_reads} x {READ_LEN}-nt uniform "
                                   "ACGT reads per GPU (BASELINE.json config 3), ASCII in, non-strict",
                       "reads_per_gpu": n_reads, "read_len": READ_LEN, "parallelism": f"{world} independent shards, no collective",
                       "l2": "inputs (10 GB) and outputs (20 GB) per step far exceed the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
