#!/usr/bin/env python
"""bench.py -- read x path likelihood evaluations per second of the HaploCart hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl reference]

A "step" is one whole sample through the hot path: reset the accumulators, run the read kernel over
the batch, collapse into the per-path vector (bit-GEMV), (N > 1: all-gather the per-rank vectors over
NCCL and sum them in rank order), argmax + log-softmax posterior + clade support, and read the result
back.  `value` times that with the batch already resident in HBM; `e2e` times the same call sequence
through the host-buffer entry point (pinned host SoA -> H2D -> kernels -> D2H) every step.

Workloads:  N = 1 -> config 2 (10^6 x 100 bp reads, 11825 x 5179 tables);
N > 1 -> config 3 shape (ancient-damaged fragments, log-normal lengths), R reads per GPU (weak scaling).
The batch (~0.6 GB) is larger than the 126 MB L2, so every step streams it from HBM.

`--impl reference` times the CPU restatement of the reference's own per-path loops (oracle, literal
mode; the reference itself cannot be compiled here) on all host cores, on a bounded sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read x path likelihood evaluations per second"
UNIT = "evals/s"
N_PATHS, N_ROWS, N_NODES = 5179, 11825, 11821
TABLE_SEED = 20260


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def workload(n_gpus, reads, config=2):
    """Config 2 (BASELINE.json configs[1]) is the bench workload at every N: the same reads-per-GPU, so the per-N values
    are one weak-scaling series.  --workload 3 selects the config-3 shape (ancient-damaged reads, log-normal lengths)."""
    if config == 2:
        return dict(name="HaploCart config 2: synthetic %d x 100bp reads per GPU x %d paths" % (reads, N_PATHS), seed=20261,
                    kw=dict(fixed_len=100, sub_rate=0.005, indel_rate=0.0005, damage=False))
    return dict(name="HaploCart config 3 shape: synthetic ancient-damaged reads (log-normal 30-150bp), %d per GPU x %d "
                     "paths" % (reads, N_PATHS), seed=20263,
                kw=dict(fixed_len=0, sub_rate=0.005, indel_rate=0.0005, damage=True))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_rate(S, alns, seconds_target, threads=0, hoisted=False):
    """Time the literal (reference-shaped) CPU restatement on a bounded sample; returns a dict."""
    from oracle import oracle as orc
    node, val = S.pmap()
    T = orc.HcTables(S.supports(), node, val, S.mappab(), min_id=1)
    G = orc.Graph(S.node_sequences(), min_id=1)
    R = orc.FlatReads(alns.arrays())
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    # calibrate on a few reads per thread, then size the sample for ~seconds_target of wall clock
    n0 = min(len(alns), 2 * cores)
    t = time.perf_counter()
    _, _, used0 = orc.hc_run(T, G, R, r0=0, r1=n0, mode=orc.MODE_LITERAL, n_threads=cores)
    dt0 = max(time.perf_counter() - t, 1e-6)
    n = int(min(len(alns), max(n0, n0 * seconds_target / dt0)))
    t = time.perf_counter()
    _, argmax, used = orc.hc_run(T, G, R, r0=0, r1=n, mode=orc.MODE_LITERAL, n_threads=cores)
    dt = time.perf_counter() - t
    out = {"value": used * N_PATHS / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "first %d reads of the same synthetic set, literal per-path loops (oracle MODE_LITERAL, long double, "
                     "-O3 -fopenmp), %.1f s wall" % (n, dt),
           "reads": used, "seconds": dt}
    if hoisted:
        # the fair optimised-CPU comparator: A_i / B_i once per mapping, then the per-path select-add
        nh = int(min(len(alns), 20 * n))
        t = time.perf_counter()
        _, _, uh = orc.hc_run(T, G, R, r0=0, r1=nh, mode=orc.MODE_HOISTED, n_threads=cores)
        dth = time.perf_counter() - t
        out["hoisted"] = {"value": uh * N_PATHS / dth, "reads": uh, "seconds": dth,
                          "what": "same oracle with A_i / B_i computed once per mapping (MODE_HOISTED)"}
    return out


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's own loop nest, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import vgan_b200 as vb
    from oracle import oracle as orc
    wl = workload(args.gpus, args.reads, args.workload)
    S = vb.SynthHaplocart(seed=TABLE_SEED, n_paths=N_PATHS, n_rows=N_ROWS, n_nodes=N_NODES)
    cores = os.cpu_count() or 1
    # bounded sample: sized so that warmup + steps end within a few minutes
    probe_reads = S.reads(4 * cores, seed=wl["seed"], **wl["kw"])
    probe = cpu_reference_rate(S, probe_reads, seconds_target=2.0)
    per_read = probe["seconds"] / max(probe["reads"], 1)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    n = int(max(cores, min(200000, budget / per_read)))
    alns = S.reads(n, seed=wl["seed"], **wl["kw"])
    node, val = S.pmap()
    T = orc.HcTables(S.supports(), node, val, S.mappab(), min_id=1)
    G = orc.Graph(S.node_sequences(), min_id=1)
    R = orc.FlatReads(alns.arrays())
    for _ in range(args.warmup):
        orc.hc_run(T, G, R, mode=orc.MODE_LITERAL, n_threads=cores)
    t = time.perf_counter()
    used = 0
    for _ in range(args.steps):
        _, _, u = orc.hc_run(T, G, R, mode=orc.MODE_LITERAL, n_threads=cores)
        used += u
    dt = time.perf_counter() - t
    value = used * N_PATHS / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f80 (long double)", "data": "synthetic",
            "config": {"workload": wl["name"], "sample_reads_per_step": n, "paths": N_PATHS},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d reads per step of the same synthetic workload, literal per-path loops "
                                       "(oracle MODE_LITERAL; the reference itself cannot be built here)" % n},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU")
    ap.add_argument("--impl", default="vgan_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU exchange of the P-vector")
    ap.add_argument("--workload", type=int, default=2, choices=[2, 3], help="SURVEY 8d config whose read shape is generated")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import vgan_b200 as vb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    wl = workload(n_gpus, args.reads, args.workload)
    S = vb.SynthHaplocart(seed=TABLE_SEED, n_paths=N_PATHS, n_rows=N_ROWS, n_nodes=N_NODES)
    alns = S.reads(args.reads, seed=wl["seed"] + 1000 * rank, **wl["kw"])
    hc = vb.Haplocart(device=local)
    hc.load(S.tables(), S.tree())
    # every kernel, copy and collective of a step runs on this stream, and so do the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hc.set_stream(stream.cuda_stream)
    t_pack = time.perf_counter()
    packed = vb.PackedBatch(S.graph, alns, pinned=True)
    t_pack = time.perf_counter() - t_pack
    dev = hc.upload(packed)
    # the e2e route ships the batch in its wire form (12- or 16-bit column records + one node id per mapping) and expands it on
    # the device; graphs with node ids above 65535 would ship the 32-bit batch itself
    try:
        wire = vb.WireBatch(packed, pinned=True)
    except RuntimeError:
        wire = None
    host_batch = wire if wire is not None else packed
    reads_local = packed.n_reads
    batch_bytes = packed.nbytes
    # algorithmic bytes of the read kernel: 4 per column + 4 per empty mapping + 7 per read + 8 per group
    pa = packed.arrays()
    alg_bytes = int(4 * pa["n_cols"].astype(np.int64).sum() + 4 * pa["n_extra"].astype(np.int64).sum() + 7 * reads_local
                    + 8 * (int(packed.batch.n_groups) + 1))
    del pa

    sharded = None
    if world > 1:
        from vgan_b200.dist import ShardedSample
        # exchange of the per-rank final_vec + fixed-order sum (vgan_b200/dist.py): peer-memory kernels, NCCL as fallback
        sharded = ShardedSample(hc, exchange=args.exchange)

    def step(host_buffers):
        if sharded is not None:
            return sharded.run([host_batch] if host_buffers else [dev], device_resident=not host_buffers)
        hc.reset()
        if host_buffers:
            if wire is not None:
                hc.update_wire(wire)
            else:
                hc.update(packed)
        else:
            hc.update_device(dev.batch)
        return hc.get_posterior(want_ll=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host_buffers, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_ms, k_n = 0.0, 0
        barrier()
        l0 = hc.launch_count()
        e0.record(stream)
        for _ in range(steps):
            res = step(host_buffers)
            ms, n = hc.read_kernel_ms()
            k_ms += ms
            k_n += n
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total, k_ms, k_n, hc.launch_count() - l0, res

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, k_ms, k_n, launches, res = timed(False, args.steps)
    for _ in range(3):
        step(True)
    e2e_steps = max(3, min(args.steps, 20))
    ms_e2e, _, _, _, res_e2e = timed(True, e2e_steps)
    clocks = sampler.finish()

    # the same sample from RAW alignments: H2D of the compact alignment arrays, reconstruction + packing on the device
    # (SURVEY 8f rank 1), kernels, result -- against the host packer's rate reported in `host_pack`
    raw_path = None
    if world == 1:
        t0 = time.perf_counter()
        raw = vb.RawBatch(alns, pinned=True)
        t_raw = time.perf_counter() - t0          # includes cudaHostAlloc of the whole block
        t_raw_pageable = []
        for _ in range(2):                        # second pass: the allocator hands back touched pages
            t0 = time.perf_counter()
            tmp = vb.RawBatch(alns, pinned=False)
            t_raw_pageable.append(time.perf_counter() - t0)
            del tmp
        hc.load_graph(S.graph)

        def raw_step():
            hc.reset()
            hc.update_raw(raw)
            return hc.get_posterior(want_ll=False)

        for _ in range(3):
            res_raw = raw_step()
        raw_steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(raw_steps):
            res_raw = raw_step()
        e1.record(stream)
        barrier()
        ms_raw = e0.elapsed_time(e1) / raw_steps
        raw_path = {"value": raw.n_reads * N_PATHS / (ms_raw * 1e-3), "unit": UNIT, "ms_per_step": ms_raw, "steps": raw_steps,
                    "h2d_bytes_per_step": raw.nbytes, "host_build_seconds": t_raw,
                    "host_build_seconds_pageable": t_raw_pageable,
                    "same_call_as_packed_route": res_raw.get("maxh_index") == res.get("maxh_index"),
                    "what": "vgl_hc_accumulate_raw: pinned raw alignments -> H2D -> device reconstruction + packing -> read kernel "
                            "-> posterior; host_build_seconds is the one-pass copy into the compact raw arrays (no reconstruction) incl. the pinned allocation; _pageable: the same into malloc'ed memory, cold and warm"}
        del raw

    # per-sample latency at the size one real sample can reach after duplicate removal (SURVEY 8d scale caveat):
    # host batch -> H2D -> kernels -> posterior on the host, one sample at a time
    latency = None
    if world == 1:
        n_lat = min(20000, len(alns))
        small = vb.PackedBatch(S.graph, alns, r0=0, r1=n_lat, pinned=True)
        for _ in range(3):
            hc.run_sample([small], want_ll=False)
        t0 = time.perf_counter()
        for _ in range(20):
            hc.run_sample([small], want_ll=False)
        latency = {"reads": small.n_reads, "ms_per_sample": (time.perf_counter() - t0) / 20 * 1e3,
                   "what": "reset + accumulate (host batch) + finalize, host wall clock, mean of 20"}
        # ... and from the alignments themselves, host marshalling included: the host packer route against the raw route
        def from_alignments(raw_route):
            hc.reset()
            if raw_route:
                hc.update_raw(vb.RawBatch(alns, r0=0, r1=n_lat))
            else:
                hc.update(vb.PackedBatch(S.graph, alns, r0=0, r1=n_lat))
            return hc.get_posterior(want_ll=False)
        for key, flag in (("ms_per_sample_from_alignments_host_packer", False), ("ms_per_sample_from_alignments_raw_route", True)):
            for _ in range(3):
                from_alignments(flag)
            t0 = time.perf_counter()
            for _ in range(10):
                from_alignments(flag)
            latency[key] = (time.perf_counter() - t0) / 10 * 1e3

    total_reads = reads_local
    if world > 1:
        t = torch.tensor([reads_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        total_reads = int(t.item())

    if rank == 0:
        peak, peak_kind = peaks()
        value = total_reads * N_PATHS * args.steps / (ms_dev * 1e-3)
        e2e = total_reads * N_PATHS * e2e_steps / (ms_e2e * 1e-3)
        kernel_ms = k_ms / max(k_n, 1)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("reads") == reads_local and tj.get("kernel") == "hc_read_kernel":
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "reads_per_gpu": reads_local, "paths": N_PATHS, "rows": N_ROWS,
                       "l2": "batch of %.0f MB per step exceeds the 126 MB L2 (no flush needed)" % (batch_bytes / 1e6),
                       "parallelism": "reads sharded over %d GPU(s); one exchange of P doubles per sample (%s)" % (
                           n_gpus, "none" if sharded is None else sharded.exchange),
                       "called_haplogroup": res.get("predicted_haplotype")},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": host_batch.nbytes,
                    "route": ("vgl_hc_accumulate_wire (wire form, %s columns, expanded on the device)" % ("12-bit" if wire.wire.col8 else "16-bit")) if wire is not None else "vgl_hc_accumulate", "d2h_bytes_per_step": 28 + 8 * hc._max_post,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "raw_path": raw_path,
            "latency": latency,
            "host_pack": {"seconds": t_pack, "reads_per_s": reads_local / t_pack, "threads": os.cpu_count(),
                          "what": "vgh_hc_pack (reconstruction + packing, OpenMP) of the step's batch; outside the timed regions"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "hc_read_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes, "batch_bytes_with_padding": batch_bytes,
                         "kernel_ms": kernel_ms, "bytes_per_read": alg_bytes / max(reads_local, 1),
                         "frac_at_survey_562B_per_read": 562.0 * reads_local / (kernel_ms * 1e-3) / 1e9 / peak},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(S, alns, args.cpu_seconds, hoisted=True)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    hc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
