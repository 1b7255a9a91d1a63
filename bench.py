#!/usr/bin/env python
"""bench.py -- read x path likelihood evaluations per second of the HaploCart hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl reference]

A "step" is one whole sample through the hot path: reset the accumulators, run the read kernel over the batch, collapse
into the per-path vector (bit-GEMV), (N > 1: exchange the per-rank vectors over NVLink peer memory and sum them in rank
order), argmax + log-softmax posterior + clade support, and read the result back.

  value   the batch already resident in HBM when the timed region starts (device time, CUDA events); the loop keeps two
          samples in flight (vgl_hc_finalize_begin / _end: the host reads step s's result while step s + 1 runs), every
          result read on the host inside the timed region; `ms_per_step_one_result_at_a_time` is the same loop with a
          synchronous get_posterior per step;
  e2e     the same sample from what the reference's call site holds -- the FLAT ALIGNMENT FIELDS of vector<AlignmentInfo*>
          (vgh_alignments, host memory) -- through the reference-facing call vgh_hc_feed (host marshalling into page-locked
          chunks, H2D, device-side reconstruct_graph_sequence + packing, read kernel) to the posterior on the host, every
          step.  `e2e_prepacked` keeps the round-1 figure (a batch packed ahead of time in its wire form).

Workload: config 2 at every N (10^6 x 100 bp reads per GPU, 11825 x 5179 tables): one weak-scaling series.
The same JSON line carries, as extra blocks, the other BASELINE.json configs where the driver runs them:
  N > 1: `config3` -- 50 M ancient-damaged reads in total (strong scaling: 50 M / N per GPU), value + e2e;
  N = 1: `config4` (euka, the reference's own euka_db.clade / .bins tables) and `config5` (soibean, P = 399), each with its
         own roofline.
The batch (~0.4 GB) is larger than the 126 MB L2, so every step streams it from HBM.

`--impl reference` times the REFERENCE'S OWN CODE (oracle/_ref/libvgan_ref.so: src/update_likelihood.cpp,
process_mapping.cpp, get_p_obs_base.cpp compiled unchanged over a header shim) on all host cores, on a bounded sample
per step; the oracle port stands in only where that library is absent.
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
CONFIG3_TOTAL_READS = 50_000_000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def workload(reads, config=2):
    if config == 2:
        return dict(name="HaploCart config 2: synthetic %d x 100bp reads per GPU x %d paths" % (reads, N_PATHS), seed=20261,
                    kw=dict(fixed_len=100, sub_rate=0.005, indel_rate=0.0005, damage=False))
    return dict(name="HaploCart config 3: synthetic ancient-damaged reads (dhigh5p/dhigh3p profiles, log-normal 30-150bp) x %d "
                     "paths" % N_PATHS, seed=20263,
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) where it was built, the oracle port otherwise
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_inputs(S, alns):
    from oracle import oracle as orc
    node, val = S.pmap()
    T = orc.HcTables(S.supports(), node, val, S.mappab(), min_id=1)
    G = orc.Graph(S.node_sequences(), min_id=1)
    R = orc.FlatReads(alns.arrays())
    return T, G, R


def _cpu_run(kind, T, G, R, r0, r1, cores):
    """One pass of the reference's hot loop over reads [r0, r1) on `cores` threads -> (reads used, seconds of the loop)."""
    from oracle import oracle as orc
    if kind == "reference":
        from oracle import ref
        tm = {}
        _, _, used = ref.hc_run(T, G, R, r0, r1, n_threads=cores, timing=tm)
        return used, tm["loop_seconds"]
    t = time.perf_counter()
    _, _, used = orc.hc_run(T, G, R, r0=r0, r1=r1, mode=orc.MODE_LITERAL, n_threads=cores)
    return used, time.perf_counter() - t


def cpu_kind():
    try:
        from oracle import ref
        if ref.available():
            ref.lib()
            return "reference"
    except Exception:
        pass
    return "port"


CPU_WHAT = {"reference": "the reference's own Haplocart::update / process_mapping / get_p_obs_base (oracle/_ref/libvgan_ref.so: "
                         "/root/reference/src compiled unchanged over a header shim) in the OpenMP read loop of src/HaploCart.cpp:408-421",
            "port": "literal per-path loops of the oracle port (oracle/vgan_oracle.cpp MODE_LITERAL; oracle/_ref not present)"}


def cpu_reference_rate(S, alns, seconds_target, hoisted=False):
    """Time the reference's CPU path on a bounded sample of the same workload; returns the cpu_baseline dict."""
    from oracle import oracle as orc
    kind = cpu_kind()
    T, G, R = _cpu_inputs(S, alns)
    cores = os.cpu_count() or 1
    n0 = min(len(alns), 2 * cores)
    used0, dt0 = _cpu_run(kind, T, G, R, 0, n0, cores)
    n = int(min(len(alns), max(n0, n0 * seconds_target / max(dt0, 1e-6))))
    used, dt = _cpu_run(kind, T, G, R, 0, n, cores)
    out = {"value": used * N_PATHS / dt, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": "first %d reads of the same synthetic set, %.1f s: %s (long double, -O3 -fopenmp)" % (n, dt, CPU_WHAT[kind]),
           "reads": used, "seconds": dt}
    if hoisted:
        # the fair optimised-CPU comparator: A_i / B_i once per mapping, then the per-path select-add
        nh = int(min(len(alns), 20 * n))
        t = time.perf_counter()
        _, _, uh = orc.hc_run(T, G, R, r0=0, r1=nh, mode=orc.MODE_HOISTED, n_threads=cores)
        dth = time.perf_counter() - t
        out["hoisted"] = {"value": uh * N_PATHS / dth, "reads": uh, "seconds": dth, "kind": "port",
                          "what": "oracle port with A_i / B_i computed once per mapping (MODE_HOISTED), same threads"}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import vgan_b200 as vb
    kind = cpu_kind()
    wl = workload(args.reads, 2)
    S = vb.SynthHaplocart(seed=TABLE_SEED, n_paths=N_PATHS, n_rows=N_ROWS, n_nodes=N_NODES)
    cores = os.cpu_count() or 1
    probe_reads = S.reads(4 * cores, seed=wl["seed"], **wl["kw"])
    T, G, R = _cpu_inputs(S, probe_reads)
    used, dt = _cpu_run(kind, T, G, R, 0, len(probe_reads), cores)
    per_read = dt / max(used, 1)
    budget = 150.0 / max(args.steps + args.warmup, 1)       # the whole run ends within a few minutes
    n = int(max(cores, min(200000, budget / per_read)))
    alns = S.reads(n, seed=wl["seed"], **wl["kw"])
    T, G, R = _cpu_inputs(S, alns)
    for _ in range(args.warmup):
        _cpu_run(kind, T, G, R, 0, n, cores)
    used, secs = 0, 0.0
    for _ in range(args.steps):
        u, d = _cpu_run(kind, T, G, R, 0, n, cores)
        used += u
        secs += d
    value = used * N_PATHS / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f80 (long double)", "data": "synthetic",
            "config": {"workload": wl["name"], "sample_reads_per_step": n, "paths": N_PATHS},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d reads per step of the same synthetic workload: %s" % (n, CPU_WHAT[kind])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# sibling configs (N = 1 only): euka on the reference's own clade / bin tables, soibean at P = 399
# ---------------------------------------------------------------------------------------------------------------------
def _event_ms(torch, stream, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def bench_config4(torch, stream, n_reads, peak):
    """euka (BASELINE.json configs[3]): clades and bins verbatim from the reference's euka_db.clade / euka_db.bins."""
    import numpy as np
    import vgan_b200 as vb
    gold = os.path.join(ROOT, "tests", "golden")
    S = vb.SynthEuka.from_tables(os.path.join(gold, "euka_db.clade"), os.path.join(gold, "euka_db.bins"), seed=20264)
    alns = S.reads(n_reads, seed=20265)
    ek = vb.Euka()
    ek.set_stream(stream.cuda_stream)
    ek.load(S.tables(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF))
    ek.load_graph(S.graph)
    t0 = time.perf_counter()
    pb = vb.EukaPackedBatch(S.graph, alns, pinned=True)
    pack_s = time.perf_counter() - t0
    cols = int(np.diff(alns.arrays()["seq_off"]).sum())
    ek.reset()
    ek.run(pb)               # leaves the batch in HBM

    def resident():          # the kernel again on the resident batch, nothing crosses PCIe
        ek.reset()
        return ek.rerun(pb.n_span)

    from vgan_b200.host import pinned_empty
    n_all = len(alns)
    feeder = vb.Feeder(n_threads=0)      # auto: three quarters of OMP_NUM_THREADS
    outs = {"clade": pinned_empty((n_all,), np.int32), "in_lik": pinned_empty((n_all,), np.float64),
            "not_lik": pinned_empty((n_all,), np.float64), "like": pinned_empty((n_all,), np.float64),
            "pass": pinned_empty((n_all,), np.uint8)}

    def from_alignments():   # flat alignment fields -> page-locked chunks -> device reconstruction + kernel -> per-read results
        ek.reset()
        return feeder.feed_euka(ek, alns, out=outs)

    for _ in range(3):
        resident()
    ms_dev, _ = _event_ms(torch, stream, resident, 10)
    k_ms = ek.kernel_ms()
    for _ in range(2):
        from_alignments()
    st0 = feeder.stats()
    ms_e2e, res = _event_ms(torch, stream, from_alignments, 3)
    st1 = feeder.stats()
    raw_bytes = (st1["h2d_bytes"] - st0["h2d_bytes"]) // 3
    # algorithmic bytes: 2 B per aligned column as stored (symbol pair + quality) + 10 B header in, 20 B out per read
    alg = 2 * cols + 30 * pb.n_reads
    out = {"workload": "euka config 4: %d synthetic aeDNA reads x 335 clades / 3929 bins taken verbatim from the reference's euka_db.clade "
                       "/ euka_db.bins (node sequences synthetic, damage dhigh5p / dhigh3p)" % pb.n_reads,
           "metric": "reads per second (one clade per read: evaluations = reads)", "reads": pb.n_reads, "aligned_columns": cols,
           "value": pb.n_reads / (ms_dev * 1e-3), "unit": "reads/s", "ms_per_step": ms_dev,
           "e2e": {"value": pb.n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": raw_bytes,
                   "d2h_bytes_per_step": 33 * pb.n_reads,
                   "host_marshal_seconds_per_step": (st1["build_seconds"] - st0["build_seconds"]) / 3,
                   "host_wait_seconds_per_step": (st1["wait_seconds"] - st0["wait_seconds"]) / 3,
                   "general_builder_chunks": st1["slow_chunks"] - st0["slow_chunks"],
                   "route": "flat alignment fields -> vgh_euka_feed: host marshalling into page-locked chunks overlapped with "
                            "vgl_euka_run_raw (H2D, device reconstruction + packing, euka kernel, per-read results into page-locked "
                            "host memory)"},
           "host_pack_seconds": pack_s, "detected_reads": int(res["pass"].sum()),
           "roofline": {"bound": "hbm", "kernel": "euka_read_kernel", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                        "traffic": None}}
    ek.close()
    return out


def bench_config5(torch, stream, n_reads, peak, P=399):
    """soibean (BASELINE.json configs[4]): the R x P likelihood matrix that feeds the multi-source MCMC."""
    import numpy as np
    import vgan_b200 as vb
    from vgan_b200.host import pinned_empty
    Sh = vb.SynthHaplocart(seed=20266, n_paths=P, n_rows=N_ROWS, n_nodes=N_NODES)
    al = Sh.reads(n_reads, seed=20267, fixed_len=0, sub_rate=0.01, indel_rate=0.001, damage=True)
    sb = vb.Soibean()
    sb.set_stream(stream.cuda_stream)
    sb.load(vb.SoibeanTables(Sh.supports(), 0, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7))
    sb.load_graph(Sh.graph)
    pk = vb.SoibeanPackedBatch(Sh.graph, al, pinned=True)
    po, pu = pinned_empty((pk.n_span, P), np.float64), pinned_empty((pk.n_span,), np.uint8)

    sb.analyse_into(pk, po, pu)   # leaves the batch in HBM

    def resident():      # the kernels again on the resident batch, the matrix left in HBM for the MCMC kernel: no R x P copy
        return sb.rerun(pk.n_span)

    def to_host():       # matrix returned into page-locked memory (what convertMapsToVector hands the host MCMC)
        return sb.analyse_into(pk, po, pu)

    for _ in range(2):
        resident()
    ms_dev, _ = _event_ms(torch, stream, resident, 5)
    k_ms = sb.kernel_ms()
    to_host()
    ms_e2e, _ = _event_ms(torch, stream, to_host, 3)
    # from the flat alignment fields (vgh_soibean_feed): marshalling, device reconstruction + packing, matrix back, chunk by chunk
    feeder = vb.Feeder(n_threads=0)
    assert pk.n_span == len(al)
    feeder.feed_soibean(sb, al, out=po, used=pu)
    ms_feed, _ = _event_ms(torch, stream, lambda: feeder.feed_soibean(sb, al, out=po, used=pu), 3)
    sb.analyse_into(pk, po, pu)   # the whole batch resident again for the MCMC kernel below
    # one MCMC proposal scored on the resident batch
    f7 = vb.soibean_freqs7(0.31, 0.19, 0.2, 0.3)
    rng = np.random.default_rng(5)
    props = [[(int(rng.integers(P)), int(rng.integers(P)), float(rng.uniform(0.05, 0.95)), float(rng.uniform(0.001, 0.05)), w)
              for w in (0.6, 0.4)] for _ in range(64)]
    sb.mcmc_loglike(props, f7, 0.01)
    t0 = time.perf_counter()
    sb.mcmc_loglike(props, f7, 0.01)
    mcmc_ms = (time.perf_counter() - t0) * 1e3 / len(props)
    alg_out = 8 * P * pk.n_reads
    out = {"workload": "soibean config 5: %d synthetic damaged reads x %d species-tree paths (PENALTY 7, dhigh profiles)" % (pk.n_reads, P),
           "reads": pk.n_reads, "paths": P, "value": pk.n_reads * P / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev,
           "e2e": {"value": pk.n_reads * P / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": pk.nbytes,
                   "d2h_bytes_per_step": alg_out + pk.n_reads,
                   "route": "packed host batch -> H2D -> soibean kernels -> R x P matrix into page-locked host memory"},
           "e2e_from_alignments": {"value": pk.n_reads * P / (ms_feed * 1e-3), "unit": UNIT, "ms_per_step": ms_feed,
                                   "route": "flat alignment fields -> vgh_soibean_feed (marshalling overlapped with H2D, device "
                                            "reconstruction + packing, kernels, R x P matrix into page-locked host memory)"},
           "mcmc_ms_per_proposal_batched_64": mcmc_ms,
           "roofline": {"bound": "hbm", "kernel": "soibean_path_kernel", "achieved": alg_out / (k_ms * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg_out / (k_ms * 1e-3) / 1e9 / peak, "kernel_ms": k_ms,
                        "algorithmic_bytes_per_launch": alg_out, "traffic": None,
                        "what": "P * 8 B written per read (SURVEY 8d: the R x P matrix is the dominant stream)"}}
    sb.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU (config 2)")
    ap.add_argument("--impl", default="vgan_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / 5 blocks")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU exchange of the P-vector")
    ap.add_argument("--workload", type=int, default=2, choices=[2, 3], help="SURVEY 8d config whose read shape the main line uses")
    ap.add_argument("--config3-tile", type=int, default=1000000, help="config 3: distinct reads generated per GPU (re-scored to reach 50 M / N)")
    ap.add_argument("--sibling-reads", type=int, default=1000000)
    ap.add_argument("--feeder-threads", type=int, default=0, help="host threads of the e2e feeder (0: the rank's share of the cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    # the host side is sharded too: every rank marshals with its share of the cores
    host_threads = max(1, (os.cpu_count() or 1) // max(local_world, 1))
    os.environ["OMP_NUM_THREADS"] = str(host_threads)

    import numpy as np
    import torch
    import vgan_b200 as vb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    wl = workload(args.reads, args.workload)
    S = vb.SynthHaplocart(seed=TABLE_SEED, n_paths=N_PATHS, n_rows=N_ROWS, n_nodes=N_NODES)
    alns = S.reads(args.reads, seed=wl["seed"] + 1000 * rank, **wl["kw"])
    hc = vb.Haplocart(device=local)
    hc.load(S.tables(), S.tree())
    hc.load_graph(S.graph)
    # every kernel, copy and collective of a step runs on this stream, and so do the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hc.set_stream(stream.cuda_stream)
    t_pack = time.perf_counter()
    packed = vb.PackedBatch(S.graph, alns, pinned=True, n_threads=host_threads)
    t_pack = time.perf_counter() - t_pack
    dev = hc.upload(packed)
    try:
        wire = vb.WireBatch(packed, pinned=True)
    except RuntimeError:
        wire = None
    reads_local = packed.n_reads
    batch_bytes = packed.nbytes
    # algorithmic bytes of the read kernel: 4 per column + 4 per empty mapping + 7 per read + 8 per group
    pa = packed.arrays()
    alg_bytes = int(4 * pa["n_cols"].astype(np.int64).sum() + 4 * pa["n_extra"].astype(np.int64).sum() + 7 * reads_local
                    + 8 * (int(packed.batch.n_groups) + 1))
    del pa
    feeder_threads = args.feeder_threads or max(1, host_threads * 3 // 4)   # the submitting thread and the driver need cores too
    feeder = vb.Feeder(n_threads=feeder_threads)

    sharded = None
    if world > 1:
        from vgan_b200.dist import ShardedSample
        sharded = ShardedSample(hc, exchange=args.exchange)

    def finish():
        if sharded is not None:
            return sharded.finish(want_ll=False)
        return hc.get_posterior(want_ll=False)

    def step_resident():
        hc.reset()
        hc.update_device(dev.batch)
        return finish()

    def step_e2e():            # from the flat alignment fields, host marshalling inside the step
        hc.reset()
        feeder.feed(hc, alns)
        return finish()

    def step_prepacked():      # round-1 e2e: a batch packed (and turned into its wire form) ahead of time
        hc.reset()
        if wire is not None:
            hc.update_wire(wire)
        else:
            hc.update(packed)
        return finish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_ms, k_n = 0.0, 0
        barrier()
        l0 = hc.launch_count()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
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

    # The device-resident loop runs two samples deep (vgl_hc_finalize_begin / _end): a step's posterior kernel and result copy
    # are enqueued, the next step's reset + read kernel follow on the same stream, and the host reads each result one step
    # later -- every step's result still reaches the host inside the timed region; the GPU does not idle while it does.
    pipelined = sharded is None or sharded.exchange == "peer"

    def timed_pipelined(steps, accumulate=None):
        accumulate = accumulate or (lambda: hc.update_device(dev.batch))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = hc.launch_count()
        e0.record(stream)
        pending, out = None, None
        for _ in range(steps):
            hc.reset()
            accumulate()
            ticket = sharded.finish_begin() if sharded is not None else hc.posterior_begin()
            if pending is not None:
                out = hc.posterior_end(pending)
            pending = ticket
        out = hc.posterior_end(pending)
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total, hc.launch_count() - l0, out

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    # one result at a time (reset ... get_posterior per step): the read kernel's own time comes from this loop's events
    ms_sync, k_ms, k_n, launches, res = timed(step_resident, args.steps)
    if pipelined:
        ms_dev, launches, res_p = timed_pipelined(args.steps)
        pipelined_same_call = bool(res_p["maxh_index"] == res["maxh_index"])
    else:
        ms_dev, pipelined_same_call = ms_sync, None
    for _ in range(3):
        step_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    st0 = feeder.stats()
    if pipelined:   # (as the device-resident loop: the host marshals step s + 1 while the GPU finishes step s and its result comes back)
        ms_e2e, _, res_e2e = timed_pipelined(e2e_steps, lambda: feeder.feed(hc, alns))
    else:
        ms_e2e, _, _, _, res_e2e = timed(step_e2e, e2e_steps)
    st1 = feeder.stats()
    for _ in range(3):
        step_prepacked()
    ms_pre, _, _, _, res_pre = timed(step_prepacked, e2e_steps)
    clocks = sampler.finish()
    same_call = res_e2e.get("maxh_index") == res.get("maxh_index") == res_pre.get("maxh_index")

    # sharded result == rank-order sum of the ranks' own vectors, identical on every rank (checked on the device-resident step)
    parity_ok = None
    if world > 1:
        def gathered(vec):
            g = torch.from_numpy(np.ascontiguousarray(vec)).cuda()
            allg = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(allg, g)
            return [x.cpu().numpy() for x in allg]
        # (a) one run: this rank's own vector, then the exchange of exactly those vectors (vgl_hc_paths first, so the library takes
        #     its two-kernel exchange): the result must be the rank-order sum of the parts, bit for bit, on every rank
        hc.reset()
        hc.update_device(dev.batch)
        mine = torch.zeros(N_PATHS, dtype=torch.float64, device="cuda")
        hc.paths(mine.data_ptr())
        got_a = sharded.finish(want_ll=True)["final_vec"]
        torch.cuda.synchronize()
        want = np.zeros(N_PATHS)
        for part in gathered(mine.cpu().numpy()):
            want = want + part
        all_a = gathered(got_a)
        ok_a = all(np.array_equal(all_a[0], x) for x in all_a) and np.array_equal(got_a, want)
        # (b) the fused exchange (the timed path): identical on every rank bit for bit; equal to (a) up to the run-to-run rounding of
        #     the node vectors (fp64 reductions arrive in a different order in every run)
        hc.reset()
        hc.update_device(dev.batch)
        got_b = sharded.finish(want_ll=True)["final_vec"]
        all_b = gathered(got_b)
        ok_b = all(np.array_equal(all_b[0], x) for x in all_b) and bool(np.max(np.abs(got_b - want) / np.abs(want)) < 1e-12)
        parity_ok = bool(ok_a and ok_b)

    # config 3 (N > 1): 50 M damaged reads in total, 50 M / N per GPU (strong scaling); every GPU holds `tile` distinct reads and
    # scores them ceil(share / tile) times per step
    config3 = None
    if world > 1 and not args.no_extras:
        wl3 = workload(0, 3)
        share = CONFIG3_TOTAL_READS // world
        tile = min(args.config3_tile, share)
        al3 = S.reads(tile, seed=wl3["seed"] + 1000 * rank, **wl3["kw"])
        pk3 = vb.PackedBatch(S.graph, al3, pinned=True, n_threads=host_threads)
        dev3 = hc.upload(pk3)
        reps = max(1, int(round(share / max(pk3.n_reads, 1))))

        def step3_resident():
            hc.reset()
            for _ in range(reps):
                hc.update_device(dev3.batch)
            return finish()

        def step3_e2e():
            hc.reset()
            for _ in range(reps):
                feeder.feed(hc, al3)
            return finish()

        step3_resident()
        ms3, k3_ms, k3_n, _, res3 = timed(step3_resident, 3)
        step3_e2e()
        ms3e, _, _, _, _ = timed(step3_e2e, 1)
        t = torch.tensor([pk3.n_reads * reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        tot3 = int(t.item())
        config3 = {"workload": wl3["name"], "total_reads": tot3, "reads_per_gpu": pk3.n_reads * reps, "scaling": "strong",
                   "distinct_reads_per_gpu": pk3.n_reads, "passes_per_step": reps,
                   "value": tot3 * N_PATHS * 3 / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / 3,
                   "read_kernel_ms_per_launch": k3_ms / max(k3_n, 1),
                   "e2e": {"value": tot3 * N_PATHS / (ms3e * 1e-3), "unit": UNIT, "ms_per_step": ms3e,
                           "route": "vgh_hc_feed per pass (flat alignment fields, host marshalling inside)"},
                   "mean_columns_per_read": float(pk3.arrays()["n_cols"].astype(np.int64).sum()) / max(pk3.n_reads, 1),
                   "called_haplogroup": res3.get("predicted_haplotype")}
        dev3.free()
        del pk3, al3

    # per-sample latency at the size one real sample can reach after duplicate removal (SURVEY 8d scale caveat)
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
                   "what": "reset + accumulate (packed host batch) + finalize, host wall clock, mean of 20"}

        def from_alignments(route):
            hc.reset()
            if route == "feed":
                feeder.feed(hc, alns, 0, n_lat)
            elif route == "raw":
                hc.update_raw(vb.RawBatch(alns, r0=0, r1=n_lat))
            else:
                hc.update(vb.PackedBatch(S.graph, alns, r0=0, r1=n_lat))
            return hc.get_posterior(want_ll=False)
        for key, route in (("ms_per_sample_from_alignments_host_packer", "pack"), ("ms_per_sample_from_alignments_raw_route", "raw"),
                           ("ms_per_sample_from_alignments_feeder", "feed")):
            for _ in range(3):
                from_alignments(route)
            t0 = time.perf_counter()
            for _ in range(10):
                from_alignments(route)
            latency[key] = (time.perf_counter() - t0) / 10 * 1e3

        # several samples per pass (vgl_hc_samples_begin / _finalize_samples): 64 samples of <= 2 x 10^4 reads each
        n_s, per = 64, min(20000, len(alns) // 64)
        if per >= 1000:
            small_packed = [vb.PackedBatch(S.graph, alns, r0=k * per, r1=(k + 1) * per, pinned=True) for k in range(n_s)]
            small_dev = [hc.upload(b) for b in small_packed]
            small_wire = [vb.WireBatch(b, pinned=True) for b in small_packed] if wire is not None else None

            def one_by_one():
                out = []
                for d in small_dev:
                    hc.reset()
                    hc.update_device(d.batch)
                    out.append(hc.get_posterior(want_ll=False)["maxh_index"])
                return out

            def two_in_flight():
                out, pending = [], None
                for d in small_dev:
                    hc.reset()
                    hc.update_device(d.batch)
                    t = hc.posterior_begin()
                    if pending is not None:
                        out.append(hc.posterior_end(pending)["maxh_index"])
                    pending = t
                out.append(hc.posterior_end(pending)["maxh_index"])
                return out

            def in_one_pass(route):
                hc.samples_begin(n_s)
                for k in range(n_s):
                    hc.sample_select(k)
                    if route == "resident":
                        hc.update_device(small_dev[k].batch)
                    elif route == "wire":
                        hc.update_wire(small_wire[k])
                    else:
                        feeder.feed(hc, alns, k * per, (k + 1) * per)
                return [r["maxh_index"] for r in hc.finalize_samples()]

            def clock(fn, reps):
                fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    res = fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / reps, res

            t_one, calls_one = clock(one_by_one, 5)
            t_two, calls_two = clock(two_in_flight, 5)
            t_res, calls_res = clock(lambda: in_one_pass("resident"), 5)
            latency["samples"] = {"samples": n_s, "reads_per_sample": per,
                                  "samples_per_s_one_call_each_resident": n_s / t_one,
                                  "samples_per_s_two_in_flight_resident": n_s / t_two,
                                  "samples_per_s_one_pass_resident": n_s / t_res,
                                  "same_calls": calls_one == calls_res == calls_two,
                                  "what": "64 samples: reset + accumulate + finalize per sample vs one vgl_hc_samples_begin / "
                                          "_finalize_samples pass (device-resident batches; host wall clock)"}
            if small_wire is not None:
                t_wire, calls_wire = clock(lambda: in_one_pass("wire"), 3)
                latency["samples"]["samples_per_s_one_pass_wire_from_host"] = n_s / t_wire
                latency["samples"]["wire_bytes_per_sample"] = small_wire[0].nbytes
            t_feed, calls_feed = clock(lambda: in_one_pass("feed"), 3)
            latency["samples"]["samples_per_s_one_pass_from_alignments"] = n_s / t_feed
            latency["samples"]["same_calls"] = bool(calls_one == calls_res == calls_feed == calls_two)
            for d in small_dev:
                d.free()
            del small_dev, small_packed, small_wire

        # file -> result: a GAM stream of one real-sample size, host reader + feeder against the device decoder; and the
        # duplicate removal that runs in front of the hot loop (src/HaploCart.cpp:390)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers as TH
            from test_gpu_recon import _alns_to_dicts
            n_io = min(20000, len(alns))
            sub = S.reads(n_io, seed=wl["seed"], **wl["kw"])
            gam_bytes = TH.encode_gam([dict(a, read_paired=False) for a in _alns_to_dicts(sub)], group=1000)

            def via_host():
                a = vb.Alignments.from_gam_bytes(gam_bytes)
                hc.reset()
                feeder.feed(hc, a)
                return hc.get_posterior(want_ll=False)["maxh_index"]

            def via_device():
                hc.reset()
                hc.update_gam(vb.GamFrames(data=gam_bytes))
                return hc.get_posterior(want_ll=False)["maxh_index"]

            io = {"reads": len(sub), "gam_bytes": len(gam_bytes)}
            for key, fn in (("ms_gam_bytes_to_call_host_parser", via_host), ("ms_gam_bytes_to_call_device_decoder", via_device)):
                fn()
                t0 = time.perf_counter()
                for _ in range(5):
                    call = fn()
                io[key] = (time.perf_counter() - t0) / 5 * 1e3
                io[key + "_call"] = int(call)
            hc.dedup(alns)
            t0 = time.perf_counter()
            mask, kept = hc.dedup(alns)
            io["ms_device_dedup_all_reads"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            hmask, hkept = vb.host.dedup(alns)
            io["ms_host_dedup_all_reads"] = (time.perf_counter() - t0) * 1e3
            io["dedup_reads"], io["dedup_kept"], io["dedup_same_mask"] = len(alns), int(kept), bool(np.array_equal(mask, hmask))
            latency["io"] = io
        except Exception as e:  # the io block must not take the headline down
            latency["io"] = {"error": repr(e)}

    total_reads = reads_local
    if world > 1:
        t = torch.tensor([reads_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        total_reads = int(t.item())

    peak, peak_kind = peaks()
    extras = {}
    if world == 1 and not args.no_extras:
        hc_free = [dev]
        for name, fn in (("config4", bench_config4), ("config5", bench_config5)):
            try:
                extras[name] = fn(torch, stream, args.sibling_reads, peak)
            except Exception as e:  # a sibling block must not take the headline down
                extras[name] = {"error": repr(e)}

    if rank == 0:
        value = total_reads * N_PATHS * args.steps / (ms_dev * 1e-3)
        e2e = total_reads * N_PATHS * e2e_steps / (ms_e2e * 1e-3)
        e2e_pre = total_reads * N_PATHS * e2e_steps / (ms_pre * 1e-3)
        kernel_ms = k_ms / max(k_n, 1)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        # which read kernel the timed steps launched: the log-term kernel when the batch's meta allows it (qualities <= 63,
        # mapping-quality segments), else the general one
        read_kernel_name = "hc_read_logterm_kernel" if hc.logterm_launches() > 0 else "hc_read_kernel"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                ent = tj.get("kernels", {}).get(read_kernel_name)
                if tj.get("reads") == reads_local and ent:
                    traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
            except Exception:
                pass
        nsteps = max(e2e_steps, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "ms_per_step_one_result_at_a_time": ms_sync / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "reads_per_gpu": reads_local, "paths": N_PATHS, "rows": N_ROWS,
                       "l2": "batch of %.0f MB per step exceeds the 126 MB L2 (no flush needed)" % (batch_bytes / 1e6),
                       "steps_in_flight": 2 if pipelined else 1, "same_call_as_one_result_at_a_time": pipelined_same_call,
                       "parallelism": "reads sharded over %d GPU(s); one exchange of P doubles per sample (%s)" % (
                           n_gpus, "none" if sharded is None else sharded.exchange),
                       "host_threads_per_rank": host_threads,
                       "called_haplogroup": res.get("predicted_haplotype")},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) // nsteps,
                    "d2h_bytes_per_step": 28 + 8 * hc._max_post, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "route": "flat alignment fields (vgh_alignments, host) -> vgh_hc_feed: host marshalling into page-locked chunks "
                             "overlapped with H2D + device reconstruct_graph_sequence + packing + read kernel -> posterior on the host"
                             + (" (two samples in flight: step s + 1 is marshalled while step s's last chunk and posterior finish)" if pipelined else ""),
                    "host_marshal_threads": feeder_threads,
                    "host_marshal_seconds_per_step": (st1["build_seconds"] - st0["build_seconds"]) / nsteps,
                    "host_wait_seconds_per_step": (st1["wait_seconds"] - st0["wait_seconds"]) / nsteps,
                    "general_builder_chunks": st1["slow_chunks"] - st0["slow_chunks"],
                    "same_call_as_resident_and_prepacked": bool(same_call)},
            "e2e_prepacked": {"value": e2e_pre, "unit": UNIT, "ms_per_step": ms_pre / e2e_steps,
                              "h2d_bytes_per_step": (wire.nbytes if wire is not None else packed.nbytes),
                              "route": "batch packed and turned into its wire form AHEAD of the timed region (round-1 e2e)"},
            "latency": latency,
            "host_pack": {"seconds": t_pack, "reads_per_s": reads_local / t_pack, "threads": host_threads,
                          "what": "vgh_hc_pack (reconstruction + packing, OpenMP) of the step's batch; outside the timed regions"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": read_kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes, "batch_bytes_with_padding": batch_bytes,
                         "kernel_ms": kernel_ms, "bytes_per_read": alg_bytes / max(reads_local, 1),
                         "frac_at_survey_562B_per_read": 562.0 * reads_local / (kernel_ms * 1e-3) / 1e9 / peak},
        }
        if parity_ok is not None:
            line["parity_ok"] = parity_ok
        if config3 is not None:
            line["config3"] = config3
        line.update(extras)
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
