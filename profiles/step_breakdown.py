#!/usr/bin/env python
"""Where one bench step spends its time: each phase of the device-resident step timed on its own with CUDA events
(mean of 50 repetitions after warm-up).  Usage: python profiles/step_breakdown.py [reads]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vgan_b200 as vb
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
wl = bench.workload(1, reads, 2)
S = vb.SynthHaplocart(seed=bench.TABLE_SEED, n_paths=bench.N_PATHS, n_rows=bench.N_ROWS, n_nodes=bench.N_NODES)
alns = S.reads(reads, seed=wl["seed"], **wl["kw"])
hc = vb.Haplocart(0)
hc.load(S.tables(), S.tree())
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
hc.set_stream(stream.cuda_stream)
dev = hc.upload(vb.PackedBatch(S.graph, alns, pinned=True))

def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def full():
    hc.reset(); hc.update_device(dev.batch); hc.get_posterior(want_ll=False)

out = {"reads": reads,
       "reset_us": timeit(hc.reset),
       "read_kernel_us": timeit(lambda: hc.update_device(dev.batch)),
       "paths_us": timeit(lambda: hc.paths()),
       "posterior_us": timeit(lambda: hc.get_posterior(want_ll=False)),
       "full_step_us": timeit(full)}
print("BREAKDOWN", json.dumps(out))
