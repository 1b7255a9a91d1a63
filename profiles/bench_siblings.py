#!/usr/bin/env python
"""Throughput of the euka and soibean sibling kernels (reduced read counts), end to end
through the C ABI with host batches (H2D and the per-read results' D2H inside the timed region).
Usage: python profiles/bench_siblings.py [euka_reads] [soibean_reads]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vgan_b200 as vb

n_euka = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_soi = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
out = {}

S = vb.SynthEuka(seed=20264, n_clades=335)
alns = S.reads(n_euka, seed=20265)
t = time.perf_counter()
pb = vb.EukaPackedBatch(S.graph, alns, pinned=True)
pack_s = time.perf_counter() - t
ek = vb.Euka()
ek.load(S.tables(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF))
cols = int(np.diff(alns.arrays()["seq_off"]).sum())
for _ in range(2):
    ek.reset(); ek.run(pb)
t = time.perf_counter()
for _ in range(5):
    ek.reset(); res = ek.run(pb)
dt = (time.perf_counter() - t) / 5
# the same from raw alignments: reconstruction + packing on the device
t = time.perf_counter()
raw = vb.RawBatch(alns, pinned=True, euka_rule=True)
raw_build_s = time.perf_counter() - t
ek.load_graph(S.graph)
for _ in range(2):
    ek.reset(); ek.run_raw(raw)
t = time.perf_counter()
for _ in range(5):
    ek.reset(); res_raw = ek.run_raw(raw)
dt_raw = (time.perf_counter() - t) / 5
assert np.array_equal(res_raw["in_lik"], res["in_lik"])
out["euka_raw"] = {"reads": raw.n_reads, "reads_per_s": raw.n_reads / dt_raw, "ms": dt_raw * 1e3, "raw_bytes": raw.nbytes,
                   "host_build_s": raw_build_s,
                   "what": "vgl_euka_run_raw: raw alignments -> H2D -> device reconstruction + packing -> euka kernel -> per-read D2H"}
out["euka"] = {"reads": pb.n_reads, "clades": 335, "reads_per_s": pb.n_reads / dt, "read_bases_per_s": cols / dt, "ms": dt * 1e3,
               "batch_bytes": pb.nbytes, "host_pack_s": pack_s,
               "what": "config 4 shape: vgl_euka_run on a host batch incl. H2D and per-read D2H; evaluations = reads"}
ek.close()

for P in (31, 127, 399):
    Sh = vb.SynthHaplocart(seed=20266, n_paths=P, n_rows=11825, n_nodes=11821)
    al = Sh.reads(n_soi, seed=20267, fixed_len=0, sub_rate=0.01, indel_rate=0.001, damage=True)
    sb = vb.Soibean()
    sb.load(vb.SoibeanTables(Sh.supports(), 0, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7))
    pk = vb.SoibeanPackedBatch(Sh.graph, al, pinned=True)
    sb.analyse(pk)
    t = time.perf_counter()
    for _ in range(3):
        got, used = sb.analyse(pk)
    dt = (time.perf_counter() - t) / 3
    # the same into page-locked output arrays (allocated once, outside the timed region)
    from vgan_b200.host import pinned_empty
    import ctypes as C
    from vgan_b200._lib import gpu, check
    po, pu = pinned_empty((pk.n_span, P), np.float64), pinned_empty((pk.n_span,), np.uint8)
    run = lambda: check(gpu.vgl_soibean_run(sb.ctx, C.byref(pk.batch), pk.n_span, po.ctypes.data_as(C.c_void_p), pu.ctypes.data_as(C.c_void_p)))
    run()
    t = time.perf_counter()
    for _ in range(3):
        run()
    dtp = (time.perf_counter() - t) / 3
    assert np.array_equal(po, got)
    out["soibean_P%d" % P] = {"reads": pk.n_reads, "paths": P, "evals_per_s": pk.n_reads * P / dt, "reads_per_s": pk.n_reads / dt,
                              "ms": dt * 1e3, "out_bytes": int(got.nbytes), "ms_pinned_output": dtp * 1e3,
                              "evals_per_s_pinned_output": pk.n_reads * P / dtp,
                              "what": "config 5 shape: vgl_soibean_run incl. H2D and the R x P matrix D2H into pageable numpy "
                                      "(ms) or into page-locked memory (ms_pinned_output)"}
    if P == 127:
        # MCMC proposals scored on the resident batch: one call per proposal (what a single chain
        # does) and 256 proposals per call (chains / proposals batched); 2 sources per proposal
        f7 = vb.soibean_freqs7(0.31, 0.19, 0.2, 0.3)
        rng = np.random.default_rng(5)
        def proposal():
            return [(int(rng.integers(P)), int(rng.integers(P)), float(rng.uniform(0.05, 0.95)), float(rng.uniform(0.001, 0.05)), w) for w in (0.6, 0.4)]
        one = [proposal()]
        sb.mcmc_loglike(one, f7, 0.01)
        t = time.perf_counter()
        for _ in range(20):
            sb.mcmc_loglike(one, f7, 0.01)
        dt1 = (time.perf_counter() - t) / 20
        many = [proposal() for _ in range(256)]
        sb.mcmc_loglike(many, f7, 0.01)
        t = time.perf_counter()
        for _ in range(3):
            sb.mcmc_loglike(many, f7, 0.01)
        dtm = (time.perf_counter() - t) / 3
        out["soibean_mcmc_P127"] = {"reads": pk.n_reads, "sources": 2, "ms_per_proposal_single": dt1 * 1e3,
                                    "ms_per_proposal_batched_256": dtm * 1e3 / 256, "read_proposals_per_s_batched": pk.n_reads * 256 / dtm,
                                    "what": "vgl_soibean_mcmc_loglike on the batch left resident by vgl_soibean_run, host wall clock incl. "
                                            "the proposal H2D and the log-likelihood D2H"}
    sb.close()
print("SIBLINGS", json.dumps(out))
