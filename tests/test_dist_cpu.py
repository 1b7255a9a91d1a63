"""N > 1 host logic on CPU: two `gloo` ranks shard the reads, each computes its share of final_vec (here with the
CPU oracle standing in for the device, since no GPU exists in this container), the vectors are all-gathered and
summed in rank order exactly as vgan_b200.dist does for NCCL, and the call must equal the single-process call."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_reads, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vgan_b200 as vb
    from vgan_b200 import dist as vd
    from oracle import oracle as orc
    import helpers as H
    S = vb.SynthHaplocart(seed=51, n_paths=120, n_rows=610, n_nodes=601)
    alns = S.reads(n_reads, seed=52)
    T, G, R = H.oracle_side(S, alns.arrays())
    lo, hi = vd.shard_range(n_reads, rank, world)
    final, _, used = orc.hc_run(T, G, R, r0=lo, r1=hi)
    local = torch.from_numpy(final.astype(np.float64))
    parts = vd.gather_paths(local)
    total = vd.sum_in_rank_order(parts.numpy())
    used_t = torch.tensor([used], dtype=torch.int64)
    dist.all_reduce(used_t)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.concatenate([total, [float(used_t.item()), lo, hi]]))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vgan_b200 as vb
    from oracle import oracle as orc
    import helpers as H
    n_reads, world = 41, 2
    mp.spawn(_worker, args=(world, _free_port(), n_reads, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0[:-3], r1[:-3])                    # every rank holds the identical, fixed-order sum
    assert (r0[-2], r0[-1], r1[-2], r1[-1]) == (0, 21, 21, 41)  # block partition covers the reads once
    S = vb.SynthHaplocart(seed=51, n_paths=120, n_rows=610, n_nodes=601)
    alns = S.reads(n_reads, seed=52)
    T, G, R = H.oracle_side(S, alns.arrays())
    final, argmax, used = orc.hc_run(T, G, R)
    assert r0[-3] == used
    assert int(np.argmax(r0[:-3])) == argmax
    assert H.rel_err(r0[:-3], final.astype(np.float64)) < 1e-13


def _mcmc_worker(rank, world, port, n_reads, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vgan_b200 import dist as vd
    import test_soibean as ts
    from oracle import oracle as orc
    S, alns, member, og, R = ts._mcmc_inputs(P=31, n=n_reads, seed=95)
    lo, hi = vd.shard_range(n_reads, rank, world)
    local = []
    for prop in ts.PROPOSALS2:  # the oracle stands in for the device: this rank's reads only
        _, per = orc.soibean_mcmc_loglike(og, R, member, prop, ts.F7, 0.01, 0, 7)
        local.append(per[lo:hi].sum())
    total = vd.sum_proposal_loglikes(np.array(local))
    np.save(os.path.join(out_dir, "mcmc%d.npy" % rank), total)
    dist.destroy_process_group()


def test_two_rank_mcmc_proposal_sum(tmp_path):
    """soibean MCMC over sharded reads: one double per proposal and rank, added in rank order on every rank."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_soibean as ts
    from oracle import oracle as orc
    n_reads, world = 90, 2
    mp.spawn(_mcmc_worker, args=(world, _free_port(), n_reads, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "mcmc0.npy"), np.load(tmp_path / "mcmc1.npy")
    assert np.array_equal(r0, r1)
    S, alns, member, og, R = ts._mcmc_inputs(P=31, n=n_reads, seed=95)
    want = [orc.soibean_mcmc_loglike(og, R, member, prop, ts.F7, 0.01, 0, 7)[0] for prop in ts.PROPOSALS2]
    assert np.allclose(r0, want, rtol=1e-13, atol=0)


def _euka_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vgan_b200 import dist as vd
    rng = np.random.default_rng(100 + rank)               # this rank's per-clade sums (the device's, in production)
    count = rng.integers(0, 50, 7)
    cov = rng.random(19)
    like_n = rng.integers(0, 50, 7)
    like_ls = -rng.random(7) * 30
    np.save(os.path.join(out_dir, "euka_in%d.npy" % rank), np.concatenate([count, cov, like_n, like_ls]))
    tot = vd.sum_euka_results(count, cov, like_n, like_ls)
    np.save(os.path.join(out_dir, "euka_out%d.npy" % rank), np.concatenate([np.ravel(t).astype(np.float64) for t in tot]))
    dist.destroy_process_group()


def test_two_rank_euka_sums(tmp_path):
    """euka over sharded reads: per-clade counts all-reduced, coverage and log sums added in rank order."""
    world = 2
    mp.spawn(_euka_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ins = [np.load(tmp_path / ("euka_in%d.npy" % r)) for r in range(world)]
    outs = [np.load(tmp_path / ("euka_out%d.npy" % r)) for r in range(world)]
    assert np.array_equal(outs[0], outs[1])
    want = ins[0].copy()
    want += ins[1]
    assert np.array_equal(outs[0], want)                      # rank-order sum (two ranks: one addition per entry)


def test_shard_range_partitions():
    from vgan_b200 import dist as vd
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            cuts = [vd.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[k][1] == cuts[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
