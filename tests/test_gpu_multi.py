"""Two ranks on two GPUs of one box: the reads are sharded, the per-rank vectors are exchanged through NVLink peer
memory (the library's own kernels) or NCCL, and the call, the log-likelihoods and the posteriors must equal the
single-GPU run bit for bit on the call and to rounding on the sums.  Skipped with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_reads, exchange, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import vgan_b200 as vb
    from vgan_b200 import dist as vd
    S = vb.SynthHaplocart(seed=51, n_paths=700, n_rows=1210, n_nodes=1201)
    alns = S.reads(n_reads, seed=52, sub_rate=0.01, indel_rate=0.002)
    lo, hi = vd.shard_range(n_reads, rank, world)
    hc = vb.Haplocart(device=rank)
    hc.load(S.tables(), S.tree())
    sh = vd.ShardedSample(hc, exchange=exchange)
    packed = vb.PackedBatch(S.graph, alns, r0=lo, r1=hi, pinned=True)
    res = None
    for _ in range(3):  # several samples in a row exercise the two buffer generations
        res = sh.run([packed], want_ll=True)
    np.save(os.path.join(out_dir, "ll_%s_%d.npy" % (exchange, rank)), res["final_vec"])
    with open(os.path.join(out_dir, "call_%s_%d.txt" % (exchange, rank)), "w") as f:
        f.write("%d %s %r\n" % (res["maxh_index"], sh.exchange, res["confidence_vec"].tolist()))
    dist.barrier()
    dist.destroy_process_group()
    hc.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_ranks_equal_one(tmp_path, exchange):
    import torch.multiprocessing as mp
    import vgan_b200 as vb
    n_reads = 3000
    mp.spawn(_worker, args=(2, _free_port(), n_reads, exchange, str(tmp_path)), nprocs=2, join=True)
    S = vb.SynthHaplocart(seed=51, n_paths=700, n_rows=1210, n_nodes=1201)
    alns = S.reads(n_reads, seed=52, sub_rate=0.01, indel_rate=0.002)
    hc = vb.Haplocart(0)
    hc.load(S.tables(), S.tree())
    one = hc.run_sample([vb.PackedBatch(S.graph, alns)])
    hc.close()
    lls = [np.load(os.path.join(str(tmp_path), "ll_%s_%d.npy" % (exchange, r))) for r in range(2)]
    assert np.array_equal(lls[0], lls[1])                                   # every rank holds the same vector
    assert np.max(np.abs(lls[0] - one["final_vec"]) / np.abs(one["final_vec"])) < 1e-12
    for r in range(2):
        idx, used, conf = open(os.path.join(str(tmp_path), "call_%s_%d.txt" % (exchange, r))).read().split(" ", 2)
        assert int(idx) == one["maxh_index"] and used == exchange
        assert np.allclose(eval(conf), one["confidence_vec"], rtol=1e-9)
