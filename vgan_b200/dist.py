"""Multi-GPU plumbing for the HaploCart path: one process per GPU, `torch.distributed` for the single
exchange step.

The path shards by reads: every rank runs the read kernel and the node x path pass on its own share of the
sample with replicated tables and ends up with its own final_vec[P].  The ranks all-gather those vectors
(41 KB each for the HaploCart graph -- pure latency on NVLink) and every rank adds them IN RANK ORDER before
the posterior kernel.  A ring all-reduce would add different chunks of the vector in different rank orders,
which could break exact ties between identically supported paths; the fixed order keeps the call
bit-identical on every rank and independent of the collective algorithm.

`torch` is used only here and in bench.py: for the process group and for device memory that NCCL can see.
"""
import numpy as np


def shard_range(n_reads, rank, world):
    """Contiguous block of reads for `rank` (block partition)."""
    base, extra = divmod(int(n_reads), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sum_in_rank_order(parts):
    """Fixed-order sum of per-rank vectors (host mirror of vgl_hc_set_paths_from_parts)."""
    parts = np.asarray(parts, dtype=np.float64)
    total = np.zeros(parts.shape[1], dtype=np.float64)
    for k in range(parts.shape[0]):
        total += parts[k]
    return total


def gather_paths(local_vec, group=None):
    """All-gather this rank's final_vec (a 1-D float64 torch tensor on the backend's device) into [world, P]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = torch.empty((world, local_vec.numel()), dtype=local_vec.dtype, device=local_vec.device)
    dist.all_gather_into_tensor(parts.view(-1), local_vec.contiguous(), group=group)
    return parts


def sum_proposal_loglikes(local_loglike, group=None, device=None):
    """soibean MCMC on sharded reads: every rank scores the proposals on its own reads
    (`Soibean.mcmc_loglike`), the [n_proposals] vectors are all-gathered -- one double per proposal and rank -- and added
    in rank order, so every rank accepts / rejects on the bit-identical number.  Returns a numpy vector."""
    import torch
    local = torch.as_tensor(np.ascontiguousarray(local_loglike, dtype=np.float64))
    if device is not None:
        local = local.to(device)
    parts = gather_paths(local, group)
    return sum_in_rank_order(parts.cpu().numpy())


def sum_euka_results(clade_count, bin_cov, like_n=None, like_logsum=None, group=None, device=None):
    """euka on sharded reads: the per-read outputs stay with the rank that scored the reads; what the detection
    filter and the abundance MCMC need are per-clade sums -- read counts (`Clade::count`), bin coverage, and the two sums behind
    `get_proposal_likelihood` (number of scored reads and sum of log(clade_like) per clade: `Euka.like_sums`).  Integer counts
    are all-reduced; the floating-point vectors are all-gathered and added in rank order, so every rank holds bit-identical
    totals.  Returns numpy arrays (clade_count, bin_cov[, like_n, like_logsum])."""
    import torch
    import torch.distributed as dist

    def dev(a, dtype):
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=dtype))
        return t.to(device) if device is not None else t

    ints = [dev(clade_count, np.int64)] + ([dev(like_n, np.int64)] if like_n is not None else [])
    for t in ints:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    floats = [dev(np.ravel(bin_cov), np.float64)] + ([dev(like_logsum, np.float64)] if like_logsum is not None else [])
    sums = [sum_in_rank_order(gather_paths(t, group).cpu().numpy()) if t.numel() else np.zeros(0) for t in floats]
    out = [ints[0].cpu().numpy(), sums[0].reshape(np.shape(bin_cov))]
    if like_n is not None:
        out.append(ints[1].cpu().numpy())
    if like_logsum is not None:
        out.append(sums[1])
    return tuple(out)


class ShardedSample:
    """One sample spread over the ranks of a process group.  Each rank owns a `Haplocart` on its own GPU."""

    def __init__(self, hc, group=None, exchange="peer"):
        import torch
        import torch.distributed as dist
        self.hc, self.group = hc, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local = torch.zeros(hc.n_paths, dtype=torch.float64, device="cuda")
        self.parts = torch.zeros(self.world * hc.n_paths, dtype=torch.float64, device="cuda")
        # One stream orders the kernels, the all-gather and the final sum.  The legacy default stream has handle 0, which
        # vgl_set_stream reads as "use the context's own stream": take a real stream unless the caller already runs on one.
        self.stream = torch.cuda.current_stream()
        if self.stream.cuda_stream == 0:
            self.stream = torch.cuda.Stream()
        hc.set_stream(self.stream.cuda_stream)
        # The exchange step: "peer" = the library's own kernels over NVLink peer memory (CUDA IPC between the ranks of one
        # node); "nccl" = all-gather + vgl_hc_set_paths_from_parts.  Both add the vectors in rank order.
        self.exchange = exchange
        if exchange == "peer":
            # every rank takes part in every collective whatever happened locally: a rank whose export failed sends None
            try:
                mine = hc.peer_export(self.world)
            except Exception as e:  # no exchange buffer on this rank (allocation / IPC failure)
                self.exchange_error = repr(e)
                mine = None
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            ok = 0
            if all(h is not None for h in handles):
                try:
                    hc.peer_attach(self.rank, self.world, handles)
                    ok = 1
                except Exception as e:  # no peer access between these devices: use the collective
                    self.exchange_error = repr(e)
            t = torch.tensor([ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            if int(t.item()) == 0:
                self.exchange = "nccl"

    def run(self, local_batches, device_resident=False, want_ll=False):
        import torch
        import torch.distributed as dist
        hc = self.hc
        with torch.cuda.stream(self.stream):
            hc.reset()
            for b in local_batches:
                if device_resident:
                    hc.update_device(b.batch)
                elif hasattr(b, "wire"):
                    hc.update_wire(b)
                else:
                    hc.update(b)
            return self.finish(want_ll=want_ll)

    def finish_begin(self):
        """The exchange step + posterior, enqueued (Haplocart.posterior_begin): returns a ticket for Haplocart.posterior_end."""
        import torch
        with torch.cuda.stream(self.stream):
            if self.exchange != "peer":
                raise RuntimeError("finish_begin needs the peer exchange (the NCCL fallback synchronises per sample)")
            self.hc.exchange_paths()
            return self.hc.posterior_begin()

    def finish(self, want_ll=False):
        """The exchange step + posterior for the sample accumulated on this rank since hc.reset() (every rank calls it)."""
        import torch
        import torch.distributed as dist
        hc = self.hc
        with torch.cuda.stream(self.stream):
            if self.exchange == "peer":
                hc.exchange_paths()
            else:
                hc.paths(self.local.data_ptr())
                dist.all_gather_into_tensor(self.parts, self.local, group=self.group)
                hc.set_paths_from_parts(self.parts.data_ptr(), self.world)
            return hc.get_posterior(want_ll=want_ll)
