"""euka per-read likelihood path over the C ABI (reference: the readGAM3 lambda, src/readGAM_Euka.h:103-549).

`Euka` mirrors the call site in Euka::run (src/Euka.cpp:534-537): it takes the alignments of a sample and returns, per
read, the clade, the two log-likelihoods, clade_like and the pass flag, plus the per-clade counts and bin coverage that
the detection filters and the abundance MCMC consume.  No CPU path: construction fails without a CUDA device."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, gpu, host
from .host import Alignments, Graph, _as_array


class EukaTables:
    """Clade distances, bins and damage profiles in the form vgl_euka_set_tables wants."""

    def __init__(self, clade_dist, bin_off, bin_lo, bin_hi, max_node, deam5p_prof="", deam3p_prof="", minimum_mq=29):
        self.clade_dist = np.ascontiguousarray(clade_dist, dtype=np.float64)
        self.bin_off = np.ascontiguousarray(bin_off, dtype=np.int32)
        self.bin_lo = np.ascontiguousarray(bin_lo, dtype=np.int32)
        self.bin_hi = np.ascontiguousarray(bin_hi, dtype=np.int32)
        self.n_clades = self.clade_dist.size
        self.n_bins = int(self.bin_off[-1])
        self.handle = C.c_void_p()
        check(host.vgh_euka_tables_build(self.n_clades, self.clade_dist.ctypes.data_as(C.c_void_p),
                                         self.bin_off.ctypes.data_as(C.c_void_p), self.bin_lo.ctypes.data_as(C.c_void_p),
                                         self.bin_hi.ctypes.data_as(C.c_void_p), int(max_node), (deam5p_prof or "").encode(),
                                         (deam3p_prof or "").encode(), int(minimum_mq), C.byref(self.handle)))
        self.view = host.vgh_euka_tables_view(self.handle).contents

    def deam(self):
        v = self.view
        return _as_array(v.deam, v.k5 * v.k3 * 16, np.float64).reshape(v.k5, v.k3, 4, 4)

    def __del__(self):
        if self.handle:
            host.vgh_euka_tables_destroy(self.handle)
            self.handle = None


class EukaPackedBatch:
    def __init__(self, graph, alns, r0=0, r1=None, n_threads=0, pinned=False):
        r1 = len(alns) if r1 is None else r1
        self.n_span = int(r1 - r0)
        self.handle = C.c_void_p()
        check(host.vgh_euka_pack(graph.handle, C.byref(alns.view), int(r0), int(r1), int(n_threads), int(pinned), C.byref(self.handle)))
        self.batch = host.vgh_euka_packed_batch(self.handle).contents

    @property
    def n_reads(self):
        return int(self.batch.n_reads)

    @property
    def nbytes(self):
        return int(gpu.vgl_euka_batch_bytes(C.byref(self.batch)))

    def __del__(self):
        if self.handle:
            host.vgh_euka_packed_destroy(self.handle)
            self.handle = None


class Euka:
    def __init__(self, device=0):
        self.ctx = C.c_void_p()
        check(gpu.vgl_create(int(device), C.byref(self.ctx)))
        self.tables = None

    def close(self):
        if self.ctx:
            gpu.vgl_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, tables):
        check(gpu.vgl_euka_set_tables(self.ctx, C.byref(tables.view)))
        self.tables = tables

    def reset(self):
        check(gpu.vgl_euka_reset(self.ctx))

    def run(self, packed):
        """Score one packed batch; per-read arrays are indexed by the read's position in the packed range."""
        n = packed.n_span
        out = {"clade": np.full(n, -1, dtype=np.int32), "in_lik": np.zeros(n), "not_lik": np.zeros(n), "like": np.zeros(n),
               "pass": np.zeros(n, dtype=np.uint8)}
        check(gpu.vgl_euka_run(self.ctx, C.byref(packed.batch), n, out["clade"].ctypes.data_as(C.c_void_p),
                               out["in_lik"].ctypes.data_as(C.c_void_p), out["not_lik"].ctypes.data_as(C.c_void_p),
                               out["like"].ctypes.data_as(C.c_void_p), out["pass"].ctypes.data_as(C.c_void_p)))
        out["not_like"] = 1 - out["like"]  # clade_not_like (src/readGAM_Euka.h:492)
        return out

    def load_graph(self, graph):
        """Node sequences for run_raw (reconstruction on the device)."""
        mid, n = C.c_int64(), C.c_int64()
        off, seq = C.c_void_p(), C.c_void_p()
        check(_lib.host.vgh_graph_view(graph.handle, C.byref(mid), C.byref(n), C.byref(off), C.byref(seq)))
        check(gpu.vgl_hc_set_graph(self.ctx, mid.value, n.value, off, seq))

    def run_raw(self, raw):
        """As run(), from a RawBatch(euka_rule=True): reconstruction and packing happen on the device."""
        n = raw.n_span
        out = {"clade": np.full(n, -1, dtype=np.int32), "in_lik": np.zeros(n), "not_lik": np.zeros(n), "like": np.zeros(n),
               "pass": np.zeros(n, dtype=np.uint8)}
        check(gpu.vgl_euka_run_raw(self.ctx, C.byref(raw.raw), n, out["clade"].ctypes.data_as(C.c_void_p),
                                   out["in_lik"].ctypes.data_as(C.c_void_p), out["not_lik"].ctypes.data_as(C.c_void_p),
                                   out["like"].ctypes.data_as(C.c_void_p), out["pass"].ctypes.data_as(C.c_void_p)))
        out["not_like"] = 1 - out["like"]
        return out

    def set_stream(self, cuda_stream_ptr):
        check(gpu.vgl_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def rerun(self, n_out=0, want=False):
        """Score again the batch the last run() / run_raw() left in HBM (no host -> device traffic).  want=True returns the
        per-read arrays as run() does; otherwise nothing leaves the device (the per-clade sums accumulate as usual)."""
        if not want:
            check(gpu.vgl_euka_rerun(self.ctx, int(n_out), None, None, None, None, None))
            return None
        out = {"clade": np.full(n_out, -1, dtype=np.int32), "in_lik": np.zeros(n_out), "not_lik": np.zeros(n_out), "like": np.zeros(n_out),
               "pass": np.zeros(n_out, dtype=np.uint8)}
        check(gpu.vgl_euka_rerun(self.ctx, int(n_out), out["clade"].ctypes.data_as(C.c_void_p), out["in_lik"].ctypes.data_as(C.c_void_p),
                                 out["not_lik"].ctypes.data_as(C.c_void_p), out["like"].ctypes.data_as(C.c_void_p),
                                 out["pass"].ctypes.data_as(C.c_void_p)))
        out["not_like"] = 1 - out["like"]
        return out

    def kernel_ms(self):
        """Device time of the last euka read kernel (CUDA events on its stream)."""
        ms = C.c_float()
        check(gpu.vgl_sibling_kernel_ms(self.ctx, C.byref(ms)))
        return float(ms.value)

    def results(self):
        count = np.zeros(self.tables.n_clades, dtype=np.int64)
        cov = np.zeros(max(self.tables.n_bins, 1), dtype=np.float64)
        check(gpu.vgl_euka_results(self.ctx, count.ctypes.data_as(C.c_void_p), cov.ctypes.data_as(C.c_void_p)))
        return count, cov[:self.tables.n_bins]

    def like_sums(self):
        """Per clade, since reset(): the number of reads scored for it and the sum of log(clade_like) over them -- all that
        euka's abundance MCMC reads (MCMC::get_proposal_likelihood, src/MCMC.cpp:1175-1213)."""
        n = np.zeros(self.tables.n_clades, dtype=np.int64)
        s = np.zeros(self.tables.n_clades, dtype=np.float64)
        check(gpu.vgl_euka_like_sums(self.ctx, n.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)))
        return n, s

    @staticmethod
    def proposal_loglike(clade_ids, frac, like_n, like_logsum):
        """get_proposal_likelihood of one proposal: clade_ids = clade_list_id, frac = proposal_vec."""
        ids = np.ascontiguousarray(clade_ids, dtype=np.int32)
        fr = np.ascontiguousarray(frac, dtype=np.float64)
        n = np.ascontiguousarray(like_n, dtype=np.int64)
        s = np.ascontiguousarray(like_logsum, dtype=np.float64)
        return float(gpu.vgl_euka_proposal_loglike(int(ids.size), ids.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p),
                                                   n.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)))

    def launch_count(self):
        return int(gpu.vgl_launch_count(self.ctx))


class SynthEuka:
    """Synthetic euka-shaped clades / bins / graph."""

    def __init__(self, seed=20264, n_clades=335, genome_len=16000, _handle=None):
        self.handle = C.c_void_p() if _handle is None else _handle
        if _handle is None:
            check(host.vgh_synth_euka_create(int(seed), int(n_clades), int(genome_len), C.byref(self.handle)))
        self._init_views()

    @classmethod
    def from_tables(cls, clade_file, bins_file, seed=20264):
        """clades (dist, snode..enode) and bins verbatim from the reference's `<db>.clade` (6 tokens
        per line: id name dist noPaths snode enode, src/load.cpp:108-157) and `<db>.bins` (name, then start end entropy
        triples, src/load.cpp:71-95); only the node sequences are synthetic."""
        dist, snode, enode, names = [], [], [], []
        for line in open(clade_file):
            t = line.split()
            if len(t) == 6:
                names.append(t[1]); dist.append(float(t[2])); snode.append(int(t[4])); enode.append(int(t[5]))
        bin_off, lo, hi = [0], [], []
        for line in open(bins_file):
            t = line.split()
            if not t:
                continue
            for j in range(1, len(t) - 2, 3):
                lo.append(int(float(t[j]))); hi.append(int(float(t[j + 1])))   # stoi("1836.0") == 1836 in the reference
            bin_off.append(len(lo))
        if len(bin_off) != len(dist) + 1:
            raise ValueError("clade and bins files list different numbers of clades")
        a = lambda x, dt: np.ascontiguousarray(x, dtype=dt)
        d, s, e, bo, bl, bh = a(dist, np.float64), a(snode, np.int32), a(enode, np.int32), a(bin_off, np.int32), a(lo, np.int32), a(hi, np.int32)
        h = C.c_void_p()
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        check(host.vgh_synth_euka_create_from_tables(int(seed), len(dist), p(d), p(s), p(e), p(bo), p(bl), p(bh), C.byref(h)))
        obj = cls(_handle=h)
        obj.clade_names = names
        return obj

    def _init_views(self):
        self.n_clades = int(host.vgh_synth_euka_n_clades(self.handle))
        self.graph = Graph(C.c_void_p(host.vgh_synth_euka_graph(self.handle)), owner=self)
        self.dist = _as_array(host.vgh_synth_euka_dist(self.handle), self.n_clades, np.float64)
        self.bin_off = _as_array(host.vgh_synth_euka_bin_off(self.handle), self.n_clades + 1, np.int32)
        nb = int(self.bin_off[-1])
        self.bin_lo = _as_array(host.vgh_synth_euka_bin_lo(self.handle), nb, np.int32)
        self.bin_hi = _as_array(host.vgh_synth_euka_bin_hi(self.handle), nb, np.int32)
        off, seq = C.c_void_p(), C.c_void_p()
        self.n_nodes = int(host.vgh_synth_euka_graph_seq(self.handle, C.byref(off), C.byref(seq)))
        self._off = _as_array(off, self.n_nodes + 1, np.int64)
        self._seq = _as_array(seq, int(self._off[-1]), np.uint8)

    def node_sequences(self):
        blob = self._seq.tobytes().decode()
        return [blob[self._off[i]:self._off[i + 1]] for i in range(self.n_nodes)]

    def tables(self, deam5p_prof="", deam3p_prof="", minimum_mq=29):
        return EukaTables(self.dist, self.bin_off, self.bin_lo, self.bin_hi, self.n_nodes, deam5p_prof, deam3p_prof, minimum_mq)

    def reads(self, n_reads, seed=20265, fixed_len=0, indel_rate=0.002, damage=True, n_rate=0.002, softclip_rate=0.05):
        h = C.c_void_p()
        check(host.vgh_synth_euka_reads(self.handle, int(seed), int(n_reads), int(fixed_len), float(indel_rate), int(bool(damage)),
                                        float(n_rate), float(softclip_rate), C.byref(h)))
        return Alignments(h)

    def __del__(self):
        if self.handle:
            host.vgh_synth_euka_destroy(self.handle)
            self.handle = None


# The reference's own damage profiles share/vgan/damageProfiles/dhigh5p.prof / dhigh3p.prof, reproduced as text so that
# tests and the benchmark do not need the reference tree (12 tab-separated rates per position, header first).
_HEADER = "A>C\tA>G\tA>T\tC>A\tC>G\tC>T\tG>A\tG>C\tG>T\tT>A\tT>C\tT>G\n"


def _prof(col, vals):
    return _HEADER + "".join("\t".join(("%g" % v) if k == col else "0" for k in range(12)) + "\n" for v in vals)


DHIGH5P_PROF = _prof(5, [0.329405, 0.221745, 0.187678, 0.161196, 0.144011])
DHIGH3P_PROF = _prof(6, [0.32891, 0.223405, 0.188599, 0.164419, 0.146352])
