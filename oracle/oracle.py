"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/vgan_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under vgan_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvgan_oracle.so")

MODE_LITERAL = 0
MODE_HOISTED = 1


def build(force=False):
    src = os.path.join(_HERE, "vgan_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"] + (["-s"] if force else []))
    return _LIB_PATH


class _Graph(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("min_id", C.c_int64), ("seq_off", C.c_void_p), ("seq", C.c_void_p)]


class _Reads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("qual_off", C.c_void_p),
                ("qual", C.c_void_p), ("mapq", C.c_void_p), ("identity", C.c_void_p), ("map_off", C.c_void_p),
                ("map_node", C.c_void_p), ("map_offset", C.c_void_p), ("map_rev", C.c_void_p),
                ("edit_off", C.c_void_p), ("edit_from", C.c_void_p), ("edit_to", C.c_void_p),
                ("edit_seq_off", C.c_void_p), ("edit_seq", C.c_void_p)]


class _HcTables(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_paths", C.c_int32), ("supports", C.c_void_p), ("n_pmap", C.c_int64),
                ("pmap_node", C.c_void_p), ("pmap_val", C.c_void_p), ("n_mappab", C.c_int64), ("mappab", C.c_void_p),
                ("background_error_prob", C.c_double), ("is_consensus_fasta", C.c_int32), ("min_id", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_hc_region_match_prob.restype = C.c_double
        _lib.orc_hc_background_freq.restype = C.c_double
        _lib.orc_hc_p_seq_error.restype = C.c_double
        assert _lib.orc_sizeof_long_double() == np.dtype(np.longdouble).itemsize
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Graph:
    """Node sequences of a graph with contiguous ids (stands in for bdsg::ODGI in the oracle)."""

    def __init__(self, node_seqs, min_id=None):
        # node_seqs: dict id -> str, or list of str for ids min_id..
        if isinstance(node_seqs, dict):
            ids = sorted(node_seqs)
            min_id = ids[0] if min_id is None else min_id
            max_id = ids[-1]
            seqs = [node_seqs.get(i, "") for i in range(min_id, max_id + 1)]
        else:
            seqs = list(node_seqs)
            min_id = 1 if min_id is None else min_id
        self.min_id = int(min_id)
        self.n_nodes = len(seqs)
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        self.seq_off = np.zeros(self.n_nodes + 1, dtype=np.int64)
        np.cumsum(lens, out=self.seq_off[1:])
        self.seq = np.frombuffer(("".join(seqs)).encode() + b"\0", dtype=np.uint8).copy()
        self.c = _Graph(self.n_nodes, self.min_id, _ptr(self.seq_off), _ptr(self.seq))


class Reads:
    """Flattened alignments.  `alns` is a list of dicts as produced by oracle.gam_py.read_gam."""

    def __init__(self, alns):
        n = len(alns)
        self.n = n
        seq = b"".join(a["sequence"] for a in alns)
        qual = b"".join(a["quality"] for a in alns)
        self.seq_off = np.zeros(n + 1, dtype=np.int64)
        self.qual_off = np.zeros(n + 1, dtype=np.int64)
        self.map_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(a["sequence"]) for a in alns], out=self.seq_off[1:])
        np.cumsum([len(a["quality"]) for a in alns], out=self.qual_off[1:])
        np.cumsum([len(a["mappings"]) for a in alns], out=self.map_off[1:])
        self.seq = np.frombuffer(seq + b"\0", dtype=np.uint8).copy()
        self.qual = np.frombuffer(qual + b"\0", dtype=np.uint8).copy()
        self.mapq = np.array([a["mapping_quality"] for a in alns], dtype=np.int32)
        self.identity = np.array([a["identity"] for a in alns], dtype=np.float64)
        maps = [m for a in alns for m in a["mappings"]]
        self.map_node = np.array([m["node_id"] for m in maps], dtype=np.int64)
        self.map_offset = np.array([m["offset"] for m in maps], dtype=np.int64)
        self.map_rev = np.array([1 if m["is_reverse"] else 0 for m in maps], dtype=np.uint8)
        self.edit_off = np.zeros(len(maps) + 1, dtype=np.int64)
        np.cumsum([len(m["edits"]) for m in maps], out=self.edit_off[1:])
        edits = [e for m in maps for e in m["edits"]]
        self.edit_from = np.array([e["from_length"] for e in edits], dtype=np.int32)
        self.edit_to = np.array([e["to_length"] for e in edits], dtype=np.int32)
        self.edit_seq_off = np.zeros(len(edits) + 1, dtype=np.int64)
        np.cumsum([len(e["sequence"]) for e in edits], out=self.edit_seq_off[1:])
        self.edit_seq = np.frombuffer(b"".join(e["sequence"] for e in edits) + b"\0", dtype=np.uint8).copy()
        # keep zero-length arrays addressable
        for name in ("map_node", "map_offset", "map_rev", "edit_from", "edit_to"):
            if getattr(self, name).size == 0:
                setattr(self, name, np.zeros(1, dtype=getattr(self, name).dtype))
        self.c = _Reads(n, _ptr(self.seq_off), _ptr(self.seq), _ptr(self.qual_off), _ptr(self.qual), _ptr(self.mapq),
                        _ptr(self.identity), _ptr(self.map_off), _ptr(self.map_node), _ptr(self.map_offset),
                        _ptr(self.map_rev), _ptr(self.edit_off), _ptr(self.edit_from), _ptr(self.edit_to),
                        _ptr(self.edit_seq_off), _ptr(self.edit_seq))


class FlatReads(Reads):
    """The same flattened layout, taken from ready-made numpy arrays (keys as in Reads' attributes)."""

    def __init__(self, a):
        self.n = int(a["n_reads"])

        def pad(x, dt):
            x = np.ascontiguousarray(x, dtype=dt)
            return x if x.size else np.zeros(1, dtype=dt)

        self.seq_off = pad(a["seq_off"], np.int64)
        self.qual_off = pad(a["qual_off"], np.int64)
        self.map_off = pad(a["map_off"], np.int64)
        self.edit_off = pad(a["edit_off"], np.int64)
        self.edit_seq_off = pad(a["edit_seq_off"], np.int64)
        self.seq = pad(a["seq"], np.uint8)
        self.qual = pad(a["qual"], np.uint8)
        self.edit_seq = pad(a["edit_seq"], np.uint8)
        self.mapq = pad(a["mapq"], np.int32)
        self.identity = pad(a["identity"], np.float64)
        self.map_node = pad(a["map_node"], np.int64)
        self.map_offset = pad(a["map_offset"], np.int64)
        self.map_rev = pad(a["map_rev"], np.uint8)
        self.edit_from = pad(a["edit_from"], np.int32)
        self.edit_to = pad(a["edit_to"], np.int32)
        self.c = _Reads(self.n, _ptr(self.seq_off), _ptr(self.seq), _ptr(self.qual_off), _ptr(self.qual), _ptr(self.mapq),
                        _ptr(self.identity), _ptr(self.map_off), _ptr(self.map_node), _ptr(self.map_offset),
                        _ptr(self.map_rev), _ptr(self.edit_off), _ptr(self.edit_from), _ptr(self.edit_to),
                        _ptr(self.edit_seq_off), _ptr(self.edit_seq))


class HcTables:
    """HaploCart tables as the reference holds them in memory."""

    def __init__(self, supports, pmap_node, pmap_val, mappab, bep=1e-4, is_fasta=False, min_id=1):
        self.supports = np.ascontiguousarray(supports, dtype=np.uint8)
        assert self.supports.ndim == 2
        self.pmap_node = np.ascontiguousarray(pmap_node, dtype=np.int64)
        self.pmap_val = np.ascontiguousarray(pmap_val, dtype=np.int32)
        self.mappab = np.ascontiguousarray(mappab, dtype=np.float64)
        self.n_rows, self.n_paths = self.supports.shape
        self.c = _HcTables(self.n_rows, self.n_paths, _ptr(self.supports), self.pmap_node.size, _ptr(self.pmap_node),
                           _ptr(self.pmap_val), self.mappab.size, _ptr(self.mappab), float(bep), int(bool(is_fasta)),
                           int(min_id))


def reconstruct(graph, reads, r, cap=8192):
    g = np.zeros(cap, dtype=np.uint8)
    p = np.zeros(cap, dtype=np.uint8)
    s = np.zeros(cap, dtype=np.int32)
    ng, npp, ns = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib().orc_reconstruct(C.byref(graph.c), C.byref(reads.c), C.c_int64(r), _ptr(g), _ptr(p), _ptr(s), cap,
                               C.byref(ng), C.byref(npp), C.byref(ns))
    if rc:
        raise RuntimeError("orc_reconstruct failed: %d" % rc)
    return g[:ng.value].tobytes().decode(), p[:npp.value].tobytes().decode(), s[:ns.value].tolist()


def hc_read(tables, graph, reads, r, mode=MODE_HOISTED):
    ll = np.zeros(tables.n_paths, dtype=np.longdouble)
    nm = int(reads.map_off[r + 1] - reads.map_off[r])
    A = np.zeros(max(nm, 1), dtype=np.longdouble)
    B = np.zeros(max(nm, 1), dtype=np.float64)
    rc = lib().orc_hc_read(C.byref(tables.c), C.byref(graph.c), C.byref(reads.c), C.c_int64(r), mode, _ptr(ll),
                           _ptr(A) if mode == MODE_HOISTED else None, _ptr(B) if mode == MODE_HOISTED else None)
    if rc:
        raise RuntimeError("orc_hc_read failed: %d" % rc)
    return ll, A[:nm], B[:nm]


def hc_run(tables, graph, reads, r0=0, r1=None, mode=MODE_HOISTED, n_threads=1):
    r1 = reads.n if r1 is None else r1
    final = np.zeros(tables.n_paths, dtype=np.longdouble)
    am = C.c_int32()
    used = C.c_int64()
    rc = lib().orc_hc_run(C.byref(tables.c), C.byref(graph.c), C.byref(reads.c), C.c_int64(r0), C.c_int64(r1), mode,
                          n_threads, _ptr(final), C.byref(am), C.byref(used))
    if rc:
        raise RuntimeError("orc_hc_run failed: %d" % rc)
    return final, am.value, used.value


def hc_posterior(final_vec, path_names, parents_txt, children_txt, predicted_idx, cap=4096):
    fv = np.ascontiguousarray(final_vec, dtype=np.longdouble)
    idx = np.zeros(cap, dtype=np.int32)
    conf = np.zeros(cap, dtype=np.float64)
    total = np.zeros(1, dtype=np.longdouble)
    n = lib().orc_hc_posterior(_ptr(fv), fv.size, ("\n".join(path_names)).encode(), parents_txt.encode(),
                               children_txt.encode(), int(predicted_idx), cap, _ptr(idx), _ptr(conf), _ptr(total))
    if n < 0:
        raise RuntimeError("orc_hc_posterior failed: %d" % n)
    return idx[:n].tolist(), conf[:n].copy(), total[0]


def qscore_vec():
    v = np.zeros(100)
    lib().orc_hc_qscore_vec(_ptr(v))
    return v


def incorrect_mapping_vec():
    v = np.zeros(100)
    lib().orc_hc_incorrect_mapping_vec(_ptr(v))
    return v


def region_match_prob(pb):
    return lib().orc_hc_region_match_prob(int(pb))


# ---- euka ------------------------------------------------------------------------------------------------------
class _EukaTables(C.Structure):
    _fields_ = [("n_clades", C.c_int32), ("clade_dist", C.c_void_p), ("bin_off", C.c_void_p), ("bin_lo", C.c_void_p),
                ("bin_hi", C.c_void_p), ("deam5p_prof", C.c_char_p), ("deam3p_prof", C.c_char_p), ("minimum_mq", C.c_uint32)]


def euka_run(clade_dist, bin_off, bin_lo, bin_hi, graph, reads, deam5p_prof="", deam3p_prof="", minimum_mq=29):
    """The readGAM3 lambda over every read.  Returns a dict of per-read arrays plus clade_count / bin_cov."""
    dist = np.ascontiguousarray(clade_dist, dtype=np.float64)
    boff = np.ascontiguousarray(bin_off, dtype=np.int64)
    blo = np.ascontiguousarray(bin_lo, dtype=np.int32)
    bhi = np.ascontiguousarray(bin_hi, dtype=np.int32)
    T = _EukaTables(dist.size, _ptr(dist), _ptr(boff), _ptr(blo), _ptr(bhi), (deam5p_prof or "").encode(),
                    (deam3p_prof or "").encode(), int(minimum_mq))
    n = reads.n
    out = {"clade": np.zeros(n, dtype=np.int32), "in_lik": np.zeros(n), "not_lik": np.zeros(n), "like": np.zeros(n),
           "not_like": np.zeros(n), "pass": np.zeros(n, dtype=np.uint8)}
    count = np.zeros(dist.size, dtype=np.int64)
    cov = np.zeros(max(int(boff[-1]), 1), dtype=np.float64)
    rc = lib().orc_euka_run(C.byref(T), C.byref(graph.c), C.byref(reads.c), _ptr(out["clade"]), _ptr(out["in_lik"]),
                            _ptr(out["not_lik"]), _ptr(out["like"]), _ptr(out["not_like"]), _ptr(out["pass"]), _ptr(count),
                            _ptr(cov))
    if rc:
        raise RuntimeError("orc_euka_run failed: %d" % rc)
    out["clade_count"] = count
    out["bin_cov"] = cov[:int(boff[-1])]
    return out


def euka_proposal_loglike(clade, like, not_like, n_clades, clade_ids, frac):
    """MCMC::get_proposal_likelihood (src/MCMC.cpp:1175-1213) over the per-read outputs of euka_run."""
    clade = np.ascontiguousarray(clade, dtype=np.int32)
    like = np.ascontiguousarray(like, dtype=np.float64)
    nl = np.ascontiguousarray(not_like, dtype=np.float64)
    ids = np.ascontiguousarray(clade_ids, dtype=np.int32)
    fr = np.ascontiguousarray(frac, dtype=np.float64)
    f = lib().orc_euka_proposal_loglike
    f.restype = C.c_longdouble
    return float(f(C.c_int64(clade.size), C.c_int32(n_clades), _ptr(clade), _ptr(like), _ptr(nl), C.c_int32(ids.size), _ptr(ids), _ptr(fr)))


def euka_deam_matrix(deam5p_prof, deam3p_prof, L, l):
    m = np.zeros(16)
    rc = lib().orc_euka_deam_matrix((deam5p_prof or "").encode(), (deam3p_prof or "").encode(), int(L), int(l), _ptr(m))
    if rc:
        raise RuntimeError("orc_euka_deam_matrix failed: %d" % rc)
    return m.reshape(4, 4)


# ---- soibean ---------------------------------------------------------------------------------------------------
def soibean_run(graph, reads, member, min_id=0, penalty=7, deam5p_prof="", deam3p_prof=""):
    """analyse_GAM's pathMap for every read: (matrix [R, P] in pathNames order, used [R])."""
    member = np.ascontiguousarray(member, dtype=np.uint8)
    n_rows, P = member.shape
    out = np.zeros((reads.n, P), dtype=np.float64)
    used = np.zeros(reads.n, dtype=np.uint8)
    rc = lib().orc_soibean_run(C.byref(graph.c), C.byref(reads.c), _ptr(member), C.c_int64(n_rows), int(min_id), int(P), int(penalty),
                               (deam5p_prof or "").encode(), (deam3p_prof or "").encode(), _ptr(out), _ptr(used))
    if rc:
        raise RuntimeError("orc_soibean_run failed: %d" % rc)
    return out, used


def soibean_mcmc_loglike(graph, reads, member, sources, freqs7, con, min_id=0, penalty=7, deam5p_prof="", deam3p_prof=""):
    """Log-likelihood of one MCMC proposal (read loop of MCMC::run_tree_proportion, src/MCMC.cpp:738-993).
    sources: list of (child path, parent path, pos_branch, dist, proportion).  Returns (logLike, per-read contributions)."""
    member = np.ascontiguousarray(member, dtype=np.uint8)
    n_rows, P = member.shape
    K = len(sources)
    child = np.array([s[0] for s in sources], dtype=np.int32)
    parent = np.array([s[1] for s in sources], dtype=np.int32)
    posb = np.array([s[2] for s in sources], dtype=np.float64)
    dist = np.array([s[3] for s in sources], dtype=np.float64)
    prop = np.array([s[4] for s in sources], dtype=np.float64)
    f7 = np.ascontiguousarray(freqs7, dtype=np.float64)
    assert f7.shape == (7,)
    out = C.c_double(0.0)
    per_read = np.zeros(reads.n, dtype=np.float64)
    rc = lib().orc_soibean_mcmc_loglike(C.byref(graph.c), C.byref(reads.c), _ptr(member), C.c_int64(n_rows), int(min_id), int(P), int(penalty),
                                        (deam5p_prof or "").encode(), (deam3p_prof or "").encode(), int(K), _ptr(child), _ptr(parent),
                                        _ptr(posb), _ptr(dist), _ptr(prop), _ptr(f7), C.c_double(con), C.byref(out), _ptr(per_read))
    if rc:
        raise RuntimeError("orc_soibean_mcmc_loglike failed: %d" % rc)
    return out.value, per_read


def soibean_base_loglike(read_base, graph_base, ll, freqs7, t, branch_len):
    """computeBaseLogLike (src/MCMC.h:111-296) for one supported column."""
    f7 = np.ascontiguousarray(freqs7, dtype=np.float64)
    out = C.c_double(0.0)
    rc = lib().orc_soibean_base_loglike(C.c_char(read_base.encode()), C.c_char(graph_base.encode()), C.c_double(ll), _ptr(f7), C.c_double(t),
                                        C.c_double(branch_len), C.byref(out))
    if rc:
        raise RuntimeError("orc_soibean_base_loglike failed: %d" % rc)
    return out.value


def soibean_freqs7(a, c, g, t):
    """params.freqs['A','C','G','T','R','Y','M'] as src/soibean.cpp:629-640 derives them from the base frequencies."""
    R, Y = a + g, c + t
    M = 1 / (2 * (((22) * (a * g)) + ((22) * (c * t)) + (a * c + (a * t) + (g * c + (g * t)))))
    return np.array([a, c, g, t, R, Y, M], dtype=np.float64)


def dedup(reads, paired=None):
    """The reference's O(n^2) duplicate marking (src/rmdup.cpp:20-110)."""
    mask = np.zeros(reads.n, dtype=np.uint8)
    pz = None if paired is None else np.ascontiguousarray(paired, dtype=np.uint8)
    rc = lib().orc_dedup(C.byref(reads.c), None if pz is None else _ptr(pz), _ptr(mask))
    if rc:
        raise RuntimeError("orc_dedup failed: %d" % rc)
    return mask.astype(bool)
