"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libvgan_ref.so: the REFERENCE'S OWN hot-path
sources (update_likelihood.cpp, process_mapping.cpp, get_p_obs_base.cpp, get_posterior.cpp, vgan_utils.h,
miscfunc.h, haplocart_functions.cpp, rmdup.cpp, damage.cpp, ...) compiled unchanged from /root/reference/src against
the header shim in oracle/ref_shim/ (recipe: oracle/ref_shim/Makefile).  It pins the oracle's arithmetic.

Only tests/, tests/golden/make_ref_golden.py and bench.py's --impl reference / cpu_baseline legs may import this module.
The library is built in the authoring container (where /root/reference exists) and travels to the GPU box as a built
file; `available()` is False when neither the built library nor the reference sources are present.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libvgan_ref.so")
_SHIM = os.path.join(_HERE, "ref_shim")
REF_ROOT = os.environ.get("VGAN_REFERENCE", "/root/reference")

_lib = None


def build(force=False):
    """Compile the reference's sources where they lie (needs REF_ROOT); a no-op when they are absent."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src")):
        return _LIB_PATH if os.path.exists(_LIB_PATH) else None
    cmd = ["make", "-s", "-C", _SHIM, "REF=" + REF_ROOT] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return _LIB_PATH


def available():
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REF_ROOT, "src"))


def lib():
    global _lib
    if _lib is None:
        build()
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("oracle/_ref/libvgan_ref.so is not built and %s is absent" % REF_ROOT)
        _lib = C.CDLL(_LIB_PATH)
        for f in ("ref_hc_p_seq_error", "ref_hc_background_freq", "ref_hc_region_match_prob"):
            getattr(_lib, f).restype = C.c_double
        _lib.ref_damage_create.restype = C.c_void_p
        assert _lib.ref_sizeof_long_double() == np.dtype(np.longdouble).itemsize
    return _lib


_ptr = orc._ptr


def reconstruct(graph, reads, r, cap=8192):
    """reconstruct_graph_sequence (src/vgan_utils.h:6-79) -> (graph_seq, read_seq, sizes)."""
    g = np.zeros(cap, dtype=np.uint8)
    p = np.zeros(cap, dtype=np.uint8)
    s = np.zeros(cap, dtype=np.int32)
    ng, npp, ns = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib().ref_reconstruct(C.byref(graph.c), C.byref(reads.c), C.c_int64(r), _ptr(g), _ptr(p), _ptr(s), cap,
                               C.byref(ng), C.byref(npp), C.byref(ns))
    if rc:
        raise RuntimeError("ref_reconstruct failed: %d" % rc)
    return g[:ng.value].tobytes().decode(), p[:npp.value].tobytes().decode(), s[:ns.value].tolist()


def quality_pad(reads, r0=None, r1=None):
    """Zero bytes to append to a quality string so that the reference's out-of-range quality reads
    (src/update_likelihood.cpp:40-44, quirk Q4) are defined: the window of the last mapping ends below
    2 * |aligned read|, and |aligned read| <= sum over the read's edits of from_length + to_length."""
    lo = 0 if r0 is None else int(r0)
    hi = reads.n if r1 is None else int(r1)
    if hi <= lo:
        return 8
    e_of_read = reads.edit_off[reads.map_off[lo:hi + 1]]
    span = (reads.edit_from.astype(np.int64) + reads.edit_to.astype(np.int64))
    csum = np.concatenate([[0], np.cumsum(span[int(e_of_read[0]):int(e_of_read[-1])])])
    per_read = csum[e_of_read[1:] - e_of_read[0]] - csum[e_of_read[:-1] - e_of_read[0]]
    return int(2 * (per_read.max() if per_read.size else 0) + 8)


def hc_read(tables, graph, reads, r, pad_quality=None):
    """Haplocart::update (src/update_likelihood.cpp:5-53) for one read from the zero vector -> long double[P]."""
    ll = np.zeros(tables.n_paths, dtype=np.longdouble)
    pad = quality_pad(reads, r, r + 1) if pad_quality is None else int(pad_quality)
    rc = lib().ref_hc_read(C.byref(tables.c), C.byref(graph.c), C.byref(reads.c), C.c_int64(r), pad, _ptr(ll))
    if rc:
        raise RuntimeError("ref_hc_read failed: %d" % rc)
    return ll


def hc_run(tables, graph, reads, r0=0, r1=None, n_threads=1, pad_quality=None, timing=None):
    """The read loop of Haplocart::run (src/HaploCart.cpp:408-424) over the reference's update().
    `timing` (a dict) receives "loop_seconds": the OpenMP loop + max_element alone, tables and reads already built."""
    r1 = reads.n if r1 is None else r1
    final = np.zeros(tables.n_paths, dtype=np.longdouble)
    am = C.c_int32()
    used = C.c_int64()
    pad = quality_pad(reads, r0, r1) if pad_quality is None else int(pad_quality)
    secs = C.c_double(0.0)
    rc = lib().ref_hc_run(C.byref(tables.c), C.byref(graph.c), C.byref(reads.c), C.c_int64(r0), C.c_int64(r1), pad,
                          int(n_threads), _ptr(final), C.byref(am), C.byref(used), C.byref(secs))
    if rc:
        raise RuntimeError("ref_hc_run failed: %d" % rc)
    if timing is not None:
        timing["loop_seconds"] = secs.value
    return final, am.value, used.value


def hc_posterior(final_vec, path_names, parents_txt, children_txt, predicted_idx, cap=4096):
    """Haplocart::get_posterior (src/get_posterior.cpp:87-127).  Returns (clade names as the function wrote them,
    confidences as written (6 significant digits), full-precision confidences from the same member functions, total).
    The first two are None when the function itself crashed on its undefined parent_vec[-1] read (quirk Q8)."""
    fv = np.ascontiguousarray(final_vec, dtype=np.longdouble)
    names = C.create_string_buffer(1 << 20)
    conf_text = np.zeros(cap)
    conf_full = np.zeros(cap)
    total = np.zeros(1, dtype=np.longdouble)
    fd, tmp = tempfile.mkstemp(prefix="vgan_ref_post_")
    os.close(fd)
    try:
        n_text = C.c_int32(0)
        n = lib().ref_hc_posterior(_ptr(fv), fv.size, ("\n".join(path_names)).encode(), parents_txt.encode(),
                                   children_txt.encode(), int(predicted_idx), tmp.encode(), cap, names, len(names),
                                   _ptr(conf_text), C.byref(n_text), _ptr(conf_full), _ptr(total))
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    if n < 0:
        raise RuntimeError("ref_hc_posterior failed: %d" % n)
    if n_text.value < 0:
        # get_posterior itself crashed (in a forked child) on its parent_vec[-1] read, quirk Q8: no text capture
        return None, None, conf_full[:n].copy(), total[0]
    clades = names.value.decode().split("\n")[:-1]
    return clades, conf_text[:n_text.value].copy(), conf_full[:n].copy(), total[0]


def qscore_vec():
    v = np.zeros(100)
    lib().ref_hc_qscore_vec(_ptr(v))
    return v


def incorrect_mapping_vec():
    v = np.zeros(100)
    lib().ref_hc_incorrect_mapping_vec(_ptr(v))
    return v


def region_match_prob(pb):
    return lib().ref_hc_region_match_prob(int(pb))


def p_seq_error(q):
    return lib().ref_hc_p_seq_error(int(q))


def background_freq(b):
    return lib().ref_hc_background_freq(C.c_char(b.encode() if isinstance(b, str) else b))


def dedup(reads, paired=None):
    """Dup_Remover::remove_duplicates_internal (src/rmdup.cpp:68-110) -> is_dup[R]."""
    out = np.zeros(max(reads.n, 1), dtype=np.uint8)
    pr = None if paired is None else np.ascontiguousarray(paired, dtype=np.uint8)
    rc = lib().ref_dedup(C.byref(reads.c), None if pr is None else _ptr(pr), _ptr(out))
    if rc:
        raise RuntimeError("ref_dedup failed: %d" % rc)
    return out[:reads.n].astype(bool)


def facade_flatten(reads, paired=None):
    """The PRODUCT's vgan_b200::FlatAlignments::assign (vgan_b200/csrc/host/vgan_haplocart.hpp) instantiated with the reference's
    own AlignmentInfo (src/AlignmentInfo.h; objects filled as the readGAM lambda fills them, src/readGAM.h:41-46): the flat arrays
    it yields, as a dict keyed like oracle.FlatReads / host.Alignments.arrays()."""
    L = lib()
    pr = None if paired is None else np.ascontiguousarray(paired, dtype=np.uint8)
    cnt = np.zeros(5, dtype=np.int64)
    nulls = [None] * 16
    rc = L.ref_facade_flatten(C.byref(reads.c), None if pr is None else _ptr(pr), _ptr(cnt), *nulls)
    if rc:
        raise RuntimeError("ref_facade_flatten failed: %d" % rc)
    n, nm, ne, nseq, neseq = (int(x) for x in cnt)
    out = {"seq_off": np.zeros(n + 1, np.int64), "seq": np.zeros(max(nseq, 1), np.uint8), "qual_off": np.zeros(n + 1, np.int64),
           "qual": np.zeros(max(int(reads.qual_off[reads.n]), 1), np.uint8), "mapq": np.zeros(max(n, 1), np.int32),
           "identity": np.zeros(max(n, 1), np.float64), "paired": np.zeros(max(n, 1), np.uint8), "map_off": np.zeros(n + 1, np.int64),
           "map_node": np.zeros(max(nm, 1), np.int64), "map_offset": np.zeros(max(nm, 1), np.int64), "map_rev": np.zeros(max(nm, 1), np.uint8),
           "edit_off": np.zeros(nm + 1, np.int64), "edit_from": np.zeros(max(ne, 1), np.int32), "edit_to": np.zeros(max(ne, 1), np.int32),
           "edit_seq_off": np.zeros(ne + 1, np.int64), "edit_seq": np.zeros(max(neseq, 1), np.uint8)}
    order = ["seq_off", "seq", "qual_off", "qual", "mapq", "identity", "paired", "map_off", "map_node", "map_offset", "map_rev", "edit_off",
             "edit_from", "edit_to", "edit_seq_off", "edit_seq"]
    rc = L.ref_facade_flatten(C.byref(reads.c), None if pr is None else _ptr(pr), _ptr(cnt), *[_ptr(out[k]) for k in order])
    if rc:
        raise RuntimeError("ref_facade_flatten failed: %d" % rc)
    cut = {"seq": nseq, "qual": int(out["qual_off"][-1]), "mapq": n, "identity": n, "paired": n, "map_node": nm, "map_offset": nm,
           "map_rev": nm, "edit_from": ne, "edit_to": ne, "edit_seq": neseq}
    return {k: (v[:cut[k]] if k in cut else v) for k, v in out.items()}


class Damage:
    """Damage::initDeamProbabilities (src/damage.cpp:41-323) on two .prof files ('' = no damage)."""

    def __init__(self, deam5p="", deam3p=""):
        self.h = lib().ref_damage_create((deam5p or "").encode(), (deam3p or "").encode())
        if not self.h:
            raise RuntimeError("ref_damage_create failed")

    def matrix(self, L, l):
        out = np.zeros(16)
        rc = lib().ref_damage_matrix(C.c_void_p(self.h), int(L), int(l), _ptr(out))
        if rc:
            raise RuntimeError("ref_damage_matrix failed: %d" % rc)
        return out.reshape(4, 4)

    def rows(self, L):
        return lib().ref_damage_rows(C.c_void_p(self.h), int(L))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_damage_destroy(C.c_void_p(self.h))
            self.h = None


# ---- euka: the readGAM3 lambda (src/readGAM_Euka.h:67-577), the reference's own ----------------------------------
def _write_tmp(text, suffix):
    fd, path = tempfile.mkstemp(prefix="vgan_ref_", suffix=suffix)
    with os.fdopen(fd, "w") as fh:
        fh.write(text)
    return path


def euka_run(clade_dist, bin_off, bin_lo, bin_hi, graph, reads, deam5p_prof="", deam3p_prof="", minimum_mq=29,
             bin_entropy=None):
    """Same arguments as oracle.euka_run (profiles as TEXT); the tables reach the reference through its own loaders
    (Euka::load_clade_info / load_clade_chunks, Damage::initDeamProbabilities) from temporary files.
    Returns per-read clade / like / not_like and per-clade count / per-bin coverage."""
    dist = np.ascontiguousarray(clade_dist, dtype=np.float64)
    boff = np.ascontiguousarray(bin_off, dtype=np.int64)
    blo = np.ascontiguousarray(bin_lo, dtype=np.int32)
    bhi = np.ascontiguousarray(bin_hi, dtype=np.int32)
    nc = dist.size
    ent = np.full(blo.size, 1.5) if bin_entropy is None else np.asarray(bin_entropy, dtype=np.float64)
    clade_txt = "".join("%d clade%d %s 1 %d %d\n" % (c, c, repr(float(dist[c])), int(blo[boff[c]]) if boff[c + 1] > boff[c] else 0,
                                                     int(bhi[boff[c + 1] - 1]) if boff[c + 1] > boff[c] else 0) for c in range(nc))
    bins_txt = "".join("clade%d " % c + " ".join("%d %d %s" % (int(blo[j]), int(bhi[j]), repr(float(ent[j])))
                                                  for j in range(int(boff[c]), int(boff[c + 1]))) + "\n" for c in range(nc))
    files = [_write_tmp(clade_txt, ".clade"), _write_tmp(bins_txt, ".bins")]
    p5 = _write_tmp(deam5p_prof, ".prof") if deam5p_prof else ""
    p3 = _write_tmp(deam3p_prof, ".prof") if deam3p_prof else ""
    files += [f for f in (p5, p3) if f]
    n = reads.n
    out = {"clade": np.zeros(max(n, 1), dtype=np.int32), "like": np.zeros(max(n, 1)), "not_like": np.zeros(max(n, 1)),
           "pass": np.zeros(max(n, 1), dtype=np.uint8)}
    count = np.zeros(nc, dtype=np.int64)
    cov = np.zeros(max(int(boff[-1]), 1), dtype=np.float64)
    try:
        rc = lib().ref_euka_run(files[0].encode(), files[1].encode(), p5.encode(), p3.encode(), C.byref(graph.c), C.byref(reads.c),
                                quality_pad(reads), C.c_uint32(int(minimum_mq)), C.c_int32(nc), _ptr(out["clade"]), _ptr(out["like"]), _ptr(out["not_like"]),
                                _ptr(out["pass"]), _ptr(count), _ptr(cov), C.c_int64(cov.size))
    finally:
        for f in files:
            os.remove(f)
    if rc:
        raise RuntimeError("ref_euka_run failed: %d" % rc)
    for k in ("clade", "like", "not_like", "pass"):
        out[k] = out[k][:n]
    out["clade_count"] = count
    out["bin_cov"] = cov[:int(boff[-1])]
    return out


# ---- soibean: analyse_GAM (src/getLCAfromGAM.h:31-700), the reference's own ------------------------------------------
def soibean_run(graph, reads, member, min_id=0, penalty=7, deam5p_prof="", deam3p_prof=""):
    """Same arguments as oracle.soibean_run (profiles as TEXT) -> (matrix [R, P], used [R])."""
    member = np.ascontiguousarray(member, dtype=np.uint8)
    n_rows, P = member.shape
    out = np.zeros((max(reads.n, 1), P), dtype=np.float64)
    used = np.zeros(max(reads.n, 1), dtype=np.uint8)
    p5 = _write_tmp(deam5p_prof, ".prof") if deam5p_prof else ""
    p3 = _write_tmp(deam3p_prof, ".prof") if deam3p_prof else ""
    try:
        rc = lib().ref_soibean_run(C.byref(graph.c), C.byref(reads.c), _ptr(member), C.c_int64(n_rows), int(min_id), int(P), int(penalty),
                                   p5.encode(), p3.encode(), quality_pad(reads), _ptr(out), _ptr(used))
    finally:
        for f in (p5, p3):
            if f:
                os.remove(f)
    if rc:
        raise RuntimeError("ref_soibean_run failed: %d" % rc)
    return out[:reads.n], used[:reads.n]


def soibean_base_loglike(read_base, graph_base, ll, freqs7, t, branch_len):
    """MCMC::computeBaseLogLike (src/MCMC.h:111-296) for one supported column."""
    f7 = np.ascontiguousarray(freqs7, dtype=np.float64)
    out = C.c_double(0.0)
    rc = lib().ref_soibean_base_loglike(C.c_char(read_base.encode()), C.c_char(graph_base.encode()), C.c_double(ll), _ptr(f7),
                                        C.c_double(t), C.c_double(branch_len), C.byref(out))
    if rc:
        raise RuntimeError("ref_soibean_base_loglike failed: %d" % rc)
    return out.value
