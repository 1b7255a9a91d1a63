"""Python handles over libvgan_host.so (include/vgan_host.h): graph sequences, GAM alignments,
alignment reconstruction, the SoA batch packer, HaploCart tables, the haplogroup tree and the
synthetic workload generator.  All work happens in the C++ library; this file only owns handles."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, host


def _as_array(ptr, n, dtype):
    """Copy n items of dtype from a raw pointer into a numpy array."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr if isinstance(ptr, int) else ptr.value)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class _PinnedBlock:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(_lib.gpu.vgl_host_alloc(C.byref(self.ptr), max(int(nbytes), 1)))

    def __del__(self):
        if self.ptr:
            _lib.gpu.vgl_host_free(self.ptr)
            self.ptr = C.c_void_p()


def pinned_empty(shape, dtype):
    """A numpy array in page-locked host memory (vgl_host_alloc): device-to-host results land in it at PCIe speed."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    blk = _PinnedBlock(n * dt.itemsize)
    buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(blk.ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    arr = arr.view(_PinnedArray)
    arr._block = blk
    return arr


class _PinnedArray(np.ndarray):
    """ndarray view that keeps its pinned block alive."""

    def __array_finalize__(self, obj):
        self._block = getattr(obj, "_block", None)


class Graph:
    """Node sequences with contiguous ids (what the hot path asks of bdsg::ODGI::get_sequence)."""

    def __init__(self, handle, owner=None):
        self.handle = handle
        self._owner = owner  # keeps a synth object alive when the graph belongs to it

    @classmethod
    def from_sequences(cls, node_seqs, min_id=None):
        if isinstance(node_seqs, dict):
            ids = sorted(node_seqs)
            min_id = ids[0] if min_id is None else min_id
            seqs = [node_seqs.get(i, "") for i in range(min_id, ids[-1] + 1)]
        else:
            seqs = list(node_seqs)
            min_id = 1 if min_id is None else min_id
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        blob = "".join(seqs).encode()
        h = C.c_void_p()
        check(host.vgh_graph_create(int(min_id), len(seqs), off.ctypes.data_as(C.c_void_p), blob, C.byref(h)))
        g = cls(h)
        g._owned = True
        return g

    def __del__(self):
        if getattr(self, "_owned", False) and self.handle:
            host.vgh_graph_destroy(self.handle)
            self.handle = None


class Alignments:
    """A set of alignments held by the C++ library (from a GAM stream or the synthetic generator)."""

    def __init__(self, handle):
        self.handle = handle
        self.view = host.vgh_gam_view(handle).contents

    @classmethod
    def from_gam_file(cls, path, drop_unmapped=True):
        h = C.c_void_p()
        check(host.vgh_gam_read_file(str(path).encode(), int(drop_unmapped), C.byref(h)))
        return cls(h)

    @classmethod
    def from_gam_bytes(cls, data, drop_unmapped=True):
        h = C.c_void_p()
        check(host.vgh_gam_parse(data, len(data), int(drop_unmapped), C.byref(h)))
        return cls(h)

    def __len__(self):
        return int(self.view.n_reads)

    def __del__(self):
        if self.handle:
            host.vgh_gam_destroy(self.handle)
            self.handle = None

    def arrays(self):
        """The flat arrays as numpy copies (used by tests to hand the same reads to the oracle)."""
        v = self.view
        n = int(v.n_reads)
        seq_off = _as_array(v.seq_off, n + 1, np.int64)
        qual_off = _as_array(v.qual_off, n + 1, np.int64)
        map_off = _as_array(v.map_off, n + 1, np.int64)
        nm = int(map_off[-1]) if n else 0
        edit_off = _as_array(v.edit_off, nm + 1, np.int64)
        ne = int(edit_off[-1]) if nm else 0
        edit_seq_off = _as_array(v.edit_seq_off, ne + 1, np.int64)
        return {
            "n_reads": n,
            "seq_off": seq_off, "seq": _as_array(v.seq, int(seq_off[-1]), np.uint8),
            "qual_off": qual_off, "qual": _as_array(v.qual, int(qual_off[-1]), np.uint8),
            "mapq": _as_array(v.mapq, n, np.int32), "identity": _as_array(v.identity, n, np.float64),
            "paired": _as_array(v.paired, n, np.uint8),
            "map_off": map_off, "map_node": _as_array(v.map_node, nm, np.int64),
            "map_offset": _as_array(v.map_offset, nm, np.int64), "map_rev": _as_array(v.map_rev, nm, np.uint8),
            "edit_off": edit_off, "edit_from": _as_array(v.edit_from, ne, np.int32),
            "edit_to": _as_array(v.edit_to, ne, np.int32), "edit_seq_off": edit_seq_off,
            "edit_seq": _as_array(v.edit_seq, int(edit_seq_off[-1]) if ne else 0, np.uint8),
        }


def reconstruct(graph, alns, r, cap=8192):
    """reconstruct_graph_sequence of read r -> (graph_seq, read_seq, sizes) (reference src/vgan_utils.h:6-79)."""
    g = np.zeros(cap, dtype=np.uint8)
    p = np.zeros(cap, dtype=np.uint8)
    s = np.zeros(cap, dtype=np.int32)
    ng, npp, ns = C.c_int32(), C.c_int32(), C.c_int32()
    check(host.vgh_reconstruct(graph.handle, C.byref(alns.view), int(r), g.ctypes.data_as(C.c_void_p),
                               p.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), cap, C.byref(ng),
                               C.byref(npp), C.byref(ns)))
    return g[:ng.value].tobytes().decode(), p[:npp.value].tobytes().decode(), s[:ns.value].tolist()


def dedup(alns):
    """Duplicate mask of Dup_Remover::remove_duplicates_internal (reference src/rmdup.cpp:20-110) in one hash pass."""
    mask = np.zeros(len(alns), dtype=np.uint8)
    kept = C.c_int64()
    check(host.vgh_dedup(C.byref(alns.view), mask.ctypes.data_as(C.c_void_p), C.byref(kept)))
    return mask.astype(bool), int(kept.value)


class PackedBatch:
    """One vgl_hc_batch in host memory (pinned when asked)."""

    def __init__(self, graph, alns, r0=0, r1=None, n_threads=0, pinned=False):
        r1 = len(alns) if r1 is None else r1
        self.handle = C.c_void_p()
        check(host.vgh_hc_pack(graph.handle, C.byref(alns.view), int(r0), int(r1), int(n_threads), int(pinned),
                               C.byref(self.handle)))
        self.batch = host.vgh_hc_packed_batch(self.handle).contents
        self.skipped = int(host.vgh_hc_packed_skipped(self.handle))

    @property
    def n_reads(self):
        return int(self.batch.n_reads)

    @property
    def nbytes(self):
        return int(_lib.gpu.vgl_hc_batch_bytes(C.byref(self.batch)))

    def arrays(self):
        b = self.batch
        G, ncb, nmb = int(b.n_groups), int(b.n_col_blks), int(b.n_map_blks)
        L, CB, MB = _lib.LANES, _lib.COL_BLK, _lib.MAP_BLK
        return {
            "grp_col_blk": _as_array(b.grp_col_blk, G + 1, np.uint32), "grp_map_blk": _as_array(b.grp_map_blk, G + 1, np.uint32),
            "col": _as_array(b.col, ncb * L * CB, np.uint32).reshape(ncb, L, CB),
            "map": _as_array(b.map, nmb * L * MB, np.uint32).reshape(nmb, L, MB),
            "n_cols": _as_array(b.n_cols, G * L, np.uint16), "n_close": _as_array(b.n_close, G * L, np.uint16),
            "n_extra": _as_array(b.n_extra, G * L, np.uint16), "mapq": _as_array(b.mapq, G * L, np.uint8),
        }

    def lanes(self):
        """Per-read records, un-blocked: a list of dicts (col records, extra records, n_cols, n_close, mapq), one per
        non-empty lane, in the packer's (shape-sorted) order."""
        a = self.arrays()
        out = []
        for g in range(int(self.batch.n_groups)):
            c0, c1 = int(a["grp_col_blk"][g]), int(a["grp_col_blk"][g + 1])
            m0, m1 = int(a["grp_map_blk"][g]), int(a["grp_map_blk"][g + 1])
            for l in range(_lib.LANES):
                k = g * _lib.LANES + l
                nc, nm = int(a["n_cols"][k]), int(a["n_extra"][k])
                if nc == 0 and nm == 0:
                    continue    # an unused lane (the last group of every mapping-quality segment is filled up with them)
                cols = a["col"][c0:c1, l, :].reshape(-1)
                maps = a["map"][m0:m1, l, :].reshape(-1)
                assert not cols[nc:].any() and not maps[nm:].any()  # padding records are zero
                out.append({"slot": k, "col": cols[:nc].tolist(), "extra": maps[:nm].tolist(), "n_cols": nc,
                            "n_close": int(a["n_close"][k]), "mapq": int(a["mapq"][k])})
        return out

    def __del__(self):
        if self.handle:
            host.vgh_hc_packed_destroy(self.handle)
            self.handle = None


class WireBatch:
    """A packed batch re-encoded for the host -> device transfer (vgl_hc_wire): 16-bit column records (12-bit ones -- col8 / flag4 /
    n_owned -- when every quality byte is <= 63) plus one 16-bit node id
    per non-empty mapping instead of 32-bit records."""

    def __init__(self, packed, n_threads=0, pinned=False):
        self.handle = C.c_void_p()
        check(host.vgh_hc_wire_from_packed(packed.handle, int(n_threads), int(pinned), C.byref(self.handle)))
        self.wire = host.vgh_hc_wire_view(self.handle).contents

    @property
    def n_reads(self):
        return int(self.wire.n_reads)

    @property
    def nbytes(self):
        return int(_lib.gpu.vgl_hc_wire_bytes(C.byref(self.wire)))

    def __del__(self):
        if self.handle:
            host.vgh_hc_wire_destroy(self.handle)
            self.handle = None


class RawBatch:
    """The raw alignments of reads [r0, r1) in the compact form the device-side reconstruction reads (vgl_hc_raw)."""

    def __init__(self, alns, r0=0, r1=None, pinned=False, euka_rule=False):
        """euka_rule: keep the reads readGAM3 keeps (identity != 0) instead of HaploCart's identity >= 1e-10."""
        r1 = len(alns) if r1 is None else r1
        self.handle = C.c_void_p()
        self.n_span = int(r1 - r0)
        build = host.vgh_euka_raw_build if euka_rule else host.vgh_hc_raw_build
        check(build(C.byref(alns.view), int(r0), int(r1), int(pinned), C.byref(self.handle)))
        self.raw = host.vgh_hc_raw_view(self.handle).contents
        self.skipped = int(host.vgh_hc_raw_skipped(self.handle))

    @property
    def n_reads(self):
        return int(self.raw.n_reads)

    @property
    def nbytes(self):
        return int(_lib.gpu.vgl_hc_raw_bytes(C.byref(self.raw)))

    def __del__(self):
        if self.handle:
            host.vgh_hc_raw_destroy(self.handle)
            self.handle = None


class GamFrames:
    """vgh_gam_frame: a GAM stream inflated and its framing walked -- offset and length of every serialized vg.Alignment --
    for the device-side decoder (Haplocart.update_gam / decode_gam); the messages are not parsed on the host."""

    def __init__(self, data=None, path=None):
        if path is not None:
            with open(path, "rb") as fh:
                data = fh.read()
        self._data = bytes(data)
        self.handle = C.c_void_p()
        src = C.cast(C.c_char_p(self._data), C.c_void_p) if self._data else None     # the bytes object's own buffer, no copy
        check(host.vgh_gam_frame(src, len(self._data), C.byref(self.handle)))
        b, off, ln = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb = C.c_uint64()
        self.n_msgs = int(host.vgh_gam_frames_view(self.handle, C.byref(b), C.byref(nb), C.byref(off), C.byref(ln)))
        self.bytes_ptr, self.n_bytes, self.off_ptr, self.len_ptr = b, int(nb.value), off, ln

    def __del__(self):
        if getattr(self, "handle", None):
            host.vgh_gam_frames_destroy(self.handle)
            self.handle = None


class Feeder:
    """vgh_hc_feeder: the drop-in for the OpenMP read loop (src/HaploCart.cpp:408-421) from flat alignment fields -- chunks of
    reads marshalled by `n_threads` host threads into two page-locked blocks while the GPU works on the chunk before."""

    def __init__(self, n_threads=0, chunk_reads=131072):
        self.handle = C.c_void_p()
        check(host.vgh_hc_feeder_create(int(n_threads), int(chunk_reads), C.byref(self.handle)))

    def feed(self, hc, alns, r0=0, r1=None):
        """Add reads [r0, r1) of `alns` to the sample open on `hc` (between reset and get_posterior)."""
        r1 = len(alns) if r1 is None else r1
        check(host.vgh_hc_feed(self.handle, hc.ctx, C.byref(alns.view), int(r0), int(r1)))

    def feed_euka(self, ek, alns, r0=0, r1=None, out=None):
        """The reads of [r0, r1) through vgl_euka_run_raw chunk by chunk (vgh_euka_feed): the per-read results of Euka.run, indexed
        by the read's position in [r0, r1).  `out` may hold caller-provided (page-locked) arrays under the same keys."""
        r1 = len(alns) if r1 is None else r1
        n = int(r1 - r0)
        if out is None:
            out = {"clade": np.full(n, -1, dtype=np.int32), "in_lik": np.zeros(n), "not_lik": np.zeros(n), "like": np.zeros(n),
                   "pass": np.zeros(n, dtype=np.uint8)}
        p = lambda k: out[k].ctypes.data_as(C.c_void_p) if out.get(k) is not None else None
        check(host.vgh_euka_feed(self.handle, ek.ctx, C.byref(alns.view), int(r0), int(r1), p("clade"), p("in_lik"), p("not_lik"),
                                 p("like"), p("pass")))
        return out

    def feed_soibean(self, sb, alns, r0=0, r1=None, out=None, used=None):
        """The reads of [r0, r1) through vgl_soibean_run_raw chunk by chunk (vgh_soibean_feed): (matrix [n, P], used [n])."""
        r1 = len(alns) if r1 is None else r1
        n, P = int(r1 - r0), int(sb.tables.n_paths)
        out = np.zeros((n, P), dtype=np.float64) if out is None else out
        used = np.zeros(n, dtype=np.uint8) if used is None else used
        check(host.vgh_soibean_feed(self.handle, sb.ctx, C.byref(alns.view), int(r0), int(r1), P, out.ctypes.data_as(C.c_void_p),
                                    used.ctypes.data_as(C.c_void_p)))
        return out, used

    def marshal(self, alns, r0=0, r1=None, sibling=False):
        """One chunk marshalled into the feeder's first block (inspection / tests): a vgl_hc_raw, or None when the chunk
        needs the general builder.  sibling=True: the euka / soibean form (identity != 0 kept; seq_len and read_index filled)."""
        r1 = len(alns) if r1 is None else r1
        out = C.POINTER(_lib.vgl_hc_raw)()
        fn = host.vgh_euka_feeder_marshal if sibling else host.vgh_hc_feeder_marshal
        check(fn(self.handle, C.byref(alns.view), int(r0), int(r1), C.byref(out)))
        return out.contents if out else None

    def stats(self):
        b, w = C.c_double(), C.c_double()
        n, c, sl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        host.vgh_hc_feeder_stats(self.handle, C.byref(b), C.byref(w), C.byref(n), C.byref(c), C.byref(sl))
        return {"build_seconds": b.value, "wait_seconds": w.value, "h2d_bytes": n.value, "chunks": c.value, "slow_chunks": sl.value}

    def close(self):
        if self.handle:
            host.vgh_hc_feeder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


class HcTables:
    """HaploCart tables in the form vgl_hc_set_tables wants, built from the reference's in-memory inputs."""

    def __init__(self, supports, pmap_node, pmap_val, mappab, bep=1e-4, is_fasta=False):
        supports = np.ascontiguousarray(supports, dtype=np.uint8)
        pmap_node = np.ascontiguousarray(pmap_node, dtype=np.int64)
        pmap_val = np.ascontiguousarray(pmap_val, dtype=np.int32)
        mappab = np.ascontiguousarray(mappab, dtype=np.float64)
        self.n_rows, self.n_paths = supports.shape
        self.handle = C.c_void_p()
        check(host.vgh_hc_tables_build(self.n_rows, self.n_paths, supports.ctypes.data_as(C.c_void_p), pmap_node.size,
                                       pmap_node.ctypes.data_as(C.c_void_p), pmap_val.ctypes.data_as(C.c_void_p),
                                       mappab.size, mappab.ctypes.data_as(C.c_void_p), float(bep), int(bool(is_fasta)),
                                       C.byref(self.handle)))
        self.view = host.vgh_hc_tables_view(self.handle).contents

    def __del__(self):
        if self.handle:
            host.vgh_hc_tables_destroy(self.handle)
            self.handle = None


class HcTree:
    def __init__(self, path_names, parents_txt, children_txt):
        self.path_names = list(path_names)
        self.handle = C.c_void_p()
        check(host.vgh_hc_tree_build(len(self.path_names), ("\n".join(self.path_names)).encode(), parents_txt.encode(),
                                     children_txt.encode(), C.byref(self.handle)))
        self.n_clades = int(host.vgh_hc_tree_n_clades(self.handle))
        self.row_words = int(host.vgh_hc_tree_row_words(self.handle))
        self.anc_off = _as_array(host.vgh_hc_tree_anc_off(self.handle), len(self.path_names) + 1, np.int32)
        self.anc_clade = _as_array(host.vgh_hc_tree_anc_clade(self.handle), int(self.anc_off[-1]), np.int32)

    def clade_bits(self):
        return _as_array(host.vgh_hc_tree_clade_bits(self.handle), max(self.n_clades, 1) * self.row_words,
                         np.uint32).reshape(max(self.n_clades, 1), self.row_words)

    def clade_name(self, c):
        return host.vgh_hc_tree_clade_name(self.handle, int(c)).decode()

    def clades_of(self, path_idx):
        """Names of the clades reported for a call of `path_idx` (reference clade_vec[1:], get_posterior.cpp:109-112)."""
        a, b = int(self.anc_off[path_idx]), int(self.anc_off[path_idx + 1])
        return [self.clade_name(c) for c in self.anc_clade[a:b]]

    def __del__(self):
        if self.handle:
            host.vgh_hc_tree_destroy(self.handle)
            self.handle = None


class SynthHaplocart:
    """Synthetic HaploCart-shaped tables + graph (11825 rows x 5179 paths, 11821 nodes)."""

    def __init__(self, seed=20260, n_paths=5179, n_rows=11825, n_nodes=11821):
        self.n_paths, self.n_rows, self.n_nodes = n_paths, n_rows, n_nodes
        self.handle = C.c_void_p()
        check(host.vgh_synth_hc_create(int(seed), n_paths, n_rows, n_nodes, C.byref(self.handle)))
        self.graph = Graph(C.c_void_p(host.vgh_synth_hc_graph(self.handle)), owner=self)
        self.total_len = int(host.vgh_synth_hc_total_len(self.handle))
        self.path_names = host.vgh_synth_hc_names(self.handle).decode().split("\n")[:-1]
        self.parents_txt = host.vgh_synth_hc_parents(self.handle).decode()
        self.children_txt = host.vgh_synth_hc_children(self.handle).decode()

    # the reference's in-memory inputs, as numpy copies
    def supports(self):
        return _as_array(host.vgh_synth_hc_supports(self.handle), self.n_rows * self.n_paths,
                         np.uint8).reshape(self.n_rows, self.n_paths)

    def pmap(self):
        n = int(host.vgh_synth_hc_n_pmap(self.handle))
        return (_as_array(host.vgh_synth_hc_pmap_node(self.handle), n, np.int64),
                _as_array(host.vgh_synth_hc_pmap_val(self.handle), n, np.int32))

    def mappab(self):
        return _as_array(host.vgh_synth_hc_mappab(self.handle), int(host.vgh_synth_hc_n_mappab(self.handle)), np.float64)

    def node_sequences(self):
        off, seq = C.c_void_p(), C.c_void_p()
        n = int(host.vgh_synth_hc_graph_seq(self.handle, C.byref(off), C.byref(seq)))
        offs = _as_array(off, n + 1, np.int64)
        blob = _as_array(seq, int(offs[-1]), np.uint8).tobytes().decode()
        return [blob[offs[i]:offs[i + 1]] for i in range(n)]

    def tables(self, bep=1e-4, is_fasta=False):
        node, val = self.pmap()
        return HcTables(self.supports(), node, val, self.mappab(), bep=bep, is_fasta=is_fasta)

    def tree(self):
        return HcTree(self.path_names, self.parents_txt, self.children_txt)

    def reads(self, n_reads, seed=20261, fixed_len=100, sub_rate=0.005, indel_rate=0.0005, damage=False, origin_path=-1):
        h = C.c_void_p()
        check(host.vgh_synth_hc_reads(self.handle, int(seed), int(n_reads), int(fixed_len), float(sub_rate),
                                      float(indel_rate), int(bool(damage)), int(origin_path), C.byref(h)))
        return Alignments(h)

    def __del__(self):
        if self.handle:
            host.vgh_synth_hc_destroy(self.handle)
            self.handle = None
