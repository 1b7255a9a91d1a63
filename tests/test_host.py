"""CPU tests of the C++ host side (libvgan_host.so) and of the C-ABI surface.  No GPU compute here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import vgan_b200 as vb
from vgan_b200 import _lib
from oracle import oracle as orc

import helpers as H

FIXTURES = ["test_reads.gam", "J2a1a1a1.gam", "all_the_same.gam", "all_the_same_reverse.gam", "two_unique.gam"]


def _flat_equal(arrays, alns):
    """Compare the product's flat arrays with a list of alignment dicts."""
    ref = orc.Reads(alns)
    n = len(alns)
    assert arrays["n_reads"] == n
    for k in ("seq_off", "qual_off", "map_off", "edit_off", "edit_seq_off"):
        assert np.array_equal(arrays[k], getattr(ref, k)), k
    nm, ne = int(ref.map_off[-1]), int(ref.edit_off[-1])
    assert np.array_equal(arrays["seq"], ref.seq[:-1]) and np.array_equal(arrays["qual"], ref.qual[:-1])
    assert np.array_equal(arrays["mapq"], ref.mapq) and np.array_equal(arrays["identity"], ref.identity)
    assert np.array_equal(arrays["map_node"], ref.map_node[:nm]) and np.array_equal(arrays["map_offset"], ref.map_offset[:nm])
    assert np.array_equal(arrays["map_rev"], ref.map_rev[:nm])
    assert np.array_equal(arrays["edit_from"], ref.edit_from[:ne]) and np.array_equal(arrays["edit_to"], ref.edit_to[:ne])
    assert np.array_equal(arrays["edit_seq"], ref.edit_seq[:-1])


@pytest.mark.parametrize("name", FIXTURES)
def test_gam_reader_on_reference_fixtures(name):
    a = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, name))
    _flat_equal(a.arrays(), H.decoded_fixture(name))


def test_gam_reader_drops_identity_zero_and_skips_empty_groups():
    alns = [H.aln("ACGT", [(5, 0, False, [(4, 4, "")])], identity=0.0),
            H.aln("AATA", [(5, 0, False, [(4, 4, "")])], identity=0.5)]
    raw = H.encode_gam(alns[:1]) + bytes([1, 3]) + b"GAM" + H.encode_gam(alns[1:])
    a = vb.Alignments.from_gam_bytes(raw)
    assert len(a) == 1 and a.arrays()["seq"].tobytes() == b"AATA"   # readGAM keeps identity != 0 only
    assert len(vb.Alignments.from_gam_bytes(raw, drop_unmapped=False)) == 2
    with pytest.raises(_lib.VglError):
        vb.Alignments.from_gam_bytes(raw[:-3])


def test_reconstruct_known_answers():
    kat = json.load(open(os.path.join(H.GOLDEN, "reconstruct_kat.json")))["cases"]
    nodes, _ = H.toy_graph()
    g = vb.Graph.from_sequences(nodes)
    a = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, "test_reads.gam"))
    for case in kat:
        gs, rs, sizes = vb.reconstruct(g, a, case["read_index"])
        assert (gs, rs) == (case["graph_seq"], case["read_seq"]), case["case"]


def test_reconstruct_matches_oracle_on_synthetic_reads_with_indels():
    S = vb.SynthHaplocart(seed=11, n_paths=40, n_rows=310, n_nodes=301)
    A = S.reads(400, seed=12, sub_rate=0.03, indel_rate=0.02, damage=True, fixed_len=0)
    R = orc.FlatReads(A.arrays())
    og = orc.Graph(S.node_sequences(), min_id=1)
    kinds = set()
    for r in range(len(A)):
        want = orc.reconstruct(og, R, r)
        assert vb.reconstruct(S.graph, A, r) == want
        kinds |= set(want[0]) | set(want[1])
    assert {"-", "S"} <= kinds  # deletions, insertions and the offset-0 "softclip" quirk all occurred


def test_reconstruct_unrecognised_edit_repeats_previous_size():
    # an edit with from != to (both > 0) matches no branch of the reference: nothing is appended and the
    # previous aligned_length of the same mapping is pushed again (src/vgan_utils.h:26,72)
    nodes, _ = H.toy_graph()
    alns = [H.aln("GCTACGA", [(9, 0, False, [(4, 4, ""), (2, 3, "CGA")])])]
    a = vb.Alignments.from_gam_bytes(H.encode_gam(alns))
    got = vb.reconstruct(vb.Graph.from_sequences(nodes), a, 0)
    assert got == orc.reconstruct(orc.Graph(nodes), orc.Reads(alns), 0)
    assert got[2] == [4, 4]


def _repack_with_oracle(og, R, arrays):
    """What each packed read lane must contain, derived independently from the oracle's reconstruction and the
    marshalling rule of update_likelihood / process_mapping (src/update_likelihood.cpp:33-49)."""
    keep = [r for r in range(arrays["n_reads"]) if not arrays["identity"][r] < 1e-10]
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    NO_NODE = 0x1FFFF
    lanes = []
    for r in keep:
        gs, rs, sizes = orc.reconstruct(og, R, r)
        n = len(rs)
        M = int(arrays["map_off"][r + 1] - arrays["map_off"][r])
        qual = arrays["qual"][arrays["qual_off"][r]:arrays["qual_off"][r + 1]]
        q90 = [c for c in range(len(qual)) if 90 <= int(qual[c]) < 128]
        nodes = arrays["map_node"][arrays["map_off"][r]:arrays["map_off"][r + 1]].tolist()
        # per column: quality byte, owner node, flags
        # canonical quality: every byte whose error probability the reference takes as 0.25 (Q <= 2, or negative as signed
        # char) is stored as 0; the field sits in bits 3-10 (include/vgan_gpu.h)
        canon = lambda b: 0 if (b >= 128 or b <= 2) else b
        qb = [canon(int(qual[c])) << 3 if c < len(qual) else 0 for c in range(n)]
        owner, flags = [NO_NODE] * n, [0] * n
        pos, extras, starts = 0, [], set()
        for i in range(M):
            lg = min(sizes[i], n - pos)
            latched = 0x0800 if q90 and q90[0] < pos + n else 0   # a byte >= 90 inside the window [pos, pos + n)
            for j in range(lg):
                g, b = gs[pos + j], rs[j]            # the read string is indexed from 0 for every mapping (quirk Q2)
                owner[pos + j] = nodes[i]
                flags[pos + j] = latched
                if g in code and b in code:
                    flags[pos + j] |= 0x1000 | (code[b] << 1) | (0 if g == b else 0x0001)
            if lg:
                flags[pos + lg - 1] |= 0x2000        # OPEN on the mapping's last column
            else:
                extras.append(nodes[i] | (pos << 20))
            if pos < n:
                starts.add(pos)
            pos += lg
        rec = [qb[c] | flags[c] | (0x4000 if c in starts else 0) | (owner[c] << 15) for c in range(n)]
        lanes.append({"col": rec, "extra": extras, "n_cols": n, "n_close": len(starts), "cov_end": pos,
                      "mapq": int(arrays["mapq"][r])})
    return lanes, keep


def _lane_key(d):
    return (d["n_close"], d["n_cols"], tuple(d["col"]), tuple(d["extra"]), d["mapq"])


def test_packer_against_oracle_reconstruction():
    S = vb.SynthHaplocart(seed=21, n_paths=33, n_rows=310, n_nodes=301)
    A = S.reads(300, seed=22, sub_rate=0.02, indel_rate=0.01)
    arrays = A.arrays()
    R = orc.FlatReads(arrays)
    og = orc.Graph(S.node_sequences(), min_id=1)
    pb = vb.PackedBatch(S.graph, A, n_threads=3)
    want, keep = _repack_with_oracle(og, R, arrays)
    got = pb.lanes()
    assert pb.n_reads == len(keep) == len(got) and pb.skipped == len(A) - len(keep)
    # the packer groups reads by shape, so compare as multisets of lanes
    assert sorted(map(_lane_key, got)) == sorted(map(_lane_key, want))
    # ... within a mapping quality: groups never mix two mapping qualities (segments, vgl_hc_batch_meta)
    shapes = [(d["mapq"], d["n_close"], d["n_cols"]) for d in got]
    assert shapes == sorted(shapes)
    # mis-slicing (quirk Q1) happened: some read has more edits than mappings, so some columns have no owner
    assert any(d["cov_end"] < d["n_cols"] for d in want)
    assert any((c >> 15) == 0x1FFFF for d in got for c in d["col"])
    # padding: rows of a group are as long as its longest lane, unused records are zero
    a = pb.arrays()
    G = int(pb.batch.n_groups)
    meta = pb.batch.meta
    mq = sorted(set(d["mapq"] for d in got))
    assert meta.valid == 1 and meta.n_segs == len(mq) and list(meta.seg_mapq[:meta.n_segs]) == mq and meta.seg_end[meta.n_segs - 1] == G
    assert G == sum((sum(1 for d in got if d["mapq"] == q) + 31) // 32 for q in mq)   # every segment's last group is filled up
    assert meta.qual_max == max((c >> 3) & 0xff for d in got for c in d["col"])
    lane_mapq = a["mapq"].reshape(G, 32)
    lane_cols = a["n_cols"].reshape(G, 32)
    lane_x = a["n_extra"].reshape(G, 32)
    g0 = 0
    for k in range(meta.n_segs):
        g1 = int(meta.seg_end[k])
        used = (lane_cols[g0:g1] > 0) | (lane_x[g0:g1] > 0)
        assert np.all(lane_mapq[g0:g1][used] == meta.seg_mapq[k]) and np.all(lane_mapq[g0:g1][~used] == 0)
        g0 = g1
    for g in range(G):
        assert a["grp_col_blk"][g + 1] - a["grp_col_blk"][g] == (int(lane_cols[g].max()) + 3) // 4
        assert a["grp_map_blk"][g + 1] - a["grp_map_blk"][g] == (int(lane_x[g].max()) + 3) // 4
    assert pb.nbytes == (G + 1) * 8 + 512 * (int(a["grp_col_blk"][-1]) + int(a["grp_map_blk"][-1])) + 7 * 32 * G
    # a quality byte >= 90 latches the background-error mode of every mapping whose window holds it
    alns = [H.aln("GCTACGATGG", [(9, 0, False, [(10, 10, "")])], quality=bytes([30, 30, 30, 95, 30, 200, 30, 30, 99, 30]))]
    nodes, _ = H.toy_graph()
    one = vb.PackedBatch(vb.Graph.from_sequences(nodes), vb.Alignments.from_gam_bytes(H.encode_gam(alns))).lanes()[0]
    assert all(c & 0x0800 for c in one["col"]) and one["col"][0] & 0x4000 and one["col"][-1] & 0x2000
    assert [(c >> 3) & 0xff for c in one["col"]] == [30, 30, 30, 95, 30, 0, 30, 30, 99, 30]    # the negative byte 200 is stored as 0
    # mappings that own no column travel as extra records
    alns = [H.aln("GCTAC", [(9, 0, False, [(0, 0, ""), (5, 5, "")]), (9, 5, False, [(3, 3, "")])])]
    two = vb.PackedBatch(vb.Graph.from_sequences(nodes), vb.Alignments.from_gam_bytes(H.encode_gam(alns))).lanes()[0]
    assert two["extra"] == [9] and two["n_close"] == 1


def test_packer_rejects_what_the_reference_cannot_process():
    nodes, _ = H.toy_graph()
    g = vb.Graph.from_sequences(nodes)
    cases = {
        "mapping without edits": H.aln("ACGT", [(9, 0, False, [(4, 4, "")]), (10, 0, False, [])]),
        "node outside the graph": H.aln("ACGT", [(999, 0, False, [(4, 4, "")])]),
        "offset beyond the node": H.aln("ACGT", [(10, 5, False, [(1, 1, "")])]),
        "mapping quality out of range": H.aln("GCTA", [(9, 0, False, [(4, 4, "")])], mapq=100),
        "quality longer than the alignment": H.aln("GCTA", [(9, 0, False, [(4, 4, "")])], quality=bytes([30] * 9)),
        "substitution shorter than its span": H.aln("GCTA", [(9, 0, False, [(4, 4, "GC")])]),
        "alignment too long for the mapping record": H.aln("A" * 5000, [(9, 0, False, [(0, 5000, "A" * 5000)])]),
    }
    for why, bad in cases.items():
        a = vb.Alignments.from_gam_bytes(H.encode_gam([bad]))
        with pytest.raises(_lib.VglError):
            vb.PackedBatch(g, a)
    ok = vb.PackedBatch(g, vb.Alignments.from_gam_bytes(H.encode_gam([H.aln("GCTA", [(9, 0, False, [(4, 4, "")])])])))
    assert ok.n_reads == 1
    empty = vb.PackedBatch(g, vb.Alignments.from_gam_bytes(b""))
    assert empty.n_reads == 0


def test_tables_match_oracle_tables():
    S = vb.SynthHaplocart(seed=31, n_paths=70, n_rows=310, n_nodes=301)
    T = S.tables(bep=2e-4)
    v = T.view
    assert np.array_equal(np.array(v.qscore[:]), orc.qscore_vec())
    assert np.array_equal(np.array(v.p_incorrect_map[:]), orc.incorrect_mapping_vec())
    assert [orc.lib().orc_hc_background_freq(C.c_char(b)) for b in b"ACGT"] == list(v.bg_freq[:])
    for pb in list(range(0, 17000, 7)) + [56, 57, 372, 373, 576, 577, 16383, 16384, 16569, 16570]:
        assert _lib.host.vgh_hc_region_match_prob(pb) == orc.region_match_prob(pb), pb
    sup = S.supports()
    bits = vb.host._as_array(v.support_bits, v.n_rows * v.row_words, np.uint32).reshape(v.n_rows, v.row_words)
    unpacked = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(v.n_rows, -1)[:, :v.n_paths]
    assert np.array_equal(unpacked, sup)
    node, val = S.pmap()
    mapp = S.mappab()
    nm = vb.host._as_array(v.node_mappability, v.n_rows, np.float64)
    nk = vb.host._as_array(v.node_match_prob, v.n_rows, np.float64)
    valid = vb.host._as_array(v.node_valid, v.n_rows, np.uint8)
    assert valid.sum() == len(set(node.tolist()))
    for n, pb in zip(node, val):
        assert nm[n] == mapp[pb] and nk[n] == orc.region_match_prob(int(pb))


def test_tree_builder_against_plain_python():
    S = vb.SynthHaplocart(seed=41, n_paths=150, n_rows=310, n_nodes=301)
    tr = S.tree()

    def rel(txt):
        out = {}
        for line in txt.split("\n"):
            t = line.split()
            if t and t[0] not in out:
                out[t[0]] = [x for x in t[1:] if "[" not in x]
        return out
    parents, children = rel(S.parents_txt), rel(S.children_txt)
    col = {n: i for i, n in enumerate(S.path_names)}
    bits = tr.clade_bits()
    for p, name in enumerate(S.path_names):
        pv = parents.get(name, [])
        want = [x for j, x in enumerate(pv) if j == 0 or x != pv[j - 1]]
        assert tr.clades_of(p) == want
    for c in range(tr.n_clades):
        desc, level = set(), {tr.clade_name(c)}
        while level:
            nxt = {k for x in level for k in children.get(x, [])}
            desc |= nxt
            level = nxt
        want = np.zeros(len(S.path_names), dtype=bool)
        for d in desc:
            if d in col:
                want[col[d]] = True
        got = ((bits[c][:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:len(S.path_names)].astype(bool)
        assert np.array_equal(got, want), tr.clade_name(c)
    assert "mt-MRCA" in [tr.clade_name(c) for c in range(tr.n_clades)]


def test_synthetic_generator_shapes_and_determinism():
    S = vb.SynthHaplocart()  # config 2 shapes
    assert (S.n_rows, S.n_paths, S.n_nodes) == (11825, 5179, 11821)      # reference src/test.cpp:605-655
    seqs = S.node_sequences()
    assert len(seqs) == 11821 and max(map(len, seqs)) <= 8 and 15000 < S.total_len < 18500
    sup = S.supports()
    assert sup.shape == (11825, 5179) and sup[0].sum() == 0 and sup[11822:].sum() == 0
    a1 = S.reads(2000, seed=5).arrays()
    a2 = vb.SynthHaplocart().reads(2000, seed=5).arrays()
    assert all(np.array_equal(a1[k], a2[k]) for k in a1 if k != "n_reads")
    nm = np.diff(a1["map_off"])
    assert 40 < nm.mean() < 56 and set(np.unique(a1["mapq"])) <= {0, 1, 3, 60}
    # every read walks a real path: its nodes are all supported by at least one common path
    r = 17
    nodes = a1["map_node"][a1["map_off"][r]:a1["map_off"][r + 1]]
    assert sup[nodes].all(axis=0).any()


def _declared(header):
    txt = open(os.path.join(os.path.dirname(H.GOLDEN), "..", "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vg[lh]_[a-z0-9_]+)\s*\(", txt)))


def test_libraries_export_every_declared_symbol():
    for header, lib, table in (("vgan_gpu.h", _lib.gpu, _lib.GPU_SYMBOLS), ("vgan_host.h", _lib.host, _lib.HOST_SYMBOLS)):
        names = _declared(header)
        assert len(names) > 10
        for n in names:
            assert hasattr(lib, n), "%s declared in %s but not exported" % (n, header)
            assert n in table, "%s has no ctypes signature" % n
    assert _lib.gpu.vgl_abi_version() == 2


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.VglError) as ei:
        vb.Haplocart(device=0)
    assert ei.value.code == 6  # VGL_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(H.GOLDEN), "..", "vgan_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "vgan_oracle" not in txt, f


def test_dedup_matches_the_reference_quadratic_scan():
    S = vb.SynthHaplocart(seed=61, n_paths=20, n_rows=130, n_nodes=121)   # tiny graph: many shared start coordinates
    A = S.reads(3000, seed=62)
    arr = A.arrays()
    want = orc.dedup(orc.FlatReads(arr), arr["paired"])
    got, kept = vb.host.dedup(A)
    assert np.array_equal(got, want) and kept == int((~want).sum())
    assert 0 < kept < 600                                               # ~2 x 220 distinct (node, offset) starts at most
    # strand is ignored, the offset is not
    a = [H.aln("GCTA", [(9, 0, False, [(4, 4, "")])]), H.aln("TAGC", [(9, 0, True, [(4, 4, "")])]),
         H.aln("CTAC", [(9, 1, False, [(4, 4, "")])]), H.aln("GCTA", [(9, 0, False, [(4, 4, "")])])]
    X = vb.Alignments.from_gam_bytes(H.encode_gam(a))
    assert vb.host.dedup(X)[0].tolist() == [False, True, False, True]
    assert np.array_equal(orc.dedup(orc.Reads(a)), [False, True, False, True])
    # a real fixture: all_the_same.gam is 128 copies of one alignment
    import os
    F = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, "all_the_same.gam"))
    m, kept = vb.host.dedup(F)
    assert kept == 1 and m[1:].all() and np.array_equal(m, orc.dedup(orc.FlatReads(F.arrays())))


def test_wire_form_of_a_packed_batch():
    """The wire form holds the 32-bit records minus the node field, and the nodes once per mapping, last mapping first."""
    S = vb.SynthHaplocart(seed=21, n_paths=33, n_rows=310, n_nodes=301)
    A = S.reads(200, seed=23, sub_rate=0.02, indel_rate=0.01)
    pb = vb.PackedBatch(S.graph, A)
    wb = vb.WireBatch(pb)
    w = wb.wire
    G = int(w.n_groups)
    gc = vb.host._as_array(w.grp_col_blk, G + 1, np.uint32)
    gn = vb.host._as_array(w.grp_node_blk, G + 1, np.uint32)
    col = vb.host._as_array(w.col, int(w.n_col_blks) * 256, np.uint16).reshape(-1, 32, 8)
    node = vb.host._as_array(w.node, int(w.n_node_blks) * 256, np.uint16).reshape(-1, 32, 8)
    assert gc[-1] == w.n_col_blks and gn[-1] == w.n_node_blks
    lanes = pb.lanes()
    for k, d in enumerate(lanes):
        g, l = divmod(d["slot"], 32)
        c16 = col[gc[g]:gc[g + 1], l, :].reshape(-1)
        n16 = node[gn[g]:gn[g + 1], l, :].reshape(-1)
        recs = d["col"]
        assert [int(x) & 0x7fff for x in c16[:len(recs)]] == [r & 0x7fff for r in recs] and not c16[len(recs):].any()
        assert [bool(x & 0x8000) for x in c16[:len(recs)]] == [(r >> 15) != 0x1FFFF for r in recs]
        opens = [r >> 15 for r in reversed(recs) if r & 0x2000]
        assert n16[:len(opens)].tolist() == opens and not n16[len(opens):].any()
    assert wb.nbytes < 0.8 * pb.nbytes
    # narrow columns (12 bits): present because every quality is <= 63, and they decode to the same 16-bit records
    assert w.col8 and w.flag4 and w.n_owned
    c8 = vb.host._as_array(w.col8, int(w.n_col_blks) * 256, np.uint8).reshape(-1, 32, 8)
    f4 = vb.host._as_array(w.flag4, int(w.n_col_blks) * 32, np.uint32).reshape(-1, 32)
    own = vb.host._as_array(w.n_owned, G * 32, np.uint16)
    for k, d in enumerate(lanes):
        g, l = divmod(d["slot"], 32)
        byte = c8[gc[g]:gc[g + 1], l, :].reshape(-1).astype(np.uint32)
        word = f4[gc[g]:gc[g + 1], l]
        nib = np.array([(int(word[c // 8]) >> (4 * (c % 8))) & 0xF for c in range(byte.size)], dtype=np.uint32)
        c = np.arange(byte.size)
        rec = ((byte & 0x3F) << 3) | (nib & 1) | ((byte >> 6) << 1) | ((nib & 8) << 9) | ((nib & 2) << 12) | ((nib & 4) << 12) | \
              np.where(c < own[d["slot"]], 0x8000, 0).astype(np.uint32)
        assert np.array_equal(rec.astype(np.uint16), col[gc[g]:gc[g + 1], l, :].reshape(-1)), k
    narrow_bytes = wb.nbytes
    w.col8 = None
    assert wb.nbytes > narrow_bytes * 1.12                     # the 16-bit columns are what is counted now
    # a quality byte above 63 (here: a latching one) keeps the whole batch on 16-bit columns
    hi = [H.aln("ACGTACGT", [(1, 0, False, [(8, 8, "")])], quality=bytes([30, 30, 95, 30, 30, 30, 30, 30]))]
    g1 = vb.Graph.from_sequences({1: "ACGTACGTAC"})
    wb2 = vb.WireBatch(vb.PackedBatch(g1, vb.Alignments.from_gam_bytes(H.encode_gam(hi))))
    assert not wb2.wire.col8 and not wb2.wire.flag4 and not wb2.wire.n_owned


def _bgzf(data, block=1500):
    """BGZF as htslib / vg write it: gzip members with a 'BC' extra subfield carrying the member size, then the empty EOF member."""
    import struct, zlib
    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(payload) + 8
        out += b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += payload + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


def test_gam_reader_inflates_bgzf_members_in_parallel_and_plain_gzip_serially():
    import gzip
    alns = [H.aln("ACGT" * (3 + k % 5), [(1 + k % 7, k % 3, bool(k & 1), [(4, 4, ""), (1, 1, "T"), (7 + k % 5, 7 + k % 5, "")])], mapq=k % 61)
            for k in range(700)]
    raw = H.encode_gam(alns, group=64)
    want = vb.Alignments.from_gam_bytes(raw).arrays()
    forms = {"bgzf": _bgzf(raw), "bgzf big blocks": _bgzf(raw, block=60000), "one gzip member": gzip.compress(raw),
             "two gzip members": gzip.compress(raw[:5000]) + gzip.compress(raw[5000:])}
    assert len(forms["bgzf"]) > 20 * 18                        # many members
    for name, blob in forms.items():
        got = vb.Alignments.from_gam_bytes(blob).arrays()
        assert set(got) == set(want), name
        for k in want:
            assert np.array_equal(got[k], want[k]), (name, k)
    # a flipped payload byte fails the member's CRC / size check instead of yielding other reads
    broken = bytearray(forms["bgzf"])
    broken[40] ^= 0x55
    with pytest.raises(Exception):
        vb.Alignments.from_gam_bytes(bytes(broken))


def test_raw_arrays_ship_one_offset_per_read_when_only_first_mappings_start_inside_a_node():
    S = vb.SynthHaplocart(seed=21, n_paths=33, n_rows=310, n_nodes=301)
    A = S.reads(300, seed=29, sub_rate=0.02, indel_rate=0.01)
    raw = vb.RawBatch(A)
    r = raw.raw
    assert r.first_offset16 and r.map_offset16 and r.map_n_edits8 and r.edit_slen8
    n, nm = int(r.n_reads), int(r.n_maps)
    moff = vb.host._as_array(r.map_off, n + 1, np.uint32)
    off = vb.host._as_array(r.map_offset, nm, np.uint32)
    first = vb.host._as_array(r.first_offset16, n, np.uint16)
    rebuilt = np.zeros(nm, dtype=np.uint32)                    # what the device does: zero, then one store per read
    rebuilt[moff[:-1][moff[1:] > moff[:-1]]] = first[moff[1:] > moff[:-1]]
    assert np.array_equal(rebuilt, off) and off[moff[:-1]].any()
    with_first = raw.nbytes
    r.first_offset16 = None
    assert raw.nbytes - with_first == 2 * nm - 2 * n
    # a mapping that starts inside its node further down the read: one offset per mapping again
    odd = [H.aln("ACGTAC", [(1, 2, False, [(3, 3, "")]), (2, 1, False, [(3, 3, "")])])]
    r2 = vb.RawBatch(vb.Alignments.from_gam_bytes(H.encode_gam(odd))).raw
    assert not r2.first_offset16 and r2.map_offset16


def test_feeder_chunk_holds_the_same_arrays_as_the_general_builder():
    """vgh_hc_feeder_marshal (the narrowing copies, per-thread ranges, runs between skipped reads) against vgh_hc_raw_build on
    the same reads, for several thread counts and a chunk that holds reads the hot loop skips (identity < 1e-10)."""
    S = vb.SynthHaplocart(seed=21, n_paths=33, n_rows=310, n_nodes=301)
    A = S.reads(3000, seed=31, sub_rate=0.02, indel_rate=0.01, fixed_len=0, damage=True)
    assert (A.arrays()["identity"] < 1e-10).any()
    as_arr = vb.host._as_array
    for r0, r1 in [(0, 3000), (17, 1301), (5, 6)]:
        general = vb.RawBatch(A, r0, r1)          # keeps the arrays `want` points into alive
        want = general.raw
        n, nm, ne = int(want.n_reads), int(want.n_maps), int(want.n_edits)
        w_eoff = as_arr(want.edit_off, nm + 1, np.uint32)
        w_soff = as_arr(want.edit_seq_off, ne + 1, np.uint32)
        for threads in (1, 2, 5):
            f = vb.Feeder(n_threads=threads, chunk_reads=4096)
            got = f.marshal(A, r0, r1)
            assert got is not None and (int(got.n_reads), int(got.n_maps), int(got.n_edits)) == (n, nm, ne)
            for name, count, dt in [("qual_off", n + 1, np.uint32), ("map_off", n + 1, np.uint32), ("str_off", n + 1, np.uint32),
                                    ("mapq", n, np.uint8), ("map_node16", nm, np.uint16), ("first_offset16", n, np.uint16),
                                    ("map_n_edits8", nm, np.uint8), ("edit_slen8", ne, np.uint8), ("edit_from8", ne, np.uint8),
                                    ("edit_to8", ne, np.uint8)]:
                assert np.array_equal(as_arr(getattr(got, name), count, dt), as_arr(getattr(want, name), count, dt)), (name, threads)
            nq = int(as_arr(want.qual_off, n + 1, np.uint32)[-1])
            assert np.array_equal(as_arr(got.qual, nq, np.uint8), as_arr(want.qual, nq, np.uint8))
            assert np.array_equal(as_arr(got.edit_seq, int(w_soff[-1]), np.uint8), as_arr(want.edit_seq, int(w_soff[-1]), np.uint8))
            # the wide offset arrays are not built: their end points stand in for them
            assert list(as_arr(got.edit_off, 2, np.uint32)) == [0, int(w_eoff[-1])]
            assert list(as_arr(got.edit_seq_off, 2, np.uint32)) == [0, int(w_soff[-1])]
            assert got.meta.valid == 1 and got.meta.qual_max == want.meta.qual_max
            assert list(got.meta.mapq_mask) == list(want.meta.mapq_mask)
            f.close()


def test_batch_format_limits_are_enforced_at_the_limit_itself():
    """include/vgan_gpu.h: <= 4095 aligned columns per read (VGL_HC_MAP_MAX_POS), node ids <= 131070 (VGL_HC_MAX_NODE, the
    17-bit node field; 131071 is VGL_HC_NO_NODE).  The value at the limit packs, the next one is refused with a status."""
    big = "ACGT" * 1024                                     # 4096 bases
    g = vb.Graph.from_sequences({7: big, 131070: "ACGTACGT", 131071: "ACGTACGT"})

    def pack(a):
        return vb.PackedBatch(g, vb.Alignments.from_gam_bytes(H.encode_gam([a])))

    at = pack(H.aln(big[:4095], [(7, 0, False, [(4095, 4095, "")])], quality=bytes([30] * 4095)))
    assert at.n_reads == 1 and int(at.lanes()[0]["n_cols"]) == 4095
    with pytest.raises(_lib.VglError, match="does not fit the batch records"):
        pack(H.aln(big, [(7, 0, False, [(4096, 4096, "")])], quality=bytes([30] * 4096)))
    ok = pack(H.aln("ACGTACGT", [(131070, 0, False, [(8, 8, "")])]))
    cols = ok.lanes()[0]["col"]
    assert len(cols) == 8 and all((int(c) >> _lib.COL_NODE_SHIFT) == 131070 for c in cols)
    with pytest.raises(_lib.VglError, match="does not fit the column record"):
        pack(H.aln("ACGTACGT", [(131071, 0, False, [(8, 8, "")])]))


def test_feeder_sibling_form_and_32_bit_node_ids():
    """The euka / soibean form of the feeder's chunk (identity != 0 kept, |sequence| and the read's position per read) against
    vgh_euka_raw_build; and node ids above 32767 (euka's graphs), which make the feeder switch to 32-bit node ids once."""
    S = vb.SynthHaplocart(seed=21, n_paths=33, n_rows=310, n_nodes=301)
    A = S.reads(2000, seed=37, sub_rate=0.02, indel_rate=0.01, fixed_len=0, damage=True)
    as_arr = vb.host._as_array
    f = vb.Feeder(n_threads=3, chunk_reads=4096)
    general = vb.RawBatch(A, 3, 1900, euka_rule=True)
    want = general.raw
    got = f.marshal(A, 3, 1900, sibling=True)
    n, nm, ne = int(want.n_reads), int(want.n_maps), int(want.n_edits)
    assert got is not None and (int(got.n_reads), int(got.n_maps), int(got.n_edits)) == (n, nm, ne)
    for name, count, dt in [("seq_len", n, np.uint16), ("read_index", n, np.uint32), ("map_node16", nm, np.uint16), ("mapq", n, np.uint8),
                            ("first_offset16", n, np.uint16), ("map_n_edits8", nm, np.uint8), ("edit_slen8", ne, np.uint8),
                            ("map_off", n + 1, np.uint32), ("str_off", n + 1, np.uint32)]:
        assert np.array_equal(as_arr(getattr(got, name), count, dt), as_arr(getattr(want, name), count, dt)), name
    # 32-bit node ids
    big = {40000: "ACGTACGTACGTACGTACGT", 40001: "TTGCATTGCATTGCATTGCA"}
    reads = [H.aln(big[40000], [(40000, 0, False, [(20, 20, "")])]),
             H.aln(big[40001][2:], [(40001, 2, True, [(18, 18, "")])]),
             H.aln(big[40000] + big[40001], [(40000, 0, False, [(20, 20, "")]), (40001, 0, False, [(20, 20, "")])])]
    B = vb.Alignments.from_gam_bytes(H.encode_gam(reads))
    wide_general = vb.RawBatch(B, euka_rule=True)
    assert not wide_general.raw.map_node16
    g2 = f.marshal(B, 0, 3, sibling=True)
    assert g2 is not None and not g2.map_node16 and g2.map_node
    assert np.array_equal(as_arr(g2.map_node, 4, np.uint32), as_arr(wide_general.raw.map_node, 4, np.uint32))
    assert list(as_arr(g2.map_node, 4, np.uint32)) == [40000, 40001 | 0x80000000, 40000, 40001]
    # the switch is sticky: the next chunk with small ids still travels with 32-bit ids, same values
    g3 = f.marshal(A, 3, 1900, sibling=True)
    assert g3 is not None and not g3.map_node16
    w16 = as_arr(want.map_node16, nm, np.uint16).astype(np.uint32)
    assert np.array_equal(as_arr(g3.map_node, nm, np.uint32), (w16 & 0x7fff) | ((w16 >> 15) << 31))
    f.close()


def test_crafted_bgzf_sizes_do_not_take_the_process_down():
    """A BGZF-looking stream whose member trailers claim ~4 GiB each: rejected with an error code, no huge
    allocation, no exception across the C ABI."""
    import struct
    member = (b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 27) +
              b"\x03\x00" + struct.pack("<II", 0, 0xFFFFFFF0))
    assert len(member) == 28
    with pytest.raises(RuntimeError):
        vb.Alignments.from_gam_bytes(member * 64)
