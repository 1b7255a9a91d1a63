"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances.  BASELINE.json asks for per-path log-likelihoods and posteriors within 1e-6 relative
(fp64 accumulate).  The device computes in fp64 throughout, so the tests hold it to 1e-11 relative on
log-likelihoods (observed ~1e-15) and 1e-6 relative on posteriors; calls (argmax) must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

import vgan_b200 as vb
from vgan_b200 import _lib
from oracle import oracle as orc

import helpers as H

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-11
POST_RTOL = 1e-6


@pytest.fixture(scope="module")
def small():
    S = vb.SynthHaplocart(seed=101, n_paths=300, n_rows=1210, n_nodes=1201)
    return S


def _oracle(S, alns, bep=1e-4, is_fasta=False, node_seqs=None, mode=orc.MODE_HOISTED):
    T, G, R = H.oracle_side(S, alns.arrays(), node_seqs=node_seqs, bep=bep, is_fasta=is_fasta)
    return orc.hc_run(T, G, R, mode=mode)


def _check(res, final, argmax, S=None):
    ref = np.asarray(final, dtype=np.float64)
    assert res["maxh_index"] == argmax
    assert H.rel_err(res["final_vec"], ref) < LL_RTOL
    if S is not None:
        idx, conf, total = orc.hc_posterior(final, S.path_names, S.parents_txt, S.children_txt, argmax)
        assert res["total_ll"] == pytest.approx(float(total), rel=1e-12)
        assert len(res["confidence_vec"]) == len(conf)
        assert np.allclose(res["confidence_vec"], conf, rtol=POST_RTOL, atol=1e-300)
        pv = [S.path_names[argmax]] + [x for j, x in enumerate(_parents(S, argmax)) if j in idx[1:]]
        assert res["clade_vec"] == pv


def _parents(S, p):
    for line in S.parents_txt.split("\n"):
        t = line.split()
        if t and t[0] == S.path_names[p]:
            return [x for x in t[1:] if "[" not in x]
    return []


@pytest.mark.parametrize("kw", [
    dict(n=400, sub_rate=0.005, indel_rate=0.0005),
    dict(n=400, sub_rate=0.03, indel_rate=0.02),
    dict(n=300, sub_rate=0.01, indel_rate=0.002, damage=True, fixed_len=0),
    dict(n=5, sub_rate=0.0, indel_rate=0.0, origin_path=123),
])
def test_synthetic_samples_match_oracle(small, kw):
    kw = dict(kw)
    n = kw.pop("n")
    alns = small.reads(n, seed=202, **kw)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    final, argmax, used = _oracle(small, alns)
    _check(res, final, argmax, small)
    hc.close()


@pytest.mark.parametrize("mode", ["0", "1", "2"])
@pytest.mark.parametrize("fasta", [False, True])
def test_large_class_tables_take_the_other_kernel_layouts(small, monkeypatch, mode, fasta):
    """VGL_HC_TABLE_MODE pins the GENERAL read kernel (products of per-column factors, one log per mapping) in one of its three
    table layouts: per-(mapping quality, class) constants in shared memory (0), the class table itself in shared (1) or
    global (2) memory -- what runs when a batch cannot take the log-term kernel; all must agree with the oracle."""
    monkeypatch.setenv("VGL_HC_TABLE_MODE", mode)
    alns = small.reads(300, seed=212, sub_rate=0.02, indel_rate=0.005)
    reads = alns
    if not fasta:  # add the hand-made corner cases as well
        extra = vb.Alignments.from_gam_bytes(H.encode_gam(_edge_reads(small)))
        hc = vb.Haplocart()
        hc.load(small.tables(bep=2e-3), small.tree())
        res = hc.run_sample([vb.PackedBatch(small.graph, extra), vb.PackedBatch(small.graph, alns)])
        f1, _, _ = _oracle(small, extra, bep=2e-3)
        f2, _, _ = _oracle(small, alns, bep=2e-3)
        final = np.asarray(f1, dtype=np.longdouble) + np.asarray(f2, dtype=np.longdouble)
        _check(res, final, int(np.argmax(final)), small)
        hc.close()
        return
    hc = vb.Haplocart()
    hc.load(small.tables(bep=1e-3, is_fasta=True), small.tree())
    res = hc.run_sample([vb.PackedBatch(small.graph, reads)])
    final, argmax, _ = _oracle(small, reads, bep=1e-3, is_fasta=True)
    _check(res, final, argmax, small)
    hc.close()


def test_literal_reference_shaped_oracle_agrees_too(small):
    alns = small.reads(40, seed=203)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    final, argmax, _ = _oracle(small, alns, mode=orc.MODE_LITERAL)
    _check(res, final, argmax, small)
    hc.close()


def test_posterior_is_not_degenerate_for_a_handful_of_reads(small):
    alns = small.reads(2, seed=204, origin_path=200, sub_rate=0.0, indel_rate=0.0)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    final, argmax, _ = _oracle(small, alns)
    _check(res, final, argmax, small)
    assert 0 < res["confidence_vec"][0] < 0.9          # several haplogroups share the evidence
    assert res["confidence_vec"][-1] == pytest.approx(1.0, rel=1e-9)  # mt-MRCA covers every path
    assert np.all(np.diff(res["confidence_vec"][1:]) >= -1e-12)       # clade support grows towards the root
    hc.close()


def test_consensus_fasta_mode_and_custom_background_error(small):
    alns = small.reads(150, seed=205, sub_rate=0.01)
    for bep, fasta in ((1e-4, True), (3e-3, True), (3e-3, False)):
        hc = vb.Haplocart()
        hc.load(small.tables(bep=bep, is_fasta=fasta), small.tree())
        res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
        final, argmax, _ = _oracle(small, alns, bep=bep, is_fasta=fasta)
        _check(res, final, argmax, small)
        hc.close()


def _edge_reads(S):
    """Hand-made alignments on the synthetic graph covering the reference's odd corners."""
    seqs = S.node_sequences()
    node, _ = S.pmap()
    valid = sorted(set(node.tolist()))
    a, b, c = valid[10], valid[11], valid[13]
    sa, sb, sc = seqs[a - 1], seqs[b - 1], seqs[c - 1]
    reads = []
    # background-error latch: one real quality byte >= 90 (src/update_likelihood.cpp:42)
    r = sa + sb
    q = bytearray([35] * len(r)); q[len(r) - 1] = 93
    reads.append(H.aln(r, [(a, 0, False, [(len(sa), len(sa), "")]), (b, 0, False, [(len(sb), len(sb), "")])], quality=q))
    # quality bytes that are negative as signed char, and qualities 0/1/2/3 around the 0.25 floor
    q = bytes([200, 0, 1, 2, 3, 89, 255] * 4)[:len(r)]
    reads.append(H.aln(r, [(a, 0, False, [(len(sa), len(sa), "")]), (b, 0, False, [(len(sb), len(sb), "")])], quality=q, mapq=3))
    # empty quality string (every byte is "past the end")
    x = H.aln(r, [(a, 0, False, [(len(sa), len(sa), "")]), (b, 0, False, [(len(sb), len(sb), "")])], mapq=1)
    x["quality"] = b""
    reads.append(x)
    # N in the read and lower-case / non-ACGT bytes are skipped by isValidDNA
    rn = "N" + r[1:-1] + "a"
    reads.append(H.aln(rn, [(a, 0, False, [(1, 1, "N"), (len(sa) - 1, len(sa) - 1, "")] if len(sa) > 1 else [(1, 1, "N")]),
                            (b, 0, False, [(len(sb) - 1, len(sb) - 1, ""), (1, 1, "a")] if len(sb) > 1 else [(1, 1, "a")])], mapq=0))
    # many more edits than mappings (quirk Q1) with an insertion at node offset 0 ('S') and a deletion
    eds_c = [(0, 2, "GG"), (1, 1, "T")] + ([(len(sc) - 1, 0, "")] if len(sc) > 1 else [])
    reads.append(H.aln(sa + "GG" + "T", [(a, 0, False, [(len(sa), len(sa), "")]), (c, 0, False, eds_c)], mapq=60))
    # mapping quality at both ends of the table
    reads.append(H.aln(sa, [(a, 0, False, [(len(sa), len(sa), "")])], mapq=99))
    reads.append(H.aln(H.revcomp(sa), [(a, 0, True, [(len(sa), len(sa), "")])], mapq=0))
    # a read visiting the same node twice
    reads.append(H.aln(sa + sa, [(a, 0, False, [(len(sa), len(sa), "")]), (a, 0, False, [(len(sa), len(sa), "")])]))
    # a single-column read
    reads.append(H.aln(sa[:1], [(a, 0, False, [(1, 1, "")])]))
    # mappings that own no column (their edit-size slot is 0, quirk Q1): in the middle, sharing the start of the next ...
    full = lambda n_, s_: (n_, 0, False, [(len(s_), len(s_), "")])
    reads.append(H.aln(sa + sb + sc, [full(a, sa), (b, 0, False, [(0, 0, ""), (len(sb), len(sb), "")]), full(c, sc)]))
    # ... trailing, where the columns no mapping owns begin ...
    reads.append(H.aln(sa + sb + sc, [full(a, sa), (b, 0, False, [(len(sb), len(sb), ""), (0, 0, "")]), full(c, sc)], mapq=20))
    # ... first of the read, and past the last column
    reads.append(H.aln(sa + sb, [(a, 0, False, [(0, 0, ""), (len(sa), len(sa), "")]), full(b, sb)], mapq=41))
    reads.append(H.aln(sa + sb, [full(a, sa), full(b, sb), (c, 0, False, [(0, 0, "")])], mapq=17))
    return reads


def test_edge_case_reads_one_by_one_and_together(small):
    reads = _edge_reads(small)
    hc = vb.Haplocart()
    hc.load(small.tables(bep=2e-3), small.tree())
    for subset in [[r] for r in reads] + [reads]:
        alns = vb.Alignments.from_gam_bytes(H.encode_gam(subset))
        res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
        final, argmax, used = _oracle(small, alns, bep=2e-3)
        assert used == len(subset)
        _check(res, final, argmax, small)
    hc.close()


def test_long_reads_many_mappings(small):
    """More than 32 mappings and more than 128 columns per read exercise the round / carry logic."""
    alns = small.reads(64, seed=206, fixed_len=700, sub_rate=0.01, indel_rate=0.003)
    assert np.diff(alns.arrays()["map_off"]).max() > 200  # several passes of the per-thread mapping lists
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    final, argmax, _ = _oracle(small, alns)
    _check(res, final, argmax, small)
    hc.close()


def test_empty_batch_and_unmapped_reads(small):
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    empty = vb.Alignments.from_gam_bytes(b"")
    res = hc.run_sample([vb.PackedBatch(small.graph, empty)])
    assert np.all(res["final_vec"] == 0.0) and res["maxh_index"] == 0       # max_element of all-zero picks index 0
    assert res["confidence_vec"][0] == pytest.approx(1.0 / small.n_paths, rel=1e-12)
    seqs = small.node_sequences()
    n = int(small.pmap()[0][0])
    alns = vb.Alignments.from_gam_bytes(H.encode_gam(
        [H.aln(seqs[n - 1], [(n, 0, False, [(len(seqs[n - 1]), len(seqs[n - 1]), "")])], identity=1e-12)]), drop_unmapped=False)
    pb = vb.PackedBatch(small.graph, alns)
    assert pb.n_reads == 0 and pb.skipped == 1                               # identity < 1e-10 (src/HaploCart.cpp:411)
    hc.close()


def test_node_without_pangenome_entry_is_an_error_not_a_guess():
    small = vb.SynthHaplocart(seed=105, n_paths=50, n_rows=312, n_nodes=303)  # 303 = 3*100 + 1 + two unused ids
    seqs = small.node_sequences()
    unused = small.n_nodes  # ids beyond the last bubble are in the graph but not in parsed_pangenome_mapping
    node, _ = small.pmap()
    if unused in set(node.tolist()):
        pytest.skip("no unmapped node in this synthetic graph")
    s = seqs[unused - 1]
    alns = vb.Alignments.from_gam_bytes(H.encode_gam([H.aln(s, [(unused, 0, False, [(len(s), len(s), "")])])]))
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    hc.reset()
    hc.update(vb.PackedBatch(small.graph, alns))
    with pytest.raises(_lib.VglError) as ei:
        hc.get_posterior()
    assert ei.value.code == 4  # VGL_ERR_BAD_NODE: the reference would terminate in map::at
    hc.close()


@pytest.mark.parametrize("name", ["J2a1a1a1.gam", "all_the_same.gam", "two_unique.gam"])
def test_reference_gam_fixtures_against_synthetic_tables(name):
    """SURVEY.md 8d config 1 substitute: the reference's real-graph alignments (node ids up to 11806)
    against synthetic tables of the real shape; node sequences are the ones the alignments imply."""
    S = vb.SynthHaplocart()  # 11825 x 5179
    alns_d = H.decoded_fixture(name)
    node_seqs = H.graph_from_alignments(alns_d, S.n_nodes)
    g = vb.Graph.from_sequences(node_seqs, min_id=1)
    alns = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, name))
    hc = vb.Haplocart()
    hc.load(S.tables(), S.tree())
    res = hc.run_sample([vb.PackedBatch(g, alns)])
    final, argmax, used = _oracle(S, alns, node_seqs=node_seqs)
    assert used == len(alns_d)
    _check(res, final, argmax, S)
    hc.close()


def test_toy_graph_fixture_end_to_end():
    """The reference's toy graph (5 paths) + its 10 test reads: tables straight from the GFA P lines."""
    nodes, paths = H.toy_graph()
    n_rows = max(nodes) + 1
    sup = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            sup[nid, p] = 1
    ids = sorted(nodes)
    coord = (np.cumsum([0] + [len(nodes[i]) for i in ids])[:-1] + 1).astype(np.int32)
    mapp = np.linspace(0.3, 1.0, int(coord[-1]) + 40)
    names = [n for n, _ in paths]
    parents = "".join("%s root\n" % n for n in names)
    children = "root " + " ".join(names) + "\n"
    for fixture in ("test_reads.gam", "all_the_same_reverse.gam"):
        alns = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, fixture))
        hc = vb.Haplocart()
        hc.load(vb.HcTables(sup, np.array(ids), coord, mapp), vb.HcTree(names, parents, children))
        res = hc.run_sample([vb.PackedBatch(vb.Graph.from_sequences(nodes), alns)])
        T = orc.HcTables(sup, np.array(ids), coord, mapp, min_id=min(ids))
        final, argmax, _ = orc.hc_run(T, orc.Graph(nodes), orc.FlatReads(alns.arrays()))
        assert res["maxh_index"] == argmax and H.rel_err(res["final_vec"], final.astype(np.float64)) < LL_RTOL
        idx, conf, total = orc.hc_posterior(final, names, parents, children, argmax)
        assert np.allclose(res["confidence_vec"], conf, rtol=POST_RTOL)
        hc.close()


def test_batches_add_up_and_order_does_not_matter(small):
    alns = small.reads(600, seed=207)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    whole = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    parts = hc.run_sample([vb.PackedBatch(small.graph, alns, r0=a, r1=b) for a, b in ((0, 1), (1, 250), (250, 600))])
    rev = hc.run_sample([vb.PackedBatch(small.graph, alns, r0=a, r1=b) for a, b in ((250, 600), (0, 1), (1, 250))])
    for other in (parts, rev):
        assert other["maxh_index"] == whole["maxh_index"]
        assert H.rel_err(other["final_vec"], whole["final_vec"]) < 1e-13
    # a device-resident batch gives the same numbers as the host-buffer entry point
    dev = hc.upload(vb.PackedBatch(small.graph, alns))
    hc.reset()
    hc.update_device(dev.batch)
    again = hc.get_posterior()
    assert H.rel_err(again["final_vec"], whole["final_vec"]) < 1e-13
    hc.close()


def test_identical_columns_tie_bit_for_bit_and_first_index_wins():
    """Paths with identical support must get bit-identical sums so that max_element ties resolve to the
    lowest index exactly as in the reference (src/HaploCart.cpp:423)."""
    S = vb.SynthHaplocart(seed=303, n_paths=200, n_rows=610, n_nodes=601)
    sup = S.supports()
    alns = S.reads(500, seed=304, origin_path=77)
    node, val = S.pmap()
    T0 = orc.HcTables(sup, node, val, S.mappab(), min_id=1)
    final, argmax, _ = orc.hc_run(T0, orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()))
    sup2 = sup.copy()
    twin = (argmax + 57) % S.n_paths
    sup2[:, twin] = sup2[:, argmax]           # an exact twin of the best path
    lo, hi = min(twin, argmax), max(twin, argmax)
    hc = vb.Haplocart()
    hc.load(vb.HcTables(sup2, node, val, S.mappab()), None)
    res = hc.run_sample([vb.PackedBatch(S.graph, alns)])
    assert res["final_vec"][lo] == res["final_vec"][hi]
    assert res["maxh_index"] == lo
    T2 = orc.HcTables(sup2, node, val, S.mappab(), min_id=1)
    final2, argmax2, _ = orc.hc_run(T2, orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()))
    assert argmax2 == lo
    hc.close()


def test_config2_shape_properties():
    """BASELINE config 2 shapes (11825 x 5179, 100 bp reads) at a size the GPU does in milliseconds; the
    oracle covers a subsample, size-independent properties cover the rest."""
    S = vb.SynthHaplocart()
    alns = S.reads(200000, seed=20261)
    hc = vb.Haplocart()
    hc.load(S.tables(), S.tree())
    pb = vb.PackedBatch(S.graph, alns)
    res = hc.run_sample([pb])
    ll = res["final_vec"]
    assert np.all(np.isfinite(ll)) and np.all(ll < 0)
    # linearity: two halves sum to the whole
    h1 = hc.run_sample([vb.PackedBatch(S.graph, alns, r0=0, r1=100000)])["final_vec"]
    h2 = hc.run_sample([vb.PackedBatch(S.graph, alns, r0=100000, r1=200000)])["final_vec"]
    assert H.rel_err(h1 + h2, ll) < 1e-12
    # posterior of the whole tree is 1 and the softmax sums to 1
    assert res["confidence_vec"][-1] == pytest.approx(1.0, rel=1e-9)
    # (|ll| ~ 5e7 in fp64 leaves ~1e-8 of absolute resolution on the differences the host recomputes here)
    assert np.exp(ll - res["total_ll"]).sum() == pytest.approx(1.0, rel=1e-6)
    # run-to-run: shared-memory atomics may reorder fp64 adds, but only in the last bits
    again = hc.run_sample([pb])
    assert again["maxh_index"] == res["maxh_index"] and H.rel_err(again["final_vec"], ll) < 1e-13
    # oracle on a subsample of the same reads
    sub = 300
    r = hc.run_sample([vb.PackedBatch(S.graph, alns, r0=0, r1=sub)])
    T, G, R = H.oracle_side(S, alns.arrays())
    final, argmax, _ = orc.hc_run(T, G, R, r0=0, r1=sub, n_threads=0)
    assert r["maxh_index"] == argmax and H.rel_err(r["final_vec"], final.astype(np.float64)) < LL_RTOL
    hc.close()


def test_parts_summed_in_rank_order_equal_the_whole(small):
    """The multi-GPU exchange on one device: two 'ranks' each reduce their half of the reads to a final_vec in
    device memory; vgl_hc_set_paths_from_parts adds them in order and the posterior kernel runs on the sum."""
    from vgan_b200._lib import gpu, check
    alns = small.reads(500, seed=208)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    whole = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    P = small.n_paths
    parts = C.c_void_p()
    check(gpu.vgl_device_alloc(hc.ctx, C.byref(parts), 2 * P * 8))
    for k, (a, b) in enumerate(((0, 250), (250, 500))):
        hc.reset()
        hc.update(vb.PackedBatch(small.graph, alns, r0=a, r1=b))
        hc.paths(parts.value + k * P * 8)
    hc.reset()
    hc.set_paths_from_parts(parts.value, 2)
    both = hc.get_posterior()
    assert both["maxh_index"] == whole["maxh_index"]
    assert H.rel_err(both["final_vec"], whole["final_vec"]) < 1e-13
    assert np.allclose(both["confidence_vec"], whole["confidence_vec"], rtol=1e-9)
    check(gpu.vgl_device_free(hc.ctx, parts))
    hc.close()


# ---- the log-term read kernel (per-column log terms from shared-memory tables; the default whenever the batch allows it) ----
def _capped(reads, cap=63):
    """The corner-case reads with every quality byte brought into 0..cap (so that nothing latches and the batch qualifies)."""
    out = []
    for r in reads:
        r = dict(r)
        r["quality"] = bytes((b % (cap + 1)) for b in r["quality"])
        out.append(r)
    return out


def test_logterm_kernel_is_taken_and_agrees_with_the_general_kernel_and_the_oracle(small, monkeypatch):
    alns = small.reads(900, seed=221, sub_rate=0.01, indel_rate=0.003, fixed_len=0, damage=True)
    for bep, fasta in ((1e-4, False), (2e-3, True)):
        hc = vb.Haplocart()
        hc.load(small.tables(bep=bep, is_fasta=fasta), small.tree())
        n0 = hc.logterm_launches()
        fast = hc.run_sample([vb.PackedBatch(small.graph, alns)])
        assert hc.logterm_launches() == n0 + 1                      # qualities <= 41, a handful of mapping qualities: it fits
        final, argmax, _ = _oracle(small, alns, bep=bep, is_fasta=fasta)
        _check(fast, final, argmax, small)
        hc.close()
        monkeypatch.setenv("VGL_HC_TABLE_MODE", "0")
        hg = vb.Haplocart()
        hg.load(small.tables(bep=bep, is_fasta=fasta), small.tree())
        general = hg.run_sample([vb.PackedBatch(small.graph, alns)])
        assert hg.logterm_launches() == 0
        monkeypatch.delenv("VGL_HC_TABLE_MODE")
        assert general["maxh_index"] == fast["maxh_index"] and H.rel_err(fast["final_vec"], general["final_vec"]) < 1e-12
        hg.close()


def test_logterm_kernel_on_the_corner_reads(small):
    reads = _capped(_edge_reads(small))
    hc = vb.Haplocart()
    hc.load(small.tables(bep=2e-3), small.tree())
    for subset in [[r] for r in reads] + [reads]:
        alns = vb.Alignments.from_gam_bytes(H.encode_gam(subset))
        n0 = hc.logterm_launches()
        res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
        assert hc.logterm_launches() == n0 + 1
        final, argmax, used = _oracle(small, alns, bep=2e-3)
        assert used == len(subset)
        _check(res, final, argmax, small)
    # one quality byte of 64 in the batch: the general kernel takes over, same numbers
    hi = [dict(r) for r in reads]
    q = bytearray(hi[0]["quality"]); q[0] = 64
    hi[0]["quality"] = bytes(q)
    alns = vb.Alignments.from_gam_bytes(H.encode_gam(hi))
    n0 = hc.logterm_launches()
    res = hc.run_sample([vb.PackedBatch(small.graph, alns)])
    assert hc.logterm_launches() == n0
    final, argmax, _ = _oracle(small, alns, bep=2e-3)
    _check(res, final, argmax, small)
    hc.close()


def test_logterm_kernel_long_reads_and_every_route(small):
    alns = small.reads(300, seed=222, fixed_len=700, sub_rate=0.01, indel_rate=0.003)
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    hc.load_graph(small.graph)
    final, argmax, _ = _oracle(small, alns)
    packed = vb.PackedBatch(small.graph, alns)
    resident = hc.upload(packed)     # (stays alive until the kernels that read it have run)
    routes = {"packed": lambda: hc.update(packed), "wire": lambda: hc.update_wire(vb.WireBatch(packed)),
              "raw": lambda: hc.update_raw(vb.RawBatch(alns)), "device": lambda: hc.update_device(resident.batch),
              "feeder": lambda: vb.Feeder(n_threads=2, chunk_reads=128).feed(hc, alns)}
    for name, run in routes.items():
        n0 = hc.logterm_launches()
        hc.reset()
        run()
        res = hc.get_posterior()
        assert hc.logterm_launches() > n0, name                     # every producer hands the value ranges over
        _check(res, final, argmax, small)
    hc.close()


# ---- several samples per pass (vgl_hc_samples_begin / _sample_select / _finalize_samples) ------------------------------------
def test_samples_pass_equals_separate_runs_bit_for_bit(small):
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    hc.load_graph(small.graph)
    samples = [small.reads(n, seed=400 + k, sub_rate=0.01, indel_rate=0.002, origin_path=(17 * k) % small.n_paths)
               for k, n in enumerate([300, 1, 64, 700, 33, 2, 150, 90, 410])]
    packed = [vb.PackedBatch(small.graph, a) for a in samples]
    alone = [hc.run_sample([pb]) for pb in packed]
    assert len(set(r["maxh_index"] for r in alone)) > 3                      # the samples really differ
    feeder = vb.Feeder(n_threads=2, chunk_reads=256)
    hc.samples_begin(len(samples))
    for k, (a, pb) in enumerate(zip(samples, packed)):
        hc.sample_select(k)
        if k % 3 == 0:
            hc.update(pb)                                                    # packed host batch
        elif k % 3 == 1:
            feeder.feed(hc, a)                                               # flat alignment fields, device-side packing
        else:
            hc.update(vb.PackedBatch(small.graph, a, r0=0, r1=len(a) // 2))  # two calls into one sample
            hc.update(vb.PackedBatch(small.graph, a, r0=len(a) // 2, r1=len(a)))
    got = hc.finalize_samples(want_ll=True)
    for k, (g, w) in enumerate(zip(got, alone)):
        assert g["status"] == 0 and g["maxh_index"] == w["maxh_index"], k
        assert H.rel_err(g["final_vec"], w["final_vec"]) < 1e-13, k
        assert g["clade_vec"] == w["clade_vec"] and np.allclose(g["confidence_vec"], w["confidence_vec"], rtol=1e-12, atol=1e-300)
        assert g["total_ll"] == pytest.approx(w["total_ll"], rel=1e-13)
    final, argmax, _ = _oracle(small, samples[3])
    _check(got[3], final, argmax, small)
    # the context is back in single-sample mode
    again = hc.run_sample([packed[0]])
    assert again["maxh_index"] == alone[0]["maxh_index"] and H.rel_err(again["final_vec"], alone[0]["final_vec"]) < 1e-13
    hc.close()


def test_samples_pass_reports_a_bad_sample_without_losing_the_others():
    S = vb.SynthHaplocart(seed=105, n_paths=50, n_rows=312, n_nodes=303)
    seqs = S.node_sequences()
    unused = S.n_nodes
    node, _ = S.pmap()
    if unused in set(node.tolist()):
        pytest.skip("no unmapped node in this synthetic graph")
    s = seqs[unused - 1]
    bad = vb.Alignments.from_gam_bytes(H.encode_gam([H.aln(s, [(unused, 0, False, [(len(s), len(s), "")])])]))
    good = S.reads(50, seed=3)
    hc = vb.Haplocart()
    hc.load(S.tables(), S.tree())
    want = hc.run_sample([vb.PackedBatch(S.graph, good)])
    hc.samples_begin(3)
    for k, a in enumerate([good, bad, good]):
        hc.sample_select(k)
        hc.update(vb.PackedBatch(S.graph, a))
    got = hc.finalize_samples(want_ll=True)
    assert [g["status"] for g in got] == [0, 4, 0]                            # VGL_ERR_BAD_NODE for the middle sample only
    for g in (got[0], got[2]):
        assert g["maxh_index"] == want["maxh_index"] and H.rel_err(g["final_vec"], want["final_vec"]) < 1e-13
    hc.close()


def test_node_table_beyond_the_shared_memory_limit_is_refused_with_a_status():
    """the node -> class table of the read kernel lives in shared memory (2 B per row next to 20 KB of small tables
    and the queue rings, 227 KB per CTA).  HaploCart's graph has 11 825 rows; 40 000 still load, 120 000 are refused by
    vgl_hc_set_tables with VGL_ERR_INVALID and a message that names the limit -- nothing is launched."""
    rng = np.random.default_rng(7)

    def tables(n_rows, n_paths=40):
        sup = (rng.random((n_rows, n_paths)) < 0.5).astype(np.uint8)
        node = np.arange(1, n_rows + 1, dtype=np.int64)
        val = rng.integers(1, 16000, n_rows).astype(np.int32)
        return vb.HcTables(sup, node, val, np.full(16569, 0.9))

    hc = vb.Haplocart()
    hc.load(tables(40000))
    with pytest.raises(_lib.VglError, match="shared memory"):
        hc.load(tables(120000))
    hc.close()


def test_posterior_begin_end_keeps_two_samples_in_flight(small):
    """vgl_hc_finalize_begin / _end: a sample's posterior is enqueued, the next sample is accumulated behind it on the same stream,
    and each result comes back from its own slot -- the same calls, vectors and posteriors as one get_posterior() per sample."""
    hc = vb.Haplocart()
    hc.load(small.tables(), small.tree())
    batches = [vb.PackedBatch(small.graph, small.reads(300 + 40 * k, seed=700 + k, sub_rate=0.01, indel_rate=0.002, origin_path=3 + 11 * k))
               for k in range(5)]
    want = [hc.run_sample([b]) for b in batches]
    got, pending = [], None
    for b in batches:
        hc.reset()
        hc.update(b)
        t = hc.posterior_begin()
        if pending is not None:
            got.append(hc.posterior_end(pending))
        pending = t
    got.append(hc.posterior_end(pending))
    assert len({w["maxh_index"] for w in want}) > 1
    for g, w in zip(got, want):
        assert g["maxh_index"] == w["maxh_index"] and g["clade_vec"] == w["clade_vec"]
        assert np.allclose(g["confidence_vec"], w["confidence_vec"], rtol=1e-9, atol=1e-300)
        assert abs(g["total_ll"] - w["total_ll"]) <= 1e-12 * abs(w["total_ll"])
    # a third result cannot be begun while two are in flight; a ticket is good once
    hc.reset(); hc.update(batches[0]); t0 = hc.posterior_begin()
    hc.reset(); hc.update(batches[1]); t1 = hc.posterior_begin()
    hc.reset(); hc.update(batches[2])
    with pytest.raises(_lib.VglError, match="in flight"):
        hc.posterior_begin()
    assert hc.posterior_end(t0)["maxh_index"] == want[0]["maxh_index"]
    assert hc.posterior_end(t1)["maxh_index"] == want[1]["maxh_index"]
    with pytest.raises(_lib.VglError, match="no result pending"):
        hc.posterior_end(t1)
    assert hc.get_posterior(want_ll=False)["maxh_index"] == want[2]["maxh_index"]
    hc.close()
