"""CPU tests of the oracle itself: pinned on the reference's own known answers where they exist."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.gam_py import read_gam

import helpers as H


def test_reconstruction_known_answers_from_reference_suite():
    # the reference's `reconstruction` suite, src/test.cpp:855-994: 10 exact (graph_seq, read_seq) pairs
    kat = json.load(open(os.path.join(H.GOLDEN, "reconstruct_kat.json")))["cases"]
    nodes, _ = H.toy_graph()
    g = orc.Graph(nodes)
    alns = read_gam(os.path.join(H.GOLDEN, "test_reads.gam"))
    assert len(alns) == 10
    R = orc.Reads(alns)
    for case in kat:
        gs, rs, sizes = orc.reconstruct(g, R, case["read_index"])
        assert gs == case["graph_seq"], case["case"]
        assert rs == case["read_seq"], case["case"]
        n_edits = sum(len(m["edits"]) for m in alns[case["read_index"]]["mappings"])
        assert len(sizes) == n_edits  # one size per EDIT (src/vgan_utils.h:72)
        assert sum(sizes) == len(gs)


def test_gam_reader_matches_committed_decoding():
    for name in ["test_reads.gam", "J2a1a1a1.gam", "all_the_same.gam", "all_the_same_reverse.gam", "two_unique.gam"]:
        got = read_gam(os.path.join(H.GOLDEN, name))
        for a in got:
            a.pop("name")
        assert got == H.decoded_fixture(name)


def test_fixture_statistics_match_survey_appendix_b():
    a = read_gam(os.path.join(H.GOLDEN, "J2a1a1a1.gam"))
    assert len(a) == 81
    nm = [len(x["mappings"]) for x in a]
    assert min(nm) == 17 and max(nm) == 67 and abs(np.mean(nm) - 43.9) < 0.05
    assert all(len(x["sequence"]) == 100 and x["identity"] != 0 for x in a)


def test_phred_tables():
    qs = orc.qscore_vec()
    assert qs[0] == 0.25 and qs[1] == 0.25 and qs[2] == 0.25  # quirk Q7: HaploCart's table has qs[2] = 0.25
    assert qs[3] == pytest.approx(10 ** -0.3, rel=1e-15)
    assert qs[40] == pytest.approx(1e-4, rel=1e-14)
    pw = orc.incorrect_mapping_vec()
    assert pw[0] == 1.0 and pw[60] == pytest.approx(1e-6, rel=1e-14)


def test_region_rates():
    # quirk Q6: the protein-coding rate is 0 because of the integer divisions -> match probability exactly 1
    assert orc.region_match_prob(3307) == 1.0 and orc.region_match_prob(15886) == 1.0
    assert orc.region_match_prob(100) == pytest.approx((1 - 30 * 1.64273e-7) ** 8, rel=1e-15)   # HVS I
    assert orc.region_match_prob(1) == pytest.approx((1 - 30 * 2.29640e-8) ** 8, rel=1e-15)     # HVS II
    assert orc.region_match_prob(16500) == pytest.approx((1 - 30 * 1.54555e-8) ** 8, rel=1e-15)
    assert orc.region_match_prob(600) == pytest.approx((1 - 30 * 6.91285e-9) ** 8, rel=1e-15)   # tRNA
    assert orc.region_match_prob(2000) == pytest.approx((1 - 30 * 6.91285e-9) ** 8, rel=1e-15)  # rRNA
    assert orc.region_match_prob(0) == pytest.approx((1 - 30 * 2.48537e-8) ** 8, rel=1e-15)     # elsewhere
    assert orc.region_match_prob(4401) == pytest.approx((1 - 30 * 2.48537e-8) ** 8, rel=1e-15)  # gap between tRNAs


def _toy_tables(n_paths=5, bep=1e-4, is_fasta=False):
    nodes, paths = H.toy_graph()
    n_rows = max(nodes) + 1
    sup = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            sup[nid, p] = 1
    ids = sorted(nodes)
    coord = np.cumsum([0] + [len(nodes[i]) for i in ids])[:-1]
    T = orc.HcTables(sup, np.array(ids), (coord + 1).astype(np.int32), np.full(int(coord[-1]) + 40, 1.0), bep=bep,
                     is_fasta=is_fasta, min_id=min(ids))
    return T, orc.Graph(nodes), nodes, paths


def test_literal_and_hoisted_modes_agree_bit_for_bit():
    T, g, _, _ = _toy_tables()
    R = orc.Reads(read_gam(os.path.join(H.GOLDEN, "test_reads.gam")))
    for r in range(10):
        a, _, _ = orc.hc_read(T, g, R, r, mode=orc.MODE_LITERAL)
        b, A, B = orc.hc_read(T, g, R, r, mode=orc.MODE_HOISTED)
        assert np.array_equal(a, b)
        assert len(A) == len(B) == int(R.map_off[r + 1] - R.map_off[r])
    fa, ia, na = orc.hc_run(T, g, R, mode=orc.MODE_LITERAL)
    fb, ib, nb = orc.hc_run(T, g, R, mode=orc.MODE_HOISTED)
    assert np.array_equal(fa, fb) and ia == ib and na == nb == 10


def test_hand_computed_perfect_match_read():
    """Read 0 of the toy fixture: two mappings (23 + 12 columns), perfect match, Q=40, mapq 60.
    Recomputed here with plain Python floats following SURVEY.md 8a's per-mapping formulas."""
    import math
    T, g, nodes, paths = _toy_tables()
    alns = read_gam(os.path.join(H.GOLDEN, "test_reads.gam"))
    R = orc.Reads(alns)
    ll, A, B = orc.hc_read(T, g, R, 0)
    read = alns[0]["sequence"].decode()
    G, ps, sizes = orc.reconstruct(g, R, 0)
    assert sizes == [23, 12] and G == ps == read
    bg = {"A": 0.27532, "C": 0.30044, "G": 0.16644, "T": 0.25780}
    qs40 = 10 ** (-40 * 0.1)
    pcm = (1 - 10 ** (-60 * 0.1)) * 1.0
    ids = sorted(nodes)
    coord = dict(zip(ids, np.cumsum([0] + [len(nodes[i]) for i in ids])[:-1] + 1))
    expect_A, expect_B = [], []
    pos = 0
    for i, m in enumerate(alns[0]["mappings"]):
        K = orc.region_match_prob(int(coord[m["node_id"]]))
        a = 0.0
        for j in range(sizes[i]):
            gb, rb = G[pos + j], ps[j]  # quirk Q2: the read is indexed from 0 for every mapping
            eps = qs40 if gb == rb else 1 - qs40
            a += math.log((1 - pcm) * bg[rb] + pcm * K * (1 - eps))
        expect_A.append(a)
        # quirk Q4: |ps| qualities starting at pos; bytes past the end count as 0 -> log(0.25)
        n_in = len(read) - pos
        expect_B.append(n_in * math.log(qs40) + pos * math.log(0.25))
        pos += sizes[i]
    assert np.allclose(np.array(A, dtype=float), expect_A, rtol=1e-13)
    assert np.allclose(B, expect_B, rtol=1e-13)
    for p, (_, steps) in enumerate(paths):
        on = {nid for nid, _ in steps}
        want = sum(expect_A[i] if m["node_id"] in on else expect_B[i] for i, m in enumerate(alns[0]["mappings"]))
        assert float(ll[p]) == pytest.approx(want, rel=1e-13)


def test_identity_filter_and_error_paths():
    T, g, nodes, _ = _toy_tables()
    alns = read_gam(os.path.join(H.GOLDEN, "test_reads.gam"))
    alns[3]["identity"] = 0.0
    R = orc.Reads(alns)
    _, _, used = orc.hc_run(T, g, R)
    assert used == 9  # identity < 1e-10 is skipped (src/HaploCart.cpp:411)
    bad = [H.aln("ACGT", [(9999, 0, False, [(4, 4, "")])])]
    with pytest.raises(RuntimeError):
        orc.hc_run(T, g, orc.Reads(bad))


def test_background_error_latch_and_fasta_mode():
    T, g, _, _ = _toy_tables(bep=1e-3)  # the fixture's Q=40 equals the default bep of 1e-4
    alns = read_gam(os.path.join(H.GOLDEN, "test_reads.gam"))
    base, _, _ = orc.hc_read(T, g, orc.Reads(alns), 1)
    q = bytearray(alns[1]["quality"])
    q[5] = 95  # any real byte >= 90 switches the whole read to the background error probability
    alns[1]["quality"] = bytes(q)
    latched, _, _ = orc.hc_read(T, g, orc.Reads(alns), 1)
    assert not np.array_equal(base, latched)
    Tf, gf, _, _ = _toy_tables(bep=1e-3, is_fasta=True)
    fasta, _, _ = orc.hc_read(Tf, gf, orc.Reads(alns), 1)
    assert not np.array_equal(fasta, latched)


def test_posterior_small_tree():
    names = ["A", "A1", "A2", "A1a", "B"]
    parents = "A root\nA1 A root\nA2 A A root\nA1a A1 A root [x]\nB root\n"
    children = "root A B\nA A1 A2\nA1 A1a\n"
    ll = np.array([-10.0, -9.0, -12.0, -9.5, -11.0], dtype=np.longdouble)
    idx, conf, total = orc.hc_posterior(ll, names, parents, children, 1)
    e = np.exp(ll.astype(float))
    assert float(total) == pytest.approx(np.log(e.sum()), rel=1e-14)
    assert idx == [-1, 0, 1]
    assert conf[0] == pytest.approx(e[1] / e.sum(), rel=1e-13)
    assert conf[1] == pytest.approx((e[1] + e[2] + e[3]) / e.sum(), rel=1e-13)  # strict descendants of A
    assert conf[2] == pytest.approx(1.0, rel=1e-13)                             # root covers every path
    # consecutive duplicate parents are reported once (A2's list is "A A root")
    idx2, conf2, _ = orc.hc_posterior(ll, names, parents, children, 2)
    assert idx2 == [-1, 0, 2] and len(conf2) == 3
