"""The oracle pinned on the REFERENCE'S OWN code: oracle/_ref/libvgan_ref.so is the reference's hot-path sources
(update_likelihood.cpp + process_mapping.cpp, get_p_obs_base.cpp, get_posterior.cpp, vgan_utils.h, miscfunc.h,
haplocart_functions.cpp, rmdup.cpp, damage.cpp) compiled unchanged from /root/reference/src against a header shim
(oracle/ref_shim/).  These CPU tests hold the oracle to it BIT FOR BIT in `long double`, and hold both to the
committed golden vectors the reference produced (tests/golden/ref_hc_golden.json, made by make_ref_golden.py).

Where the reference itself is undefined (quality bytes read past the end of the string, quirk Q4) the driver pads
the quality string with zero bytes so the run is defined and equals the semantics the oracle documents."""
import json
import os

import numpy as np
import pytest

import vgan_b200 as vb
from oracle import oracle as orc
from oracle import ref
from oracle.gam_py import read_gam

import helpers as H

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")


def _toy_tables(bep=1e-4, is_fasta=False, mapp=None):
    nodes, paths = H.toy_graph()
    n_rows = max(nodes) + 1
    sup = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            sup[nid, p] = 1
    ids = sorted(nodes)
    coord = np.cumsum([0] + [len(nodes[i]) for i in ids])[:-1]
    mp = np.linspace(0.3, 1.0, int(coord[-1]) + 40) if mapp is None else mapp
    T = orc.HcTables(sup, np.array(ids), (coord + 1).astype(np.int32), mp, bep=bep, is_fasta=is_fasta, min_id=min(ids))
    return T, orc.Graph(nodes)


@needs_ref
def test_reference_reconstruction_reproduces_its_own_known_answers():
    """The reference's reconstruct_graph_sequence, compiled here over the shim, against its own 10 KATs
    (src/test.cpp:855-994): this is what pins the shim's path_string / edit_is_* / get_sequence."""
    kat = json.load(open(os.path.join(H.GOLDEN, "reconstruct_kat.json")))["cases"]
    nodes, _ = H.toy_graph()
    g = orc.Graph(nodes)
    R = orc.Reads(read_gam(os.path.join(H.GOLDEN, "test_reads.gam")))
    for case in kat:
        gs, rs, sizes = ref.reconstruct(g, R, case["read_index"])
        assert gs == case["graph_seq"] and rs == case["read_seq"], case["case"]
        assert (gs, rs, sizes) == orc.reconstruct(g, R, case["read_index"])


@needs_ref
def test_tables_equal_the_reference_functions():
    assert np.array_equal(orc.qscore_vec(), ref.qscore_vec())
    assert np.array_equal(orc.incorrect_mapping_vec(), ref.incorrect_mapping_vec())
    for pb in range(0, 16700):
        assert orc.region_match_prob(pb) == ref.region_match_prob(pb), pb
    for q in range(-5, 130):
        assert orc.lib().orc_hc_p_seq_error(q) == ref.p_seq_error(q)
    for b in "ACGTNacgt-S":
        assert orc.lib().orc_hc_background_freq(orc.C.c_char(b.encode())) == ref.background_freq(b)


def _same(T, G, R, r0=0, r1=None):
    fo, io_, no = orc.hc_run(T, G, R, r0, r1, mode=orc.MODE_LITERAL)
    fh, ih, nh = orc.hc_run(T, G, R, r0, r1, mode=orc.MODE_HOISTED)
    fr, ir, nr = ref.hc_run(T, G, R, r0, r1)
    assert np.array_equal(fo, fr), float(np.max(np.abs(fo - fr)))   # long double, bit for bit
    assert np.array_equal(fh, fr)
    assert io_ == ir == ih and no == nr == nh
    return fr, ir


@needs_ref
@pytest.mark.parametrize("bep,fasta", [(1e-4, False), (1e-3, False), (1e-3, True)])
def test_toy_graph_reads_bit_for_bit(bep, fasta):
    T, G = _toy_tables(bep=bep, is_fasta=fasta)
    for name in ("test_reads.gam", "all_the_same_reverse.gam"):
        R = orc.Reads(read_gam(os.path.join(H.GOLDEN, name)))
        for r in range(R.n):
            a, _, _ = orc.hc_read(T, G, R, r, mode=orc.MODE_LITERAL)
            assert np.array_equal(a, ref.hc_read(T, G, R, r)), (name, r)
        _same(T, G, R)


@needs_ref
def test_quality_latch_on_the_reference():
    """A real quality byte >= 90 switches to the background error probability (src/update_likelihood.cpp:42)."""
    T, G = _toy_tables(bep=1e-3)
    alns = read_gam(os.path.join(H.GOLDEN, "test_reads.gam"))
    for pos in (0, 5, 30):
        q = bytearray(alns[1]["quality"]); q[pos] = 95
        alns[1]["quality"] = bytes(q)
        R = orc.Reads(alns)
        a, _, _ = orc.hc_read(T, G, R, 1, mode=orc.MODE_LITERAL)
        assert np.array_equal(a, ref.hc_read(T, G, R, 1)), pos


@pytest.fixture(scope="module")
def small():
    return vb.SynthHaplocart(seed=101, n_paths=300, n_rows=1210, n_nodes=1201)


@needs_ref
@pytest.mark.parametrize("kw", [
    dict(n=120, sub_rate=0.005, indel_rate=0.0005),
    dict(n=80, sub_rate=0.03, indel_rate=0.02),
    dict(n=80, sub_rate=0.01, indel_rate=0.002, damage=True, fixed_len=0),
    dict(n=16, sub_rate=0.01, indel_rate=0.003, fixed_len=400),
])
def test_synthetic_samples_bit_for_bit(small, kw):
    kw = dict(kw)
    n = kw.pop("n")
    alns = small.reads(n, seed=202, **kw)
    for bep, fasta in ((1e-4, False), (2e-3, True)):
        T, G, R = H.oracle_side(small, alns.arrays(), bep=bep, is_fasta=fasta)
        _same(T, G, R)


def _defined_edge_reads(S):
    """The hand-made corner reads of the GPU suite, minus the one with quality bytes that are negative as
    signed char (they index qscore_vec out of range in the reference, Q4b: nothing to pin there)."""
    from test_gpu_haplocart import _edge_reads
    reads = _edge_reads(S)
    return [r for r in reads if not any(b >= 128 for b in r["quality"])]


@needs_ref
def test_edge_case_reads_bit_for_bit(small):
    reads = _defined_edge_reads(small)
    assert len(reads) >= 11
    for subset in [[r] for r in reads] + [reads]:
        alns = vb.Alignments.from_gam_bytes(H.encode_gam(subset))
        T, G, R = H.oracle_side(small, alns.arrays(), bep=2e-3)
        _same(T, G, R)


@needs_ref
@pytest.mark.parametrize("name", ["J2a1a1a1.gam", "all_the_same.gam", "two_unique.gam"])
def test_reference_gam_fixtures_bit_for_bit(name):
    """The reference's real-graph alignments against small synthetic tables over the real id range (the literal
    per-path loops cost P per mapping, so P is kept at 64 here)."""
    S = vb.SynthHaplocart(seed=7, n_paths=64, n_rows=11825, n_nodes=11821)
    alns_d = H.decoded_fixture(name)[:24]
    node_seqs = H.graph_from_alignments(alns_d, S.n_nodes)
    alns = vb.Alignments.from_gam_bytes(H.encode_gam(alns_d))
    T, G, R = H.oracle_side(S, alns.arrays(), node_seqs=node_seqs)
    _same(T, G, R)


@needs_ref
def test_posterior_equals_get_posterior(small):
    alns = small.reads(60, seed=209)
    T, G, R = H.oracle_side(small, alns.arrays())
    final, argmax = _same(T, G, R)
    for call in (argmax, 0, small.n_paths - 1, 17):
        idx, conf, total = orc.hc_posterior(final, small.path_names, small.parents_txt, small.children_txt, call)
        clades, conf_text, conf_full, total_ref = ref.hc_posterior(final, small.path_names, small.parents_txt,
                                                                   small.children_txt, call)
        assert total == total_ref                                      # long double, bit for bit
        assert np.array_equal(conf, conf_full[:len(conf)]) and len(conf) == len(conf_full)
        parents = []
        for line in small.parents_txt.split("\n"):
            t = line.split()
            if t and t[0] == small.path_names[call]:
                parents = [x for x in t[1:] if "[" not in x]
        names = [small.path_names[call]] + [parents[j] for j in idx[1:]]
        # the function itself, end to end through write_posterior_log (6 significant digits in its output file)
        if clades is not None:   # None: the function crashed on its own parent_vec[-1] read (Q8) in the forked child
            assert clades == names
            assert np.allclose(conf_text, conf, rtol=2e-5, atol=1e-300)


@needs_ref
def test_posterior_small_tree_on_the_reference():
    names = ["A", "A1", "A2", "A1a", "B"]
    parents = "A root\nA1 A root\nA2 A A root\nA1a A1 A root [x]\nB root\n"
    children = "root A B\nA A1 A2\nA1 A1a\nA2\nA1a\nB\n"
    ll = np.array([-10.0, -9.0, -12.0, -9.5, -11.0], dtype=np.longdouble)
    for call in range(5):
        idx, conf, total = orc.hc_posterior(ll, names, parents, children, call)
        clades, conf_text, conf_full, total_ref = ref.hc_posterior(ll, names, parents, children, call)
        assert total == total_ref and np.array_equal(conf, conf_full)
        assert clades is None or len(clades) == len(idx)


@needs_ref
def test_dedup_equals_dup_remover(small):
    alns = small.reads(300, seed=211)
    a = alns.arrays()
    R = orc.FlatReads(a)
    want = ref.dedup(R)
    got = np.zeros(R.n, dtype=np.uint8)
    assert orc.lib().orc_dedup(orc.C.byref(R.c), None, orc._ptr(got)) == 0
    assert np.array_equal(got.astype(bool), want)
    assert 0 < want.sum() < R.n
    from vgan_b200 import host as vbh
    mask, kept = vbh.dedup(alns)
    assert np.array_equal(mask, want) and kept == int((~want).sum())


@needs_ref
def test_product_facade_flattens_the_references_own_alignment_objects(small):
    """The drop-in boundary compiled against the reference's class: vgan_b200::FlatAlignments::assign (the product's C++ facade,
    vgan_b200/csrc/host/vgan_haplocart.hpp) instantiated with vector<AlignmentInfo*> -- the reference's own src/AlignmentInfo.h,
    objects filled as the readGAM lambda fills them (src/readGAM.h:41-46) -- yields exactly the flat fields that went in: synthetic
    reads with indels and damage, both strands, and the reference's GAM fixtures."""
    import os
    import vgan_b200 as vb
    cases = [small.reads(400, seed=223, sub_rate=0.02, indel_rate=0.01, fixed_len=0, damage=True).arrays()]
    for name in ("test_reads.gam", "J2a1a1a1.gam", "all_the_same_reverse.gam"):
        path = os.path.join(H.GOLDEN, name)
        if os.path.exists(path):
            cases.append(vb.Alignments.from_gam_file(path).arrays())
    for a in cases:
        R = orc.FlatReads(a)
        paired = (np.arange(R.n) % 3 == 0).astype(np.uint8)
        got = ref.facade_flatten(R, paired)
        for key in ("seq_off", "seq", "qual_off", "qual", "mapq", "identity", "map_off", "map_node", "map_offset", "map_rev", "edit_off",
                    "edit_from", "edit_to", "edit_seq_off", "edit_seq"):
            want = np.asarray(a[key])
            assert np.array_equal(np.asarray(got[key]).astype(want.dtype), want), key
        assert np.array_equal(got["paired"], paired)
        assert (a["map_rev"] != 0).any() or len(cases) > 1


@needs_ref
@pytest.mark.parametrize("profiles", [("dhigh5p.prof", "dhigh3p.prof"), ("none.prof", "none.prof"), ("dhigh5p.prof", "")])
def test_damage_matrices_equal_init_deam_probabilities(profiles):
    """The product's collapsed (k5 x k3) damage table and the oracle's, against Damage::initDeamProbabilities."""
    d5, d3 = [os.path.join(H.GOLDEN, "damageProfiles", p) if p else "" for p in profiles]
    D = ref.Damage(d5, d3)
    t5, t3 = [open(f).read() if f else "" for f in (d5, d3)]
    from oracle.oracle import euka_deam_matrix
    for L in (15, 16, 23, 40, 100, 999):
        for l in sorted(set([0, 1, 2, 3, 4, 5, 6, L // 2, L - 6, L - 5, L - 4, L - 3, L - 2, L - 1])):
            if 0 <= l < L:
                assert np.array_equal(euka_deam_matrix(t5, t3, L, l), D.matrix(L, l)), (L, l)


def test_oracle_reproduces_the_committed_reference_vectors():
    """Runs everywhere (no _ref needed): per-read log-likelihood vectors, sums, calls and posteriors that the
    reference's own code produced in the authoring container (tests/golden/make_ref_golden.py)."""
    gold = json.load(open(os.path.join(H.GOLDEN, "ref_hc_golden.json")))
    for case in gold["cases"]:
        T, G, R, tree = H.ref_golden_inputs(case)
        final, argmax, used = orc.hc_run(T, G, R, mode=orc.MODE_HOISTED)
        want = H.ld_from_pairs(case["final_vec"])
        assert np.array_equal(final, want), case["name"]
        assert argmax == case["argmax"] and used == case["used"]
        for r, pairs in case["per_read"].items():
            a, _, _ = orc.hc_read(T, G, R, int(r), mode=orc.MODE_LITERAL)
            assert np.array_equal(a, H.ld_from_pairs(pairs)), (case["name"], r)
        if tree is not None:
            names, parents, children = tree
            idx, conf, total = orc.hc_posterior(final, names, parents, children, argmax)
            assert total == H.ld_from_pairs([case["total_ll"]])[0]
            assert np.array_equal(conf, np.array(case["confidence_full"]))
            assert len(idx) == len(case["clades"])


# ---- siblings: the reference's own euka lambda (readGAM3, src/readGAM_Euka.h:67-577) and soibean analyse_GAM
#      (src/getLCAfromGAM.h:31-700), compiled unchanged and run over injected alignments ---------------------------------
def _euka_inputs():
    S = vb.SynthEuka(seed=11, n_clades=40, genome_len=3000)
    alns = S.reads(600, seed=12, n_rate=0.01, softclip_rate=0.2, indel_rate=0.01)
    return S, orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()), alns


@needs_ref
@pytest.mark.parametrize("p5,p3", [(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF), ("", ""), (vb.DHIGH5P_PROF, "")])
def test_euka_lambda_bit_for_bit(p5, p3, capfd):
    S, og, R, _keep = _euka_inputs()
    want = orc.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, og, R, p5, p3)
    got = ref.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, og, R, p5, p3)
    m = want["clade"] >= 0
    assert m.sum() >= 590 and np.array_equal(want["clade"], got["clade"])
    assert np.array_equal(want["like"][m], got["like"][m]) and np.array_equal(want["not_like"][m], got["not_like"][m])   # doubles, bit for bit
    assert np.array_equal(want["clade_count"], got["clade_count"]) and want["clade_count"].sum() > 30
    assert np.array_equal(want["bin_cov"], got["bin_cov"])


@needs_ref
def test_euka_lambda_symbol_cases_on_the_toy_graph(capfd):
    nodes, _ = H.toy_graph()
    og = orc.Graph(nodes)
    n9, n13 = nodes[9], nodes[13]
    reads = [
        H.aln("TTTTTTT" + n9[:2] + "N" + n9[3:7] + "Y" + n9[9:], [(9, 0, False, [(0, 7, "TTTTTTT"), (2, 2, ""), (1, 1, "N"), (4, 4, ""), (1, 1, "Y"), (1, 0, ""), (9, 9, "")])],
              quality=bytes([10, 20, 30, 40, 41, 35, 2] * 4)[:24], mapq=37),
        H.aln(H.revcomp(n13), [(13, 0, True, [(32, 32, "")])], quality=bytes([33] * 32), mapq=60),
        H.aln(n13[:20] + "ACGT" + n13[20:], [(13, 0, False, [(20, 20, ""), (0, 4, "ACGT"), (12, 12, "")])], quality=bytes([25] * 10), mapq=30),
        H.aln(n13, [(13, 0, False, [(32, 32, "")])], quality=bytes([99, 0, 1, 2, 3] * 7)[:32], mapq=0),   # (a byte >= 100 indexes qscore_vec out of range in the reference)
    ]
    dist, boff, blo, bhi = [0.12, 0.07], [0, 2, 3], [1, 10, 12], [9, 11, 40]
    R = orc.Reads(reads)
    want = orc.euka_run(dist, boff, blo, bhi, og, R, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    got = ref.euka_run(dist, boff, blo, bhi, og, R, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    assert np.array_equal(want["clade"], got["clade"]) and list(got["clade"]) == [0, 1, 1, 1]
    assert np.array_equal(want["like"], got["like"]) and np.array_equal(want["not_like"], got["not_like"])
    assert np.array_equal(want["clade_count"], got["clade_count"]) and np.array_equal(want["bin_cov"], got["bin_cov"])


@needs_ref
@pytest.mark.parametrize("P,p5,p3,penalty", [(31, "", "", 7), (127, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7), (63, vb.DHIGH5P_PROF, "", 3)])
def test_soibean_analyse_gam_bit_for_bit(P, p5, p3, penalty, capfd):
    S = vb.SynthHaplocart(seed=81 + P, n_paths=P, n_rows=610, n_nodes=601)
    alns = S.reads(200, seed=82 + P, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    g, R, member = orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()), S.supports()
    want, used_w = orc.soibean_run(g, R, member, 0, penalty, p5, p3)
    got, used = ref.soibean_run(g, R, member, 0, penalty, p5, p3)
    assert np.array_equal(used_w, used) and used.sum() >= 195
    assert np.array_equal(want, got)                       # doubles, bit for bit


@needs_ref
def test_soibean_analyse_gam_on_the_toy_reads(capfd):
    nodes, paths = H.toy_graph()
    member = np.zeros((max(nodes) + 1, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            member[nid, p] = 1
    R = orc.Reads(read_gam(os.path.join(H.GOLDEN, "test_reads.gam")))     # match / mismatch / indel / softclip x both strands
    want, used_w = orc.soibean_run(orc.Graph(nodes), R, member, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    got, used = ref.soibean_run(orc.Graph(nodes), R, member, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    assert used.all() and np.array_equal(used_w, used) and np.array_equal(want, got)


@needs_ref
def test_soibean_base_loglike_equals_compute_base_log_like():
    """MCMC::computeBaseLogLike (src/MCMC.h:111-296), the per-column HKY term of the soibean MCMC likelihood."""
    F7 = orc.soibean_freqs7(0.31, 0.19, 0.2, 0.3)
    n = 0
    for rb in "ACGT":
        for gb in "ACGT":
            for t in (0.0, 1e-3, 0.05, 0.5):
                for bl in (0.0, 0.01, 0.3):
                    def call(f):
                        try:
                            return f(rb, gb, -0.3, F7, t, bl)
                        except RuntimeError:     # the reference throws on a nan / inf term; the oracle reports a range error
                            return None
                    a, b = call(orc.soibean_base_loglike), call(ref.soibean_base_loglike)
                    assert a == b, (rb, gb, t, bl, a, b)
                    n += a is not None
    assert n >= 100          # the rest (branch length 0 -> log 0) throw in the reference and are range errors in the oracle


def test_oracle_reproduces_the_committed_sibling_vectors():
    """Runs everywhere: euka per-read (clade, like, not_like) and a soibean R x P matrix the reference's own code produced."""
    gold = json.load(open(os.path.join(H.GOLDEN, "ref_sibling_golden.json")))
    S, og, R, _keep = _euka_inputs()
    e = gold["euka"]
    want = orc.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, og, R, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    assert want["clade"].tolist() == e["clade"]
    m = want["clade"] >= 0
    assert np.array_equal(want["like"][m], np.array(e["like"])[m]) and np.array_equal(want["not_like"][m], np.array(e["not_like"])[m])
    assert want["clade_count"].tolist() == e["clade_count"] and np.array_equal(want["bin_cov"], np.array(e["bin_cov"]))
    s = gold["soibean"]
    S2 = vb.SynthHaplocart(seed=s["seed"], n_paths=s["P"], n_rows=610, n_nodes=601)
    alns = S2.reads(s["n"], seed=s["seed"] + 1, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    mat, used = orc.soibean_run(orc.Graph(S2.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()), S2.supports(), 0, 7,
                                vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    assert np.array_equal(mat, np.array(s["matrix"])) and used.tolist() == s["used"]
