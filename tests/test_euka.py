"""euka sibling kernel: CPU tests of the host side, GPU parity against the oracle."""
import numpy as np
import pytest

import vgan_b200 as vb
from vgan_b200 import _lib
from oracle import oracle as orc

import helpers as H


def _oracle(S, alns, p5="", p3="", minimum_mq=29, node_seqs=None):
    og = orc.Graph(S.node_sequences() if node_seqs is None else node_seqs, min_id=1)
    return orc.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, og, orc.FlatReads(alns.arrays()), p5, p3, minimum_mq)


def test_damage_matrices_match_the_reference_construction():
    S = vb.SynthEuka(seed=3, n_clades=4, genome_len=1000)
    for p5, p3 in ((vb.DHIGH5P_PROF, vb.DHIGH3P_PROF), ("", ""), (vb.DHIGH5P_PROF, ""), ("", vb.DHIGH3P_PROF)):
        d = S.tables(p5, p3).deam()
        k5, k3 = d.shape[:2]
        assert (k5, k3) == (5 if p5 else 1, 5 if p3 else 1)
        for L in (15, 16, 23, 40, 100, 1000):
            for l in {0, 1, 2, 3, 4, 5, L // 2, L - 6, L - 5, L - 2, L - 1}:
                if 0 <= l < L:
                    want = orc.euka_deam_matrix(p5, p3, L, l)
                    assert np.array_equal(d[min(l, k5 - 1), min(L - l - 1, k3 - 1)], want), (L, l)
        assert np.allclose(d.sum(axis=3), 1.0, atol=1e-15)
    # the row with the smaller diagonal wins, per original base: a C at the 5' end takes the 5' row, a G the 3' row
    d = S.tables(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF).deam()
    assert d[0, 0][1, 3] == pytest.approx(0.329405) and d[0, 0][2, 0] == pytest.approx(0.32891)
    with pytest.raises(_lib.VglError):
        S.tables("header\n0\t1\n", "")  # 2 fields rather than 12


def test_packer_symbols_and_quality_by_aligned_column():
    nodes, _ = H.toy_graph()
    g = vb.Graph.from_sequences(nodes)
    # node 9 = GCTACGATGGATGAGTCT; an insertion at node offset 0 is written 'S' by the reconstruction (quirk Q3)
    a = H.aln("TTGCNACGAYGGATGAGT", [(9, 0, False, [(0, 2, "TT"), (2, 2, ""), (1, 1, "N"), (4, 4, ""), (1, 1, "Y"), (1, 0, ""), (7, 7, "")])],
              quality=bytes(range(30, 48)), mapq=37)
    alns = vb.Alignments.from_gam_bytes(H.encode_gam([a]))
    gs, rs, _ = vb.reconstruct(g, alns, 0)
    assert gs == "SSGCTACGATGGATGAGT" and rs == "TTGCNACG-AYGATGAGT"  # the gap lands at the from_length coordinate (vgan_utils.h:67)
    pb = vb.EukaPackedBatch(g, alns)
    b = pb.batch
    cols = vb.host._as_array(b.col, int(b.n_col_blks) * 256, np.uint16).reshape(-1, 32, 8)[:, 0, :].reshape(-1)[:len(gs)]
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5, "Y": 6, "S": 7}
    qual = list(range(30, 48)) + [0] * 4
    for m, rec in enumerate(cols.tolist()):
        assert rec & 0xff == qual[m]                      # the quality string is indexed by aligned column
        assert (rec >> 8) & 15 == code[gs[m]] and rec >> 12 == code[rs[m]]
    assert vb.host._as_array(b.seq_len, 32, np.uint16)[0] == 18 and vb.host._as_array(b.mapq, 32, np.uint8)[0] == 37
    # fragments outside [15, 1000] index the reference's damage table out of range: rejected
    short = vb.Alignments.from_gam_bytes(H.encode_gam([H.aln("GCTACGA", [(9, 0, False, [(7, 7, "")])])]))
    with pytest.raises(_lib.VglError):
        vb.EukaPackedBatch(g, short)


def test_oracle_hand_checked_column():
    """One perfect-match column without damage: log((1-d)(1-e) + d*e/3*(sum of tt ratios)) by hand."""
    nodes, _ = H.toy_graph()
    og = orc.Graph(nodes)
    read = "GCTACGATGGATGAGTCT"
    a = H.aln(read, [(9, 0, False, [(18, 18, "")])], quality=bytes([30] * 18), mapq=60)
    out = orc.euka_run([0.1], [0, 1], [1], [40], og, orc.Reads([a]))
    d, e = 0.1, 1e-3
    tt = {"A": {"C": .02381, "G": .95238, "T": .02381}, "C": {"A": .02381, "G": .02381, "T": .95238},
          "G": {"A": .95238, "C": .02381, "T": .02381}, "T": {"A": .02381, "C": .95238, "G": .02381}}
    want_in = sum(np.log((1 - d) * (1 - e) + sum(d * tt[b][x] * e / 3 for x in tt[b])) for b in read)
    assert out["in_lik"][0] == pytest.approx(want_in, rel=1e-13)
    assert out["not_lik"][0] == pytest.approx(18 * np.log(1 - 0.25536), rel=1e-14)
    lse = np.logaddexp(out["in_lik"][0], out["not_lik"][0])
    assert out["like"][0] == pytest.approx((1 - 1e-6) * np.exp(out["in_lik"][0] - lse), rel=1e-13)
    assert out["pass"][0] == 1 and out["clade_count"][0] == 1 and out["bin_cov"][0] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("p5,p3", [(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF), ("", ""), (vb.DHIGH5P_PROF, "")])
def test_synthetic_reads_match_oracle(p5, p3):
    S = vb.SynthEuka(seed=11, n_clades=40, genome_len=3000)
    alns = S.reads(3000, seed=12, n_rate=0.01, softclip_rate=0.2, indel_rate=0.01)
    want = _oracle(S, alns, p5, p3)
    ek = vb.Euka()
    ek.load(S.tables(p5, p3))
    got = ek.run(vb.EukaPackedBatch(S.graph, alns))
    count, cov = ek.results()
    mapped = alns.arrays()["identity"] != 0
    assert np.array_equal(got["clade"][mapped], want["clade"][mapped]) and np.all(got["clade"][~mapped] == -1)
    assert H.rel_err(got["in_lik"], want["in_lik"]) < 1e-11 and H.rel_err(got["not_lik"], want["not_lik"]) < 1e-12
    assert np.allclose(got["like"], want["like"], rtol=1e-9, atol=1e-300)
    assert np.allclose(got["not_like"][mapped], want["not_like"][mapped], rtol=1e-9)
    # the pass filter is a threshold on in - not: identical except for reads within rounding of the threshold
    margin = np.abs(want["in_lik"] - want["not_lik"] - 1.0)
    assert np.array_equal(got["pass"][margin > 1e-9], want["pass"][margin > 1e-9])
    if np.all(margin > 1e-9):
        assert np.array_equal(count, want["clade_count"])
        assert np.allclose(cov, want["bin_cov"], rtol=1e-12) and np.array_equal(cov == 0, want["bin_cov"] == 0)
    assert want["pass"].sum() > 100 and set(np.unique(want["clade"])) >= set(range(40))
    # abundance MCMC (MCMC::get_proposal_likelihood, src/MCMC.cpp:1175-1213): the per-clade sums the read kernel keeps are the
    # reference's clade_like vectors reduced the way the proposal likelihood reads them (a mapping quality of 0 gives
    # clade_like = 0 and log 0 = -inf, on both sides)
    like_n, like_logsum = ek.like_sums()
    assert np.array_equal(like_n, np.bincount(want["clade"][mapped], minlength=40))
    with np.errstate(divide="ignore"):
        logs = np.log(want["like"])
    finite = 0
    for c in range(40):
        sel = mapped & (want["clade"] == c)
        ref = logs[sel].sum()
        if np.isfinite(ref):
            finite += 1
            assert abs(like_logsum[c] - ref) < 1e-9 * abs(ref)
        else:
            assert like_logsum[c] == -np.inf
    listed, frac = [3, 17, 29], [0.2, 0.5, 0.3]
    ll_want = orc.euka_proposal_loglike(np.where(mapped, want["clade"], -1), want["like"], want["not_like"], 40, listed, frac)
    ll_got = vb.Euka.proposal_loglike(listed, frac, like_n, like_logsum)
    assert (ll_got == ll_want == -np.inf) or abs(ll_got - ll_want) < 1e-9 * abs(ll_want)
    ek.close()


@pytest.mark.gpu
def test_symbol_cases_and_strands():
    """Hand-made alignments on the toy graph: N, gap, rare code, soft clips (every third counts as a match), both strands,
    quality indexed by aligned column past the end of the quality string."""
    nodes, _ = H.toy_graph()
    g = vb.Graph.from_sequences(nodes)
    og = orc.Graph(nodes)
    n9, n13 = nodes[9], nodes[13]
    reads = [
        H.aln("TTTTTTT" + n9[:2] + "N" + n9[3:7] + "Y" + n9[9:], [(9, 0, False, [(0, 7, "TTTTTTT"), (2, 2, ""), (1, 1, "N"), (4, 4, ""), (1, 1, "Y"), (1, 0, ""), (9, 9, "")])],
              quality=bytes([10, 20, 30, 40, 41, 35, 2] * 4)[:24], mapq=37),
        H.aln(H.revcomp(n13), [(13, 0, True, [(32, 32, "")])], quality=bytes([33] * 32), mapq=60),
        H.aln(n13[:20] + "ACGT" + n13[20:], [(13, 0, False, [(20, 20, ""), (0, 4, "ACGT"), (12, 12, "")])], quality=bytes([25] * 10), mapq=30),
        H.aln(n13, [(13, 0, False, [(32, 32, "")])], quality=bytes([200, 0, 1, 2, 3] * 7)[:32], mapq=0),
    ]
    alns = vb.Alignments.from_gam_bytes(H.encode_gam(reads))
    dist, boff, blo, bhi = [0.12, 0.07], [0, 2, 3], [1, 10, 12], [9, 11, 40]
    want = orc.euka_run(dist, boff, blo, bhi, og, orc.Reads(reads), vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    ek = vb.Euka()
    ek.load(vb.EukaTables(dist, boff, blo, bhi, max(nodes), vb.DHIGH5P_PROF, vb.DHIGH3P_PROF))
    got = ek.run(vb.EukaPackedBatch(g, alns))
    assert np.array_equal(got["clade"], want["clade"]) and list(want["clade"]) == [0, 1, 1, 1]
    assert H.rel_err(got["in_lik"], want["in_lik"]) < 1e-12 and H.rel_err(got["not_lik"], want["not_lik"]) < 1e-13
    assert np.allclose(got["like"], want["like"], rtol=1e-10) and np.array_equal(got["pass"], want["pass"])
    count, cov = ek.results()
    assert np.array_equal(count, want["clade_count"]) and np.allclose(cov, want["bin_cov"], rtol=1e-14)
    ek.close()


def test_oracle_proposal_likelihood_is_counts_and_log_sums():
    """get_proposal_likelihood as written: index 0 of every clade_like vector is the 0.0 it was created with and is skipped,
    `1/334` is an integer division, so a proposal's value is sum_j n_j log(frac_j) + sum log(clade_like)."""
    rng = np.random.default_rng(3)
    n, C_ = 500, 6
    clade = rng.integers(-1, C_, size=n).astype(np.int32)
    like = rng.uniform(0.01, 1.0, size=n)
    not_like = 1 - like
    ids, frac = [4, 0, 2], [0.5, 0.2, 0.3]
    got = orc.euka_proposal_loglike(clade, like, not_like, C_, ids, frac)
    want = sum((clade == c).sum() * np.log(f) + np.log(like[clade == c]).sum() for c, f in zip(ids, frac))
    assert abs(got - want) < 1e-12 * abs(want)
    # the clade_not_like term really is gone
    assert orc.euka_proposal_loglike(clade, like, not_like * 7, C_, ids, frac) == got
    # the product's host-side evaluation from the per-clade sums (no device needed for this one)
    like_n = np.bincount(clade[clade >= 0], minlength=C_)
    like_logsum = np.array([np.log(like[clade == c]).sum() for c in range(C_)])
    assert abs(vb.Euka.proposal_loglike(ids, frac, like_n, like_logsum) - got) < 1e-12 * abs(got)


@pytest.mark.gpu
def test_config4_tables_taken_verbatim_from_the_reference(capfd):
    """335 clades / 3929 bins exactly as share/vgan/euka_dir/euka_db.clade / .bins list them
    (copied to tests/golden/), synthetic node sequences; packed and raw routes against the oracle, then the resident re-run."""
    import os
    S = vb.SynthEuka.from_tables(os.path.join(H.GOLDEN, "euka_db.clade"), os.path.join(H.GOLDEN, "euka_db.bins"))
    assert S.n_clades == 335 and int(S.bin_off[-1]) == 3929 and S.n_nodes == 6926126
    alns = S.reads(4000, seed=20265)
    # (the oracle's graph view of 6.9 M nodes: build it from the flat arrays rather than from a Python list of strings)
    g = orc.Graph.__new__(orc.Graph)
    g.min_id, g.n_nodes = 1, S.n_nodes
    g.seq_off = np.ascontiguousarray(S._off, dtype=np.int64)
    g.seq = np.concatenate([S._seq, np.zeros(1, dtype=np.uint8)])
    g.c = orc._Graph(g.n_nodes, g.min_id, orc._ptr(g.seq_off), orc._ptr(g.seq))
    want = orc.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, g, orc.FlatReads(alns.arrays()), vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    ek = vb.Euka()
    ek.load(S.tables(vb.DHIGH5P_PROF, vb.DHIGH3P_PROF))
    ek.load_graph(S.graph)
    mapped = alns.arrays()["identity"] != 0
    margin = np.abs(want["in_lik"] - want["not_lik"] - 1.0)
    feeder = vb.Feeder(n_threads=2, chunk_reads=1024)          # four chunks; node ids up to 6.9 M travel as 32 bits
    for route in ("packed", "raw", "feeder", "rerun"):
        ek.reset()
        if route == "packed":
            got = ek.run(vb.EukaPackedBatch(S.graph, alns))
        elif route == "raw":
            got = ek.run_raw(vb.RawBatch(alns, euka_rule=True))
        elif route == "feeder":
            got = feeder.feed_euka(ek, alns)
            assert feeder.stats()["chunks"] == 4 and feeder.stats()["slow_chunks"] == 0
        else:                                                      # (the feeder left only its last chunk resident)
            ek.reset()
            ek.run_raw(vb.RawBatch(alns, euka_rule=True))
            ek.reset()
            got = ek.rerun(len(alns), want=True)
        assert np.array_equal(got["clade"][mapped], want["clade"][mapped]), route
        assert H.rel_err(got["in_lik"], want["in_lik"]) < 1e-11 and H.rel_err(got["not_lik"], want["not_lik"]) < 1e-12
        assert np.allclose(got["like"], want["like"], rtol=1e-9, atol=1e-300)
        assert np.array_equal(got["pass"][margin > 1e-9], want["pass"][margin > 1e-9])
        count, cov = ek.results()
        if np.all(margin > 1e-9):
            assert np.array_equal(count, want["clade_count"]) and np.allclose(cov, want["bin_cov"], rtol=1e-12)
        assert ek.kernel_ms() > 0
    assert len(set(want["clade"][mapped].tolist())) > 200      # reads land in clades in proportion to their node ranges
    ek.close()
