"""soibean sibling kernel: per-read x per-path matrix against the oracle."""
import numpy as np
import pytest

import vgan_b200 as vb
from vgan_b200 import _lib
from oracle import oracle as orc

import helpers as H


def _inputs(P, seed, n_reads, **kw):
    S = vb.SynthHaplocart(seed=seed, n_paths=P, n_rows=610, n_nodes=601)
    alns = S.reads(n_reads, seed=seed + 1, **kw)
    return S, alns


def test_packer_edit_records():
    S, alns = _inputs(31, 71, 200, fixed_len=60, sub_rate=0.03, indel_rate=0.01)
    arr = alns.arrays()
    og = orc.Graph(S.node_sequences(), min_id=1)
    R = orc.FlatReads(arr)
    pb = vb.SoibeanPackedBatch(S.graph, alns)
    b = pb.batch
    G = int(b.n_groups)
    maps = vb.host._as_array(b.map, int(b.n_map_blks) * 128, np.uint32).reshape(-1, 32, 4)
    blk = vb.host._as_array(b.grp_map_blk, G + 1, np.uint32)
    n_maps = vb.host._as_array(b.n_maps, G * 32, np.uint16)
    ridx = vb.host._as_array(b.read_index, G * 32, np.uint32)
    seen_extra = False
    for lane in range(pb.n_reads):
        g, l = divmod(lane, 32)
        r = int(ridx[lane])
        _, _, sizes = orc.reconstruct(og, R, r)
        M = int(arr["map_off"][r + 1] - arr["map_off"][r])
        recs = maps[blk[g]:blk[g + 1], l, :].reshape(-1)[:n_maps[lane]]
        assert len(recs) == len(sizes)                       # one record per EDIT (getLCAfromGAM.h:144)
        nodes = arr["map_node"][arr["map_off"][r]:arr["map_off"][r + 1]]
        for i, rec in enumerate(recs.tolist()):
            assert rec >> 20 == sizes[i]
            if len(sizes) != M and i >= M:
                assert rec & 0xFFFFF == 0xFFFFF              # surplus entries get no support (:156-160)
                seen_extra = True
            else:
                assert rec & 0xFFFFF == nodes[i]
    assert seen_extra


@pytest.mark.gpu
@pytest.mark.parametrize("P,p5,p3,penalty", [(31, "", "", 7), (127, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7), (399, vb.DHIGH5P_PROF, "", 3)])
def test_path_matrix_matches_oracle(P, p5, p3, penalty):
    S, alns = _inputs(P, 81 + P, 1500, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    member = S.supports()                                    # row index == node id
    want, used_w = orc.soibean_run(orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()), member, 0, penalty, p5, p3)
    sb = vb.Soibean()
    sb.load(vb.SoibeanTables(member, 0, p5, p3, penalty))
    got, used = sb.analyse(vb.SoibeanPackedBatch(S.graph, alns))
    assert np.array_equal(used, used_w) and used.sum() >= 1490
    assert H.rel_err(got, want) < 1e-11
    # the per-read most probable paths (AlignmentInfo::mostProbPath) agree wherever the maximum is not a rounding-level tie
    top_w = np.sort(want, axis=1)
    clear = (top_w[:, -1] - top_w[:, -2] > 1e-9) & (used_w == 1)
    assert np.array_equal(got.argmax(axis=1)[clear], want.argmax(axis=1)[clear]) and clear.sum() > 0
    # the feeder route (vgh_soibean_feed): flat alignment fields -> page-locked chunks -> device reconstruction; several chunks
    sb.load_graph(S.graph)
    f = vb.Feeder(n_threads=2, chunk_reads=400)
    fed, fed_used = f.feed_soibean(sb, alns)
    assert f.stats()["chunks"] == 4 and f.stats()["slow_chunks"] == 0
    assert np.array_equal(fed_used, used_w) and H.rel_err(fed, want) < 1e-11
    f.close()
    sb.close()


@pytest.mark.gpu
def test_reverse_strand_and_symbol_cases_on_the_toy_graph():
    nodes, paths = H.toy_graph()
    n_rows = max(nodes) + 1
    member = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            member[nid, p] = 1
    import os
    alns = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, "test_reads.gam"))   # match / mismatch / indel / softclip x both strands
    want, used_w = orc.soibean_run(orc.Graph(nodes), orc.FlatReads(alns.arrays()), member, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    sb = vb.Soibean()
    sb.load(vb.SoibeanTables(member, 0, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7))
    got, used = sb.analyse(vb.SoibeanPackedBatch(vb.Graph.from_sequences(nodes), alns))
    assert used.all() and H.rel_err(got, want) < 1e-12
    sb.close()


# ---- MCMC proposal likelihood (src/MCMC.cpp:738-993, src/MCMC.h:111-316) -----------------------------
F7 = orc.soibean_freqs7(0.31, 0.19, 0.2, 0.3)


def test_oracle_base_loglike_against_the_closed_form():
    """computeBaseLogLike restated independently: HKY with kappa = 1/22 == 0 (integer division, src/MCMC.h:66)."""
    import math
    a, c, g, t, R, Y, mu = F7
    assert abs(mu - 1 / (2 * (22 * a * g + 22 * c * t + a * c + a * t + g * c + g * t))) < 1e-18
    tt, con = 0.013, 0.01
    f = dict(A=a, C=c, G=g, T=t)

    def hky(rb, ref):
        fam = R if rb in "AG" else Y
        A = 1 - fam                                           # 1 + fam * (kappa - 1) with kappa = 0
        jut1 = f[rb] + f[rb] * (1 / fam - 1) * math.exp(-mu * tt)
        if rb == ref:
            p = jut1 + (fam - f[rb]) / fam * math.exp(-mu * tt * A)
        elif {rb, ref} in ({"A", "G"}, {"C", "T"}):
            p = abs(jut1 - f[rb] / fam * math.exp(-mu * tt * A))
        else:
            p = f[rb] * (1 - math.exp(-mu * tt))
        return max(p, 1e-8)

    for read_b in "ACGT":
        for ref in "ACGT":
            want = math.log(sum(hky(rb, ref) * ((1 - con) if rb == read_b else con / 3) for rb in "ACGT")) - 0.125
            got = orc.soibean_base_loglike(read_b, ref, -0.125, F7, tt, con)
            assert abs(got - want) < 1e-13, (read_b, ref)
    # a match is worth about log(1 - con), a mismatch costs several nats
    assert abs(orc.soibean_base_loglike("A", "A", 0.0, F7, 0.01, 0.01) - (-0.0109016)) < 1e-6
    assert orc.soibean_base_loglike("G", "A", 0.0, F7, 0.01, 0.01) < -5


def _mcmc_inputs(P=31, n=400, seed=91):
    S, alns = _inputs(P, seed, n, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    return S, alns, S.supports(), orc.Graph(S.node_sequences(), min_id=1), orc.FlatReads(alns.arrays())


def test_oracle_mcmc_properties():
    S, alns, member, og, R = _mcmc_inputs()
    p5, p3 = vb.DHIGH5P_PROF, vb.DHIGH3P_PROF
    want, used = orc.soibean_run(og, R, member, 0, 7, p5, p3)
    # child == parent in the middle of the branch: both sides are the same number, the weighted average returns it
    ll, per = orc.soibean_mcmc_loglike(og, R, member, [(5, 5, 0.5, 0.02, 1.0)], F7, 0.01, 0, 7, p5, p3)
    assert np.all(per[used == 1] < 0) and np.all(per[used == 0] == 0) and abs(ll - per.sum()) < 1e-9 * abs(ll)
    # every supported regular column adds an HKY term <= 0, nothing else changes: never above pathMap
    assert np.all(per <= want[:, 5] + 1e-9)
    # a path that goes through none of the read's nodes: no column is supported, the proposal's value is pathMap itself
    member0 = member.copy()
    member0[:, 5] = 0
    want0, _ = orc.soibean_run(og, R, member0, 0, 7, p5, p3)
    _, per0 = orc.soibean_mcmc_loglike(og, R, member0, [(5, 5, 0.5, 0.02, 1.0)], F7, 0.01, 0, 7, p5, p3)
    assert np.allclose(per0, want0[:, 5], rtol=1e-12, atol=0) and np.all(per0[used == 1] < per[used == 1])
    # two sources, proportions (1, 0): log(0) removes the second one
    ll1, per1 = orc.soibean_mcmc_loglike(og, R, member, [(5, 9, 0.3, 0.02, 1.0)], F7, 0.01, 0, 7, p5, p3)
    ll2, per2 = orc.soibean_mcmc_loglike(og, R, member, [(5, 9, 0.3, 0.02, 1.0), (11, 2, 0.6, 0.01, 0.0)], F7, 0.01, 0, 7, p5, p3)
    assert np.allclose(per1, per2, rtol=1e-12, atol=0)


PROPOSALS = [
    [(5, 9, 0.3, 0.02, 1.0)],
    [(5, 5, 0.5, 0.0, 1.0)],          # root: no branch, t becomes 1e-5
    [(12, 3, 1.0, 0.05, 1.0)],        # at the end of the branch: log(1 - pos_branch) = -inf on the parent's side
    [(12, 3, 0.0, 0.05, 1.0)],
]
PROPOSALS2 = [
    [(5, 9, 0.3, 0.02, 0.6), (11, 2, 0.9, 0.0, 0.4)],
    [(0, 1, 0.5, 0.01, 0.01), (30, 29, 0.25, 0.2, 0.99)],
]
PROPOSALS3 = [[(5, 9, 0.3, 0.02, 0.2), (11, 2, 0.9, 0.0, 0.3), (20, 21, 0.45, 0.03, 0.5)]]


@pytest.mark.gpu
@pytest.mark.parametrize("route", ["packed", "raw"])
@pytest.mark.parametrize("p5,p3,penalty", [("", "", 7), (vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 3)])
def test_mcmc_proposal_likelihood_matches_oracle(route, p5, p3, penalty):
    S, alns, member, og, R = _mcmc_inputs(P=31, n=1200, seed=93)
    sb = vb.Soibean()
    sb.load(vb.SoibeanTables(member, 0, p5, p3, penalty))
    with pytest.raises(RuntimeError):                         # nothing resident yet
        sb.mcmc_loglike(PROPOSALS, F7, 0.01)
    if route == "packed":
        sb.analyse(vb.SoibeanPackedBatch(S.graph, alns))
    else:
        sb.load_graph(S.graph)
        sb.analyse_raw(vb.RawBatch(alns, euka_rule=True))
    n = len(alns)
    for props in (PROPOSALS, PROPOSALS2, PROPOSALS3):
        got, got_per = sb.mcmc_loglike(props, F7, 0.01, n_out=n, per_read=True)
        for q, prop in enumerate(props):
            want, want_per = orc.soibean_mcmc_loglike(og, R, member, prop, F7, 0.01, 0, penalty, p5, p3)
            assert H.rel_err(got_per[q], want_per) < 1e-11
            assert abs(got[q] - want) < 1e-11 * abs(want)
        again = sb.mcmc_loglike(props, F7, 0.01)              # fixed-order reduction: bit-identical from run to run
        assert np.array_equal(again, got)
    sb.close()


def test_oracle_mcmc_on_the_toy_graph_fixture():
    """The reference's reconstruction fixture (match / mismatch / indel / softclip on both strands) through the MCMC
    likelihood: finite, never above the pre-computed pathMap value, and the mixture of a source with itself is that source."""
    import os
    nodes, paths = H.toy_graph()
    n_rows = max(nodes) + 1
    member = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            member[nid, p] = 1
    alns = vb.Alignments.from_gam_file(os.path.join(H.GOLDEN, "test_reads.gam"))
    og, R = orc.Graph(nodes), orc.FlatReads(alns.arrays())
    want, used = orc.soibean_run(og, R, member, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    P = len(paths)
    for c in range(P):
        ll, per = orc.soibean_mcmc_loglike(og, R, member, [(c, (c + 1) % P, 0.35, 0.04, 1.0)], F7, 0.01, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
        assert np.all(np.isfinite(per)) and np.all(per[used == 1] < 0)
        # log(0.35 e^a + 0.65 e^b) <= max(a, b) <= max of the two pathMap values
        assert np.all(per <= np.maximum(want[:, c], want[:, (c + 1) % P]) + 1e-9)
        # two sources that are the same source, proportions adding up to one: the single-source value of the mixture form
        _, one = orc.soibean_mcmc_loglike(og, R, member, [(c, c, 0.5, 0.04, 1.0)], F7, 0.01, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
        _, two = orc.soibean_mcmc_loglike(og, R, member, [(c, c, 0.5, 0.04, 0.3), (c, c, 0.5, 0.04, 0.7)], F7, 0.01, 0, 7, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
        assert np.allclose(one, two, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_rerun_on_the_resident_batch_gives_the_same_matrix():
    S, alns = _inputs(127, 95, 900, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    sb = vb.Soibean()
    sb.load(vb.SoibeanTables(S.supports(), 0, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF, 7))
    pk = vb.SoibeanPackedBatch(S.graph, alns)
    got, used = sb.analyse(pk)
    again, used2 = np.full_like(got, 7.0), np.zeros_like(used)
    sb.rerun(pk.n_span, again, used2)
    assert np.array_equal(again, got) and np.array_equal(used2, used) and sb.kernel_ms() > 0
    sb.rerun(pk.n_span)                      # matrix left on the device: the MCMC kernel still scores the same batch
    f7 = vb.soibean_freqs7(0.31, 0.19, 0.2, 0.3)
    prop = [[(5, 2, 0.4, 0.01, 1.0)]]
    assert np.isfinite(sb.mcmc_loglike(prop, f7, 0.01)[0])
    sb.close()
