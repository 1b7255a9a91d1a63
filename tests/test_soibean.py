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
