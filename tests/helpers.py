"""Shared test scaffolding: builds identical inputs for the product (C ABI) and the oracle (checker)."""
import json
import os
import struct

import numpy as np

from oracle import oracle as orc
from oracle.gam_py import read_gfa

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- a minimal GAM writer, so hand-made alignments reach the product through its own GAM reader ----
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(no, wt, payload):
    key = _varint((no << 3) | wt)
    if wt == 0:
        return key + _varint(payload)
    if wt == 1:
        return key + payload
    return key + _varint(len(payload)) + payload


def encode_alignment(a):
    path = b""
    for m in a["mappings"]:
        pos = _field(1, 0, m["node_id"])
        if m.get("offset", 0):
            pos += _field(2, 0, m["offset"])
        if m.get("is_reverse", False):
            pos += _field(4, 0, 1)
        body = _field(1, 2, pos)
        for e in m["edits"]:
            eb = b""
            if e["from_length"]:
                eb += _field(1, 0, e["from_length"])
            if e["to_length"]:
                eb += _field(2, 0, e["to_length"])
            if e["sequence"]:
                eb += _field(3, 2, e["sequence"])
            body += _field(2, 2, eb)
        path += _field(2, 2, body)
    msg = _field(1, 2, a["sequence"]) + _field(2, 2, path)
    if a.get("quality"):
        msg += _field(4, 2, a["quality"])
    if a.get("mapping_quality", 0):
        msg += _field(5, 0, a["mapping_quality"])
    msg += _field(16, 1, struct.pack("<d", a.get("identity", 1.0)))
    if a.get("read_paired", False):
        msg += _field(20, 0, 1)
    return msg


def encode_gam(alns, group=100):
    out = bytearray()
    for i in range(0, len(alns), group):
        chunk = alns[i:i + group]
        out += _varint(len(chunk) + 1) + _varint(3) + b"GAM"
        for a in chunk:
            m = encode_alignment(a)
            out += _varint(len(m)) + m
    return bytes(out)


def aln(seq, mappings, quality=None, mapq=60, identity=1.0):
    """mappings: list of (node_id, offset, is_reverse, [(from, to, seq), ...])"""
    seq = seq.encode() if isinstance(seq, str) else seq
    if quality is None:
        quality = bytes([40] * len(seq))
    return {"sequence": seq, "quality": bytes(quality), "mapping_quality": mapq, "identity": identity, "read_paired": False,
            "mappings": [{"node_id": n, "offset": o, "is_reverse": r,
                          "edits": [{"from_length": f, "to_length": t, "sequence": s.encode() if isinstance(s, str) else s}
                                    for (f, t, s) in eds]} for (n, o, r, eds) in mappings]}


def decoded_fixture(name):
    """Alignments of a golden GAM as decoded by the oracle-side reader when the fixture was made."""
    with open(os.path.join(GOLDEN, name + ".decoded.json")) as fh:
        raw = json.load(fh)
    out = []
    for a in raw:
        out.append({"sequence": a["sequence"].encode(), "quality": bytes(a["quality"]), "mapping_quality": a["mapping_quality"],
                    "identity": a["identity"], "read_paired": a["read_paired"],
                    "mappings": [{"node_id": m["node_id"], "offset": m["offset"], "is_reverse": m["is_reverse"],
                                  "edits": [{"from_length": e[0], "to_length": e[1], "sequence": e[2].encode()} for e in m["edits"]]}
                                 for m in a["mappings"]]})
    return out


def toy_graph():
    nodes, paths = read_gfa(os.path.join(GOLDEN, "target_graph.gfa"))
    return nodes, paths


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def graph_from_alignments(alns, n_nodes):
    """Node sequences implied by the match edits of a set of alignments (nodes never matched stay 'N').
    Lets the reference's real-graph GAM fixtures run against synthetic tables."""
    need = {}
    for a in alns:
        for m in a["mappings"]:
            tot = m["offset"] + sum(e["from_length"] for e in m["edits"])
            need[m["node_id"]] = max(need.get(m["node_id"], 1), tot)
    seqs = {n: ["N"] * need.get(n, 1) for n in range(1, n_nodes + 1)}
    for a in alns:
        rp = 0
        read = a["sequence"].decode()
        for m in a["mappings"]:
            off = m["offset"]
            ln = len(seqs[m["node_id"]])
            for e in m["edits"]:
                f, t, s = e["from_length"], e["to_length"], e["sequence"]
                if f == t and not s:
                    for k in range(f):
                        b = read[rp + k]
                        if m["is_reverse"]:
                            seqs[m["node_id"]][ln - 1 - (off + k)] = revcomp(b)
                        else:
                            seqs[m["node_id"]][off + k] = b
                off += f
                rp += t
    return ["".join(seqs[n]) for n in range(1, n_nodes + 1)]


def oracle_side(S, alns_arrays, node_seqs=None, bep=1e-4, is_fasta=False):
    """(tables, graph, reads) for the oracle from a vgan_b200.SynthHaplocart and flat alignment arrays."""
    node, val = S.pmap()
    T = orc.HcTables(S.supports(), node, val, S.mappab(), bep=bep, is_fasta=is_fasta, min_id=1)
    G = orc.Graph(S.node_sequences() if node_seqs is None else node_seqs, min_id=1)
    R = orc.FlatReads(alns_arrays)
    return T, G, R


def rel_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    denom = np.where(want == 0, 1.0, np.abs(want))
    return float(np.max(np.abs(got - want) / denom)) if want.size else 0.0


# ---- golden vectors produced by the reference's own code (tests/golden/make_ref_golden.py) ----
def ld_pairs(x):
    """long double array -> [[hi, lo], ...] of doubles, exact for the x87 64-bit mantissa."""
    x = np.asarray(x, dtype=np.longdouble)
    hi = x.astype(np.float64)
    lo = (x - hi.astype(np.longdouble)).astype(np.float64)
    return [[float(a), float(b)] for a, b in zip(hi, lo)]


def ld_from_pairs(pairs):
    p = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
    return p[:, 0].astype(np.longdouble) + p[:, 1].astype(np.longdouble)


def toy_tables(bep=1e-4, is_fasta=False):
    """HaploCart tables for the reference's toy graph: membership from the GFA P lines, coordinates cumulative."""
    nodes, paths = toy_graph()
    n_rows = max(nodes) + 1
    sup = np.zeros((n_rows, len(paths)), dtype=np.uint8)
    for p, (_, steps) in enumerate(paths):
        for nid, _ in steps:
            sup[nid, p] = 1
    ids = sorted(nodes)
    coord = np.cumsum([0] + [len(nodes[i]) for i in ids])[:-1]
    mp = np.linspace(0.3, 1.0, int(coord[-1]) + 40)
    T = orc.HcTables(sup, np.array(ids), (coord + 1).astype(np.int32), mp, bep=bep, is_fasta=is_fasta, min_id=min(ids))
    names = [n for n, _ in paths]
    parents = "".join("%s root\n" % n for n in names)
    children = "root " + " ".join(names) + "\n" + "".join("%s\n" % n for n in names)
    return T, orc.Graph(nodes), (names, parents, children)


def ref_golden_inputs(case):
    """Rebuild the inputs of one case of tests/golden/ref_hc_golden.json from its descriptor (seeds / fixture names)."""
    import vgan_b200 as vb
    from oracle.gam_py import read_gam
    kind = case["kind"]
    bep, fasta = case["bep"], case["fasta"]
    if kind == "toy":
        T, G, tree = toy_tables(bep=bep, is_fasta=fasta)
        return T, G, orc.Reads(read_gam(os.path.join(GOLDEN, case["fixture"]))), tree
    S = vb.SynthHaplocart(**case["synth"])
    tree = (S.path_names, S.parents_txt, S.children_txt)
    if kind == "synth":
        alns = S.reads(case["n"], seed=case["read_seed"], **case["read_kw"])
        T, G, R = oracle_side(S, alns.arrays(), bep=bep, is_fasta=fasta)
    elif kind == "edge":
        from test_ref_pin import _defined_edge_reads
        alns = vb.Alignments.from_gam_bytes(encode_gam(_defined_edge_reads(S)))
        T, G, R = oracle_side(S, alns.arrays(), bep=bep, is_fasta=fasta)
    elif kind == "gamfix":
        alns_d = decoded_fixture(case["fixture"])[:case["n"]]
        node_seqs = graph_from_alignments(alns_d, S.n_nodes)
        alns = vb.Alignments.from_gam_bytes(encode_gam(alns_d))
        T, G, R = oracle_side(S, alns.arrays(), node_seqs=node_seqs, bep=bep, is_fasta=fasta)
    else:
        raise ValueError(kind)
    R._keep = alns
    return T, G, R, tree
