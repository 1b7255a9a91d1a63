"""Writes tests/golden/ref_hc_golden.json: log-likelihood vectors, calls and posteriors computed by the
REFERENCE'S OWN CODE (oracle/_ref/libvgan_ref.so = /root/reference/src/{update_likelihood,process_mapping,
get_p_obs_base,get_posterior,...}.cpp compiled unchanged over oracle/ref_shim/).  Run in the authoring container
only -- /root/reference does not exist on the GPU box, which is why the vectors are committed.

Inputs are not stored: every case is a descriptor (fixture name or generator seeds) that tests/helpers.py
(`ref_golden_inputs`) turns back into the same alignments and tables.  long double values are stored as exact
[hi, lo] pairs of doubles.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

import helpers as H  # noqa: E402
from oracle import ref  # noqa: E402

SMALL = dict(seed=101, n_paths=300, n_rows=1210, n_nodes=1201)
REAL_IDS = dict(seed=7, n_paths=64, n_rows=11825, n_nodes=11821)
CASES = [
    dict(name="toy_test_reads", kind="toy", fixture="test_reads.gam", bep=1e-4, fasta=False),
    dict(name="toy_test_reads_fasta", kind="toy", fixture="test_reads.gam", bep=1e-3, fasta=True),
    dict(name="toy_reverse", kind="toy", fixture="all_the_same_reverse.gam", bep=1e-4, fasta=False),
    dict(name="synth_plain", kind="synth", synth=SMALL, n=150, read_seed=202, read_kw=dict(sub_rate=0.005, indel_rate=0.0005),
         bep=1e-4, fasta=False),
    dict(name="synth_noisy", kind="synth", synth=SMALL, n=100, read_seed=203, read_kw=dict(sub_rate=0.03, indel_rate=0.02),
         bep=1e-4, fasta=False),
    dict(name="synth_damaged_variable_length", kind="synth", synth=SMALL, n=100, read_seed=204,
         read_kw=dict(sub_rate=0.01, indel_rate=0.002, damage=True, fixed_len=0), bep=1e-4, fasta=False),
    dict(name="synth_fasta", kind="synth", synth=SMALL, n=60, read_seed=205, read_kw=dict(sub_rate=0.01), bep=3e-3, fasta=True),
    dict(name="edge_reads", kind="edge", synth=SMALL, bep=2e-3, fasta=False),
    dict(name="gam_J2a1a1a1", kind="gamfix", synth=REAL_IDS, fixture="J2a1a1a1.gam", n=24, bep=1e-4, fasta=False),
    dict(name="gam_two_unique", kind="gamfix", synth=REAL_IDS, fixture="two_unique.gam", n=16, bep=1e-4, fasta=False),
]


def main():
    out = []
    for case in CASES:
        T, G, R, tree = H.ref_golden_inputs(case)
        final, argmax, used = ref.hc_run(T, G, R)
        c = dict(case)
        c["final_vec"] = H.ld_pairs(final)
        c["argmax"], c["used"] = int(argmax), int(used)
        pick = sorted(set([0, R.n // 2, R.n - 1]))
        c["per_read"] = {str(r): H.ld_pairs(ref.hc_read(T, G, R, r)) for r in pick}
        names, parents, children = tree
        clades, conf_text, conf_full, total = ref.hc_posterior(final, names, parents, children, argmax)
        c["clades"] = clades
        c["confidence_text"] = None if conf_text is None else [float(x) for x in conf_text]
        c["confidence_full"] = [float(x) for x in conf_full]
        c["total_ll"] = H.ld_pairs([total])[0]
        out.append(c)
        print(case["name"], "reads", R.n, "argmax", argmax, names[argmax], "conf", conf_full[:3])
    with open(os.path.join(HERE, "ref_hc_golden.json"), "w") as fh:
        json.dump({"source": "reference HaploCart hot path compiled unchanged (oracle/ref_shim/Makefile); "
                             "quality strings zero-padded so the out-of-range reads of update_likelihood.cpp:40-44 are defined",
                   "cases": out}, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()


def siblings():
    """euka per-read outputs and one soibean R x P matrix from the reference's own readGAM3 lambda / analyse_GAM."""
    import vgan_b200 as vb
    from oracle import oracle as orc
    from test_ref_pin import _euka_inputs
    S, og, R, _keep = _euka_inputs()
    e = ref.euka_run(S.dist, S.bin_off, S.bin_lo, S.bin_hi, og, R, vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    euka = {"clade": e["clade"].tolist(), "like": e["like"].tolist(), "not_like": e["not_like"].tolist(),
            "clade_count": e["clade_count"].tolist(), "bin_cov": e["bin_cov"].tolist()}
    seed, P, n = 112, 31, 120
    S2 = vb.SynthHaplocart(seed=seed, n_paths=P, n_rows=610, n_nodes=601)
    alns = S2.reads(n, seed=seed + 1, fixed_len=0, sub_rate=0.02, indel_rate=0.01, damage=True)
    mat, used = ref.soibean_run(orc.Graph(S2.node_sequences(), min_id=1), orc.FlatReads(alns.arrays()), S2.supports(), 0, 7,
                                vb.DHIGH5P_PROF, vb.DHIGH3P_PROF)
    with open(os.path.join(HERE, "ref_sibling_golden.json"), "w") as fh:
        json.dump({"source": "reference readGAM3 lambda (src/readGAM_Euka.h) and analyse_GAM (src/getLCAfromGAM.h) compiled unchanged "
                             "(oracle/ref_shim/Makefile); quality strings zero-padded",
                   "euka": euka, "soibean": {"seed": seed, "P": P, "n": n, "matrix": mat.tolist(), "used": used.tolist()}}, fh,
                  separators=(",", ":"))


if __name__ == "__main__":
    siblings()
