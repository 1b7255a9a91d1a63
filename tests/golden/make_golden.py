"""Regenerates tests/golden/ from the reference checkout (run in the authoring container only:
/root/reference does not exist on the GPU box, which is why the outputs are committed).

  * copies the reference's own GAM / GFA test inputs (binary test data, not source code);
  * extracts the 10 `reconstruction` known-answer strings from the reference's test suite
    (src/test.cpp:855-994) into reconstruct_kat.json;
  * decodes each GAM fixture with the ORACLE-side Python reader into *.decoded.json so that the
    C++ reader of the product can be compared against it without the reference tree.
"""
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.gam_py import read_gam  # noqa: E402

FIXTURES = [
    "test/reconstructInputSeq/test_reads.gam",
    "test/reconstructInputSeq/target_graph.gfa",
    "test/input_files/alignments/J2a1a1a1.gam",
    "test/input_files/alignments/all_the_same.gam",
    "test/input_files/alignments/all_the_same_reverse.gam",
    "test/input_files/alignments/two_unique.gam",
]
# data tables the synthetic configs are specified on: damage profiles and the euka clade / bin lists
DATA = [
    ("share/vgan/damageProfiles/dhigh5p.prof", "damageProfiles/dhigh5p.prof"),
    ("share/vgan/damageProfiles/dhigh3p.prof", "damageProfiles/dhigh3p.prof"),
    ("share/vgan/damageProfiles/none.prof", "damageProfiles/none.prof"),
    ("share/vgan/euka_dir/euka_db.clade", "euka_db.clade"),
    ("share/vgan/euka_dir/euka_db.bins", "euka_db.bins"),
]


def main():
    for rel in FIXTURES:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(HERE, os.path.basename(rel)))
        os.chmod(os.path.join(HERE, os.path.basename(rel)), 0o644)

    for rel, dst in DATA:
        os.makedirs(os.path.dirname(os.path.join(HERE, dst)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)

    src = open(os.path.join(REF, "src/test.cpp")).read()
    suite = src[src.index("BOOST_AUTO_TEST_SUITE(reconstruction)"):]
    suite = suite[:suite.index("BOOST_AUTO_TEST_SUITE_END()")]
    kats = []
    for case in re.split(r"BOOST_AUTO_TEST_CASE\(", suite)[1:]:
        name = case[:case.index(")")]
        idx = int(re.search(r"read_info->at\((\d+)\)->path", case).group(1))
        read_seq = re.search(r'BOOST_CHECK_EQUAL\(read_seq,\s*"([^"]*)"\)', case).group(1)
        graph_seq = re.search(r'BOOST_CHECK_EQUAL\(graph_seq,\s*"([^"]*)"\)', case).group(1)
        kats.append({"case": name, "read_index": idx, "graph_seq": graph_seq, "read_seq": read_seq})
    assert len(kats) == 10
    with open(os.path.join(HERE, "reconstruct_kat.json"), "w") as fh:
        json.dump({"source": "reference src/test.cpp:855-994 (suite `reconstruction`)", "cases": kats}, fh, indent=1)

    for rel in FIXTURES:
        if not rel.endswith(".gam"):
            continue
        alns = read_gam(os.path.join(REF, rel))
        out = []
        for a in alns:
            out.append({
                "sequence": a["sequence"].decode(), "quality": list(a["quality"]), "mapping_quality": a["mapping_quality"],
                "identity": a["identity"], "read_paired": a["read_paired"],
                "mappings": [{"node_id": m["node_id"], "offset": m["offset"], "is_reverse": m["is_reverse"],
                              "edits": [[e["from_length"], e["to_length"], e["sequence"].decode()] for e in m["edits"]]}
                             for m in a["mappings"]],
            })
        with open(os.path.join(HERE, os.path.basename(rel) + ".decoded.json"), "w") as fh:
            json.dump(out, fh, separators=(",", ":"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
