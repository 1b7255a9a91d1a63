// A stand-in for the patched region of src/HaploCart.cpp: builds vgan-shaped inputs (AlignmentInfo-like structs whose
// `path` mimics the vg::Path accessors), drives vgan_b200::HaplocartGpu exactly as INTEGRATION.md shows, and prints the
// call and its clade posteriors.  tests/test_cpp_facade.py compares the output with the Python path and the oracle.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../vgan_b200/csrc/host/vgan_haplocart.hpp"

// ---- minimal mimics of the protobuf accessors the facade touches ----
struct Position { int64_t node_id_, offset_; bool rev_; int64_t node_id() const { return node_id_; } int64_t offset() const { return offset_; } bool is_reverse() const { return rev_; } };
struct Edit { int32_t f_, t_; std::string s_; int32_t from_length() const { return f_; } int32_t to_length() const { return t_; } const std::string& sequence() const { return s_; } };
struct Mapping { Position p_; std::vector<Edit> e_; const Position& position() const { return p_; } const std::vector<Edit>& edit() const { return e_; } };
struct Path { std::vector<Mapping> m_; const std::vector<Mapping>& mapping() const { return m_; } };
struct AlignmentInfo { std::string seq, quality_scores; Path path; int32_t mapping_quality; bool is_paired; double identity; };

int main(int argc, char** argv) {
    const int n_reads = argc > 1 ? atoi(argv[1]) : 200;
    const bool compile_only = argc > 2 && std::string(argv[2]) == "--no-gpu";
    const bool on_device = argc > 2 && std::string(argv[2]) == "--device-reconstruction";  // raw alignments, packed on the GPU
    const bool streamed = argc > 2 && std::string(argv[2]) == "--streamed";                // the feeder: chunks overlapped with the GPU
    const bool samples = argc > 2 && std::string(argv[2]) == "--samples";                  // the same sample three times in one pass
    const bool in_flight = argc > 2 && std::string(argv[2]) == "--in-flight";              // two samples in flight (begin / end_posterior)
    vgh_synth_hc* S = nullptr;
    if (vgh_synth_hc_create(101, 300, 1210, 1201, &S)) return 2;
    vgh_gam* gam = nullptr;
    if (vgh_synth_hc_reads(S, 202, n_reads, 100, 0.005, 0.0005, 0, -1, &gam)) return 2;
    const vgh_alignments* v = vgh_gam_view(gam);
    // rebuild pointer-rich AlignmentInfo objects, the form vgan's readGAM returns
    std::vector<AlignmentInfo*> algnvector;
    for (int64_t r = 0; r < v->n_reads; ++r) {
        auto* ai = new AlignmentInfo();
        ai->seq.assign(v->seq + v->seq_off[r], v->seq + v->seq_off[r + 1]);
        ai->quality_scores.assign(v->qual + v->qual_off[r], v->qual + v->qual_off[r + 1]);
        ai->mapping_quality = v->mapq[r];
        ai->identity = v->identity[r];
        ai->is_paired = false;
        for (int64_t m = v->map_off[r]; m < v->map_off[r + 1]; ++m) {
            Mapping mp{{v->map_node[m], v->map_offset[m], v->map_rev[m] != 0}, {}};
            for (int64_t e = v->edit_off[m]; e < v->edit_off[m + 1]; ++e)
                mp.e_.push_back(Edit{v->edit_from[e], v->edit_to[e], std::string(v->edit_seq + v->edit_seq_off[e], v->edit_seq + v->edit_seq_off[e + 1])});
            ai->path.m_.push_back(mp);
        }
        algnvector.push_back(ai);
    }
    vgan_b200::FlatAlignments flat;
    flat.assign(algnvector);
    if (flat.view().n_reads != v->n_reads) return 3;
    if (compile_only) { printf("flattened %lld reads\n", (long long)v->n_reads); return 0; }

    try {
        vgan_b200::HaplocartGpu hc(0);
        const int64_t* off; const char* seq;
        const int64_t n_nodes = vgh_synth_hc_graph_seq(S, &off, &seq);
        std::vector<std::string> node_seqs;
        for (int64_t k = 0; k < n_nodes; ++k) node_seqs.emplace_back(seq + off[k], seq + off[k + 1]);
        hc.load_graph(1, node_seqs);
        std::vector<int64_t> pn(vgh_synth_hc_pmap_node(S), vgh_synth_hc_pmap_node(S) + vgh_synth_hc_n_pmap(S));
        std::vector<int32_t> pv(vgh_synth_hc_pmap_val(S), vgh_synth_hc_pmap_val(S) + vgh_synth_hc_n_pmap(S));
        std::vector<double> mp(vgh_synth_hc_mappab(S), vgh_synth_hc_mappab(S) + vgh_synth_hc_n_mappab(S));
        hc.load_tables(1210, 300, vgh_synth_hc_supports(S), pn, pv, mp, 1e-4, false);
        std::vector<std::string> names;
        { std::string all = vgh_synth_hc_names(S); size_t a = 0; for (size_t k = 0; k < all.size(); ++k) if (all[k] == '\n') { names.push_back(all.substr(a, k - a)); a = k + 1; } }
        hc.load_tree(names, vgh_synth_hc_parents(S), vgh_synth_hc_children(S));
        hc.reset();
        // duplicate removal on the device: the same mask as the host pass
        {
            const std::vector<uint8_t> dup = hc.duplicates(algnvector);
            std::vector<uint8_t> host_dup(dup.size());
            int64_t kept = 0;
            const vgh_alignments fv = flat.view();
            vgan_b200::check(vgh_dedup(&fv, host_dup.data(), &kept), "host dedup");
            if (dup != host_dup) { fprintf(stderr, "device and host duplicate masks differ\n"); return 4; }
        }
        vgan_b200::Posterior post;
        if (samples) {
            hc.begin_samples(3);
            for (int k = 0; k < 3; ++k) { hc.select_sample(k); hc.update_streamed(algnvector, 2); }
            const std::vector<vgan_b200::Posterior> all = hc.finish_samples(true);
            if (all[0].maxh_index != all[2].maxh_index) { fprintf(stderr, "samples of one pass disagree\n"); return 4; }
            post = all[1];
        } else if (in_flight) {
            // sample after sample with the posterior of one still in flight while the next is accumulated
            hc.reset();
            hc.update_streamed(algnvector, 2);
            const int t0 = hc.begin_posterior();
            hc.reset();
            hc.update_streamed(algnvector, 2);
            const int t1 = hc.begin_posterior();
            const vgan_b200::Posterior first = hc.end_posterior(t0), second = hc.end_posterior(t1);
            if (first.maxh_index != second.maxh_index || first.confidence_vec.size() != second.confidence_vec.size()) {
                fprintf(stderr, "samples in flight disagree\n");
                return 4;
            }
            hc.reset();
            hc.update_streamed(algnvector, 2);
            post = hc.get_posterior(true);
            if (post.maxh_index != first.maxh_index) { fprintf(stderr, "in-flight and synchronous results disagree\n"); return 4; }
        } else {
            hc.reset();
            if (on_device) hc.update_on_device(algnvector);
            else if (streamed) hc.update_streamed(algnvector, 3);
            else hc.update(algnvector, 4);
            post = hc.get_posterior(true);
        }
        printf("call %d %s total %.17g\n", post.maxh_index, names[post.maxh_index].c_str(), post.total_ll);
        for (size_t k = 0; k < post.confidence_vec.size(); ++k)
            printf("clade %s %.17g\n", k == 0 ? names[post.maxh_index].c_str() : hc.clade_name(post.clade_ids[k - 1]), post.confidence_vec[k]);
        for (int p = 0; p < 300; ++p) printf("ll %d %.17g\n", p, post.final_vec[p]);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    for (auto* ai : algnvector) delete ai;
    vgh_gam_destroy(gam);
    vgh_synth_hc_destroy(S);
    return 0;
}
