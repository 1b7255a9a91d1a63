// vgan_haplocart.hpp -- the C++ facade a vgan maintainer includes from src/HaploCart.cpp.
//
// Header-only glue between the reference's own data structures and the two C ABIs
// (include/vgan_host.h, include/vgan_gpu.h).  It mirrors the three call sites the B200 path replaces
// (file:line relative to the reference):
//
//   reference                                                          here
//   -----------------------------------------------------------------  ------------------------------------
//   tables loaded once            src/HaploCart.cpp:279-286            HaplocartGpu::load_tables / load_tree
//   omp loop: update() + critical sum  src/HaploCart.cpp:408-421      HaplocartGpu::update(algnvector)
//   max_element                   src/HaploCart.cpp:423-424            HaplocartGpu::get_posterior -> maxh_index
//   Haplocart::get_posterior      src/get_posterior.cpp:87-127         HaplocartGpu::get_posterior -> clade_vec,
//                                                                       confidence_vec (feed write_posterior_log)
//
// Error behaviour follows the reference: its hot-path functions are `noexcept` and everything else throws
// std::runtime_error("[HaploCart] Error, ...") (src/HaploCart.cpp:110,387); every non-zero status from the
// C ABI becomes such an exception here.  No exception ever crosses the C ABI itself.
//
// The alignment type is a template parameter so that this header compiles both against vgan's
// AlignmentInfo (src/AlignmentInfo.h:16-46, whose `path` is a vg::Path protobuf) and against plain structs
// in tests: it needs `.seq`, `.quality_scores`, `.mapping_quality`, `.identity`, `.is_paired` and
// `.path.mapping()` yielding objects with `.position().node_id() / .offset() / .is_reverse()` and `.edit()`
// yielding objects with `.from_length() / .to_length() / .sequence()`.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/vgan_gpu.h"
#include "../../../include/vgan_host.h"

namespace vgan_b200 {

inline void check(int rc, const char* what) {
    if (rc == 0) return;
    const char* msg = rc >= 100 ? vgh_last_error() : vgl_last_error();
    throw std::runtime_error(std::string("[HaploCart] Error, ") + what + ": " + msg);
}

// Flat copy of a vector<AlignmentInfo*> in the layout vgh_alignments describes.
struct FlatAlignments {
    std::vector<int64_t> seq_off{0}, qual_off{0}, map_off{0}, edit_off{0}, edit_seq_off{0};
    std::string seq, qual, edit_seq;
    std::vector<int32_t> mapq, edit_from, edit_to;
    std::vector<double> identity;
    std::vector<uint8_t> paired, map_rev;
    std::vector<int64_t> map_node, map_offset;

    template <typename AlignmentInfoPtrVector>
    void assign(const AlignmentInfoPtrVector& algnvector) {
        *this = FlatAlignments();
        for (const auto* ai : algnvector) {
            seq += ai->seq;
            seq_off.push_back((int64_t)seq.size());
            qual += ai->quality_scores;
            qual_off.push_back((int64_t)qual.size());
            mapq.push_back((int32_t)ai->mapping_quality);
            identity.push_back((double)ai->identity);
            paired.push_back(ai->is_paired ? 1 : 0);
            for (const auto& m : ai->path.mapping()) {
                map_node.push_back((int64_t)m.position().node_id());
                map_offset.push_back((int64_t)m.position().offset());
                map_rev.push_back(m.position().is_reverse() ? 1 : 0);
                for (const auto& e : m.edit()) {
                    edit_from.push_back((int32_t)e.from_length());
                    edit_to.push_back((int32_t)e.to_length());
                    edit_seq += e.sequence();
                    edit_seq_off.push_back((int64_t)edit_seq.size());
                }
                edit_off.push_back((int64_t)edit_from.size());
            }
            map_off.push_back((int64_t)map_node.size());
        }
    }

    vgh_alignments view() const {
        vgh_alignments v{};
        v.n_reads = (int64_t)mapq.size();
        v.seq_off = seq_off.data(); v.seq = seq.data();
        v.qual_off = qual_off.data(); v.qual = qual.data();
        v.mapq = mapq.data(); v.identity = identity.data(); v.paired = paired.data();
        v.map_off = map_off.data(); v.map_node = map_node.data(); v.map_offset = map_offset.data(); v.map_rev = map_rev.data();
        v.edit_off = edit_off.data(); v.edit_from = edit_from.data(); v.edit_to = edit_to.data();
        v.edit_seq_off = edit_seq_off.data(); v.edit_seq = edit_seq.data();
        return v;
    }
};

struct Posterior {
    int maxh_index = -1;                  // std::max_element index (first maximum)
    double total_ll = 0.0;                // log-sum-exp over all paths
    std::vector<double> final_vec;        // per-path log-likelihoods (empty unless asked for)
    std::vector<int32_t> clade_ids;       // clades listed for the call (indices into the tree's clade table)
    std::vector<double> confidence_vec;   // [0] the call, [1 + k] its k-th listed clade
};

class HaplocartGpu {
public:
    // Must be constructed AFTER the process has fork()ed its mapping children (src/HaploCart.cpp:313-378): the
    // CUDA context is created here.
    explicit HaplocartGpu(int device = 0) { check(vgl_create(device, &ctx_), "cannot create the GPU context"); }
    ~HaplocartGpu() {
        if (feeder_) vgh_hc_feeder_destroy(feeder_);
        if (tree_) vgh_hc_tree_destroy(tree_);
        if (tables_) vgh_hc_tables_destroy(tables_);
        if (graph_) vgh_graph_destroy(graph_);
        if (ctx_) vgl_destroy(ctx_);
    }
    HaplocartGpu(const HaplocartGpu&) = delete;
    HaplocartGpu& operator=(const HaplocartGpu&) = delete;

    // Node sequences of the graph (what reconstruct_graph_sequence asks of bdsg::ODGI::get_sequence).
    // node_seqs[k] is the forward sequence of node min_id + k.
    void load_graph(int64_t min_id, const std::vector<std::string>& node_seqs) {
        std::vector<int64_t> off{0};
        std::string blob;
        for (const auto& s : node_seqs) { blob += s; off.push_back((int64_t)blob.size()); }
        if (graph_) vgh_graph_destroy(graph_);
        graph_ = nullptr;
        check(vgh_graph_create(min_id, (int64_t)node_seqs.size(), off.data(), blob.data(), &graph_), "cannot build the node table");
        check(vgl_hc_set_graph(ctx_, min_id, (int64_t)node_seqs.size(), off.data(), blob.data()), "cannot upload the node table");
    }

    // The tables exactly as the reference holds them: NodeInfo::pathsgo rows (row index == node id),
    // pangenome_map entries (value as stored: file coordinate + 1), mappabilities, -e, FASTA mode.
    void load_tables(int32_t n_rows, int32_t n_paths, const uint8_t* pathsgo_rows, const std::vector<int64_t>& pmap_node,
                     const std::vector<int32_t>& pmap_val, const std::vector<double>& mappabilities,
                     double background_error_prob, bool is_consensus_fasta) {
        if (tables_) vgh_hc_tables_destroy(tables_);
        tables_ = nullptr;
        check(vgh_hc_tables_build(n_rows, n_paths, pathsgo_rows, (int64_t)pmap_node.size(), pmap_node.data(), pmap_val.data(),
                                  (int64_t)mappabilities.size(), mappabilities.data(), background_error_prob,
                                  is_consensus_fasta ? 1 : 0, &tables_),
              "cannot build the HaploCart tables");
        check(vgl_hc_set_tables(ctx_, vgh_hc_tables_view(tables_)), "cannot upload the HaploCart tables");
        n_paths_ = n_paths;
    }

    // parents.txt / children.txt contents and the path names in column order (src/load.cpp:44-58,303-345).
    void load_tree(const std::vector<std::string>& path_names, const std::string& parents_txt, const std::string& children_txt) {
        std::string names;
        for (const auto& n : path_names) { names += n; names += '\n'; }
        if (tree_) vgh_hc_tree_destroy(tree_);
        tree_ = nullptr;
        check(vgh_hc_tree_build((int32_t)path_names.size(), names.c_str(), parents_txt.c_str(), children_txt.c_str(), &tree_),
              "cannot build the haplogroup tree");
        check(vgl_hc_set_tree(ctx_, vgh_hc_tree_n_clades(tree_), vgh_hc_tree_clade_bits(tree_), vgh_hc_tree_anc_off(tree_),
                              vgh_hc_tree_anc_clade(tree_)),
              "cannot upload the haplogroup tree");
        max_post_ = 1;
        const int32_t* off = vgh_hc_tree_anc_off(tree_);
        for (size_t p = 0; p < path_names.size(); ++p) max_post_ = std::max(max_post_, 1 + off[p + 1] - off[p]);
    }

    // fill(final_vec, 0): start of a sample (src/HaploCart.cpp:481).
    void reset() { check(vgl_hc_reset(ctx_), "reset"); }

    // The whole `#pragma omp parallel for` of src/HaploCart.cpp:408-421 for one vector of alignments; may be called
    // several times per sample.  n_threads only drives the host packer (the reference's -t).
    template <typename AlignmentInfoPtrVector>
    void update(const AlignmentInfoPtrVector& algnvector, int n_threads = 0) {
        flat_.assign(algnvector);
        const vgh_alignments v = flat_.view();
        vgh_hc_packed* packed = nullptr;
        check(vgh_hc_pack(graph_, &v, 0, v.n_reads, n_threads, /*pinned=*/1, &packed), "cannot pack the alignments");
        const int rc = vgl_hc_accumulate(ctx_, vgh_hc_packed_batch(packed));
        vgh_hc_packed_destroy(packed);
        check(rc, "likelihood kernel");
    }

    // Same, with reconstruct_graph_sequence and the packing done on the GPU: the host only copies the alignments' fields
    // into the compact raw arrays (one linear pass, no per-read string building).
    template <typename AlignmentInfoPtrVector>
    void update_on_device(const AlignmentInfoPtrVector& algnvector) {
        flat_.assign(algnvector);
        const vgh_alignments v = flat_.view();
        vgh_hc_raw* raw = nullptr;
        check(vgh_hc_raw_build(&v, 0, v.n_reads, /*pinned=*/1, &raw), "cannot marshal the alignments");
        const int rc = vgl_hc_accumulate_raw(ctx_, vgh_hc_raw_view(raw));  // returns once the host arrays have been copied
        vgh_hc_raw_destroy(raw);
        check(rc, "likelihood kernel");
    }

    // Same again, and the route to prefer: the reads travel in chunks, `n_threads` host threads marshal one chunk into
    // page-locked memory while the GPU copies, reconstructs and scores the chunk before (vgh_hc_feed).  The two blocks live as
    // long as this object: no allocation per sample.
    template <typename AlignmentInfoPtrVector>
    void update_streamed(const AlignmentInfoPtrVector& algnvector, int n_threads = 0) {
        flat_.assign(algnvector);
        const vgh_alignments v = flat_.view();
        if (!feeder_) check(vgh_hc_feeder_create(n_threads, 0, &feeder_), "cannot create the feeder");
        check(vgh_hc_feed(feeder_, ctx_, &v, 0, v.n_reads), "likelihood kernel");
    }

    // Dup_Remover::remove_duplicates_internal (src/rmdup.cpp:68-110, called at src/HaploCart.cpp:390) on the device:
    // is_dup[i] takes the place of the reference's vector<bool>.
    template <typename AlignmentInfoPtrVector>
    std::vector<uint8_t> duplicates(const AlignmentInfoPtrVector& algnvector) {
        flat_.assign(algnvector);
        const vgh_alignments v = flat_.view();
        std::vector<uint64_t> keys((size_t)v.n_reads);
        std::vector<uint8_t> is_dup((size_t)v.n_reads);
        check(vgh_dedup_keys(&v, keys.data()), "duplicate removal");
        uint64_t kept = 0;
        check(vgl_dedup(ctx_, (uint32_t)v.n_reads, keys.data(), v.paired, is_dup.data(), &kept), "duplicate removal");
        return is_dup;
    }

    // Several samples per pass (the reference loops over samples, src/HaploCart.cpp:303): begin_samples(S), then for every
    // sample select_sample(s) + update*(...), then finish_samples() -- one collapse, one posterior kernel, one copy back.
    void begin_samples(int n_samples) {
        check(vgl_hc_samples_begin(ctx_, n_samples), "begin_samples");
        n_samples_ = n_samples;
    }
    void select_sample(int s) { check(vgl_hc_sample_select(ctx_, s), "select_sample"); }
    std::vector<Posterior> finish_samples(bool want_final_vec = false) {
        const size_t S = (size_t)n_samples_;
        std::vector<double> ll(want_final_vec ? S * (size_t)n_paths_ : 0), total(S), post(S * (size_t)max_post_);
        std::vector<int32_t> arg(S), n_post(S);
        check(vgl_hc_finalize_samples(ctx_, want_final_vec ? ll.data() : nullptr, arg.data(), total.data(), post.data(), max_post_, n_post.data(),
                                      nullptr),
              "posterior kernel");
        std::vector<Posterior> out(S);
        for (size_t k = 0; k < S; ++k) {
            Posterior& r = out[k];
            r.maxh_index = arg[k];
            r.total_ll = total[k];
            if (want_final_vec) r.final_vec.assign(ll.begin() + (long)(k * (size_t)n_paths_), ll.begin() + (long)((k + 1) * (size_t)n_paths_));
            r.confidence_vec.assign(post.begin() + (long)(k * (size_t)max_post_), post.begin() + (long)(k * (size_t)max_post_) + n_post[k]);
            if (tree_) {
                const int32_t* off = vgh_hc_tree_anc_off(tree_);
                const int32_t* cl = vgh_hc_tree_anc_clade(tree_);
                r.clade_ids.assign(cl + off[arg[k]], cl + off[arg[k] + 1]);
            }
        }
        return out;
    }

    // get_posterior in two halves, for a loop over samples that keeps the GPU busy: begin_posterior() enqueues the kernels and the
    // copy of the result and returns a ticket; the next sample's reset() / update*() may follow at once; end_posterior(ticket)
    // waits for that result alone.  At most two in flight.
    int begin_posterior() {
        int32_t t = -1;
        check(vgl_hc_finalize_begin(ctx_, &t), "begin_posterior");
        return t;
    }
    Posterior end_posterior(int ticket) {
        Posterior p;
        std::vector<double> post((size_t)max_post_ > 0 ? (size_t)max_post_ : 1);
        int32_t n_post = 0;
        check(vgl_hc_finalize_end(ctx_, ticket, &p.maxh_index, &p.total_ll, post.data(), (int32_t)post.size(), &n_post), "end_posterior");
        p.confidence_vec.assign(post.begin(), post.begin() + n_post);
        if (tree_) {
            const int32_t* off = vgh_hc_tree_anc_off(tree_);
            const int32_t* cl = vgh_hc_tree_anc_clade(tree_);
            p.clade_ids.assign(cl + off[p.maxh_index], cl + off[p.maxh_index + 1]);
        }
        return p;
    }

    // max_element + Haplocart::get_posterior.
    Posterior get_posterior(bool want_final_vec = false) {
        Posterior r;
        if (want_final_vec) r.final_vec.resize((size_t)n_paths_);
        std::vector<double> post((size_t)max_post_);
        int32_t arg = -1, n_post = 0;
        check(vgl_hc_finalize(ctx_, want_final_vec ? r.final_vec.data() : nullptr, &arg, &r.total_ll, post.data(), max_post_, &n_post),
              "posterior kernel");
        r.maxh_index = arg;
        r.confidence_vec.assign(post.begin(), post.begin() + n_post);
        if (tree_) {
            const int32_t* off = vgh_hc_tree_anc_off(tree_);
            const int32_t* cl = vgh_hc_tree_anc_clade(tree_);
            r.clade_ids.assign(cl + off[arg], cl + off[arg + 1]);
        }
        return r;
    }
    const char* clade_name(int32_t clade) const { return vgh_hc_tree_clade_name(tree_, clade); }
    vgl_ctx* context() { return ctx_; }

private:
    vgl_ctx* ctx_ = nullptr;
    vgh_graph* graph_ = nullptr;
    vgh_hc_tables* tables_ = nullptr;
    vgh_hc_tree* tree_ = nullptr;
    vgh_hc_feeder* feeder_ = nullptr;
    FlatAlignments flat_;
    int32_t n_paths_ = 0, max_post_ = 1, n_samples_ = 0;
};

}  // namespace vgan_b200
