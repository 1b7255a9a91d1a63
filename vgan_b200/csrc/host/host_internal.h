// Internal definitions shared by the translation units of libvgan_host.so.  Not installed.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/vgan_host.h"

int vgh_fail(int code, const char* fmt, ...);
const char* vgh_error_text();

// ------------------------------------------------------------------------------------------------
// graph
// ------------------------------------------------------------------------------------------------
struct vgh_graph {
    int64_t min_id = 0;
    int64_t n_nodes = 0;
    std::vector<int64_t> off;
    std::string seq;

    bool has(int64_t id) const { return id >= min_id && id < min_id + n_nodes; }
    int64_t length(int64_t id) const { return off[(size_t)(id - min_id) + 1] - off[(size_t)(id - min_id)]; }
    // base at oriented offset `k` of node `id` (what get_sequence(get_handle(id, rev))[k] returns)
    char base(int64_t id, bool rev, int64_t k) const {
        const int64_t a = off[(size_t)(id - min_id)], n = length(id);
        if (!rev) return seq[(size_t)(a + k)];
        const char c = seq[(size_t)(a + n - 1 - k)];
        switch (c) {
            case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
            case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
            default: return c;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// GAM reader
// ------------------------------------------------------------------------------------------------
struct vgh_gam {
    std::vector<int64_t> seq_off{0}, qual_off{0}, map_off{0}, edit_off{0}, edit_seq_off{0};
    std::string seq, qual, edit_seq;
    std::vector<int32_t> mapq, edit_from, edit_to;
    std::vector<double> identity;
    std::vector<uint8_t> paired, map_rev;
    std::vector<int64_t> map_node, map_offset;
    vgh_alignments view{};

    void finish() {
        // keep every pointer addressable even when empty
        if (mapq.empty()) mapq.reserve(1);
        view.n_reads = (int64_t)mapq.size();
        view.seq_off = seq_off.data(); view.seq = seq.data();
        view.qual_off = qual_off.data(); view.qual = qual.data();
        view.mapq = mapq.data(); view.identity = identity.data(); view.paired = paired.data();
        view.map_off = map_off.data(); view.map_node = map_node.data(); view.map_offset = map_offset.data();
        view.map_rev = map_rev.data();
        view.edit_off = edit_off.data(); view.edit_from = edit_from.data(); view.edit_to = edit_to.data();
        view.edit_seq_off = edit_seq_off.data(); view.edit_seq = edit_seq.data();
    }
};


// reconstruct_graph_sequence (reference src/vgan_utils.h:6-79) for read r: aligned graph string, aligned read string,
// one size per edit.  Defined in vgan_host.cpp.
struct Recon {
    std::string graph_seq, read_seq;
    std::vector<int32_t> sizes;
};
int vgh_recon(const vgh_graph& G, const vgh_alignments& a, int64_t r, Recon& out);
