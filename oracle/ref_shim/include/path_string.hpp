// TEST INFRASTRUCTURE ONLY -- shim for vg::algorithms::path_string (vg src/algorithms/... path_string.hpp).
// Published behaviour: the READ as spelled by the path -- per edit, a match contributes the node's own
// bases from the running offset (orientation aware), a substitution or insertion contributes
// edit.sequence, a deletion nothing.  Pinned through the reference's own 10
// reconstruction KATs, which call this via reconstruct_graph_sequence (src/vgan_utils.h:18).
#pragma once
#include "vg/vg.pb.h"
#include "vg/io/edit.hpp"
#include "bdsg/odgi.hpp"
namespace vg { namespace algorithms {
template <class G>
inline std::string path_string(const G& graph, const vg::Path& path) {
    std::string seq;
    for (int i = 0; i < path.mapping_size(); ++i) {
        const vg::Mapping& m = path.mapping(i);
        const std::string node_seq = graph.get_sequence(graph.get_handle(m.position().node_id(), m.position().is_reverse()));
        size_t f = (size_t)m.position().offset();
        for (int j = 0; j < m.edit_size(); ++j) {
            const vg::Edit& e = m.edit(j);
            if (vg::io::edit_is_match(e)) seq.append(node_seq.substr(f, (size_t)e.from_length()));
            else if (vg::io::edit_is_sub(e) || vg::io::edit_is_insertion(e)) seq.append(e.sequence());
            f += (size_t)e.from_length();
        }
    }
    return seq;
}
}}
