// TEST INFRASTRUCTURE ONLY -- shim for bdsg::ODGI: node sequences by id, orientation aware
// (get_sequence of a reverse handle is the reverse complement, as libhandlegraph specifies;
// pinned through the reference's reconstruction KATs on both strands, src/test.cpp:855-994).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <stdexcept>
#include <functional>
namespace handlegraph {
struct handle_t { int64_t id; bool rev; };
struct path_handle_t { int64_t id; };
struct step_handle_t { int64_t a, b; };
typedef int64_t nid_t;
class HandleGraph { public: virtual ~HandleGraph() {} };
class PathHandleGraph : public HandleGraph {};
}
namespace bdsg {
using handlegraph::handle_t;
using handlegraph::path_handle_t;
using handlegraph::step_handle_t;
class ODGI : public handlegraph::PathHandleGraph {
    int64_t min_id_ = 1;
    std::vector<std::string> seq_;
    std::vector<std::string> path_names_;
    static char comp(char c) {
        switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                     case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return c; }
    }
public:
    void shim_set(int64_t min_id, std::vector<std::string> seqs) { min_id_ = min_id; seq_ = std::move(seqs); }
    bool has_node(int64_t id) const { return id >= min_id_ && id < min_id_ + (int64_t)seq_.size(); }
    handle_t get_handle(int64_t id, bool is_reverse = false) const {
        if (!has_node(id)) throw std::runtime_error("ref_shim ODGI: no such node " + std::to_string(id));
        return handle_t{id, is_reverse};
    }
    int64_t get_id(const handle_t& h) const { return h.id; }
    bool get_is_reverse(const handle_t& h) const { return h.rev; }
    size_t get_length(const handle_t& h) const { return seq_[(size_t)(h.id - min_id_)].size(); }
    std::string get_sequence(const handle_t& h) const {
        const std::string& s = seq_[(size_t)(h.id - min_id_)];
        if (!h.rev) return s;
        std::string r(s.rbegin(), s.rend());
        for (auto& c : r) c = comp(c);
        return r;
    }
    int64_t min_node_id() const { return min_id_; }
    int64_t max_node_id() const { return min_id_ + (int64_t)seq_.size() - 1; }
    size_t get_node_count() const { return seq_.size(); }
    size_t get_path_count() const { return path_names_.size(); }
    void deserialize(std::istream&) { throw std::runtime_error("ref_shim ODGI: deserialize not available"); }
    template <class F> bool for_each_handle(const F&, bool = false) const { return true; }
    template <class F> bool for_each_path_handle(const F&) const { return true; }
    template <class F> bool for_each_step_on_handle(const handle_t&, const F&) const { return true; }
    template <class F> bool for_each_step_in_path(const path_handle_t&, const F&) const { return true; }
    std::string get_path_name(const path_handle_t&) const { return ""; }
    path_handle_t get_path_handle_of_step(const step_handle_t&) const { return path_handle_t{0}; }
    handle_t get_handle_of_step(const step_handle_t&) const { return handle_t{0, false}; }
};
}
