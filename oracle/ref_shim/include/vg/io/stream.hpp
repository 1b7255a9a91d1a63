// TEST INFRASTRUCTURE ONLY -- shim for libvgio's stream iteration.  vg::io::for_each does not parse a
// file here: it replays the alignments the test driver injected (ref_shim::injected()), so that the
// reference's own readGAM / readGAM3 / analyse_GAM lambdas run unchanged over chosen inputs.
#pragma once
#include "vg/vg.pb.h"
#include "vg/io/edit.hpp"
namespace ref_shim {
inline std::vector<vg::Alignment>& injected() { static std::vector<vg::Alignment> v; return v; }
}
namespace vg { namespace io {
template <class T>
inline void for_each(std::istream&, const std::function<void(T&)>& fn) {
    for (auto& a : ref_shim::injected()) { T copy = a; fn(copy); }
}
template <class T>
inline void for_each(std::istream& in, const std::function<void(T&)>& fn, const std::function<void(void)>&) { for_each<T>(in, fn); }
template <class T>
inline void write_buffered(std::ostream&, std::vector<T>&, size_t) {}
}}
