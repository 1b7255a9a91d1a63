// TEST INFRASTRUCTURE ONLY -- shim.  Edit classification as published in libvgio (vg/io/edit.hpp):
//   match     from == to, no sequence        sub       from == to, sequence present
//   insertion from == 0, to > 0, sequence    deletion  from > 0, to == 0
// (pinned through the reference's own 10 reconstruction KATs, src/test.cpp:855-994.)
#pragma once
#include "vg/vg.pb.h"
namespace vg { namespace io {
inline bool edit_is_match(const vg::Edit& e) { return e.from_length() == e.to_length() && e.sequence().empty(); }
inline bool edit_is_sub(const vg::Edit& e) { return e.from_length() == e.to_length() && !e.sequence().empty(); }
inline bool edit_is_insertion(const vg::Edit& e) { return e.from_length() == 0 && e.to_length() > 0 && !e.sequence().empty(); }
inline bool edit_is_deletion(const vg::Edit& e) { return e.from_length() > 0 && e.to_length() == 0; }
}}
