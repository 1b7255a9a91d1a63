// TEST INFRASTRUCTURE ONLY -- shim.
#pragma once
#include "vg/vg.pb.h"
inline std::string pb2json(const vg::Alignment&) { return "{}"; }
