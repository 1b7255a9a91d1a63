// TEST INFRASTRUCTURE ONLY -- shim for vg's utility.hpp: get_input_file hands the callback a stream
// (unused: the shim's vg::io::for_each replays injected alignments), temp_file is a no-op.
#pragma once
#include <functional>
#include <iostream>
#include <fstream>
#include <string>
namespace vg {
inline void get_input_file(const std::string&, std::function<void(std::istream&)> fn) { fn(std::cin); }
namespace temp_file { inline void set_dir(const std::string&) {} }
}
