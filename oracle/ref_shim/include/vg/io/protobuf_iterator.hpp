// TEST INFRASTRUCTURE ONLY -- shim: iterates the injected alignments (see vg/io/stream.hpp).
#pragma once
#include "vg/io/stream.hpp"
namespace vg { namespace io {
template <class T>
class ProtobufIterator {
    size_t i_ = 0;
public:
    explicit ProtobufIterator(std::istream&) {}
    bool has_current() const { return i_ < ref_shim::injected().size(); }
    bool has_next() const { return has_current(); }
    T& operator*() { return ref_shim::injected()[i_]; }
    void advance() { ++i_; }
    void get_next() { ++i_; }
    ProtobufIterator& operator++() { ++i_; return *this; }
    T take() { return ref_shim::injected()[i_++]; }
};
}}
