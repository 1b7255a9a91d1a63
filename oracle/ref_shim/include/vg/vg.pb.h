// TEST INFRASTRUCTURE ONLY -- header shim for compiling the reference's own hot-path sources
// (read in place from /root/reference/src, never copied) without libvg / protobuf.
//
// Stands in for vg's generated vg.pb.h: the accessor surface of vg::Alignment / Path / Mapping /
// Position / Edit that the reference's path touches (field numbers and meaning per SURVEY.md
// Appendix A), with protobuf's container semantics (int size(), operator[], value copies).
// No arithmetic lives here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <tuple>
#include <functional>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <numeric>
#include <cmath>
#include <stdexcept>
#include <unordered_map>

namespace google { namespace protobuf {
typedef int64_t int64;
typedef int32_t int32;
typedef uint64_t uint64;
typedef uint32_t uint32;

template <class T>
class RepeatedPtrField {
    std::vector<T> v_;
public:
    typedef typename std::vector<T>::iterator iterator;
    typedef typename std::vector<T>::const_iterator const_iterator;
    typedef T value_type;
    int size() const { return (int)v_.size(); }
    bool empty() const { return v_.empty(); }
    // protobuf's RepeatedPtrField::operator[] / Get are unchecked; an index == size() (rmdup.cpp:50,55)
    // is undefined there.  Here it aborts loudly rather than reading garbage.
    const T& operator[](int i) const { return v_.at((size_t)i); }
    T& operator[](int i) { return v_.at((size_t)i); }
    const T& Get(int i) const { return v_.at((size_t)i); }
    const T& at(int i) const { return v_.at((size_t)i); }
    T* Add() { v_.emplace_back(); return &v_.back(); }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
    void Clear() { v_.clear(); }
};
}}  // namespace google::protobuf

namespace vg {
using google::protobuf::int32;
using google::protobuf::int64;

class Edit {
    int32 from_length_ = 0, to_length_ = 0;
    std::string sequence_;
public:
    int32 from_length() const { return from_length_; }
    int32 to_length() const { return to_length_; }
    const std::string& sequence() const { return sequence_; }
    void set_from_length(int32 v) { from_length_ = v; }
    void set_to_length(int32 v) { to_length_ = v; }
    void set_sequence(const std::string& s) { sequence_ = s; }
};

class Position {
    int64 node_id_ = 0, offset_ = 0;
    bool is_reverse_ = false;
    std::string name_;
public:
    int64 node_id() const { return node_id_; }
    int64 offset() const { return offset_; }
    bool is_reverse() const { return is_reverse_; }
    const std::string& name() const { return name_; }
    void set_node_id(int64 v) { node_id_ = v; }
    void set_offset(int64 v) { offset_ = v; }
    void set_is_reverse(bool v) { is_reverse_ = v; }
};

class Mapping {
    Position position_;
    google::protobuf::RepeatedPtrField<Edit> edit_;
    int64 rank_ = 0;
public:
    const Position& position() const { return position_; }
    Position* mutable_position() { return &position_; }
    const google::protobuf::RepeatedPtrField<Edit>& edit() const { return edit_; }
    const Edit& edit(int i) const { return edit_[i]; }
    int edit_size() const { return edit_.size(); }
    Edit* add_edit() { return edit_.Add(); }
    int64 rank() const { return rank_; }
    void set_rank(int64 r) { rank_ = r; }
};

class Path {
    std::string name_;
    google::protobuf::RepeatedPtrField<Mapping> mapping_;
public:
    const std::string& name() const { return name_; }
    void set_name(const std::string& s) { name_ = s; }
    const google::protobuf::RepeatedPtrField<Mapping>& mapping() const { return mapping_; }
    const Mapping& mapping(int i) const { return mapping_[i]; }
    int mapping_size() const { return mapping_.size(); }
    Mapping* add_mapping() { return mapping_.Add(); }
};

class Alignment {
    std::string sequence_, quality_, name_;
    Path path_;
    int32 mapping_quality_ = 0;
    double identity_ = 0;
    bool read_paired_ = false;
    int32 score_ = 0;
public:
    const std::string& sequence() const { return sequence_; }
    const std::string& quality() const { return quality_; }
    const std::string& name() const { return name_; }
    const Path& path() const { return path_; }
    Path* mutable_path() { return &path_; }
    int32 mapping_quality() const { return mapping_quality_; }
    double identity() const { return identity_; }
    bool read_paired() const { return read_paired_; }
    int32 score() const { return score_; }
    void set_sequence(const std::string& s) { sequence_ = s; }
    void set_quality(const std::string& s) { quality_ = s; }
    void set_name(const std::string& s) { name_ = s; }
    void set_mapping_quality(int32 v) { mapping_quality_ = v; }
    void set_identity(double v) { identity_ = v; }
    void set_read_paired(bool v) { read_paired_ = v; }
};

class Node {
    int64 id_ = 0; std::string sequence_;
public:
    int64 id() const { return id_; }
    const std::string& sequence() const { return sequence_; }
};
class Graph {
    google::protobuf::RepeatedPtrField<Node> node_;
    google::protobuf::RepeatedPtrField<Path> path_;
public:
    const google::protobuf::RepeatedPtrField<Node>& node() const { return node_; }
    const google::protobuf::RepeatedPtrField<Path>& path() const { return path_; }
};
}  // namespace vg
