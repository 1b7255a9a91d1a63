// TEST INFRASTRUCTURE ONLY -- shim for libgab's gzstream: plain-file streams with the same names
// (the test driver only hands the reference's loaders uncompressed text).
#pragma once
#include <fstream>
class igzstream : public std::ifstream {
public:
    igzstream() {}
    igzstream(const char* name, int mode = std::ios::in) { open(name, mode); }
    void open(const char* name, int mode = std::ios::in) { std::ifstream::open(name, (std::ios_base::openmode)mode); }
};
class ogzstream : public std::ofstream {
public:
    ogzstream() {}
    ogzstream(const char* name, int mode = std::ios::out) { open(name, mode); }
    void open(const char* name, int mode = std::ios::out) { std::ofstream::open(name, (std::ios_base::openmode)mode); }
};
