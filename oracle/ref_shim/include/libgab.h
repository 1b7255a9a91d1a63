// TEST INFRASTRUCTURE ONLY -- shim for libgab (github.com/grenaud/libgab, unpinned: the reference's
// src/Makefile:121 clones HEAD).  libgab is absent from /root/reference, so the few helpers the path
// calls are restated here from libgab's published behaviour (SURVEY.md 8c names each call site):
//   oplusnatl(x,y)      = log(e^x + e^y) = max + log1p(exp(min - max))
//   oplusInitnatl(x,y)  = y when x == 0 (the "not initialised" sentinel), else oplusnatl(x,y)
//   isValidDNA(c)       = c is one of upper-case A C G T
//   allTokens / allTokensWhiteSpaces / destringify / stringify : plain string helpers
//   dimer2indexInt      : 12 off-diagonal base pairs in .prof column order (A>C A>G A>T C>A ... T>G)
// Everything else in this build is the reference's own code, compiled unchanged.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <cassert>
#include <sys/stat.h>
#include <unistd.h>
#include <gzstream.h>

using namespace std;

template <typename T>
inline T oplusnatl(const T x, const T y) {
    if (x > y) return x + log1p(exp(y - x));
    return y + log1p(exp(x - y));
}
template <typename T>
inline T oplusInitnatl(const T x, const T y) {
    if (x == 0) return y;  // "no need to add, x is not initialised"
    return oplusnatl(x, y);
}
template <typename T>
inline T oplus(const T x, const T y) { return x + log10(1 + pow(10.0, y - x)); }
template <typename T>
inline T oplusInit(const T x, const T y) { if (x == 0) return y; return oplus(x, y); }

inline bool isValidDNA(const char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
inline bool isResolvedDNA(const char c) { return isValidDNA(c); }

inline vector<string> allTokens(const string& s, const char delim) {
    vector<string> out;
    size_t start = 0;
    for (;;) {
        size_t p = s.find(delim, start);
        if (p == string::npos) { out.push_back(s.substr(start)); break; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}
inline vector<string> allTokensWhiteSpaces(const string& s) {
    vector<string> out;
    istringstream in(s);
    string tok;
    while (in >> tok) out.push_back(tok);
    return out;
}
template <typename T>
inline T destringify(const string& s) {
    istringstream in(s);
    T v;
    in >> v;
    if (in.fail()) throw runtime_error("destringify: cannot convert " + s);
    return v;
}
template <typename T>
inline string stringify(const T& v) { ostringstream o; o << v; return o.str(); }
inline bool strBeginsWith(const string& s, const string& p) { return s.compare(0, p.size(), p) == 0; }
inline bool strEndsWith(const string& s, const string& p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline bool isFile(const string& f) { struct stat st; return stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
inline bool isDirectory(const string& f) { struct stat st; return stat(f.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool isStringInVector(const string& s, const vector<string>& v) { return find(v.begin(), v.end(), s) != v.end(); }

inline int base2int(const char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
    throw runtime_error(string("base2int: invalid base ") + c);
}
// 12 substitution types, index = row-major over (from, to) skipping the diagonal: AC AG AT CA CG CT GA GC GT TA TC TG
inline int dimer2indexInt(const int from, const int to) {
    if (from == to || from < 0 || from > 3 || to < 0 || to > 3) throw runtime_error("dimer2indexInt: invalid pair");
    return from * 3 + (to > from ? to - 1 : to);
}
inline int dimer2index(const char c1, const char c2) { return dimer2indexInt(base2int(c1), base2int(c2)); }
inline char complement(const char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; }
    return c;
}
inline string reverseComplement(const string& s) { string r(s.rbegin(), s.rend()); for (auto& c : r) c = complement(c); return r; }
inline void printprogressBarCerr(float) {}
inline string getCWD(const char*) { return "."; }
inline string getFullPath(const string& s) { return s; }

// libgab prints containers with operator<< (used by debug output in src/MCMC.cpp:850,951)
template <typename T>
inline ostream& operator<<(ostream& os, const vector<T>& v) {
    for (size_t i = 0; i < v.size(); ++i) { if (i) os << " "; os << v[i]; }
    return os;
}
