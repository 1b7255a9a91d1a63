// libvgan_host.so -- host side of the drop-in boundary (include/vgan_host.h).
// GAM reading, alignment reconstruction, SoA packing and table construction.  No likelihood
// arithmetic is done here and nothing in this file touches the oracle.
#include "host_internal.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <thread>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
thread_local char g_err[512] = "";
}  // namespace

int vgh_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
const char* vgh_error_text() { return g_err; }
#define fail vgh_fail

// vgl_hc_batch_meta of a set of reads: the largest canonical quality byte and the mapping qualities present.
namespace {
struct MetaAcc {
    uint32_t qmax = 0;
    uint32_t mask[4] = {0, 0, 0, 0};
    void quality(const uint8_t* q, size_t n) {
        uint32_t m = qmax;
        for (size_t i = 0; i < n; ++i) { const uint32_t c = VGL_HC_CANON_QUAL((uint32_t)q[i]); m = c > m ? c : m; }
        qmax = m;
    }
    void mapq(uint32_t v) { if (v < 128) mask[v >> 5] |= 1u << (v & 31); }
    void merge(const MetaAcc& o) { qmax = std::max(qmax, o.qmax); for (int k = 0; k < 4; ++k) mask[k] |= o.mask[k]; }
    void store(vgl_hc_batch_meta& m) const {
        memset(&m, 0, sizeof(m));
        m.valid = 1;
        m.qual_max = qmax;
        for (int k = 0; k < 4; ++k) m.mapq_mask[k] = mask[k];
    }
};
}  // namespace

namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (p < end && shift < 64) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false;
        return 0;
    }
    // reads a length-delimited field body
    Cursor sub() {
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Cursor{p, p, false}; }
        Cursor c{p, p + n, true};
        p += n;
        return c;
    }
    void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: if (end - p < 8) ok = false; else p += 8; break;
            case 2: sub(); break;
            case 5: if (end - p < 4) ok = false; else p += 4; break;
            default: ok = false;
        }
    }
};

// The framing of a GAM stream: groups of "count, then count length-prefixed items"; the first item of a group is the type
// tag "GAM", the others are serialized vg.Alignment messages.  A first item that does not look like a type tag at all is a
// message of a legacy (untagged) stream, as libvgio reads it.
int scan_gam_framing(const uint8_t* raw_p, size_t raw_n, std::vector<std::pair<const uint8_t*, const uint8_t*>>& msgs) {
    Cursor c{raw_p, raw_p + raw_n, true};
    while (c.ok && c.p < c.end) {
        const uint64_t count = c.varint();
        for (uint64_t k = 0; c.ok && k < count; ++k) {
            Cursor item = c.sub();
            if (!c.ok) break;
            if (k == 0) {
                const size_t len = (size_t)(item.end - item.p);
                if (len == 3 && memcmp(item.p, "GAM", 3) == 0) continue;
                bool tag_like = len > 0 && len <= 25;   // type tags are short identifiers ("VG", "GAM", "MultipathAlignment", ...)
                for (size_t i = 0; tag_like && i < len; ++i) tag_like = isalnum(item.p[i]) != 0;
                if (tag_like) return fail(VGH_ERR_FORMAT, "unexpected type tag in GAM stream");
            }
            msgs.emplace_back(item.p, item.end);
        }
    }
    if (!c.ok) return fail(VGH_ERR_FORMAT, "truncated GAM stream");
    return VGH_OK;
}

// One serialized vg.Alignment (field numbers as on the wire).
bool parse_alignment(Cursor c, int drop_unmapped, vgh_gam& out) {
    const uint8_t *seq_p = nullptr, *seq_e = nullptr, *qual_p = nullptr, *qual_e = nullptr;  // the last occurrence wins, as in protobuf
    int32_t mapq = 0;
    double identity = 0.0;
    uint8_t paired = 0;
    const size_t map_mark = out.map_node.size(), edit_mark = out.edit_from.size(), es_mark = out.edit_seq.size();
    while (c.ok && c.p < c.end) {
        const uint64_t key = c.varint();
        const int fno = (int)(key >> 3), wt = (int)(key & 7);
        if (fno == 1 && wt == 2) { Cursor s = c.sub(); seq_p = s.p; seq_e = s.end; }
        else if (fno == 4 && wt == 2) { Cursor s = c.sub(); qual_p = s.p; qual_e = s.end; }
        else if (fno == 5 && wt == 0) mapq = (int32_t)c.varint();
        else if (fno == 16 && wt == 1) { if (c.end - c.p < 8) { c.ok = false; break; } memcpy(&identity, c.p, 8); c.p += 8; }
        else if (fno == 20 && wt == 0) paired = c.varint() ? 1 : 0;
        else if (fno == 2 && wt == 2) {  // Path
            Cursor path = c.sub();
            while (path.ok && path.p < path.end) {
                const uint64_t k2 = path.varint();
                if ((k2 >> 3) == 2 && (k2 & 7) == 2) {  // Mapping
                    Cursor m = path.sub();
                    int64_t node = 0, offset = 0;
                    uint8_t rev = 0;
                    while (m.ok && m.p < m.end) {
                        const uint64_t k3 = m.varint();
                        if ((k3 >> 3) == 1 && (k3 & 7) == 2) {  // Position
                            Cursor pos = m.sub();
                            while (pos.ok && pos.p < pos.end) {
                                const uint64_t k4 = pos.varint();
                                if ((k4 >> 3) == 1 && (k4 & 7) == 0) node = (int64_t)pos.varint();
                                else if ((k4 >> 3) == 2 && (k4 & 7) == 0) offset = (int64_t)pos.varint();
                                else if ((k4 >> 3) == 4 && (k4 & 7) == 0) rev = pos.varint() ? 1 : 0;
                                else pos.skip((int)(k4 & 7));
                            }
                            if (!pos.ok) m.ok = false;
                        } else if ((k3 >> 3) == 2 && (k3 & 7) == 2) {  // Edit
                            Cursor e = m.sub();
                            int32_t from = 0, to = 0;
                            while (e.ok && e.p < e.end) {
                                const uint64_t k4 = e.varint();
                                if ((k4 >> 3) == 1 && (k4 & 7) == 0) from = (int32_t)e.varint();
                                else if ((k4 >> 3) == 2 && (k4 & 7) == 0) to = (int32_t)e.varint();
                                else if ((k4 >> 3) == 3 && (k4 & 7) == 2) { Cursor s = e.sub(); out.edit_seq.append((const char*)s.p, (const char*)s.end); }
                                else e.skip((int)(k4 & 7));
                            }
                            if (!e.ok) m.ok = false;
                            out.edit_from.push_back(from);
                            out.edit_to.push_back(to);
                            out.edit_seq_off.push_back((int64_t)out.edit_seq.size());
                        } else {
                            m.skip((int)(k3 & 7));
                        }
                    }
                    if (!m.ok) path.ok = false;
                    out.map_node.push_back(node);
                    out.map_offset.push_back(offset);
                    out.map_rev.push_back(rev);
                    out.edit_off.push_back((int64_t)out.edit_from.size());
                } else {
                    path.skip((int)(k2 & 7));
                }
            }
            if (!path.ok) c.ok = false;
        } else {
            c.skip(wt);
        }
    }
    if (!c.ok) return false;
    if (drop_unmapped && identity == 0) {  // readGAM keeps only identity != 0 (src/readGAM.h:47)
        out.map_node.resize(map_mark); out.map_offset.resize(map_mark); out.map_rev.resize(map_mark);
        out.edit_off.resize(map_mark + 1);
        out.edit_from.resize(edit_mark); out.edit_to.resize(edit_mark);
        out.edit_seq_off.resize(edit_mark + 1);
        out.edit_seq.resize(es_mark);
        return true;
    }
    if (seq_p) out.seq.append((const char*)seq_p, (const char*)seq_e);
    out.seq_off.push_back((int64_t)out.seq.size());
    if (qual_p) out.qual.append((const char*)qual_p, (const char*)qual_e);
    out.qual_off.push_back((int64_t)out.qual.size());
    out.mapq.push_back(mapq);
    out.identity.push_back(identity);
    out.paired.push_back(paired);
    out.map_off.push_back((int64_t)out.map_node.size());
    return true;
}

// BGZF (what vg writes): a series of gzip members of at most 64 KB, each carrying its own compressed size in a "BC" extra
// subfield and its uncompressed size in the trailer.  The members are found by hopping from header to header, their
// output offsets are a prefix sum, and they inflate independently -- in parallel.  Returns false (out untouched) when the
// stream is not made of such members from start to end; the caller then inflates serially.
struct BgzfMember { size_t payload, payload_len, out_off; uint32_t isize, crc; };
bool bgzf_members(const uint8_t* in, size_t n, std::vector<BgzfMember>& ms, size_t& total) {
    size_t p = 0;
    total = 0;
    while (p < n) {
        if (n - p < 18 || in[p] != 0x1f || in[p + 1] != 0x8b || in[p + 2] != 8 || !(in[p + 3] & 4)) return false;
        if (in[p + 3] & ~4) return false;  // FNAME / FCOMMENT / FHCRC: not BGZF
        const size_t xlen = (size_t)in[p + 10] | ((size_t)in[p + 11] << 8);
        if (n - p < 12 + xlen + 8) return false;
        size_t bsize = 0;
        for (size_t q = p + 12; q + 4 <= p + 12 + xlen;) {
            const size_t slen = (size_t)in[q + 2] | ((size_t)in[q + 3] << 8);
            if (in[q] == 'B' && in[q + 1] == 'C' && slen == 2 && q + 6 <= p + 12 + xlen) bsize = ((size_t)in[q + 4] | ((size_t)in[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || bsize > n - p) return false;
        const uint8_t* tail = in + p + bsize - 8;
        BgzfMember m;
        m.payload = p + 12 + xlen;
        m.payload_len = bsize - (12 + xlen) - 8;
        m.crc = (uint32_t)tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
        m.isize = (uint32_t)tail[4] | ((uint32_t)tail[5] << 8) | ((uint32_t)tail[6] << 16) | ((uint32_t)tail[7] << 24);
        // BGZF caps a member's uncompressed size at 64 KiB; a trailer that claims more (a crafted file could claim 4 GiB per
        // 28-byte member) is not BGZF.  The serial inflater then decides what the stream really holds.
        if (m.isize > 65536u) return false;
        m.out_off = total;
        total += m.isize;
        ms.push_back(m);
        p += bsize;
    }
    return !ms.empty();
}

int inflate_bgzf(const uint8_t* in, const std::vector<BgzfMember>& ms, size_t total, std::vector<uint8_t>& out) {
    out.resize(total);
    int bad = 0;
#pragma omp parallel
    {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        const bool ready = inflateInit2(&zs, -MAX_WBITS) == Z_OK;  // raw deflate: header and trailer are handled here
#pragma omp for schedule(dynamic, 16)
        for (int64_t k = 0; k < (int64_t)ms.size(); ++k) {
            const BgzfMember& m = ms[(size_t)k];
            if (!ready) { bad = 1; continue; }
            if (m.isize == 0) continue;  // the empty end-of-file member
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef*>(in + m.payload);
            zs.avail_in = (uInt)m.payload_len;
            zs.next_out = out.data() + m.out_off;
            zs.avail_out = m.isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0 ||
                (uint32_t)crc32(crc32(0L, Z_NULL, 0), out.data() + m.out_off, m.isize) != m.crc) {
#pragma omp atomic write
                bad = 1;
            }
        }
        if (ready) inflateEnd(&zs);
    }
    if (bad) return fail(VGH_ERR_FORMAT, "BGZF member corrupt (size or CRC mismatch)");
    return VGH_OK;
}

int inflate_all(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    if (n < 2 || in[0] != 0x1f || in[1] != 0x8b) {  // not gzip: plain stream
        out.assign(in, in + n);
        return VGH_OK;
    }
    {
        std::vector<BgzfMember> ms;
        size_t total = 0;
        if (bgzf_members(in, n, ms, total)) return inflate_bgzf(in, ms, total, out);
    }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) return fail(VGH_ERR_IO, "inflateInit2 failed");
    zs.next_in = const_cast<Bytef*>(in);
    zs.avail_in = (uInt)n;
    std::vector<uint8_t> buf(1 << 16);
    while (true) {
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        const int rc = inflate(&zs, Z_NO_FLUSH);
        out.insert(out.end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
        if (rc == Z_STREAM_END) {
            if (zs.avail_in == 0) break;
            if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); return fail(VGH_ERR_IO, "inflateReset failed"); }  // next BGZF member
        } else if (rc != Z_OK) {
            inflateEnd(&zs);
            return fail(VGH_ERR_FORMAT, "gzip stream corrupt (zlib %d)", rc);
        } else if (zs.avail_in == 0 && zs.avail_out != 0) {
            break;  // truncated stream: keep what was decoded
        }
    }
    inflateEnd(&zs);
    return VGH_OK;
}

}  // namespace

extern "C" {

const char* vgh_last_error(void) { return g_err; }

static int vgh_graph_create__impl(int64_t min_id, int64_t n_nodes, const int64_t* seq_off, const char* seq, vgh_graph** out) {
    if (!out || n_nodes < 0 || !seq_off || (!seq && seq_off[n_nodes] > 0)) return fail(VGH_ERR_INVALID, "bad graph arguments");
    vgh_graph* g = new vgh_graph();
    g->min_id = min_id;
    g->n_nodes = n_nodes;
    g->off.assign(seq_off, seq_off + n_nodes + 1);
    g->seq.assign(seq, seq + seq_off[n_nodes]);
    *out = g;
    return VGH_OK;
}
void vgh_graph_destroy(vgh_graph* g) { delete g; }
int vgh_graph_view(const vgh_graph* g, int64_t* min_id, int64_t* n_nodes, const int64_t** seq_off, const char** seq) {
    if (!g || !min_id || !n_nodes || !seq_off || !seq) return fail(VGH_ERR_INVALID, "bad arguments");
    *min_id = g->min_id;
    *n_nodes = g->n_nodes;
    *seq_off = g->off.data();
    *seq = g->seq.data();
    return VGH_OK;
}

static int vgh_gam_parse__impl(const void* bytes, size_t n, int drop_unmapped, vgh_gam** out) {
    if (!out || (!bytes && n)) return fail(VGH_ERR_INVALID, "bad arguments");
    std::vector<uint8_t> inflated;
    const uint8_t* raw_p = (const uint8_t*)bytes;
    size_t raw_n = n;
    if (n >= 2 && raw_p[0] == 0x1f && raw_p[1] == 0x8b) {  // gzip / BGZF; a plain stream is parsed where it lies
        if (int rc = inflate_all(raw_p, n, inflated)) return rc;
        raw_p = inflated.data();
        raw_n = inflated.size();
    }
    // Pass 1 (serial, a few GB/s): walk the framing -- groups of "count, type tag, count-1 length-prefixed messages" -- and
    // note where every vg.Alignment starts and ends.  Pass 2 (parallel): threads parse contiguous ranges of messages into
    // pieces of their own; the pieces are then appended in order.  readGAM itself (vg::io::for_each, src/readGAM.h:20-68) is
    // single-threaded protobuf parsing.
    std::vector<std::pair<const uint8_t*, const uint8_t*>> msgs;
    if (int rc = scan_gam_framing(raw_p, raw_n, msgs)) return rc;
    int n_threads = 1;
#ifdef _OPENMP
    n_threads = omp_get_max_threads();
#endif
    const size_t n_msgs = msgs.size();
    const int n_parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, n_msgs / 256 + 1));
    std::vector<vgh_gam> parts((size_t)n_parts);
    bool bad = false, oom = false;
#pragma omp parallel for schedule(static, 1) num_threads(n_parts)
    for (int t = 0; t < n_parts; ++t) {
        const size_t lo = n_msgs * (size_t)t / (size_t)n_parts, hi = n_msgs * ((size_t)t + 1) / (size_t)n_parts;
        if (hi > lo) {
            // room for a typical range up front (an aligned read spends about a seventh of its bytes on each of sequence and
            // quality, >= 10 bytes per mapping and >= 6 per edit): the vectors hardly ever reallocate while the messages are
            // parsed.  Only a hint: anything larger still grows, and a failed reservation is ignored.
            vgh_gam& P = parts[(size_t)t];
            const size_t bytes = (size_t)(msgs[hi - 1].second - msgs[lo].first), nr = hi - lo;
            try {
                P.seq.reserve(bytes / 4); P.qual.reserve(bytes / 4); P.edit_seq.reserve(bytes / 16);
                P.seq_off.reserve(nr + 1); P.qual_off.reserve(nr + 1); P.map_off.reserve(nr + 1);
                P.mapq.reserve(nr); P.identity.reserve(nr); P.paired.reserve(nr);
                P.map_node.reserve(bytes / 10); P.map_offset.reserve(bytes / 10); P.map_rev.reserve(bytes / 10); P.edit_off.reserve(bytes / 10 + 1);
                P.edit_from.reserve(bytes / 6); P.edit_to.reserve(bytes / 6); P.edit_seq_off.reserve(bytes / 6 + 1);
            } catch (const std::bad_alloc&) {
            }
        }
        try {   // an exception must not leave the parallel region (std::terminate): a failed allocation becomes an error code
            for (size_t k = lo; k < hi; ++k) {
                Cursor item{msgs[k].first, msgs[k].second, true};
                if (!parse_alignment(item, drop_unmapped, parts[(size_t)t])) {
#pragma omp atomic write
                    bad = true;
                    break;
                }
            }
        } catch (...) {
#pragma omp atomic write
            oom = true;
        }
    }
    if (oom) return fail(VGH_ERR_INVALID, "out of host memory while parsing the GAM stream");
    if (bad) return fail(VGH_ERR_FORMAT, "malformed vg.Alignment message");
    vgh_gam* g = new vgh_gam();
    if (n_parts == 1) {
        *g = std::move(parts[0]);
    } else {
        // lay the pieces out one after the other: sizes first, then every piece copies itself (offset arrays rebased)
        const size_t NP = parts.size();
        std::vector<size_t> r0(NP + 1, 0), m0(NP + 1, 0), e0(NP + 1, 0), s0(NP + 1, 0), q0(NP + 1, 0), x0(NP + 1, 0);
        for (size_t t = 0; t < NP; ++t) {
            const vgh_gam& p = parts[t];
            r0[t + 1] = r0[t] + p.mapq.size();
            m0[t + 1] = m0[t] + p.map_node.size();
            e0[t + 1] = e0[t] + p.edit_from.size();
            s0[t + 1] = s0[t] + p.seq.size();
            q0[t + 1] = q0[t] + p.qual.size();
            x0[t + 1] = x0[t] + p.edit_seq.size();
        }
        const size_t nr = r0[NP], nm = m0[NP], ne = e0[NP];
        // sixteen arrays: allocating and zero-filling them one after the other was the longest part of the whole parse
#pragma omp parallel sections num_threads(n_parts)
        {
#pragma omp section
            { g->seq_off.resize(nr + 1); g->qual_off.resize(nr + 1); g->map_off.resize(nr + 1); g->mapq.resize(nr); g->paired.resize(nr); }
#pragma omp section
            { g->identity.resize(nr); g->map_rev.resize(nm); }
#pragma omp section
            g->map_node.resize(nm);
#pragma omp section
            g->map_offset.resize(nm);
#pragma omp section
            { g->edit_off.resize(nm + 1); g->edit_seq.resize(x0[NP]); }
#pragma omp section
            { g->edit_from.resize(ne); g->edit_to.resize(ne); }
#pragma omp section
            g->edit_seq_off.resize(ne + 1);
#pragma omp section
            { g->seq.resize(s0[NP]); g->qual.resize(q0[NP]); }
        }
#pragma omp parallel for schedule(static, 1) num_threads(n_parts)
        for (int t = 0; t < n_parts; ++t) {
            vgh_gam& p = parts[(size_t)t];
            const size_t R = p.mapq.size(), M = p.map_node.size(), E = p.edit_from.size();
            for (size_t k = 1; k <= R; ++k) {
                g->seq_off[r0[t] + k] = (int64_t)s0[t] + p.seq_off[k];
                g->qual_off[r0[t] + k] = (int64_t)q0[t] + p.qual_off[k];
                g->map_off[r0[t] + k] = (int64_t)m0[t] + p.map_off[k];
            }
            for (size_t k = 1; k <= M; ++k) g->edit_off[m0[t] + k] = (int64_t)e0[t] + p.edit_off[k];
            for (size_t k = 1; k <= E; ++k) g->edit_seq_off[e0[t] + k] = (int64_t)x0[t] + p.edit_seq_off[k];
            if (R) {
                memcpy(&g->mapq[r0[t]], p.mapq.data(), R * sizeof(int32_t));
                memcpy(&g->identity[r0[t]], p.identity.data(), R * sizeof(double));
                memcpy(&g->paired[r0[t]], p.paired.data(), R);
            }
            if (M) {
                memcpy(&g->map_node[m0[t]], p.map_node.data(), M * sizeof(int64_t));
                memcpy(&g->map_offset[m0[t]], p.map_offset.data(), M * sizeof(int64_t));
                memcpy(&g->map_rev[m0[t]], p.map_rev.data(), M);
            }
            if (E) {
                memcpy(&g->edit_from[e0[t]], p.edit_from.data(), E * sizeof(int32_t));
                memcpy(&g->edit_to[e0[t]], p.edit_to.data(), E * sizeof(int32_t));
            }
            if (!p.seq.empty()) memcpy(&g->seq[s0[t]], p.seq.data(), p.seq.size());
            if (!p.qual.empty()) memcpy(&g->qual[q0[t]], p.qual.data(), p.qual.size());
            if (!p.edit_seq.empty()) memcpy(&g->edit_seq[x0[t]], p.edit_seq.data(), p.edit_seq.size());
            p = vgh_gam();  // release the piece
        }
    }
    g->finish();
    *out = g;
    return VGH_OK;
}

static int vgh_gam_read_file__impl(const char* path, int drop_unmapped, vgh_gam** out) {
    if (!path || !out) return fail(VGH_ERR_INVALID, "bad arguments");
    FILE* fh = fopen(path, "rb");
    if (!fh) return fail(VGH_ERR_IO, "cannot open %s", path);
    std::vector<uint8_t> raw;
    uint8_t buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), fh)) > 0) raw.insert(raw.end(), buf, buf + k);
    fclose(fh);
    return vgh_gam_parse(raw.data(), raw.size(), drop_unmapped, out);
}
const vgh_alignments* vgh_gam_view(const vgh_gam* g) { return g ? &g->view : nullptr; }
void vgh_gam_destroy(vgh_gam* g) { delete g; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// alignment reconstruction
// ------------------------------------------------------------------------------------------------
namespace {

enum EditKind { kMatch, kSub, kIns, kDel, kOther };
inline EditKind classify(int32_t from, int32_t to, bool has_seq) {  // vg io edit_is_*
    if (from == to && !has_seq) return kMatch;
    if (from == to && has_seq) return kSub;
    if (from == 0 && to > 0 && has_seq) return kIns;
    if (from > 0 && to == 0) return kDel;
    return kOther;
}


// reconstruct_graph_sequence (reference src/vgan_utils.h:6-79), including the behaviours the
// reference's counters produce: an insertion is an 'S' run whenever the running node
// offset is 0, a '-' run otherwise; one size per edit; the '-' run of a deletion is spliced into the
// read string at the running from_length coordinate.
}  // namespace

int vgh_recon(const vgh_graph& G, const vgh_alignments& a, int64_t r, Recon& out) {
    out.graph_seq.clear();
    out.read_seq.clear();
    out.sizes.clear();
    struct Splice { size_t at; int32_t n; };
    std::vector<Splice> splices;
    int64_t from_total = 0;
    for (int64_t m = a.map_off[r]; m < a.map_off[r + 1]; ++m) {
        const int64_t node = a.map_node[m];
        const bool rev = a.map_rev[m] != 0;
        if (!G.has(node)) return fail(VGH_ERR_BAD_NODE, "read %lld: node %lld is not in the graph", (long long)r, (long long)node);
        const int64_t nlen = G.length(node);
        int64_t offset = a.map_offset[m];
        int32_t aligned = 0;  // per mapping, NOT per edit: an edit of no recognised kind repeats the previous size (vgan_utils.h:26,72)
        for (int64_t e = a.edit_off[m]; e < a.edit_off[m + 1]; ++e) {
            const int32_t from = a.edit_from[e], to = a.edit_to[e];
            const int64_t s0 = a.edit_seq_off[e], s1 = a.edit_seq_off[e + 1];
            const EditKind kind = classify(from, to, s1 > s0);
            if (kind == kMatch || kind == kSub || kind == kDel) {
                if (offset > nlen || offset < 0)
                    return fail(VGH_ERR_BAD_READ, "read %lld: offset %lld beyond node %lld (length %lld)", (long long)r,
                                (long long)offset, (long long)node, (long long)nlen);
                const int64_t take = std::min<int64_t>(from, nlen - offset);  // std::string::substr clips
                for (int64_t k = 0; k < take; ++k) out.graph_seq.push_back(G.base(node, rev, offset + k));
                aligned = (int32_t)take;
                if (kind == kMatch) out.read_seq.append(out.graph_seq, out.graph_seq.size() - (size_t)take, (size_t)take);
                else if (kind == kSub) out.read_seq.append(a.edit_seq + s0, (size_t)(s1 - s0));
                else splices.push_back(Splice{(size_t)from_total, from});
            } else if (kind == kIns) {
                out.graph_seq.append((size_t)to, offset == 0 ? 'S' : '-');
                out.read_seq.append(a.edit_seq + s0, (size_t)(s1 - s0));
                aligned = to;
            }
            offset += from;
            from_total += from;
            out.sizes.push_back(aligned);
        }
    }
    for (const Splice& s : splices) {  // applied in order, each against the string as it stands
        if (s.at > out.read_seq.size()) return fail(VGH_ERR_BAD_READ, "read %lld: deletion beyond the read string", (long long)r);
        out.read_seq.insert(s.at, (size_t)s.n, '-');
    }
    return VGH_OK;
}

static inline int reconstruct(const vgh_graph& G, const vgh_alignments& a, int64_t r, Recon& out) { return vgh_recon(G, a, r, out); }

extern "C" int vgh_reconstruct(const vgh_graph* g, const vgh_alignments* a, int64_t r, char* graph_seq, char* read_seq,
                               int32_t* sizes, int32_t cap, int32_t* n_graph, int32_t* n_read, int32_t* n_sizes) {
    if (!g || !a || r < 0 || r >= a->n_reads) return fail(VGH_ERR_INVALID, "bad arguments");
    Recon rc;
    if (int e = reconstruct(*g, *a, r, rc)) return e;
    if ((int64_t)rc.graph_seq.size() > cap || (int64_t)rc.read_seq.size() > cap || (int64_t)rc.sizes.size() > cap)
        return fail(VGH_ERR_INVALID, "buffers too small");
    memcpy(graph_seq, rc.graph_seq.data(), rc.graph_seq.size());
    memcpy(read_seq, rc.read_seq.data(), rc.read_seq.size());
    for (size_t i = 0; i < rc.sizes.size(); ++i) sizes[i] = rc.sizes[i];
    *n_graph = (int32_t)rc.graph_seq.size();
    *n_read = (int32_t)rc.read_seq.size();
    *n_sizes = (int32_t)rc.sizes.size();
    return VGH_OK;
}

// ------------------------------------------------------------------------------------------------
// duplicate removal
// ------------------------------------------------------------------------------------------------
extern "C" int vgh_dedup(const vgh_alignments* a, uint8_t* is_dup, int64_t* n_kept) {
    if (!a || !is_dup) return fail(VGH_ERR_INVALID, "bad arguments");
    struct Key { int64_t node, offset; bool operator==(const Key& o) const { return node == o.node && offset == o.offset; } };
    struct KeyHash { size_t operator()(const Key& k) const { return std::hash<int64_t>()(k.node * 1000003 + k.offset); } };
    std::unordered_map<Key, char, KeyHash> seen;  // start coordinates of the unpaired reads met so far
    seen.reserve((size_t)a->n_reads * 2 + 16);
    int64_t kept = 0;
    for (int64_t r = 0; r < a->n_reads; ++r) {
        if (a->map_off[r + 1] == a->map_off[r]) return fail(VGH_ERR_BAD_READ, "read %lld has no mapping (the reference dereferences mapping()[0])", (long long)r);
        const Key k{a->map_node[a->map_off[r]], a->map_offset[a->map_off[r]]};
        is_dup[r] = seen.count(k) ? 1 : 0;
        kept += !is_dup[r];
        const bool paired = a->paired && a->paired[r];
        if (!paired) seen.emplace(k, 1);
    }
    if (n_kept) *n_kept = kept;
    return VGH_OK;
}

extern "C" int vgh_dedup_keys(const vgh_alignments* a, uint64_t* keys) {
    if (!a || (a->n_reads && !keys)) return fail(VGH_ERR_INVALID, "bad arguments");
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t r = 0; r < a->n_reads; ++r) {
        if (a->map_off[r + 1] == a->map_off[r]) { bad = 1; keys[r] = 0; continue; }
        const int64_t node = a->map_node[a->map_off[r]], off = a->map_offset[a->map_off[r]];
        if (node < 0 || node > 0x7fffffffll || off < 0 || off > 0xffffffffll) bad = 1;
        keys[r] = ((uint64_t)node << 32) | (uint64_t)(uint32_t)off;
    }
    if (bad) return fail(VGH_ERR_BAD_READ, "a read has no mapping (the reference dereferences mapping()[0]) or start coordinates out of range");
    return VGH_OK;
}

// ------------------------------------------------------------------------------------------------
// batch packer
// ------------------------------------------------------------------------------------------------
struct vgh_hc_packed {
    vgl_hc_batch batch{};
    bool pinned = false;
    int64_t skipped = 0;
    void* block = nullptr;  // one allocation holds every array
    size_t block_bytes = 0;
};

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" {

// Meta of the kept reads `keep` of `a` (one parallel pass over their quality bytes and mapping qualities).
static void reads_meta(const vgh_alignments* a, const std::vector<int64_t>& keep, vgl_hc_batch_meta& out) {
    MetaAcc total;
#pragma omp parallel
    {
        MetaAcc mine;
#pragma omp for schedule(static) nowait
        for (int64_t k = 0; k < (int64_t)keep.size(); ++k) {
            const int64_t r = keep[(size_t)k];
            mine.quality((const uint8_t*)a->qual + a->qual_off[r], (size_t)(a->qual_off[r + 1] - a->qual_off[r]));
            mine.mapq((uint32_t)a->mapq[r]);
        }
#pragma omp critical(vgh_meta)
        total.merge(mine);
    }
    total.store(out);
}

static int vgh_hc_pack__impl(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                vgh_hc_packed** out) {
    if (!g || !a || !out || r0 < 0 || r1 > a->n_reads || r0 > r1) return fail(VGH_ERR_INVALID, "bad arguments");
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    // The reads the hot loop processes: identity >= 1e-10 (src/HaploCart.cpp:411).
    std::vector<int64_t> keep;
    keep.reserve((size_t)(r1 - r0));
    for (int64_t r = r0; r < r1; ++r)
        if (!(a->identity[r] < 1e-10)) keep.push_back(r);
    const size_t R = keep.size();
    if (R >= 0xffffffffull) return fail(VGH_ERR_INVALID, "too many reads for one batch");

    // pass 1: reconstruct every kept read once to learn its column / mapping counts
    std::vector<uint32_t> ncols(R), nmaps(R), ncloses(R), nextras(R);
    int err = VGH_OK;
    std::string err_msg;
#pragma omp parallel num_threads(n_threads)
    {
        Recon rc;
#pragma omp for schedule(static)
        for (int64_t k = 0; k < (int64_t)R; ++k) {
            const int64_t r = keep[(size_t)k];
            int e = reconstruct(*g, *a, r, rc);
            const int64_t M = a->map_off[r + 1] - a->map_off[r];
            const int64_t Lq = a->qual_off[r + 1] - a->qual_off[r];
            if (!e) {
                if (rc.graph_seq.size() != rc.read_seq.size())
                    e = fail(VGH_ERR_BAD_READ, "read %lld: aligned graph / read strings differ in length (%zu vs %zu)", (long long)r,
                             rc.graph_seq.size(), rc.read_seq.size());
                else if ((int64_t)rc.sizes.size() < M)
                    e = fail(VGH_ERR_BAD_READ, "read %lld: a mapping has no edits (the reference indexes mppg_sizes out of range)", (long long)r);
                else if (Lq > (int64_t)rc.read_seq.size())
                    e = fail(VGH_ERR_BAD_READ, "read %lld: quality string longer than the alignment", (long long)r);
                else if (rc.read_seq.size() > VGL_HC_MAP_MAX_POS || M > 65535)
                    e = fail(VGH_ERR_BAD_READ, "read %lld: alignment of %zu columns does not fit the batch records (max %u)", (long long)r,
                             rc.read_seq.size(), VGL_HC_MAP_MAX_POS);
                else if (a->mapq[r] < 0 || a->mapq[r] >= 100)
                    e = fail(VGH_ERR_BAD_READ, "read %lld: mapping quality %d outside [0,100)", (long long)r, a->mapq[r]);
                else {
                    for (int64_t i = 0; i < M && !e; ++i) {
                        const int64_t node = a->map_node[a->map_off[r] + i];
                        if (node < 0 || node > (int64_t)VGL_HC_MAX_NODE)
                            e = fail(VGH_ERR_BAD_NODE, "read %lld: node id %lld does not fit the column record", (long long)r, (long long)node);
                    }
                }
            }
            if (e) {
#pragma omp critical(vgh_pack_err)
                if (!err) { err = e; err_msg = g_err; }
                continue;
            }
            // mapping i starts at pos_i = min(sum of the earlier sizes, n); CLOSE bits sit at the distinct pos_i < n
            const size_t L = rc.read_seq.size();
            size_t pos = 0, last_close = (size_t)-1;
            uint32_t n_close = 0, n_extra = 0;
            for (int64_t i = 0; i < M; ++i) {
                const size_t lg = std::min((size_t)rc.sizes[(size_t)i], L - pos);
                if (lg == 0) ++n_extra;
                if (pos < L && pos != last_close) { ++n_close; last_close = pos; }
                pos += lg;
            }
            ncols[(size_t)k] = (uint32_t)L;
            nmaps[(size_t)k] = (uint32_t)M;
            ncloses[(size_t)k] = n_close;
            nextras[(size_t)k] = n_extra;
        }
    }
    if (err) return fail(err, "%s", err_msg.c_str());

    // Group reads into warps of 32 lanes (the sum over reads is order independent).  First by mapping quality -- a group
    // holds reads of ONE mapping quality and every mapping quality's run of groups is a segment of the batch, so that the
    // log-term read kernel stages the tables of a single mapping quality per CTA (vgl_hc_batch_meta) -- then by shape
    // (mappings, columns), so that the zero padding up to each group's longest read stays small.  A segment's last group is
    // filled up with unused lanes.  More than VGL_HC_MAX_SEGS distinct mapping qualities: no segments, shape only.
    const uint32_t kNoRead = 0xffffffffu;
    uint32_t mq_count[128];
    memset(mq_count, 0, sizeof(mq_count));
    for (size_t k = 0; k < R; ++k) mq_count[(size_t)a->mapq[keep[k]] & 127u] += 1;
    int n_distinct = 0;
    for (int q = 0; q < 128; ++q) n_distinct += mq_count[q] ? 1 : 0;
    const bool segmented = n_distinct >= 1 && n_distinct <= VGL_HC_MAX_SEGS;
    std::vector<uint32_t> sorted(R);
    for (size_t k = 0; k < R; ++k) sorted[k] = (uint32_t)k;
    std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t x, uint32_t y) {
        const uint64_t kx = ((uint64_t)(segmented ? a->mapq[keep[x]] : 0) << 48) | ((uint64_t)ncloses[x] << 24) | ncols[x];
        const uint64_t ky = ((uint64_t)(segmented ? a->mapq[keep[y]] : 0) << 48) | ((uint64_t)ncloses[y] << 24) | ncols[y];
        return kx < ky;
    });
    std::vector<uint32_t> order;  // lane -> index into keep, kNoRead for an unused lane
    vgl_hc_batch_meta seg_meta;
    memset(&seg_meta, 0, sizeof(seg_meta));
    if (segmented) {
        size_t at = 0;
        for (int q = 0; q < 128; ++q) {
            if (!mq_count[q]) continue;
            for (uint32_t i = 0; i < mq_count[q]; ++i) order.push_back(sorted[at++]);
            while (order.size() % VGL_HC_LANES) order.push_back(kNoRead);
            seg_meta.seg_mapq[seg_meta.n_segs] = (uint8_t)q;
            seg_meta.seg_end[seg_meta.n_segs] = (uint32_t)(order.size() / VGL_HC_LANES);
            seg_meta.n_segs += 1;
        }
    } else {
        order = sorted;
        while (order.size() % VGL_HC_LANES) order.push_back(kNoRead);
    }
    const size_t n_groups = order.size() / VGL_HC_LANES;
    std::vector<uint32_t> g_col(n_groups + 1, 0), g_map(n_groups + 1, 0);
    for (size_t gI = 0; gI < n_groups; ++gI) {
        uint32_t mc = 0, mm = 0;
        for (size_t l = 0; l < VGL_HC_LANES; ++l) {
            const uint32_t k = order[gI * VGL_HC_LANES + l];
            if (k == kNoRead) continue;
            mc = std::max(mc, ncols[k]);
            mm = std::max(mm, nextras[k]);
        }
        const uint64_t nc = (uint64_t)g_col[gI] + (mc + VGL_HC_COL_BLK - 1) / VGL_HC_COL_BLK;
        const uint64_t nm = (uint64_t)g_map[gI] + (mm + VGL_HC_MAP_BLK - 1) / VGL_HC_MAP_BLK;
        if (nc > 0x007fffffull || nm > 0x007fffffull) return fail(VGH_ERR_INVALID, "batch too large for 32-bit block offsets; split it");
        g_col[gI + 1] = (uint32_t)nc;
        g_map[gI + 1] = (uint32_t)nm;
    }
    const size_t col_blks = g_col[n_groups], map_blks = g_map[n_groups], lanes = n_groups * VGL_HC_LANES;
    const size_t kColBlk = (size_t)VGL_HC_LANES * VGL_HC_COL_BLK, kMapBlk = (size_t)VGL_HC_LANES * VGL_HC_MAP_BLK;  // records per block

    // one block, 256-byte aligned sub-arrays
    size_t o = 0;
    const size_t o_gc = o; o = align_up(o + (n_groups + 1) * 4, 256);
    const size_t o_gm = o; o = align_up(o + (n_groups + 1) * 4, 256);
    const size_t o_col = o; o = align_up(o + col_blks * kColBlk * 4, 256);
    const size_t o_map = o; o = align_up(o + map_blks * kMapBlk * 4, 256);
    const size_t o_nc = o; o = align_up(o + lanes * 2, 256);
    const size_t o_ncl = o; o = align_up(o + lanes * 2, 256);
    const size_t o_nx = o; o = align_up(o + lanes * 2, 256);
    const size_t o_mq = o; o = align_up(o + lanes, 256);
    vgh_hc_packed* P = new vgh_hc_packed();
    P->block_bytes = std::max<size_t>(o, 256);
    P->pinned = pinned != 0;
    if (pinned) {
        if (vgl_host_alloc(&P->block, P->block_bytes) != VGL_OK) {
            delete P;
            return fail(VGH_ERR_INVALID, "pinned allocation failed: %s", vgl_last_error());
        }
    } else {
        P->block = aligned_alloc(256, P->block_bytes);
        if (!P->block) { delete P; return fail(VGH_ERR_INVALID, "out of host memory"); }
    }
    uint8_t* base = (uint8_t*)P->block;
    uint32_t* grp_col_blk = (uint32_t*)(base + o_gc);
    uint32_t* grp_map_blk = (uint32_t*)(base + o_gm);
    uint32_t* col = (uint32_t*)(base + o_col);
    uint32_t* map = (uint32_t*)(base + o_map);
    uint16_t* n_cols = (uint16_t*)(base + o_nc);
    uint16_t* n_close = (uint16_t*)(base + o_ncl);
    uint16_t* n_extra = (uint16_t*)(base + o_nx);
    uint8_t* mapq = base + o_mq;
    memcpy(grp_col_blk, g_col.data(), (n_groups + 1) * 4);
    memcpy(grp_map_blk, g_map.data(), (n_groups + 1) * 4);

    // pass 2: fill, one group per task.  Column records pair G[c] with the read base process_mapping compares it
    // with (src/update_likelihood.cpp:33-49, src/process_mapping.cpp:47,59-66): for mapping i the graph slice
    // G[pos_i, pos_i + size_i) faces algnseq[0, size_i) -- the whole aligned read string from its start.
    auto base_code = [](char ch) -> int { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; };
#pragma omp parallel num_threads(n_threads)
    {
        Recon rc;
#pragma omp for schedule(dynamic, 16)
        for (int64_t gI = 0; gI < (int64_t)n_groups; ++gI) {
            uint32_t* gcol = col + (size_t)g_col[(size_t)gI] * kColBlk;
            uint32_t* gmap = map + (size_t)g_map[(size_t)gI] * kMapBlk;
            memset(gcol, 0, (size_t)(g_col[(size_t)gI + 1] - g_col[(size_t)gI]) * kColBlk * 4);
            memset(gmap, 0, (size_t)(g_map[(size_t)gI + 1] - g_map[(size_t)gI]) * kMapBlk * 4);
            // record c of lane l: block c / 4, then lane, then c % 4 (extra records likewise)
            auto col_at = [&](size_t c, size_t l) -> uint32_t& { return gcol[((c / VGL_HC_COL_BLK) * VGL_HC_LANES + l) * VGL_HC_COL_BLK + c % VGL_HC_COL_BLK]; };
            auto map_at = [&](size_t i, size_t l) -> uint32_t& { return gmap[((i / VGL_HC_MAP_BLK) * VGL_HC_LANES + l) * VGL_HC_MAP_BLK + i % VGL_HC_MAP_BLK]; };
            for (size_t l = 0; l < VGL_HC_LANES; ++l) {
                const size_t lane_id = (size_t)gI * VGL_HC_LANES + l;
                n_cols[lane_id] = n_close[lane_id] = n_extra[lane_id] = 0;
                mapq[lane_id] = 0;
                const uint32_t k = order[lane_id];
                if (k == kNoRead) continue;
                const int64_t r = keep[k];
                if (reconstruct(*g, *a, r, rc)) continue;  // cannot fail: pass 1 succeeded
                const size_t L = rc.read_seq.size();
                const int64_t q0 = a->qual_off[r], Lq = a->qual_off[r + 1] - q0;
                size_t first_q90 = (size_t)-1;  // the first REAL quality byte >= 90 latches the background-error mode
                for (size_t c = 0; c < L; ++c) {
                    const uint8_t qb = (int64_t)c < Lq ? (uint8_t)a->qual[q0 + (int64_t)c] : 0;
                    col_at(c, l) = (VGL_HC_CANON_QUAL((uint32_t)qb) << VGL_HC_COL_QUAL_SHIFT) | (VGL_HC_NO_NODE << VGL_HC_COL_NODE_SHIFT);
                    if ((int64_t)c < Lq && (int)(int8_t)qb >= 90 && first_q90 == (size_t)-1) first_q90 = c;
                }
                const int64_t mm = a->map_off[r];
                size_t pos = 0, n_x = 0;
                for (uint32_t i = 0; i < nmaps[k]; ++i) {
                    const size_t lg = std::min((size_t)rc.sizes[i], L - pos);
                    const uint32_t node = (uint32_t)a->map_node[mm + i];
                    // its quality window is [pos, pos + L)  (src/update_likelihood.cpp:40-44)
                    const uint32_t latched = (first_q90 != (size_t)-1 && first_q90 < pos + L) ? VGL_HC_COL_LATCHED : 0u;
                    for (size_t j = 0; j < lg; ++j) {
                        const char gb = rc.graph_seq[pos + j], rb = rc.read_seq[j];
                        const int cg = base_code(gb), cb = base_code(rb);
                        uint32_t& rec = col_at(pos + j, l);
                        rec = (rec & (0xffu << VGL_HC_COL_QUAL_SHIFT)) | (node << VGL_HC_COL_NODE_SHIFT) | latched;
                        if (cg < 4 && cb < 4) rec |= VGL_HC_COL_VALID | ((uint32_t)cb << VGL_HC_COL_BASE_SHIFT) | (gb == rb ? 0u : VGL_HC_COL_MISMATCH);
                    }
                    if (lg) col_at(pos + lg - 1, l) |= VGL_HC_COL_OPEN;
                    else map_at(n_x++, l) = VGL_HC_MAP_REC(node, (uint32_t)pos);
                    if (pos < L) col_at(pos, l) |= VGL_HC_COL_CLOSE;
                    pos += lg;
                }
                n_cols[lane_id] = (uint16_t)L;
                n_close[lane_id] = (uint16_t)ncloses[k];
                n_extra[lane_id] = (uint16_t)n_x;
                mapq[lane_id] = (uint8_t)a->mapq[r];
            }
        }
    }
    P->skipped = (r1 - r0) - (int64_t)R;
    P->batch.n_reads = (uint32_t)R;
    P->batch.n_groups = (uint32_t)n_groups;
    P->batch.n_col_blks = col_blks;
    P->batch.n_map_blks = map_blks;
    P->batch.grp_col_blk = grp_col_blk; P->batch.grp_map_blk = grp_map_blk;
    P->batch.col = col; P->batch.map = map;
    P->batch.n_cols = n_cols; P->batch.n_close = n_close; P->batch.n_extra = n_extra; P->batch.mapq = mapq;
    reads_meta(a, keep, P->batch.meta);
    P->batch.meta.n_segs = seg_meta.n_segs;
    memcpy(P->batch.meta.seg_mapq, seg_meta.seg_mapq, sizeof(seg_meta.seg_mapq));
    memcpy(P->batch.meta.seg_end, seg_meta.seg_end, sizeof(seg_meta.seg_end));
    *out = P;
    return VGH_OK;
}

const vgl_hc_batch* vgh_hc_packed_batch(const vgh_hc_packed* p) { return p ? &p->batch : nullptr; }
int64_t vgh_hc_packed_skipped(const vgh_hc_packed* p) { return p ? p->skipped : 0; }
void vgh_hc_packed_destroy(vgh_hc_packed* p) {
    if (!p) return;
    if (p->block) {
        if (p->pinned) vgl_host_free(p->block);
        else free(p->block);
    }
    delete p;
}

// ------------------------------------------------------------------------------------------------
// wire form of a packed batch
// ------------------------------------------------------------------------------------------------
}  // extern "C"

struct vgh_hc_wire {
    vgl_hc_wire view{};
    bool pinned = false;
    void* block = nullptr;
};

extern "C" {

static int vgh_hc_wire_from_packed__impl(const vgh_hc_packed* p, int n_threads, int pinned, vgh_hc_wire** out) {
    if (!p || !out) return fail(VGH_ERR_INVALID, "bad arguments");
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    const vgl_hc_batch& b = p->batch;
    const size_t G = b.n_groups;
    // per group: 16-bit column blocks (8 columns) and node blocks (8 nodes) as its longest lane needs
    std::vector<uint32_t> g_c16(G + 1, 0), g_n16(G + 1, 0);
    int err = VGH_OK;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int64_t g = 0; g < (int64_t)G; ++g) {
        const uint32_t cb0 = b.grp_col_blk[g], ncb = b.grp_col_blk[g + 1] - cb0;
        uint32_t max_cols = 0, max_nodes = 0;
        for (uint32_t l = 0; l < VGL_HC_LANES; ++l) {
            const uint32_t L = b.n_cols[(size_t)g * VGL_HC_LANES + l];
            uint32_t n_open = 0;
            for (uint32_t c = 0; c < L; ++c) {
                const uint32_t rec = b.col[(((size_t)cb0 + c / VGL_HC_COL_BLK) * VGL_HC_LANES + l) * VGL_HC_COL_BLK + c % VGL_HC_COL_BLK];
                if (rec & VGL_HC_COL_OPEN) {
                    ++n_open;
                    if (VGL_HC_COL_NODE(rec) > 0xffffu) err = VGH_ERR_INVALID;
                }
            }
            max_cols = std::max(max_cols, L);
            max_nodes = std::max(max_nodes, n_open);
        }
        (void)ncb;
        g_c16[(size_t)g + 1] = (max_cols + VGL_HC_WIRE_BLK - 1) / VGL_HC_WIRE_BLK;
        g_n16[(size_t)g + 1] = (max_nodes + VGL_HC_WIRE_BLK - 1) / VGL_HC_WIRE_BLK;
    }
    if (err) return fail(VGH_ERR_INVALID, "a node id exceeds 65535: this graph needs the 32-bit batch");
    for (size_t g = 0; g < G; ++g) { g_c16[g + 1] += g_c16[g]; g_n16[g + 1] += g_n16[g]; }
    const size_t lanes = G * VGL_HC_LANES, n_map_rec = (size_t)b.n_map_blks * VGL_HC_LANES * VGL_HC_MAP_BLK;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    const size_t o_gc = take((G + 1) * 4), o_gn = take((G + 1) * 4), o_gm = take((G + 1) * 4);
    const size_t o_col = take((size_t)g_c16[G] * 512), o_node = take((size_t)g_n16[G] * 512), o_map = take(n_map_rec * 4);
    const size_t o_nc = take(lanes * 2), o_ncl = take(lanes * 2), o_nx = take(lanes * 2), o_mq = take(lanes);
    // narrow column records (12 bits per column), used when every column of the batch qualifies
    const size_t o_c8 = take((size_t)g_c16[G] * 256), o_f4 = take((size_t)g_c16[G] * 128), o_own = take(lanes * 2);
    vgh_hc_wire* W = new vgh_hc_wire();
    W->pinned = pinned != 0;
    const size_t bytes = std::max<size_t>(o, 256);
    if (pinned) {
        if (vgl_host_alloc(&W->block, bytes) != VGL_OK) {
            delete W;
            return fail(VGH_ERR_INVALID, "pinned allocation failed: %s", vgl_last_error());
        }
    } else {
        W->block = aligned_alloc(256, bytes);
        if (!W->block) { delete W; return fail(VGH_ERR_INVALID, "out of host memory"); }
    }
    uint8_t* base = (uint8_t*)W->block;
    uint32_t* gc = (uint32_t*)(base + o_gc); uint32_t* gn = (uint32_t*)(base + o_gn); uint32_t* gm = (uint32_t*)(base + o_gm);
    uint16_t* col = (uint16_t*)(base + o_col); uint16_t* node = (uint16_t*)(base + o_node); uint32_t* map = (uint32_t*)(base + o_map);
    memcpy(gc, g_c16.data(), (G + 1) * 4);
    memcpy(gn, g_n16.data(), (G + 1) * 4);
    memcpy(gm, b.grp_map_blk, (G + 1) * 4);
    if (n_map_rec) memcpy(map, b.map, n_map_rec * 4);
    memcpy(base + o_nc, b.n_cols, lanes * 2);
    memcpy(base + o_ncl, b.n_close, lanes * 2);
    memcpy(base + o_nx, b.n_extra, lanes * 2);
    memcpy(base + o_mq, b.mapq, lanes);
    uint8_t* col8 = base + o_c8;
    uint32_t* flag4 = (uint32_t*)(base + o_f4);
    uint16_t* n_owned = (uint16_t*)(base + o_own);
    bool narrow = true;  // every quality <= 63, no LATCHED column, VALID only on owned columns, the owned columns a prefix
#pragma omp parallel for schedule(static) num_threads(n_threads) reduction(&& : narrow)
    for (int64_t g = 0; g < (int64_t)G; ++g) {
        const uint32_t cb0 = b.grp_col_blk[g];
        uint16_t* gcol = col + (size_t)g_c16[(size_t)g] * 256;
        uint16_t* gnode = node + (size_t)g_n16[(size_t)g] * 256;
        uint8_t* gc8 = col8 + (size_t)g_c16[(size_t)g] * 256;
        uint32_t* gf4 = flag4 + (size_t)g_c16[(size_t)g] * 32;
        memset(gcol, 0, (size_t)(g_c16[(size_t)g + 1] - g_c16[(size_t)g]) * 512);
        memset(gnode, 0, (size_t)(g_n16[(size_t)g + 1] - g_n16[(size_t)g]) * 512);
        memset(gc8, 0, (size_t)(g_c16[(size_t)g + 1] - g_c16[(size_t)g]) * 256);
        memset(gf4, 0, (size_t)(g_c16[(size_t)g + 1] - g_c16[(size_t)g]) * 128);
        for (uint32_t l = 0; l < VGL_HC_LANES; ++l) {
            const uint32_t L = b.n_cols[(size_t)g * VGL_HC_LANES + l];
            uint32_t k = 0;  // nodes are listed in the order a walk from the last column meets the OPEN marks
            uint32_t owned = 0;      // owned columns seen so far; they must be columns 0 .. owned - 1
            bool seen_owned = false; // walking down from the last column: once a column is owned, all below it must be
            for (uint32_t cc = L; cc > 0; --cc) {
                const uint32_t c = cc - 1;
                const uint32_t rec = b.col[(((size_t)cb0 + c / VGL_HC_COL_BLK) * VGL_HC_LANES + l) * VGL_HC_COL_BLK + c % VGL_HC_COL_BLK];
                const uint32_t nd = VGL_HC_COL_NODE(rec);
                const bool own = nd != VGL_HC_NO_NODE;
                const size_t at = ((size_t)(c / VGL_HC_WIRE_BLK) * VGL_HC_LANES + l) * VGL_HC_WIRE_BLK + c % VGL_HC_WIRE_BLK;
                gcol[at] = (uint16_t)((rec & 0x7fffu) | (own ? VGL_HC_WIRE_OWNED : 0u));
                if (rec & VGL_HC_COL_OPEN) {
                    gnode[((size_t)(k / VGL_HC_WIRE_BLK) * VGL_HC_LANES + l) * VGL_HC_WIRE_BLK + k % VGL_HC_WIRE_BLK] = (uint16_t)nd;
                    ++k;
                }
                // narrow form of the same record
                if (VGL_HC_COL_QUAL(rec) > 63u || (rec & VGL_HC_COL_LATCHED) || ((rec & VGL_HC_COL_VALID) && !own) || (seen_owned && !own)) narrow = false;
                if (own) { seen_owned = true; ++owned; }
                gc8[at] = (uint8_t)((VGL_HC_COL_QUAL(rec) & 0x3fu) | (VGL_HC_COL_BASE(rec) << 6));
                const uint32_t nib = ((rec & VGL_HC_COL_MISMATCH) ? 1u : 0u) | ((rec & VGL_HC_COL_OPEN) ? 2u : 0u) |
                                     ((rec & VGL_HC_COL_CLOSE) ? 4u : 0u) | ((rec & VGL_HC_COL_VALID) ? 8u : 0u);
                gf4[(size_t)(c / VGL_HC_WIRE_BLK) * VGL_HC_LANES + l] |= nib << (4 * (c % VGL_HC_WIRE_BLK));
            }
            n_owned[(size_t)g * VGL_HC_LANES + l] = (uint16_t)owned;
        }
    }
    vgl_hc_wire& v = W->view;
    v.n_reads = b.n_reads; v.n_groups = b.n_groups;
    v.n_col_blks = g_c16[G]; v.n_node_blks = g_n16[G]; v.n_map_blks = b.n_map_blks;
    v.grp_col_blk = gc; v.grp_node_blk = gn; v.grp_map_blk = gm; v.col = col; v.node = node; v.map = map;
    v.n_cols = (const uint16_t*)(base + o_nc); v.n_close = (const uint16_t*)(base + o_ncl); v.n_extra = (const uint16_t*)(base + o_nx);
    v.mapq = base + o_mq;
    if (narrow) { v.col8 = col8; v.flag4 = flag4; v.n_owned = n_owned; }
    v.meta = b.meta;
    *out = W;
    return VGH_OK;
}
const vgl_hc_wire* vgh_hc_wire_view(const vgh_hc_wire* w) { return w ? &w->view : nullptr; }
void vgh_hc_wire_destroy(vgh_hc_wire* w) {
    if (!w) return;
    if (w->block) {
        if (w->pinned) vgl_host_free(w->block);
        else free(w->block);
    }
    delete w;
}

// ------------------------------------------------------------------------------------------------
// raw alignments for the device-side reconstruction
// ------------------------------------------------------------------------------------------------
}  // extern "C"

struct vgh_hc_raw {
    vgl_hc_raw view{};
    bool pinned = false;
    int64_t skipped = 0;
    void* block = nullptr;
};

extern "C" {

static int raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, bool euka_rule, vgh_hc_raw** out);
static int vgh_hc_raw_build__impl(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out) {
    return raw_build(a, r0, r1, pinned, false, out);
}
static int vgh_euka_raw_build__impl(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out) {
    return raw_build(a, r0, r1, pinned, true, out);
}
static int raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, bool euka_rule, vgh_hc_raw** out) {
    if (!a || !out || r0 < 0 || r1 > a->n_reads || r0 > r1) return fail(VGH_ERR_INVALID, "bad arguments");
    std::vector<int64_t> keep;
    keep.reserve((size_t)(r1 - r0));
    for (int64_t r = r0; r < r1; ++r)
        if (euka_rule ? a->identity[r] != 0 : !(a->identity[r] < 1e-10)) keep.push_back(r);
    const size_t R = keep.size();
    // per-read extents, then exclusive prefix sums: every read's slice of every array is known before the (parallel) fill
    std::vector<uint64_t> pm(R + 1, 0), pe(R + 1, 0), pq(R + 1, 0), ps(R + 1, 0), pc(R + 1, 0);
    int err = VGH_OK;
    std::string err_msg;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)R; ++k) {
        const int64_t r = keep[(size_t)k];
        const int64_t e0 = a->edit_off[a->map_off[r]], e1 = a->edit_off[a->map_off[r + 1]];
        pm[(size_t)k + 1] = (uint64_t)(a->map_off[r + 1] - a->map_off[r]);
        pe[(size_t)k + 1] = (uint64_t)(e1 - e0);
        pq[(size_t)k + 1] = (uint64_t)(a->qual_off[r + 1] - a->qual_off[r]);
        ps[(size_t)k + 1] = (uint64_t)(a->edit_seq_off[e1] - a->edit_seq_off[e0]);
        uint64_t cap = ps[(size_t)k + 1];
        int e = VGH_OK;
        for (int64_t x = e0; x < e1; ++x) {
            const int64_t f = a->edit_from[x], t = a->edit_to[x];
            if (f < 0 || f > 65535 || t < 0 || t > 65535) { e = VGH_ERR_BAD_READ; break; }
            cap += (uint64_t)(f + t);
        }
        if (!e && (a->mapq[r] < 0 || a->mapq[r] > 255)) e = VGH_ERR_BAD_READ;
        for (int64_t m = a->map_off[r]; !e && m < a->map_off[r + 1]; ++m)
            if (a->map_node[m] < 0 || a->map_node[m] > 0x7fffffffll || a->map_offset[m] < 0 || a->map_offset[m] > 0xffffffffll) e = VGH_ERR_BAD_READ;
        pc[(size_t)k + 1] = (cap + 3) & ~(uint64_t)3;  // every read's strings start 4-byte aligned on the device
        if (e) {
#pragma omp critical(vgh_raw_err)
            if (!err) { err = e; err_msg = "read " + std::to_string(r) + ": mapping quality, node id, offset or edit length out of range"; }
        }
    }
    if (err) return fail(err, "%s", err_msg.c_str());
    for (size_t k = 0; k < R; ++k) { pm[k + 1] += pm[k]; pe[k + 1] += pe[k]; pq[k + 1] += pq[k]; ps[k + 1] += ps[k]; pc[k + 1] += pc[k]; }
    const uint64_t nm = pm[R], ne = pe[R], nq = pq[R], ns = ps[R];
    if (R >= 0x7fffffffull || nm >= 0xffffffffull || ne >= 0xffffffffull || nq >= 0xffffffffull || ns >= 0xffffffffull || pc[R] >= 0xffffffffull)
        return fail(VGH_ERR_INVALID, "batch too large for 32-bit offsets; split it");
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    const size_t o_qoff = take((R + 1) * 4), o_soff = take((R + 1) * 4), o_moff = take((R + 1) * 4), o_mq = take(R), o_q = take(nq);
    const size_t o_node = take(nm * 4), o_offs = take(nm * 4), o_eoff = take((nm + 1) * 4), o_from = take(ne * 2), o_to = take(ne * 2);
    const size_t o_esoff = take((ne + 1) * 4), o_eseq = take(ns), o_slen = take(R * 2), o_ridx = take(R * 4);
    const size_t o_off16 = take(nm * 2), o_ne8 = take(nm), o_sl8 = take(ne), o_f8 = take(ne), o_t8 = take(ne), o_n16 = take(nm * 2);
    const size_t o_fo16 = take(R * 2);
    vgh_hc_raw* P = new vgh_hc_raw();
    P->pinned = pinned != 0;
    const size_t bytes = std::max<size_t>(o, 256);
    if (pinned) {
        if (vgl_host_alloc(&P->block, bytes) != VGL_OK) {
            delete P;
            return fail(VGH_ERR_INVALID, "pinned allocation failed: %s", vgl_last_error());
        }
    } else {
        P->block = aligned_alloc(256, bytes);
        if (!P->block) { delete P; return fail(VGH_ERR_INVALID, "out of host memory"); }
    }
    uint8_t* base = (uint8_t*)P->block;
    uint32_t* qoff = (uint32_t*)(base + o_qoff); uint32_t* soff = (uint32_t*)(base + o_soff); uint32_t* moff = (uint32_t*)(base + o_moff);
    uint8_t* mq = base + o_mq; uint8_t* q = base + o_q;
    uint32_t* node = (uint32_t*)(base + o_node); uint32_t* offs = (uint32_t*)(base + o_offs); uint32_t* eoff = (uint32_t*)(base + o_eoff);
    uint16_t* from = (uint16_t*)(base + o_from); uint16_t* to = (uint16_t*)(base + o_to);
    uint32_t* esoff = (uint32_t*)(base + o_esoff); uint8_t* eseq = base + o_eseq;
    uint16_t* slen = (uint16_t*)(base + o_slen); uint32_t* ridx = (uint32_t*)(base + o_ridx);
    uint16_t* off16 = (uint16_t*)(base + o_off16); uint8_t* ne8 = base + o_ne8; uint8_t* sl8 = base + o_sl8;
    uint8_t* f8 = base + o_f8; uint8_t* t8 = base + o_t8; uint16_t* n16 = (uint16_t*)(base + o_n16);
    uint16_t* fo16 = (uint16_t*)(base + o_fo16);
    bool first_only = true;   // only a read's first mapping starts inside its node (offset != 0)
    bool narrow_node = true;  // every node id <= 32767
    bool narrow_len = true;  // every from_length / to_length <= 255
    bool narrow = true;  // every offset <= 65535, <= 255 edits per mapping, edit sequences <= 255 bytes
#pragma omp parallel for schedule(static) reduction(&& : narrow, narrow_len, narrow_node, first_only)
    for (int64_t k = 0; k < (int64_t)R; ++k) {
        const int64_t r = keep[(size_t)k];
        uint64_t cm = pm[(size_t)k], ce = pe[(size_t)k], cs = ps[(size_t)k];
        qoff[k] = (uint32_t)pq[(size_t)k]; soff[k] = (uint32_t)pc[(size_t)k]; moff[k] = (uint32_t)cm;
        mq[k] = (uint8_t)a->mapq[r];
        slen[k] = (uint16_t)std::min<int64_t>(a->seq_off[r + 1] - a->seq_off[r], 65535);
        ridx[k] = (uint32_t)(r - r0);
        memcpy(q + pq[(size_t)k], a->qual + a->qual_off[r], (size_t)(pq[(size_t)k + 1] - pq[(size_t)k]));
        fo16[k] = a->map_off[r + 1] > a->map_off[r] ? (uint16_t)a->map_offset[a->map_off[r]] : (uint16_t)0;
        for (int64_t m = a->map_off[r]; m < a->map_off[r + 1]; ++m) {
            first_only = first_only && (m == a->map_off[r] || a->map_offset[m] == 0);
            node[cm] = (uint32_t)a->map_node[m] | (a->map_rev[m] ? 0x80000000u : 0u);
            offs[cm] = (uint32_t)a->map_offset[m];
            eoff[cm] = (uint32_t)ce;
            const int64_t n_ed = a->edit_off[m + 1] - a->edit_off[m];
            off16[cm] = (uint16_t)a->map_offset[m];
            n16[cm] = (uint16_t)((a->map_node[m] & 0x7fff) | (a->map_rev[m] ? 0x8000 : 0));
            narrow_node = narrow_node && a->map_node[m] <= 32767;
            ne8[cm] = (uint8_t)n_ed;
            narrow = narrow && a->map_offset[m] <= 65535 && n_ed <= 255;
            ++cm;
            for (int64_t e = a->edit_off[m]; e < a->edit_off[m + 1]; ++e) {
                const int64_t sl = a->edit_seq_off[e + 1] - a->edit_seq_off[e];
                from[ce] = (uint16_t)a->edit_from[e];
                to[ce] = (uint16_t)a->edit_to[e];
                f8[ce] = (uint8_t)a->edit_from[e];
                t8[ce] = (uint8_t)a->edit_to[e];
                narrow_len = narrow_len && a->edit_from[e] <= 255 && a->edit_to[e] <= 255;
                esoff[ce] = (uint32_t)cs;
                sl8[ce] = (uint8_t)sl;
                narrow = narrow && sl <= 255;
                memcpy(eseq + cs, a->edit_seq + a->edit_seq_off[e], (size_t)sl);
                cs += (uint64_t)sl;
                ++ce;
            }
        }
    }
    const uint64_t cq = nq, ccap = pc[R], cm = nm, ce = ne, cs = ns;
    qoff[R] = (uint32_t)cq; soff[R] = (uint32_t)ccap; moff[R] = (uint32_t)cm; eoff[cm] = (uint32_t)ce; esoff[ce] = (uint32_t)cs;
    P->skipped = (r1 - r0) - (int64_t)R;
    vgl_hc_raw& v = P->view;
    v.n_reads = (uint32_t)R; v.n_maps = (uint32_t)cm; v.n_edits = (uint32_t)ce;
    v.qual_off = qoff; v.qual = q; v.mapq = mq; v.str_off = soff; v.map_off = moff; v.map_node = node; v.map_offset = offs;
    v.edit_off = eoff; v.edit_from = from; v.edit_to = to; v.edit_seq_off = esoff; v.edit_seq = eseq;
    v.seq_len = slen; v.read_index = ridx;
    if (narrow) { v.map_offset16 = off16; v.map_n_edits8 = ne8; v.edit_slen8 = sl8; }
    if (narrow && first_only) v.first_offset16 = fo16;  // (narrow: the first offsets fit 16 bits too)
    if (narrow_len) { v.edit_from8 = f8; v.edit_to8 = t8; }
    if (narrow_node) v.map_node16 = n16;
    if (!euka_rule) reads_meta(a, keep, v.meta);
    *out = P;
    return VGH_OK;
}
const vgl_hc_raw* vgh_hc_raw_view(const vgh_hc_raw* p) { return p ? &p->view : nullptr; }
int64_t vgh_hc_raw_skipped(const vgh_hc_raw* p) { return p ? p->skipped : 0; }
void vgh_hc_raw_destroy(vgh_hc_raw* p) {
    if (!p) return;
    if (p->block) {
        if (p->pinned) vgl_host_free(p->block);
        else free(p->block);
    }
    delete p;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// feeder: the drop-in for the OpenMP read loop of src/HaploCart.cpp:408-421 when the caller holds the flat
// alignment fields (what FlatAlignments::assign yields).  Reads go to the GPU in chunks: while the device copies,
// reconstructs and scores chunk c, the host marshals chunk c + 1 into the other of two page-locked blocks that live
// as long as the feeder (no allocation, no page faults per sample).  Only the narrow arrays of vgl_hc_raw are
// written (node ids as 16 bits, one node offset per read, edit counts / lengths / sequence lengths as one byte
// each); a chunk whose values do not fit, or that holds a read the hot loop skips (identity < 1e-10), goes through
// vgh_hc_raw_build instead.
// ------------------------------------------------------------------------------------------------
struct vgh_hc_feeder {
    int n_threads = 1;
    int64_t chunk_reads = 131072;
    void* block[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    bool pinned[2] = {false, false};
    vgl_hc_raw view[2];
    double build_s = 0, wait_s = 0;
    uint64_t bytes = 0, chunks = 0, slow_chunks = 0, reads = 0;
    bool wide_nodes = false;   // a node id above 32767 was seen: node ids travel as 32 bits from then on (euka's graphs)
};

extern "C" {

static int vgh_hc_feeder_create__impl(int n_threads, int64_t chunk_reads, vgh_hc_feeder** out) {
    if (!out) return fail(VGH_ERR_INVALID, "null output");
    vgh_hc_feeder* f = new (std::nothrow) vgh_hc_feeder();
    if (!f) return fail(VGH_ERR_INVALID, "out of host memory");
#ifdef _OPENMP
    // auto: three quarters of the OpenMP threads -- the submitting thread and the driver's own threads need cores too (measured:
    // 16 marshalling threads on 16 cores 28.5 ms per 10^6 reads, 12 threads 23.7 ms)
    f->n_threads = n_threads > 0 ? n_threads : std::max(1, omp_get_max_threads() * 3 / 4);
#endif
    if (chunk_reads > 0) f->chunk_reads = std::min<int64_t>(chunk_reads, 131072);  // one device chunk per submission
    memset(f->view, 0, sizeof(f->view));
    *out = f;
    return VGH_OK;
}
void vgh_hc_feeder_destroy(vgh_hc_feeder* f) {
    if (!f) return;
    for (int k = 0; k < 2; ++k)
        if (f->block[k]) { if (f->pinned[k]) vgl_host_free(f->block[k]); else free(f->block[k]); }
    delete f;
}
void vgh_hc_feeder_stats(const vgh_hc_feeder* f, double* build_seconds, double* wait_seconds, uint64_t* h2d_bytes, uint64_t* chunks,
                         uint64_t* slow_chunks) {
    if (!f) return;
    if (build_seconds) *build_seconds = f->build_s;
    if (wait_seconds) *wait_seconds = f->wait_s;
    if (h2d_bytes) *h2d_bytes = f->bytes;
    if (chunks) *chunks = f->chunks;
    if (slow_chunks) *slow_chunks = f->slow_chunks;
}

// The narrowing copies of the feeder, one flat loop per source array.  Compiled once per vector width the host may have
// (resolved when the library is loaded): 64-bit -> 16 / 8-bit narrowing is one instruction per vector from AVX-512 on.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define VGH_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define VGH_CLONES
#endif

// node ids as 16 bits (HaploCart's graph: ids <= 32767) or 32 bits (euka's graphs); the top bit carries is_reverse
#define VGH_FEEDER_NARROW_MAPPINGS(NAME, NodeT, ID_MASK, REV_BIT)                                                                          \
    VGH_CLONES static void NAME(const int64_t* __restrict__ node, const int64_t* __restrict__ offs, const uint8_t* __restrict__ rev,       \
                                const int64_t* __restrict__ eo, NodeT* __restrict__ on, uint8_t* __restrict__ oc, int64_t n,              \
                                uint64_t* viol, uint64_t* node_viol, uint64_t* nonzero_off) {                                             \
        uint64_t v = 0, nv = 0, nzo = 0;                                                                                                   \
        for (int64_t i = 0; i < n; ++i) {                                                                                                  \
            const uint64_t nd = (uint64_t)node[i], cnt = (uint64_t)(eo[i + 1] - eo[i]);                                                    \
            v |= cnt & ~(uint64_t)0xff;                                                                                                    \
            nv |= nd & ~(uint64_t)(ID_MASK);                                                                                               \
            on[i] = (NodeT)((nd & (ID_MASK)) | (rev[i] ? (REV_BIT) : 0ull));                                                               \
            oc[i] = (uint8_t)cnt;                                                                                                          \
            nzo += offs[i] != 0;                                                                                                           \
        }                                                                                                                                  \
        *viol |= v;                                                                                                                        \
        *node_viol |= nv;                                                                                                                  \
        *nonzero_off += nzo;                                                                                                               \
    }
VGH_FEEDER_NARROW_MAPPINGS(feeder_narrow_mappings16, uint16_t, 0x7fffull, 0x8000ull)
VGH_FEEDER_NARROW_MAPPINGS(feeder_narrow_mappings32, uint32_t, 0x7fffffffull, 0x80000000ull)
#undef VGH_FEEDER_NARROW_MAPPINGS

VGH_CLONES static void feeder_narrow_edits(const int32_t* __restrict__ fr, const int32_t* __restrict__ to, const int64_t* __restrict__ so,
                                           uint8_t* __restrict__ of, uint8_t* __restrict__ ot, uint8_t* __restrict__ os, int64_t n, uint64_t* viol) {
    uint64_t v = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t f_ = (uint32_t)fr[i], t_ = (uint32_t)to[i];
        const uint64_t sl = (uint64_t)(so[i + 1] - so[i]);
        v |= (uint64_t)((f_ | t_) & ~0xffu) | (sl & ~(uint64_t)0xff);
        of[i] = (uint8_t)f_;
        ot[i] = (uint8_t)t_;
        os[i] = (uint8_t)sl;
    }
    *viol |= v;
}

// largest quality byte as the reference sees it (signed char: bytes >= 128 are negative, canonical 0)
VGH_CLONES static int8_t feeder_max_quality(const int8_t* __restrict__ qb, int64_t n, int8_t mx) {
    for (int64_t i = 0; i < n; ++i) mx = qb[i] > mx ? qb[i] : mx;
    return mx;
}

static constexpr size_t kFeederRunReads = 128;

// Marshals the reads of [ra, rb) that the hot loop keeps (identity >= 1e-10, src/HaploCart.cpp:411) into block `slot`.
// Two parallel passes over per-thread ranges of reads: (A) how many mappings / edits / bytes each range keeps (read-level
// offsets only), (B) the copy itself.  Returns 1 when the chunk needs the general builder instead.
//   mode 0: HaploCart's rule (identity >= 1e-10 is kept, src/HaploCart.cpp:411) and the batch's value ranges (meta);
//   mode 1: euka / soibean (identity != 0 is kept, src/readGAM_Euka.h:73) plus |Alignment.sequence| and the read's position in
//           [ra, rb) per read, which their kernels need (vgl_euka_run_raw / vgl_soibean_run_raw).
static int feeder_fill(vgh_hc_feeder* f, int slot, const vgh_alignments* a, int64_t ra, int64_t rb, int mode = 0) {
    const size_t Rall = (size_t)(rb - ra);
    const auto skipped = [a, mode](int64_t r) { return mode == 0 ? a->identity[r] < 1e-10 : a->identity[r] == 0; };
    const bool wide = f->wide_nodes;
    int T = std::max(1, f->n_threads);
    T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, Rall / 512 + 1));
    struct Cnt { uint64_t r = 0, m = 0, e = 0, q = 0, s = 0; };
    std::vector<Cnt> base((size_t)T + 1);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const size_t k0 = Rall * (size_t)t / (size_t)T, k1 = Rall * (size_t)(t + 1) / (size_t)T;
        Cnt c;
        size_t kept = 0;
        for (size_t k = k0; k < k1; ++k) kept += !skipped(ra + (int64_t)k);
        if (kept == k1 - k0) {   // the usual case, nothing skipped: the counts are differences of the offsets at the range's ends
            const int64_t rs = ra + (int64_t)k0, re = ra + (int64_t)k1;
            const int64_t ma = a->map_off[rs], mb = a->map_off[re];
            const int64_t ea = a->edit_off[ma], eb = a->edit_off[mb];
            c.r = kept; c.m = (uint64_t)(mb - ma); c.e = (uint64_t)(eb - ea);
            c.q = (uint64_t)(a->qual_off[re] - a->qual_off[rs]);
            c.s = (uint64_t)(a->edit_seq_off[eb] - a->edit_seq_off[ea]);
        } else {
            for (size_t k = k0; k < k1; ++k) {
                const int64_t r = ra + (int64_t)k;
                if (skipped(r)) continue;
                const int64_t ma = a->map_off[r], mb = a->map_off[r + 1];
                const int64_t ea = a->edit_off[ma], eb = a->edit_off[mb];
                c.r += 1; c.m += (uint64_t)(mb - ma); c.e += (uint64_t)(eb - ea);
                c.q += (uint64_t)(a->qual_off[r + 1] - a->qual_off[r]);
                c.s += (uint64_t)(a->edit_seq_off[eb] - a->edit_seq_off[ea]);
            }
        }
        base[(size_t)t + 1] = c;
    }
    for (int t = 0; t < T; ++t) {
        base[(size_t)t + 1].r += base[(size_t)t].r; base[(size_t)t + 1].m += base[(size_t)t].m; base[(size_t)t + 1].e += base[(size_t)t].e;
        base[(size_t)t + 1].q += base[(size_t)t].q; base[(size_t)t + 1].s += base[(size_t)t].s;
    }
    const size_t R = (size_t)base[(size_t)T].r, nm = (size_t)base[(size_t)T].m, ne = (size_t)base[(size_t)T].e;
    const size_t nq = (size_t)base[(size_t)T].q, ns = (size_t)base[(size_t)T].s;
    if (nm >= 0xffffffffull || ne >= 0xffffffffull || nq >= 0xffffffffull || ns >= 0xffffffffull) return 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    const size_t o_qoff = take((R + 1) * 4), o_soff = take((R + 1) * 4), o_moff = take((R + 1) * 4), o_mq = take(R), o_q = take(nq);
    const size_t o_n16 = take(nm * (wide ? 4 : 2)), o_fo16 = take(R * 2), o_ne8 = take(nm), o_sl8 = take(ne), o_f8 = take(ne), o_t8 = take(ne);
    const size_t o_eseq = take(ns), o_ends = take(64);
    const size_t o_slen = take(mode ? R * 2 : 0), o_ridx = take(mode ? R * 4 : 0);
    if (o > f->cap[slot]) {
        if (f->block[slot]) { if (f->pinned[slot]) vgl_host_free(f->block[slot]); else free(f->block[slot]); }
        f->block[slot] = nullptr;
        f->cap[slot] = 0;
        const size_t want = align_up(o + o / 4, 1 << 20);
        f->pinned[slot] = vgl_host_alloc(&f->block[slot], want) == VGL_OK;
        if (!f->pinned[slot]) {   // no CUDA device in this process (host-only tests): ordinary memory holds the same bytes
            f->block[slot] = aligned_alloc(4096, want);
            if (!f->block[slot]) return fail(VGH_ERR_INVALID, "out of host memory");
        }
        f->cap[slot] = want;
    }
    uint8_t* blk = (uint8_t*)f->block[slot];
    uint32_t* qoff = (uint32_t*)(blk + o_qoff); uint32_t* soff = (uint32_t*)(blk + o_soff); uint32_t* moff = (uint32_t*)(blk + o_moff);
    uint8_t* mq = blk + o_mq; uint8_t* q = blk + o_q;
    uint16_t* n16 = (uint16_t*)(blk + o_n16); uint32_t* n32 = (uint32_t*)(blk + o_n16); uint16_t* fo16 = (uint16_t*)(blk + o_fo16);
    uint16_t* slen = (uint16_t*)(blk + o_slen); uint32_t* ridx = (uint32_t*)(blk + o_ridx);
    uint8_t* ne8 = blk + o_ne8; uint8_t* sl8 = blk + o_sl8; uint8_t* f8 = blk + o_f8; uint8_t* t8 = blk + o_t8; uint8_t* eseq = blk + o_eseq;
    // Pass B works on RUNS of consecutive kept reads (a chunk is one run unless it holds a read the hot loop skips): inside a
    // run the mappings, edits, quality bytes and edit sequences are contiguous in the source arrays, so the narrowing copies
    // are flat loops over whole arrays (the compiler vectorises them) and the range checks OR-accumulate instead of branching.
    int bad = 0, bad_node = 0;
    std::vector<MetaAcc> metas((size_t)T);
#pragma omp parallel for schedule(static, 1) num_threads(T) reduction(| : bad, bad_node)
    for (int t = 0; t < T; ++t) {
        const size_t k0 = Rall * (size_t)t / (size_t)T, k1 = Rall * (size_t)(t + 1) / (size_t)T;
        MetaAcc& meta = metas[(size_t)t];
        size_t cr = (size_t)base[(size_t)t].r, cm = (size_t)base[(size_t)t].m, ce = (size_t)base[(size_t)t].e;
        size_t cq = (size_t)base[(size_t)t].q, cs = (size_t)base[(size_t)t].s;
        uint64_t viol = 0;        // any bit set: a value does not fit its narrow field
        uint64_t node_viol = 0;   // ... a node id does not fit the node field in use
        uint64_t nonzero_off = 0, nonzero_first = 0;   // mappings with a non-zero node offset; reads whose FIRST mapping has one
        int8_t qmax = 0;          // largest quality byte as signed char (bytes >= 128 are negative: canonical 0)
        size_t k = k0;
        while (k < k1) {
            if (skipped(ra + (int64_t)k)) { ++k; continue; }
            size_t ke = k + 1;
            const size_t kcap = std::min(k1, k + kFeederRunReads);   // short runs: the per-read pass below finds the run's lines still in L2
            while (ke < kcap && !skipped(ra + (int64_t)ke)) ++ke;
            const int64_t rs = ra + (int64_t)k, re = ra + (int64_t)ke;   // the run [rs, re)
            const int64_t M0 = a->map_off[rs], M1 = a->map_off[re], E0 = a->edit_off[M0], E1 = a->edit_off[M1];
            const int64_t Q0 = a->qual_off[rs], Q1 = a->qual_off[re], S0 = a->edit_seq_off[E0], S1 = a->edit_seq_off[E1];
            if (wide)
                feeder_narrow_mappings32(a->map_node + M0, a->map_offset + M0, a->map_rev + M0, a->edit_off + M0, n32 + cm, ne8 + cm, M1 - M0,
                                         &viol, &node_viol, &nonzero_off);
            else
                feeder_narrow_mappings16(a->map_node + M0, a->map_offset + M0, a->map_rev + M0, a->edit_off + M0, n16 + cm, ne8 + cm, M1 - M0,
                                         &viol, &node_viol, &nonzero_off);
            feeder_narrow_edits(a->edit_from + E0, a->edit_to + E0, a->edit_seq_off + E0, f8 + ce, t8 + ce, sl8 + ce, E1 - E0, &viol);
            memcpy(q + cq, a->qual + Q0, (size_t)(Q1 - Q0));
            memcpy(eseq + cs, a->edit_seq + S0, (size_t)(S1 - S0));
            qmax = feeder_max_quality((const int8_t*)a->qual + Q0, Q1 - Q0, qmax);
            // per read: offsets, mapping quality, first node offset, capacity of its aligned strings
            for (int64_t r = rs; r < re; ++r) {
                const int64_t ma = a->map_off[r], mb = a->map_off[r + 1];
                const int64_t ea = a->edit_off[ma], eb = a->edit_off[mb];
                const uint32_t mqv = (uint32_t)a->mapq[r];
                viol |= (uint64_t)(mqv & ~0xffu);
                qoff[cr] = (uint32_t)(cq + (size_t)(a->qual_off[r] - Q0));
                moff[cr] = (uint32_t)(cm + (size_t)(ma - M0));
                mq[cr] = (uint8_t)mqv;
                meta.mapq(mqv);
                const uint64_t fo = mb > ma ? (uint64_t)a->map_offset[ma] : 0;
                viol |= fo & ~(uint64_t)0xffff;
                nonzero_first += fo != 0;
                fo16[cr] = (uint16_t)fo;
                uint32_t span = 0;   // sum over the read's edits of from_length + to_length (from the narrowed copies)
                const uint8_t* __restrict__ pf = f8 + ce + (size_t)(ea - E0);
                const uint8_t* __restrict__ pt = t8 + ce + (size_t)(ea - E0);
                for (int64_t i = 0; i < eb - ea; ++i) span += (uint32_t)pf[i] + (uint32_t)pt[i];
                const uint64_t cap = (uint64_t)span + (uint64_t)(a->edit_seq_off[eb] - a->edit_seq_off[ea]);
                soff[cr + 1] = (uint32_t)((cap + 3) & ~(uint64_t)3);  // prefix-summed below
                if (mode) {
                    slen[cr] = (uint16_t)std::min<int64_t>(a->seq_off[r + 1] - a->seq_off[r], 65535);
                    ridx[cr] = (uint32_t)(r - ra);
                }
                ++cr;
            }
            cm += (size_t)(M1 - M0); ce += (size_t)(E1 - E0); cq += (size_t)(Q1 - Q0); cs += (size_t)(S1 - S0);
            k = ke;
        }
        // only a read's first mapping may start inside its node (first_offset16 form)
        bad |= (viol != 0 || nonzero_off != nonzero_first) ? 1 : 0;
        bad_node |= node_viol != 0 ? 1 : 0;
        meta.qmax = std::max<uint32_t>(meta.qmax, qmax > 2 ? (uint32_t)qmax : 0u);
    }
    if (bad) return 1;
    if (bad_node) {   // a node id above the field in use: 32-bit node ids from now on (once per feeder), or the general builder
        if (wide) return 1;
        f->wide_nodes = true;
        return feeder_fill(f, slot, a, ra, rb, mode);
    }
    soff[0] = 0;
    uint64_t run = 0;
    for (size_t k = 0; k < R; ++k) { run += soff[k + 1]; soff[k + 1] = (uint32_t)run; }
    if (run >= 0xffffffffull) return 1;
    qoff[R] = (uint32_t)nq; moff[R] = (uint32_t)nm;
    // the two wide offset arrays are not built: the library only reads their end points for a one-chunk view
    uint32_t* ends = (uint32_t*)(blk + o_ends);
    ends[0] = 0; ends[1] = (uint32_t)ne;   // stands in for edit_off[0], edit_off[n_maps]       (see VGL_HC_RAW_NARROW_ONLY)
    ends[2] = 0; ends[3] = (uint32_t)ns;   // stands in for edit_seq_off[0], edit_seq_off[n_edits]
    vgl_hc_raw& v = f->view[slot];
    memset(&v, 0, sizeof(v));
    v.n_reads = (uint32_t)R; v.n_maps = (uint32_t)nm; v.n_edits = (uint32_t)ne;
    v.reserved = VGL_HC_RAW_NARROW_ONLY;
    v.qual_off = qoff; v.qual = q; v.mapq = mq; v.str_off = soff; v.map_off = moff;
    v.edit_off = ends; v.edit_seq_off = ends + 2;
    v.edit_seq = eseq;
    if (wide) v.map_node = n32; else v.map_node16 = n16;
    v.first_offset16 = fo16; v.map_n_edits8 = ne8; v.edit_slen8 = sl8; v.edit_from8 = f8; v.edit_to8 = t8;
    if (mode) {
        v.seq_len = slen; v.read_index = ridx;
    } else {
        MetaAcc all;
        for (const MetaAcc& m : metas) all.merge(m);
        all.store(v.meta);
    }
    return 0;
}

// One chunk marshalled into block 0, for inspection and tests: *out is NULL when the chunk needs the general builder.
static int feeder_marshal(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, int mode, const vgl_hc_raw** out) {
    if (!f || !a || !out || r0 < 0 || r1 > a->n_reads || r0 > r1 || r1 - r0 > f->chunk_reads) return fail(VGH_ERR_INVALID, "bad arguments");
    const int how = feeder_fill(f, 0, a, r0, r1, mode);
    if (how > 1) return how;
    *out = how == 0 ? &f->view[0] : nullptr;
    return VGH_OK;
}
static int vgh_hc_feeder_marshal__impl(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, const vgl_hc_raw** out) {
    return feeder_marshal(f, a, r0, r1, 0, out);
}
int vgh_euka_feeder_marshal(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, const vgl_hc_raw** out) {
    try {
        return feeder_marshal(f, a, r0, r1, 1, out);
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_euka_feeder_marshal: %s", e.what());
    }
}

}  // extern "C" (the pipeline below is a template)

// The pipeline shared by the three tools: chunk c + 1 is marshalled while a second thread submits chunk c (its H2D copies,
// kernels and -- for the sibling tools -- the copy of its per-read results back).  submit(view, ra, rb) returns a vgl status.
template <typename Submit>
static int feeder_run(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, int mode, Submit submit_chunk) {
    try {
        std::thread submit;
        int submit_rc = VGL_OK;
        int rc = VGH_OK;
        int64_t c = 0;
        const int64_t step = f->chunk_reads;   // (splitting a sample smaller than one chunk was measured: the extra launches and
                                               //  submissions cost more than the overlap gains -- 1.4 against 0.9-1.1 ms per 2e4 reads)
        for (int64_t ra = r0; ra < r1 && rc == VGH_OK; ra += step, ++c) {
            const int64_t rb = std::min(r1, ra + step);
            const int slot = (int)(c & 1);
            const auto t0 = std::chrono::steady_clock::now();
            const int how = feeder_fill(f, slot, a, ra, rb, mode);   // block `slot` was released when submission c - 2 was joined
            vgh_hc_raw* general = nullptr;
            if (how == 1) {
                ++f->slow_chunks;
                rc = raw_build(a, ra, rb, /*pinned=*/0, /*euka_rule=*/mode != 0, &general);
            } else if (how != 0) {
                rc = how;
            }
            const auto t1 = std::chrono::steady_clock::now();
            f->build_s += std::chrono::duration<double>(t1 - t0).count();
            if (submit.joinable()) submit.join();
            f->wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            if (submit_rc != VGL_OK) { rc = 1000 + submit_rc; }
            if (rc != VGH_OK) { if (general) vgh_hc_raw_destroy(general); break; }
            ++f->chunks;
            f->reads += (uint64_t)(rb - ra);
            if (general) {   // rare: synchronous, the block is freed right after the copy
                f->bytes += vgl_hc_raw_bytes(&general->view);
                submit_rc = submit_chunk(&general->view, ra, rb);
                vgh_hc_raw_destroy(general);
                if (submit_rc != VGL_OK) rc = 1000 + submit_rc;
            } else {
                f->bytes += vgl_hc_raw_bytes(&f->view[slot]);
                const vgl_hc_raw* v = &f->view[slot];
                submit = std::thread([v, ra, rb, &submit_rc, &submit_chunk]() { submit_rc = submit_chunk(v, ra, rb); });
            }
        }
        if (submit.joinable()) submit.join();
        if (rc == VGH_OK && submit_rc != VGL_OK) rc = 1000 + submit_rc;
        if (rc >= 1000) return fail(VGH_ERR_DEVICE, "likelihood library: %s", vgl_last_error());
        return rc;
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "feeder: %s", e.what());
    }
}

extern "C" {

static int vgh_hc_feed__impl(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1) {
    if (!f || !ctx || !a || r0 < 0 || r1 > a->n_reads || r0 > r1) return fail(VGH_ERR_INVALID, "bad arguments");
    return feeder_run(f, a, r0, r1, 0, [ctx](const vgl_hc_raw* v, int64_t, int64_t) { return vgl_hc_accumulate_raw(ctx, v); });
}

static int vgh_euka_feed__impl(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t* clade,
                               double* in_lik, double* not_lik, double* like, uint8_t* pass) {
    if (!f || !ctx || !a || r0 < 0 || r1 > a->n_reads || r0 > r1) return fail(VGH_ERR_INVALID, "bad arguments");
    return feeder_run(f, a, r0, r1, 1, [=](const vgl_hc_raw* v, int64_t ra, int64_t rb) {
        const int64_t at = ra - r0;   // read_index counts from the chunk's first read
        return vgl_euka_run_raw(ctx, v, (uint32_t)(rb - ra), clade ? clade + at : nullptr, in_lik ? in_lik + at : nullptr,
                                not_lik ? not_lik + at : nullptr, like ? like + at : nullptr, pass ? pass + at : nullptr);
    });
}

static int vgh_soibean_feed__impl(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t n_paths,
                                  double* path_ll, uint8_t* used) {
    if (!f || !ctx || !a || r0 < 0 || r1 > a->n_reads || r0 > r1 || n_paths < 1) return fail(VGH_ERR_INVALID, "bad arguments");
    return feeder_run(f, a, r0, r1, 1, [=](const vgl_hc_raw* v, int64_t ra, int64_t rb) {
        const int64_t at = ra - r0;
        return vgl_soibean_run_raw(ctx, v, (uint32_t)(rb - ra), path_ll ? path_ll + at * (int64_t)n_paths : nullptr, used ? used + at : nullptr);
    });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// HaploCart tables
// ------------------------------------------------------------------------------------------------
struct vgh_hc_tables {
    vgl_hc_tables view{};
    std::vector<uint32_t> bits;
    std::vector<double> node_mappability, node_match_prob;
    std::vector<uint8_t> node_valid;
};

namespace {

// Per-year substitution rates by rCRS region, in the precedence order of get_p_obs_base
// (reference src/get_p_obs_base.cpp:43-61).  The protein-coding entry is exactly 0 because the
// reference multiplies by the integer quotients (2/3) and (1/3).
struct Region { int lo, hi; double mu; };
const Region kRegions[] = {
    {57, 372, 1.64273e-7},                                                                   // HVS I
    {1, 56, 2.29640e-8}, {373, 576, 2.29640e-8},                                             // HVS II
    {16384, 16569, 1.54555e-8},                                                              // rest of the control region
    {3307, 4262, 0.0}, {4470, 5511, 0.0}, {5904, 7445, 0.0}, {7586, 8269, 0.0}, {8366, 9990, 0.0},
    {10059, 10403, 0.0}, {10470, 12137, 0.0}, {12337, 14673, 0.0}, {14747, 15886, 0.0},      // protein coding
    {577, 647, 6.91285e-9}, {1602, 1670, 6.91285e-9}, {3230, 3304, 6.91285e-9}, {4263, 4400, 6.91285e-9},
    {4402, 4469, 6.91285e-9}, {5512, 5579, 6.91285e-9}, {5587, 5654, 6.91285e-9}, {5657, 5728, 6.91285e-9},
    {5761, 5891, 6.91285e-9}, {7446, 7514, 6.91285e-9}, {7518, 7585, 6.91285e-9}, {8295, 8364, 6.91285e-9},
    {15888, 15953, 6.91285e-9}, {15956, 16023, 6.91285e-9},                                  // tRNA
    {648, 1601, 6.91285e-9}, {1671, 3229, 6.91285e-9},                                       // rRNA
};
const double kMuElsewhere = 2.48537e-8;

double region_rate_per_generation(int32_t pangenome_base) {
    double mu = kMuElsewhere;
    const unsigned x = (unsigned)pangenome_base;  // the reference compares as unsigned
    for (const Region& r : kRegions)
        if ((unsigned)r.lo <= x && x <= (unsigned)r.hi) { mu = r.mu; break; }
    return mu * 30;  // years -> generations
}

}  // namespace

extern "C" {

double vgh_hc_region_match_prob(int32_t pangenome_base) {
    return std::pow((1 - region_rate_per_generation(pangenome_base)), 8);
}

static int vgh_hc_tables_build__impl(int32_t n_rows, int32_t n_paths, const uint8_t* supports, int64_t n_pmap,
                        const int64_t* pmap_node, const int32_t* pmap_val, int64_t n_mappab, const double* mappab,
                        double background_error_prob, int is_consensus_fasta, vgh_hc_tables** out) {
    if (!out || n_rows <= 0 || n_paths <= 0 || !supports || (n_pmap && (!pmap_node || !pmap_val)) || (n_mappab && !mappab))
        return fail(VGH_ERR_INVALID, "bad table arguments");
    vgh_hc_tables* T = new vgh_hc_tables();
    const int W = (n_paths + 31) / 32;
    T->bits.assign((size_t)n_rows * (size_t)W, 0u);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < n_rows; ++n) {
        const uint8_t* row = supports + (size_t)n * (size_t)n_paths;
        uint32_t* dst = T->bits.data() + (size_t)n * (size_t)W;
        for (int p = 0; p < n_paths; ++p)
            if (row[p]) dst[p >> 5] |= 1u << (p & 31);
    }
    T->node_mappability.assign((size_t)n_rows, 0.0);
    T->node_match_prob.assign((size_t)n_rows, 0.0);
    T->node_valid.assign((size_t)n_rows, 0);
    for (int64_t k = 0; k < n_pmap; ++k) {
        const int64_t node = pmap_node[k];
        const int32_t pb = pmap_val[k];
        if (node < 0 || node >= n_rows) continue;  // entry for a node outside the support matrix: unreachable from the hot loop
        if (T->node_valid[(size_t)node]) continue;  // std::map::insert keeps the first entry (src/load.cpp:39)
        if (pb < 0 || pb >= n_mappab) {
            delete T;
            return fail(VGH_ERR_INVALID, "pangenome coordinate %d of node %lld is outside the mappability vector (%lld entries)", pb,
                        (long long)node, (long long)n_mappab);
        }
        T->node_valid[(size_t)node] = 1;
        T->node_mappability[(size_t)node] = mappab[pb];
        T->node_match_prob[(size_t)node] = vgh_hc_region_match_prob(pb);
    }
    vgl_hc_tables& v = T->view;
    v.n_rows = n_rows;
    v.n_paths = n_paths;
    v.row_words = W;
    v.support_bits = T->bits.data();
    v.node_mappability = T->node_mappability.data();
    v.node_match_prob = T->node_match_prob.data();
    v.node_valid = T->node_valid.data();
    for (int Q = 0; Q < 100; ++Q) {
        // get_qscore_vec (src/miscfunc.h:198-212): Q >= 2 goes through get_p_seq_error, which itself
        // returns 0.25 up to Q == 2 (:180-187)
        v.qscore[Q] = (Q > 2) ? std::pow(10, ((-1 * Q) * 0.1)) : 0.25;
        v.p_incorrect_map[Q] = std::pow(10, ((-1 * Q) * 0.1));  // src/haplocart_functions.cpp:100-107
    }
    v.bg_freq[0] = 0.27532; v.bg_freq[1] = 0.30044; v.bg_freq[2] = 0.16644; v.bg_freq[3] = 0.25780;  // :81-97
    v.background_error_prob = background_error_prob;
    v.is_consensus_fasta = is_consensus_fasta ? 1 : 0;
    *out = T;
    return VGH_OK;
}
const vgl_hc_tables* vgh_hc_tables_view(const vgh_hc_tables* t) { return t ? &t->view : nullptr; }
void vgh_hc_tables_destroy(vgh_hc_tables* t) { delete t; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// haplogroup tree
// ------------------------------------------------------------------------------------------------
struct vgh_hc_tree {
    int32_t n_paths = 0, row_words = 0;
    std::vector<std::string> clade_names;
    std::vector<uint32_t> clade_bits;
    std::vector<int32_t> anc_off, anc_clade;
};

namespace {

typedef std::map<std::string, std::vector<std::string>> Relations;

// "name rel1 rel2 ..." per line, tokens containing '[' dropped, blank lines skipped, first entry of a
// repeated name wins (std::map::insert) -- reference src/load.cpp:303-345.
Relations parse_relations(const char* text) {
    Relations rel;
    std::istringstream in(text ? text : "");
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok, name;
        std::vector<std::string> v;
        bool first = true;
        while (ls >> tok) {
            if (first) { name = tok; first = false; }
            else if (tok.find('[') == std::string::npos) v.push_back(tok);
        }
        if (first) continue;
        rel.insert(std::make_pair(name, v));
    }
    return rel;
}

}  // namespace

extern "C" {

static int vgh_hc_tree_build__impl(int32_t n_paths, const char* path_names_txt, const char* parents_txt, const char* children_txt,
                      vgh_hc_tree** out) {
    if (!out || n_paths <= 0 || !path_names_txt) return fail(VGH_ERR_INVALID, "bad tree arguments");
    std::vector<std::string> names;
    {
        std::istringstream in(path_names_txt);
        std::string line;
        while (std::getline(in, line)) names.push_back(line);
    }
    if ((int32_t)names.size() != n_paths) return fail(VGH_ERR_FORMAT, "%zu path names for %d paths", names.size(), n_paths);
    const Relations parents = parse_relations(parents_txt);
    const Relations children = parse_relations(children_txt);
    // A name may label several columns; the reference collects final_vec[idx] for every matching column.
    std::unordered_map<std::string, std::vector<int32_t>> cols_of;
    for (int32_t p = 0; p < n_paths; ++p) cols_of[names[(size_t)p]].push_back(p);

    vgh_hc_tree* T = new vgh_hc_tree();
    T->n_paths = n_paths;
    T->row_words = (n_paths + 31) / 32;
    std::unordered_map<std::string, int32_t> clade_id;
    T->anc_off.assign((size_t)n_paths + 1, 0);
    for (int32_t p = 0; p < n_paths; ++p) {
        auto it = parents.find(names[(size_t)p]);
        if (it != parents.end()) {
            const std::vector<std::string>& pv = it->second;
            for (size_t j = 0; j < pv.size(); ++j) {
                if (j > 0 && pv[j] == pv[j - 1]) continue;  // consecutive duplicates are reported once (get_posterior.cpp:110,117)
                auto cit = clade_id.find(pv[j]);
                int32_t id;
                if (cit == clade_id.end()) {
                    id = (int32_t)T->clade_names.size();
                    clade_id.insert(std::make_pair(pv[j], id));
                    T->clade_names.push_back(pv[j]);
                } else {
                    id = cit->second;
                }
                T->anc_clade.push_back(id);
            }
        }
        T->anc_off[(size_t)p + 1] = (int32_t)T->anc_clade.size();
    }
    // strict descendants of every listed clade, generation by generation (get_posterior_of_clade, :51-76)
    const size_t W = (size_t)T->row_words;
    T->clade_bits.assign(std::max<size_t>(T->clade_names.size(), 1) * W, 0u);
    for (size_t c = 0; c < T->clade_names.size(); ++c) {
        uint32_t* bits = T->clade_bits.data() + c * W;
        std::set<std::string> level{T->clade_names[c]}, seen;
        while (!level.empty()) {
            std::set<std::string> next;
            for (const std::string& name : level) {
                auto it = children.find(name);
                if (it == children.end()) continue;
                for (const std::string& ch : it->second) next.insert(ch);
            }
            std::set<std::string> fresh;
            for (const std::string& ch : next) {
                auto pit = cols_of.find(ch);
                if (pit != cols_of.end())
                    for (int32_t p : pit->second) bits[(size_t)p >> 5] |= 1u << (p & 31);
                if (seen.insert(ch).second) fresh.insert(ch);  // guards against cycles in malformed files
            }
            level.swap(fresh);
        }
    }
    *out = T;
    return VGH_OK;
}
int32_t vgh_hc_tree_n_clades(const vgh_hc_tree* t) { return t ? (int32_t)t->clade_names.size() : 0; }
int32_t vgh_hc_tree_row_words(const vgh_hc_tree* t) { return t ? t->row_words : 0; }
const uint32_t* vgh_hc_tree_clade_bits(const vgh_hc_tree* t) { return t ? t->clade_bits.data() : nullptr; }
const int32_t* vgh_hc_tree_anc_off(const vgh_hc_tree* t) { return t ? t->anc_off.data() : nullptr; }
const int32_t* vgh_hc_tree_anc_clade(const vgh_hc_tree* t) { return t ? t->anc_clade.data() : nullptr; }
const char* vgh_hc_tree_clade_name(const vgh_hc_tree* t, int32_t clade) {
    if (!t || clade < 0 || clade >= (int32_t)t->clade_names.size()) return "";
    return t->clade_names[(size_t)clade].c_str();
}
void vgh_hc_tree_destroy(vgh_hc_tree* t) { delete t; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// GAM framing for the device-side decoder (vgl_hc_accumulate_gam): inflate + framing scan only; the messages themselves are
// parsed on the GPU
// ------------------------------------------------------------------------------------------------
struct vgh_gam_frames {
    std::vector<uint8_t> inflated;      // empty when the input was a plain stream and `bytes` points into the caller's buffer
    const uint8_t* bytes = nullptr;
    size_t n_bytes = 0;
    std::vector<uint64_t> msg_off;      // byte offset of every vg.Alignment message
    std::vector<uint32_t> msg_len;
};
extern "C" {
static int vgh_gam_frame__impl(const void* bytes, size_t n, vgh_gam_frames** out) {
    if (!out || (!bytes && n)) return fail(VGH_ERR_INVALID, "bad arguments");
    vgh_gam_frames* F = new vgh_gam_frames();
    const uint8_t* raw_p = (const uint8_t*)bytes;
    if (n >= 2 && raw_p[0] == 0x1f && raw_p[1] == 0x8b) {
        if (int rc = inflate_all(raw_p, n, F->inflated)) { delete F; return rc; }
        F->bytes = F->inflated.data();
        F->n_bytes = F->inflated.size();
    } else {
        F->bytes = raw_p;   // an uncompressed stream is walked where it lies (no copy): the caller keeps it alive
        F->n_bytes = n;
    }
    std::vector<std::pair<const uint8_t*, const uint8_t*>> msgs;
    if (int rc = scan_gam_framing(F->bytes, F->n_bytes, msgs)) { delete F; return rc; }
    F->msg_off.reserve(msgs.size());
    F->msg_len.reserve(msgs.size());
    for (const auto& m : msgs) {
        if ((size_t)(m.second - m.first) > 0x7fffffffu) { delete F; return fail(VGH_ERR_FORMAT, "message too long"); }
        F->msg_off.push_back((uint64_t)(m.first - F->bytes));
        F->msg_len.push_back((uint32_t)(m.second - m.first));
    }
    *out = F;
    return VGH_OK;
}
int vgh_gam_frame(const void* bytes, size_t n, vgh_gam_frames** out) {
    try {
        return vgh_gam_frame__impl(bytes, n, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_gam_frame: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_gam_frame: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_gam_frame: unknown C++ exception");
    }
}
int64_t vgh_gam_frames_view(const vgh_gam_frames* f, const uint8_t** bytes, uint64_t* n_bytes, const uint64_t** msg_off, const uint32_t** msg_len) {
    if (!f) return -1;
    if (bytes) *bytes = f->bytes;
    if (n_bytes) *n_bytes = f->n_bytes;
    if (msg_off) *msg_off = f->msg_off.data();
    if (msg_len) *msg_len = f->msg_len.data();
    return (int64_t)f->msg_off.size();
}
void vgh_gam_frames_destroy(vgh_gam_frames* f) { delete f; }
}  // extern "C"

// No C++ exception crosses the C ABI (allocation failures on crafted inputs included): every entry point that can
// allocate runs inside a try block and reports an error code instead.
extern "C" {

int vgh_graph_create(int64_t min_id, int64_t n_nodes, const int64_t* seq_off, const char* seq, vgh_graph** out) {
    try {
        return vgh_graph_create__impl(min_id, n_nodes, seq_off, seq, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_graph_create: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_graph_create: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_graph_create: unknown C++ exception");
    }
}

int vgh_gam_parse(const void* bytes, size_t n, int drop_unmapped, vgh_gam** out) {
    try {
        return vgh_gam_parse__impl(bytes, n, drop_unmapped, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_gam_parse: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_gam_parse: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_gam_parse: unknown C++ exception");
    }
}

int vgh_gam_read_file(const char* path, int drop_unmapped, vgh_gam** out) {
    try {
        return vgh_gam_read_file__impl(path, drop_unmapped, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_gam_read_file: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_gam_read_file: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_gam_read_file: unknown C++ exception");
    }
}

int vgh_hc_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned, vgh_hc_packed** out) {
    try {
        return vgh_hc_pack__impl(g, a, r0, r1, n_threads, pinned, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_pack: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_pack: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_pack: unknown C++ exception");
    }
}

int vgh_hc_wire_from_packed(const vgh_hc_packed* p, int n_threads, int pinned, vgh_hc_wire** out) {
    try {
        return vgh_hc_wire_from_packed__impl(p, n_threads, pinned, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_wire_from_packed: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_wire_from_packed: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_wire_from_packed: unknown C++ exception");
    }
}

int vgh_hc_raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out) {
    try {
        return vgh_hc_raw_build__impl(a, r0, r1, pinned, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_raw_build: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_raw_build: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_raw_build: unknown C++ exception");
    }
}

int vgh_euka_raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out) {
    try {
        return vgh_euka_raw_build__impl(a, r0, r1, pinned, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_euka_raw_build: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_euka_raw_build: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_euka_raw_build: unknown C++ exception");
    }
}

int vgh_hc_feeder_create(int n_threads, int64_t chunk_reads, vgh_hc_feeder** out) {
    try {
        return vgh_hc_feeder_create__impl(n_threads, chunk_reads, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_create: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_create: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_create: unknown C++ exception");
    }
}

int vgh_hc_feeder_marshal(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, const vgl_hc_raw** out) {
    try {
        return vgh_hc_feeder_marshal__impl(f, a, r0, r1, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_marshal: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_marshal: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feeder_marshal: unknown C++ exception");
    }
}

int vgh_euka_feed(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t* clade, double* in_lik,
                  double* not_lik, double* like, uint8_t* pass) {
    try {
        return vgh_euka_feed__impl(f, ctx, a, r0, r1, clade, in_lik, not_lik, like, pass);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_euka_feed: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_euka_feed: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_euka_feed: unknown C++ exception");
    }
}

int vgh_soibean_feed(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t n_paths, double* path_ll,
                     uint8_t* used) {
    try {
        return vgh_soibean_feed__impl(f, ctx, a, r0, r1, n_paths, path_ll, used);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_soibean_feed: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_soibean_feed: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_soibean_feed: unknown C++ exception");
    }
}

int vgh_hc_feed(vgh_hc_feeder* f, vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1) {
    try {
        return vgh_hc_feed__impl(f, ctx, a, r0, r1);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feed: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feed: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_feed: unknown C++ exception");
    }
}

int vgh_hc_tables_build(int32_t n_rows, int32_t n_paths, const uint8_t* supports, int64_t n_pmap, const int64_t* pmap_node, const int32_t* pmap_val, int64_t n_mappab, const double* mappab, double background_error_prob, int is_consensus_fasta, vgh_hc_tables** out) {
    try {
        return vgh_hc_tables_build__impl(n_rows, n_paths, supports, n_pmap, pmap_node, pmap_val, n_mappab, mappab, background_error_prob, is_consensus_fasta, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tables_build: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tables_build: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tables_build: unknown C++ exception");
    }
}

int vgh_hc_tree_build(int32_t n_paths, const char* path_names_txt, const char* parents_txt, const char* children_txt, vgh_hc_tree** out) {
    try {
        return vgh_hc_tree_build__impl(n_paths, path_names_txt, parents_txt, children_txt, out);
    } catch (const std::bad_alloc&) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tree_build: out of host memory");
    } catch (const std::exception& e) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tree_build: %s", e.what());
    } catch (...) {
        return fail(VGH_ERR_INVALID, "vgh_hc_tree_build: unknown C++ exception");
    }
}

}  // extern "C"
