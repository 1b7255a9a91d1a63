// GAM wire decode on the device (reference: readGAM, src/readGAM.h:20-68, which hands every
// vg::Alignment of the stream to a lambda that keeps sequence, path, mapping_quality, quality, read_paired and identity).
//
// The host inflates the stream and walks its framing (vgh_gam_frame); the serialized vg.Alignment messages are parsed here,
// one thread per message, straight into the raw-alignment arrays the reconstruction kernels read (ReconRawDev): nothing of a
// message is ever materialised on the host.  Field numbers as on the wire:
//   Alignment 1 sequence | 2 path | 4 quality | 5 mapping_quality | 16 identity (fixed64) | 20 read_paired
//   Path 2 mapping (repeated);  Mapping 1 position | 2 edit (repeated);  Position 1 node_id | 2 offset | 4 is_reverse
//   Edit 1 from_length | 2 to_length | 3 sequence
// Protobuf semantics kept: unknown fields are skipped by wire type, a repeated scalar / bytes field keeps its LAST value,
// repeated messages accumulate.
//
// Two passes over the messages: G1 counts (mappings, edits, quality bytes, edit-sequence bytes, capacity of the aligned
// strings) and decides which reads the hot loop keeps (identity >= 1e-10: readGAM drops identity == 0, src/readGAM.h:47, and the
// loop skips identity < 1e-10, src/HaploCart.cpp:411); a scan turns the counts of the kept reads into offsets; G2 parses again
// and writes.  A message that does not parse sets the error flag (the host parser reports VGH_ERR_FORMAT for it).
#include <cuda_runtime.h>
#include <stdint.h>

#include "vgl_internal.h"

namespace vgl {

namespace {

struct Cur {
    const uint8_t* p;
    const uint8_t* end;
    bool ok;
    __device__ __forceinline__ uint64_t varint() {
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
    __device__ __forceinline__ Cur sub() {
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Cur{p, p, false}; }
        Cur c{p, p + n, true};
        p += n;
        return c;
    }
    __device__ __forceinline__ void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: if (end - p < 8) ok = false; else p += 8; break;
            case 2: sub(); break;
            case 5: if (end - p < 4) ok = false; else p += 4; break;
            default: ok = false;
        }
    }
};

// Walks one message; `V` receives the pieces.  The same code drives both passes.
template <class V>
__device__ __forceinline__ bool walk_alignment(Cur c, V& v) {
    while (c.ok && c.p < c.end) {
        const uint64_t key = c.varint();
        const int fno = (int)(key >> 3), wt = (int)(key & 7);
        if (fno == 1 && wt == 2) { Cur s = c.sub(); v.sequence(s.p, s.end); }
        else if (fno == 4 && wt == 2) { Cur s = c.sub(); v.quality(s.p, s.end); }
        else if (fno == 5 && wt == 0) v.mapq((int64_t)c.varint());
        else if (fno == 16 && wt == 1) {
            if (c.end - c.p < 8) { c.ok = false; break; }
            unsigned long long bits = 0;
            for (int k = 0; k < 8; ++k) bits |= (unsigned long long)c.p[k] << (8 * k);
            v.identity(__longlong_as_double((long long)bits));
            c.p += 8;
        } else if (fno == 2 && wt == 2) {  // Path
            Cur path = c.sub();
            while (path.ok && path.p < path.end) {
                const uint64_t k2 = path.varint();
                if ((k2 >> 3) == 2 && (k2 & 7) == 2) {  // Mapping
                    Cur m = path.sub();
                    int64_t node = 0, offset = 0;
                    bool rev = false;
                    v.mapping_begin();
                    while (m.ok && m.p < m.end) {
                        const uint64_t k3 = m.varint();
                        if ((k3 >> 3) == 1 && (k3 & 7) == 2) {  // Position
                            Cur pos = m.sub();
                            while (pos.ok && pos.p < pos.end) {
                                const uint64_t k4 = pos.varint();
                                if ((k4 >> 3) == 1 && (k4 & 7) == 0) node = (int64_t)pos.varint();
                                else if ((k4 >> 3) == 2 && (k4 & 7) == 0) offset = (int64_t)pos.varint();
                                else if ((k4 >> 3) == 4 && (k4 & 7) == 0) rev = pos.varint() != 0;
                                else pos.skip((int)(k4 & 7));
                            }
                            if (!pos.ok) m.ok = false;
                        } else if ((k3 >> 3) == 2 && (k3 & 7) == 2) {  // Edit
                            Cur e = m.sub();
                            int64_t from = 0, to = 0;
                            v.edit_begin();
                            while (e.ok && e.p < e.end) {
                                const uint64_t k4 = e.varint();
                                if ((k4 >> 3) == 1 && (k4 & 7) == 0) from = (int64_t)e.varint();
                                else if ((k4 >> 3) == 2 && (k4 & 7) == 0) to = (int64_t)e.varint();
                                else if ((k4 >> 3) == 3 && (k4 & 7) == 2) { Cur s = e.sub(); if (e.ok) v.edit_sequence(s.p, s.end); }
                                else e.skip((int)(k4 & 7));
                            }
                            if (!e.ok) m.ok = false;
                            v.edit_end(from, to);
                        } else {
                            m.skip((int)(k3 & 7));
                        }
                    }
                    if (!m.ok) path.ok = false;
                    v.mapping_end(node, offset, rev);
                } else {
                    path.skip((int)(k2 & 7));
                }
            }
            if (!path.ok) c.ok = false;
        } else {
            c.skip(wt);
        }
    }
    return c.ok;
}

struct CountVisitor {
    uint32_t n_maps = 0, n_edits = 0, n_qual = 0, n_eseq = 0;
    uint64_t span = 0;       // sum over the edits of from_length + to_length + |sequence|
    int64_t mq = 0;
    double ident = 0.0;
    bool range_ok = true;    // values fit the raw arrays (node 31 bits, offset 32 bits, lengths 16 bits)
    __device__ void sequence(const uint8_t*, const uint8_t*) {}
    __device__ void quality(const uint8_t* a, const uint8_t* b) { n_qual = (uint32_t)(b - a); }
    __device__ void mapq(int64_t v) { mq = v; }
    __device__ void identity(double v) { ident = v; }
    __device__ void mapping_begin() {}
    __device__ void mapping_end(int64_t node, int64_t offset, bool) {
        ++n_maps;
        range_ok = range_ok && node >= 0 && node <= 0x7fffffffll && offset >= 0 && offset <= 0xffffffffll;
    }
    __device__ void edit_begin() {}
    __device__ void edit_sequence(const uint8_t* a, const uint8_t* b) { n_eseq += (uint32_t)(b - a); span += (uint64_t)(b - a); }
    __device__ void edit_end(int64_t from, int64_t to) {
        ++n_edits;
        range_ok = range_ok && from >= 0 && from <= 65535 && to >= 0 && to <= 65535;
        span += (uint64_t)(from + to);
    }
};

struct FillVisitor {
    GamOutDev o;
    uint32_t r;                       // index among the kept reads
    uint32_t m, e, s;                 // running mapping / edit / edit-sequence positions
    const uint8_t* q_a = nullptr;
    const uint8_t* q_b = nullptr;
    uint32_t seq_len = 0;
    __device__ void sequence(const uint8_t* a, const uint8_t* b) { seq_len = (uint32_t)(b - a); }
    __device__ void quality(const uint8_t* a, const uint8_t* b) { q_a = a; q_b = b; }
    __device__ void mapq(int64_t) {}
    __device__ void identity(double) {}
    __device__ void mapping_begin() { o.edit_off[m] = e; }
    __device__ void mapping_end(int64_t node, int64_t offset, bool rev) {
        o.map_node[m] = (uint32_t)node | (rev ? 0x80000000u : 0u);
        o.map_offset[m] = (uint32_t)offset;
        ++m;
    }
    __device__ void edit_begin() { o.edit_seq_off[e] = s; }
    __device__ void edit_sequence(const uint8_t* a, const uint8_t* b) {
        for (const uint8_t* p = a; p < b; ++p) o.edit_seq[s++] = *p;
    }
    __device__ void edit_end(int64_t from, int64_t to) {
        o.edit_from[e] = (uint16_t)from;
        o.edit_to[e] = (uint16_t)to;
        ++e;
    }
};

constexpr int kErrGamFormat = 0x20000000;   // the context's "malformed alignment" bit (kErrBadRead of recon_kernels.cu)

// G1: one thread per message.  counts[5][n]: mappings, edits, quality bytes, edit-sequence bytes, string capacity (0 for reads that
// are not kept); keep[n]; mapq histogram of the kept reads; largest canonical quality byte.
__global__ void __launch_bounds__(128)
gam_count_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ msg_off, const uint32_t* __restrict__ msg_len, uint32_t n,
                 uint64_t byte_base, uint32_t* __restrict__ counts, uint8_t* __restrict__ keep, uint32_t* __restrict__ mapq_hist,
                 uint32_t* __restrict__ qual_max, int* __restrict__ err_flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = bytes + (msg_off[i] - byte_base);
    CountVisitor v;
    const bool ok = walk_alignment(Cur{p, p + msg_len[i], true}, v);
    bool kept = ok && !(v.ident < 1e-10);
    if (!ok) atomicOr(err_flag, kErrGamFormat);
    if (kept && (!v.range_ok || v.mq < 0 || v.mq > 255 || v.span > 0xfffffff0ull)) {  // does not fit the raw arrays: malformed for the path
        atomicOr(err_flag, kErrGamFormat);
        kept = false;
    }
    keep[i] = kept ? 1 : 0;
    counts[i] = kept ? 1u : 0u;                         // (row 0: the read itself)
    counts[(size_t)n + i] = kept ? v.n_maps : 0u;
    counts[(size_t)2 * n + i] = kept ? v.n_edits : 0u;
    counts[(size_t)3 * n + i] = kept ? v.n_qual : 0u;
    counts[(size_t)4 * n + i] = kept ? v.n_eseq : 0u;
    counts[(size_t)5 * n + i] = kept ? (uint32_t)((v.span + 3) & ~(uint64_t)3) : 0u;
    if (kept) {
        atomicAdd(mapq_hist + (v.mq & 127), 1u);
        // largest canonical quality byte (as the reference reads it: signed char; <= 2 counts as 0)
        const uint8_t* q = nullptr;
        uint32_t nq = 0;
        {   // the quality field's bytes: walk once more, cheaply, only the top level
            Cur c{p, p + msg_len[i], true};
            while (c.ok && c.p < c.end) {
                const uint64_t key = c.varint();
                if ((key >> 3) == 4 && (key & 7) == 2) { Cur s = c.sub(); q = s.p; nq = (uint32_t)(s.end - s.p); }
                else c.skip((int)(key & 7));
            }
        }
        uint32_t mx = 0;
        for (uint32_t k = 0; k < nq; ++k) { const uint32_t b = q[k]; const uint32_t cq = ((b & 0x80u) || b <= 2u) ? 0u : b; mx = cq > mx ? cq : mx; }
        if (mx) atomicMax(qual_max, mx);
    }
}

// exclusive scans of the six count rows, one block; row k's total lands in totals[k]
__global__ void __launch_bounds__(1024) gam_scan_kernel(uint32_t n, uint32_t* __restrict__ counts, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int t = threadIdx.x;
    for (int row = 0; row < 6; ++row) {
        uint32_t* a = counts + (size_t)row * n;
        if (t == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < n; base += 1024) {
            const uint32_t i = base + t;
            const uint32_t v = i < n ? a[i] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, x, k);
                if ((t & 31) >= k) x += o;
            }
            if ((t & 31) == 31) s_warp[t >> 5] = x;
            __syncthreads();
            if (t < 32) {
                uint32_t w = s_warp[t];
#pragma unroll
                for (int k = 1; k < 32; k <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, w, k);
                    if (t >= k) w += o;
                }
                s_warp[t] = w;
            }
            __syncthreads();
            const uint32_t excl = s_carry + x - v + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
            if (i < n) a[i] = excl;
            __syncthreads();
            if (t == 1023) s_carry = excl + v;
            __syncthreads();
        }
        if (t == 0) totals[row] = s_carry;
        __syncthreads();
    }
}

// G2: one thread per kept message: parse again and write at the scanned offsets
__global__ void __launch_bounds__(128)
gam_fill_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ msg_off, const uint32_t* __restrict__ msg_len, uint32_t n,
                uint64_t byte_base, const uint32_t* __restrict__ offs, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ totals,
                GamOutDev o, int* __restrict__ err_flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {   // the closing entries of the offset arrays
        const uint32_t R = totals[0];
        o.map_off[R] = totals[1]; o.qual_off[R] = totals[3]; o.str_off[R] = totals[5];
        o.edit_off[totals[1]] = totals[2];
        o.edit_seq_off[totals[2]] = totals[4];
    }
    if (!keep[i]) return;
    const uint8_t* p = bytes + (msg_off[i] - byte_base);
    FillVisitor v;
    v.o = o;
    v.r = offs[i];
    v.m = offs[(size_t)n + i];
    v.e = offs[(size_t)2 * n + i];
    v.s = offs[(size_t)4 * n + i];
    const uint32_t q0 = offs[(size_t)3 * n + i];
    o.map_off[v.r] = v.m;
    o.qual_off[v.r] = q0;
    o.str_off[v.r] = offs[(size_t)5 * n + i];
    // mapq / identity were validated by G1; read the scalar again from the top level
    int64_t mq = 0;
    {
        Cur c{p, p + msg_len[i], true};
        while (c.ok && c.p < c.end) {
            const uint64_t key = c.varint();
            if ((key >> 3) == 5 && (key & 7) == 0) mq = (int64_t)c.varint();
            else c.skip((int)(key & 7));
        }
    }
    o.mapq[v.r] = (uint8_t)mq;
    if (!walk_alignment(Cur{p, p + msg_len[i], true}, v)) atomicOr(err_flag, kErrGamFormat);
    for (const uint8_t* q = v.q_a; q < v.q_b; ++q) o.qual[q0 + (uint32_t)(q - v.q_a)] = *q;
    if (o.seq_len) o.seq_len[v.r] = (uint16_t)(v.seq_len > 65535u ? 65535u : v.seq_len);
    if (o.read_index) o.read_index[v.r] = i;
}

}  // namespace

cudaError_t gam_launch_count(const uint8_t* bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint32_t n, uint64_t byte_base,
                             uint32_t* counts, uint8_t* keep, uint32_t* mapq_hist, uint32_t* qual_max, uint32_t* totals, int* err_flag,
                             cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gam_count_kernel<<<(n + 127) / 128, 128, 0, s>>>(bytes, msg_off, msg_len, n, byte_base, counts, keep, mapq_hist, qual_max, err_flag);
    gam_scan_kernel<<<1, 1024, 0, s>>>(n, counts, totals);
    return cudaGetLastError();
}

cudaError_t gam_launch_fill(const uint8_t* bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint32_t n, uint64_t byte_base,
                            const uint32_t* offs, const uint8_t* keep, const uint32_t* totals, const GamOutDev& out, int* err_flag,
                            cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gam_fill_kernel<<<(n + 127) / 128, 128, 0, s>>>(bytes, msg_off, msg_len, n, byte_base, offs, keep, totals, out, err_flag);
    return cudaGetLastError();
}

}  // namespace vgl
