// Device-side alignment reconstruction and batch packing.
//
// The host packer (vgh_hc_pack) rebuilds every read's aligned graph / read strings with
// reconstruct_graph_sequence (reference src/vgan_utils.h:6-79) and marshals them the way update_likelihood hands
// them to process_mapping (src/update_likelihood.cpp:33-49); at ~2e6 reads/s on 16 threads it is 60x slower than the
// read kernel.  These kernels do the same work on the device from the raw alignments (vgl_hc_raw: quality bytes,
// mappings, edits) and a copy of the node sequences, and leave a vgl_hc_batch in HBM for hc_read_kernel:
//   R1 hc_recon_kernel   one WARP per read: graph string, read string (with the deletion splices), the size slot of
//                        every mapping (quirk Q1), column / CLOSE / empty-mapping counts, histogram of column counts
//   R2 hc_scan4096       exclusive scan of the histogram                      (counting sort of the reads by columns,
//   R3 hc_order_kernel   scatter of the read ids into their buckets            so that a group's 32 lanes are equally long)
//   R4 hc_groups_kernel  per group of 32 lanes: column blocks, extra blocks;  R5 hc_scan_groups: their prefix sums
//   R6 hc_pack_kernel    one thread per lane: the 32-bit column records of include/vgan_gpu.h, 16 bytes per store
// Every quirk of the host implementation is kept (Q1 sizes per edit, Q2 read string from 0, Q3 'S' at node offset 0,
// stale size for an edit of no recognised kind); tests/test_gpu_recon.py compares the device batch with the host
// packer's record for record.
#include <cuda_runtime.h>
#include <stdint.h>

#include "vgl_internal.h"

namespace vgl {

namespace {

constexpr int kErrBadRead = 0x20000000;   // bits of the context's error word
constexpr int kErrBadNodeRaw = 0x10000000;
constexpr int kErrBatchCapacity = 0x04000000;   // the device-built batch has no room left (vgl_hc_finalize names the way out)
constexpr int kReconHaplocart = 0, kReconSoibean = 2;  // (1 = euka)

enum EditKind { kMatch, kSub, kIns, kDel, kOther };
__device__ __forceinline__ EditKind classify(uint32_t from, uint32_t to, bool has_seq) {  // vg io edit_is_*
    if (from == to && !has_seq) return kMatch;
    if (from == to && has_seq) return kSub;
    if (from == 0 && to > 0 && has_seq) return kIns;
    if (from > 0 && to == 0) return kDel;
    return kOther;
}

__device__ __forceinline__ uint8_t complement(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;
    }
}

// R1 works a WARP per read: the lanes take the read's mappings 32 at a time, so the mapping / edit arrays are read in
// consecutive words (one thread per read touched one 32-byte sector per 4-byte field and was bound by L2 sectors: 5 ms per
// 10^6 reads) and the strings are written in neighbouring bytes.  A CTA of eight warps takes 64 consecutive reads.
constexpr int kReconWarps = 8, kReconReadsPerCta = 64;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, k);
        if (lane >= k) v += o;
    }
    return v;
}

// One mapping (src/vgan_utils.h:21-75): what its edits append to the graph string and to the read string.  WRITE = false
// measures (and stores one size per EDIT, :72); WRITE = true writes the bytes at g_pos / r_pos -- same control flow, so the
// two passes agree by construction.  Returns an error bit (WRITE = false only).
template <bool WRITE>
__device__ __forceinline__ int recon_mapping(const ReconRawDev& raw, const ReconGraphDev& G, const ReconWorkDev& W, int mode, uint32_t m,
                                             uint32_t& g_pos, uint32_t& r_pos, uint32_t& n_del, uint8_t* __restrict__ gs, uint8_t* __restrict__ rs) {
    const uint32_t mn = raw.map_node[m];
    const uint32_t node = mn & 0x7fffffffu;
    const bool rev = (mn >> 31) != 0;
    if (!WRITE && (node < G.min_id || node >= G.min_id + G.n_nodes || (mode == kReconHaplocart && node > VGL_HC_MAX_NODE))) return kErrBadNodeRaw;
    const uint32_t na = G.seq_off[node - G.min_id], nlen = G.seq_off[node - G.min_id + 1] - na;
    uint32_t offset = raw.map_offset[m];
    uint32_t aligned = 0;  // per mapping, NOT per edit: an edit of no recognised kind repeats the previous size
    const uint32_t e1 = raw.edit_off[m + 1];
    for (uint32_t e = raw.edit_off[m]; e < e1; ++e) {
        const uint32_t from = raw.edit_from[e], to = raw.edit_to[e];
        const uint32_t q0 = raw.edit_seq_off[e], q1 = raw.edit_seq_off[e + 1];
        const EditKind kind = classify(from, to, q1 > q0);
        if (kind == kMatch || kind == kSub || kind == kDel) {
            if (!WRITE && offset > nlen) return kErrBadRead;
            const uint32_t take = min(from, nlen - offset);  // std::string::substr clips
            if (WRITE) {
                for (uint32_t k = 0; k < take; ++k) {
                    // base at oriented offset of the node (what get_sequence(get_handle(id, rev))[k] returns)
                    const uint8_t c = rev ? complement(G.seq[na + nlen - 1 - (offset + k)]) : G.seq[na + offset + k];
                    gs[g_pos + k] = c;
                    if (kind == kMatch) rs[r_pos + k] = c;
                }
                if (kind == kSub)
                    for (uint32_t k = q0; k < q1; ++k) rs[r_pos + (k - q0)] = raw.edit_seq[k];
            }
            g_pos += take;
            r_pos += kind == kMatch ? take : kind == kSub ? q1 - q0 : 0u;
            aligned = take;
            n_del += kind == kDel;
        } else if (kind == kIns) {
            if (WRITE) {
                const uint8_t fill = offset == 0 ? 'S' : '-';
                for (uint32_t k = 0; k < to; ++k) gs[g_pos + k] = fill;
                for (uint32_t k = q0; k < q1; ++k) rs[r_pos + (k - q0)] = raw.edit_seq[k];
            }
            g_pos += to;
            r_pos += q1 - q0;
            aligned = to;
        }
        offset += from;
        if (!WRITE) W.sizes[e] = (uint16_t)min(aligned, 0xffffu);  // one size per EDIT (src/vgan_utils.h:72)
    }
    return 0;
}

__global__ void __launch_bounds__(kReconWarps * 32)
hc_recon_kernel(ReconRawDev raw, ReconGraphDev G, ReconWorkDev W, int mode, int* __restrict__ err_flag) {
    __shared__ uint32_t s_key[kReconReadsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < kReconReadsPerCta / kReconWarps; ++it) {
        const uint32_t slot = (uint32_t)(it * kReconWarps + warp);
        const uint32_t r = blockIdx.x * kReconReadsPerCta + slot;
        if (r >= raw.n_reads) {
            if (lane == 0) s_key[slot] = 0xffffffffu;
            continue;
        }
        const uint32_t m0 = raw.map_off[r], m1 = raw.map_off[r + 1], M = m1 - m0;
        const uint32_t s0 = raw.str_off[r], cap = raw.str_off[r + 1] - s0;
        uint8_t* gs = W.gstr + s0;
        uint8_t* rs = W.rstr + s0;
        uint32_t glen = 0, rlen = 0, n_del = 0;
        int err = 0;
        for (uint32_t mb = m0; mb < m1 && !err; mb += 32) {
            const uint32_t m = mb + (uint32_t)lane;
            uint32_t gl = 0, rl = 0, nd = 0;
            int e_mine = 0;
            if (m < m1) e_mine = recon_mapping<false>(raw, G, W, mode, m, gl, rl, nd, nullptr, nullptr);
            const uint32_t bad = __ballot_sync(0xffffffffu, e_mine != 0);
            if (bad) {   // the error the serial loop of the reference meets first
                err = __shfl_sync(0xffffffffu, e_mine, __ffs(bad) - 1);
                break;
            }
            const uint32_t gi = warp_incl_scan(gl, lane), ri = warp_incl_scan(rl, lane);
            const uint32_t g_chunk = __shfl_sync(0xffffffffu, gi, 31), r_chunk = __shfl_sync(0xffffffffu, ri, 31);
            if (glen + g_chunk > cap || rlen + r_chunk > cap) { err = kErrBadRead; break; }  // never write past the read's strings
            if (m < m1) {
                uint32_t gp = glen + gi - gl, rp = rlen + ri - rl, unused = 0;
                recon_mapping<true>(raw, G, W, mode, m, gp, rp, unused, gs, rs);
            }
            glen += g_chunk;
            rlen += r_chunk;
            n_del += __reduce_add_sync(0xffffffffu, nd);
        }
        __syncwarp();
        // deletions: a '-' run spliced into the read string at the running from_length coordinate, in order, each against the
        // string as it stands (src/vgan_utils.h:67) -- rare, one lane
        if (!err && n_del) {
            if (lane == 0) {
                uint32_t from_total = 0;
                for (uint32_t m = m0; m < m1 && !err; ++m) {
                    for (uint32_t e = raw.edit_off[m]; e < raw.edit_off[m + 1]; ++e) {
                        const uint32_t from = raw.edit_from[e], to = raw.edit_to[e];
                        const bool has_seq = raw.edit_seq_off[e + 1] > raw.edit_seq_off[e];
                        if (classify(from, to, has_seq) == kDel) {
                            if (from_total > rlen || rlen + from > cap) { err = kErrBadRead; break; }
                            for (uint32_t k = rlen; k > from_total; --k) rs[k - 1 + from] = rs[k - 1];
                            for (uint32_t k = 0; k < from; ++k) rs[from_total + k] = '-';
                            rlen += from;
                        }
                        from_total += from;
                    }
                }
            }
            rlen = __shfl_sync(0xffffffffu, rlen, 0);
            err = __shfl_sync(0xffffffffu, err, 0);
            __syncwarp();
        }
        const uint32_t Lq = raw.qual_off[r + 1] - raw.qual_off[r];
        const uint32_t e_first = M ? raw.edit_off[m0] : 0u;  // the read's edits are e_first .. e_first + E - 1
        const uint32_t E = M ? raw.edit_off[m1] - e_first : 0u;
        uint32_t n_close = 0, n_extra = 0;
        if (mode == kReconHaplocart) {
            if (!err && (glen != rlen || E < M || Lq > rlen || rlen > VGL_HC_MAP_MAX_POS || M > 65535u || raw.mapq[r] >= 100)) err = kErrBadRead;
            if (!err) {
                // mapping i starts at pos_i = min(sum of the earlier size slots, n) -- the size slot of mapping i is the size of the
                // read's i-th EDIT (quirk Q1); CLOSE bits sit at the distinct pos_i < n; a mapping with pos_{i+1} = pos_i owns no column
                uint32_t carry = 0, carry_lg = 1;
                for (uint32_t ib = 0; ib < M; ib += 32) {
                    const uint32_t i = ib + (uint32_t)lane;
                    const uint32_t sz = i < M ? (uint32_t)W.sizes[e_first + i] : 0u;
                    const uint32_t incl = warp_incl_scan(sz, lane);
                    const uint32_t pos = min(carry + incl - sz, rlen), lg = min(carry + incl, rlen) - pos;
                    uint32_t prev_lg = __shfl_up_sync(0xffffffffu, lg, 1);
                    if (lane == 0) prev_lg = carry_lg;
                    n_extra += (uint32_t)__popc(__ballot_sync(0xffffffffu, i < M && lg == 0));
                    n_close += (uint32_t)__popc(__ballot_sync(0xffffffffu, i < M && pos < rlen && prev_lg > 0));
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                    carry_lg = __shfl_sync(0xffffffffu, lg, 31);
                }
            }
        } else {
            // euka / soibean (vgh_euka_pack / vgh_soibean_pack): fragment length inside the damage table, the damage position
            // stays inside the read; the "extra" count is the number of mapping (euka) or edit (soibean) records
            const uint32_t Lseq = mode == kReconSoibean ? glen : (uint32_t)raw.seq_len[r];
            uint32_t nongap = 0;
            if (!err)
                for (uint32_t k = (uint32_t)lane; k < rlen; k += 32) nongap += rs[k] != '-';
            nongap = __reduce_add_sync(0xffffffffu, nongap);
            const uint32_t n_rec = mode == kReconSoibean ? E : M;
            if (!err && (M < 1 || glen != rlen || Lseq < 15u || Lseq > 1000u || glen > 4095u || M > 65535u || n_rec > 65535u || nongap > Lseq)) err = kErrBadRead;
            if (!err && mode == kReconSoibean) {
                if (raw.seq_len[r] < 1) err = kErrBadRead;
                bool far = false;
                for (uint32_t i = (uint32_t)lane; i < min(M, E); i += 32) far = far || (raw.map_node[m0 + i] & 0x7fffffffu) >= VGL_SOIBEAN_NO_NODE;
                if (__any_sync(0xffffffffu, far)) err = kErrBadNodeRaw;
            }
            n_extra = n_rec;
        }
        if (err) {
            rlen = 0;
            n_close = 0xffffu;  // marks the read as dead for the packer
            n_extra = 0;
        }
        if (lane == 0) {
            if (err) atomicOr(err_flag, err);
            W.n_cols[r] = (uint16_t)rlen;
            W.n_close[r] = (uint16_t)n_close;
            W.n_extra[r] = (uint16_t)n_extra;
            s_key[slot] = (uint32_t)W.slot_of_mapq[raw.mapq[r] & 127u] * 4096u + rlen;
        }
    }
    __syncthreads();
    // histogram of (mapping-quality slot, column count): one atomic per distinct value among 32 reads
    if (warp < kReconReadsPerCta / 32) {
        const uint32_t key = s_key[warp * 32 + lane];
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (key != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(W.hist + key, (uint32_t)__popc(peers));
    }
}

// exclusive scan of a slot's 4096 counters, one block of 1024 threads (4 per thread) per slot; slot s starts at lane lane_base
__global__ void __launch_bounds__(1024) hc_scan4096(const uint32_t* __restrict__ hist_all, uint32_t* __restrict__ start_all, ReconWorkDev W) {
    __shared__ uint32_t s_warp[32];
    const int t = threadIdx.x;
    const uint32_t* hist = hist_all + (size_t)blockIdx.x * 4096;
    uint32_t* start = start_all + (size_t)blockIdx.x * 4096;
    const uint32_t lane0 = W.lane_base[blockIdx.x];
    const uint32_t a = hist[4 * t], b = hist[4 * t + 1], c = hist[4 * t + 2], d = hist[4 * t + 3];
    const uint32_t mine = a + b + c + d;
    uint32_t v = mine;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, k);
        if ((t & 31) >= k) v += o;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = v;
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
    const uint32_t base = lane0 + v - mine + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
    start[4 * t] = base;
    start[4 * t + 1] = base + a;
    start[4 * t + 2] = base + a + b;
    start[4 * t + 3] = base + a + b + c;
}

__global__ void __launch_bounds__(256)
hc_order_kernel(uint32_t n_reads, const uint16_t* __restrict__ n_cols, const uint8_t* __restrict__ mapq, const uint32_t* __restrict__ start,
                uint32_t* __restrict__ cursor, uint32_t* __restrict__ order, ReconWorkDev W) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t key = (uint32_t)W.slot_of_mapq[mapq[r] & 127u] * 4096u + n_cols[r];
    // one atomic per distinct key in the warp (reads of one batch are mostly equally long: a single hot counter otherwise)
    const uint32_t peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[start[key] + base + __popc(peers & ((1u << lane) - 1u))] = r;
}

__global__ void __launch_bounds__(256)
hc_groups_kernel(uint32_t n_reads, uint32_t n_groups, const uint32_t* __restrict__ order, const uint16_t* __restrict__ n_cols,
                 const uint16_t* __restrict__ n_extra, uint32_t col_blk, uint32_t* __restrict__ grp_cb, uint32_t* __restrict__ grp_mb) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    uint32_t mc = 0, mx = 0;
    for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t r = order[g * 32 + l];
        if (r == kReconNoRead) continue;
        mc = max(mc, (uint32_t)n_cols[r]);
        mx = max(mx, (uint32_t)n_extra[r]);
    }
    grp_cb[g] = (mc + col_blk - 1) / col_blk;
    grp_mb[g] = (mx + VGL_HC_MAP_BLK - 1) / VGL_HC_MAP_BLK;
}

// in-place exclusive scan of two arrays of n + 1 entries (entry n receives the total), one block
__global__ void __launch_bounds__(1024) hc_scan_groups(uint32_t n, uint32_t* __restrict__ a, uint32_t* __restrict__ b) {
    __shared__ uint32_t s_wa[32], s_wb[32];
    __shared__ uint32_t s_ca, s_cb;
    const int t = threadIdx.x;
    if (t == 0) { s_ca = 0; s_cb = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + t;
        const uint32_t va = i < n ? a[i] : 0u, vb = i < n ? b[i] : 0u;
        uint32_t xa = va, xb = vb;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t oa = __shfl_up_sync(0xffffffffu, xa, k), ob = __shfl_up_sync(0xffffffffu, xb, k);
            if ((t & 31) >= k) { xa += oa; xb += ob; }
        }
        if ((t & 31) == 31) { s_wa[t >> 5] = xa; s_wb[t >> 5] = xb; }
        __syncthreads();
        if (t < 32) {
            uint32_t wa = s_wa[t], wb = s_wb[t];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t oa = __shfl_up_sync(0xffffffffu, wa, k), ob = __shfl_up_sync(0xffffffffu, wb, k);
                if (t >= k) { wa += oa; wb += ob; }
            }
            s_wa[t] = wa;
            s_wb[t] = wb;
        }
        __syncthreads();
        const uint32_t pa = s_ca + ((t >> 5) ? s_wa[(t >> 5) - 1] : 0u) + xa - va;
        const uint32_t pb = s_cb + ((t >> 5) ? s_wb[(t >> 5) - 1] : 0u) + xb - vb;
        if (i < n) { a[i] = pa; b[i] = pb; }
        __syncthreads();
        if (t == 0) { s_ca += s_wa[31]; s_cb += s_wb[31]; }
        __syncthreads();
    }
    if (t == 0) { a[n] = s_ca; b[n] = s_cb; }
}

__device__ __forceinline__ int base_code(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }

// One thread per lane of the batch: column records pair G[c] with the read base process_mapping compares it with
// (src/update_likelihood.cpp:33-49, src/process_mapping.cpp:47,59-66): for mapping i the graph slice
// G[pos_i, pos_i + size_i) faces algnseq[0, size_i) -- the whole aligned read string from its start.
__global__ void __launch_bounds__(128)
hc_pack_kernel(ReconRawDev raw, ReconWorkDev W, uint32_t n_groups, uint32_t col_cap_blks, uint32_t map_cap_blks,
               uint32_t* __restrict__ col, uint32_t* __restrict__ xmap, uint16_t* __restrict__ o_n_cols,
               uint16_t* __restrict__ o_n_close, uint16_t* __restrict__ o_n_extra, uint8_t* __restrict__ o_mapq,
               int* __restrict__ err_flag) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = slot >> 5, lane = slot & 31;
    if (g >= n_groups) return;
    if (W.grp_cb[n_groups] > col_cap_blks || W.grp_mb[n_groups] > map_cap_blks) {  // columns: cannot happen for sorted groups; extra records: a chunk with very many empty mappings.  Never write past the buffers
        if (slot == 0) atomicOr(err_flag, kErrBatchCapacity);
        // the read kernel still runs over the chunk: every lane empty (no extra record is ever fetched), its column blocks zero
        o_n_cols[slot] = o_n_close[slot] = o_n_extra[slot] = 0;
        o_mapq[slot] = 0;
        if (W.grp_cb[n_groups] <= col_cap_blks) {
            uint4* z = reinterpret_cast<uint4*>(col) + (size_t)W.grp_cb[g] * 32 + lane;
            for (uint32_t cb = 0; cb < W.grp_cb[g + 1] - W.grp_cb[g]; ++cb) z[(size_t)cb * 32] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const uint32_t cb0 = W.grp_cb[g], ncb = W.grp_cb[g + 1] - cb0;
    uint4* gc = reinterpret_cast<uint4*>(col) + (size_t)cb0 * 32 + lane;  // block k of this lane: gc[k * 32]
    if (W.order[slot] == kReconNoRead) {  // unused lane (the last group of a mapping-quality slot is filled up with them)
        for (uint32_t cb = 0; cb < ncb; ++cb) gc[(size_t)cb * 32] = make_uint4(0, 0, 0, 0);
        o_n_cols[slot] = o_n_close[slot] = o_n_extra[slot] = 0;
        o_mapq[slot] = 0;
        return;
    }
    const uint32_t r = W.order[slot];
    const uint32_t L = W.n_cols[r];
    const bool dead = W.n_close[r] == 0xffffu;  // rejected by the reconstruction: an empty lane
    const uint32_t m0 = raw.map_off[r], M = dead ? 0u : raw.map_off[r + 1] - m0;
    const uint32_t e_first = M ? raw.edit_off[m0] : 0u;
    const uint32_t q0 = raw.qual_off[r], Lq = raw.qual_off[r + 1] - q0;
    const uint8_t* gs = W.gstr + raw.str_off[r];
    const uint8_t* rs = W.rstr + raw.str_off[r];
    uint32_t* gx = xmap + ((size_t)W.grp_mb[g] * 32 + lane) * VGL_HC_MAP_BLK;  // extra record i: gx[(i >> 2) * 128 + (i & 3)]
    // the first REAL quality byte >= 90 latches the background-error mode (src/update_likelihood.cpp:42)
    uint32_t f90 = 0xffffffffu;
    for (uint32_t c = 0; c < Lq; ++c)
        if ((int)(int8_t)raw.qual[q0 + c] >= 90) { f90 = c; break; }
    uint32_t nxt = 0, npos = 0;  // next mapping to start and its first column
    uint32_t own_node = VGL_HC_NO_NODE, own_pos = 0, own_end = 0, latched = 0, n_x = 0;
    for (uint32_t cb = 0; cb < ncb; ++cb) {
        uint32_t rec[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = cb * 4 + j;
            if (c >= L) { rec[j] = 0; continue; }
            uint32_t v = c < Lq ? (VGL_HC_CANON_QUAL((uint32_t)raw.qual[q0 + c]) << VGL_HC_COL_QUAL_SHIFT) : 0u;
            while (nxt < M && npos == c) {  // mappings that start at this column (quirk Q1: size slot = the nxt-th edit's size)
                const uint32_t lg = min((uint32_t)W.sizes[e_first + nxt], L - npos);
                const uint32_t node = raw.map_node[m0 + nxt] & 0x7fffffffu;
                v |= VGL_HC_COL_CLOSE;
                if (lg) {
                    own_node = node;
                    own_pos = c;
                    own_end = c + lg;
                    latched = (f90 != 0xffffffffu && f90 < c + L) ? VGL_HC_COL_LATCHED : 0u;  // its quality window is [pos, pos + L)
                    npos = own_end;
                } else {
                    gx[(size_t)(n_x >> 2) * 128 + (n_x & 3)] = VGL_HC_MAP_REC(node, c);
                    ++n_x;
                }
                ++nxt;
            }
            if (c < own_end) {
                const uint8_t gb = gs[c], rb = rs[c - own_pos];
                const int cg = base_code(gb), cr = base_code(rb);
                v |= (own_node << VGL_HC_COL_NODE_SHIFT) | latched;
                if (cg < 4 && cr < 4) v |= VGL_HC_COL_VALID | ((uint32_t)cr << VGL_HC_COL_BASE_SHIFT) | (gb == rb ? 0u : VGL_HC_COL_MISMATCH);
                if (c + 1 == own_end) v |= VGL_HC_COL_OPEN;
            } else {
                v |= VGL_HC_NO_NODE << VGL_HC_COL_NODE_SHIFT;
            }
            rec[j] = v;
        }
        gc[(size_t)cb * 32] = make_uint4(rec[0], rec[1], rec[2], rec[3]);
    }
    for (; nxt < M; ++nxt) {  // mappings that start past the last column
        gx[(size_t)(n_x >> 2) * 128 + (n_x & 3)] = VGL_HC_MAP_REC(raw.map_node[m0 + nxt] & 0x7fffffffu, L);
        ++n_x;
    }
    o_n_cols[slot] = (uint16_t)L;
    o_n_close[slot] = dead ? (uint16_t)0 : W.n_close[r];
    o_n_extra[slot] = (uint16_t)n_x;
    o_mapq[slot] = raw.mapq[r];
}

// Counts (one byte each) -> offsets: out[i] = base + in[0] + ... + in[i-1] for i in [0, n], three small kernels.
// (The raw alignments ship edit counts per mapping and sequence lengths per edit instead of 32-bit offsets.)
constexpr int kScanTile = 4096;  // 256 threads x 16 elements
__global__ void __launch_bounds__(256) scan_u8_tile_sums(const uint8_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_w[8];
    const uint32_t t0 = blockIdx.x * kScanTile + threadIdx.x * 16;
    uint32_t v = 0;
    for (uint32_t k = 0; k < 16; ++k) v += t0 + k < n ? in[t0 + k] : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_w[w];
        tile_sum[blockIdx.x] = tot;
    }
}
__global__ void __launch_bounds__(1024) scan_tiles(uint32_t n_tiles, uint32_t base, uint32_t* __restrict__ tile_sum) {  // in place, exclusive, + base
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const int t = threadIdx.x;
    if (t == 0) s_carry = base;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_tiles; b0 += 1024) {
        const uint32_t i = b0 + t;
        const uint32_t v = i < n_tiles ? tile_sum[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, x, k);
            if ((t & 31) >= k) x += o;
        }
        if ((t & 31) == 31) s_w[t >> 5] = x;
        __syncthreads();
        if (t < 32) {
            uint32_t w = s_w[t];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, w, k);
                if (t >= k) w += o;
            }
            s_w[t] = w;
        }
        __syncthreads();
        if (i < n_tiles) tile_sum[i] = s_carry + ((t >> 5) ? s_w[(t >> 5) - 1] : 0u) + x - v;
        __syncthreads();
        if (t == 0) s_carry += s_w[31];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) scan_u8_final(const uint8_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ tile_off,
                                                     uint32_t* __restrict__ out) {
    __shared__ uint32_t s_w[8];
    const uint32_t t0 = blockIdx.x * kScanTile + threadIdx.x * 16;
    uint32_t e[16], mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) { e[k] = t0 + k < n ? in[t0 + k] : 0u; mine += e[k]; }
    uint32_t x = mine;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, x, k);
        if ((threadIdx.x & 31) >= (unsigned)k) x += o;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t run = tile_off[blockIdx.x] + x - mine;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) run += s_w[w];
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) {
        if (t0 + k <= n) out[t0 + k] = run;  // entry n receives the total
        run += e[k];
    }
}
__global__ void __launch_bounds__(256) widen_node16(const uint16_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(in[i] & 0x7fffu) | ((uint32_t)(in[i] >> 15) << 31);  // node | is_reverse << 31
}
// map_offset[first mapping of read r] = first16[r]; every other entry was zeroed before (vgl_hc_raw::first_offset16)
__global__ void __launch_bounds__(256) set_first_offsets(const uint32_t* __restrict__ map_off, const uint16_t* __restrict__ first16, uint32_t n,
                                                         uint32_t* __restrict__ map_offset_abs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t m0 = map_off[r];
    if (map_off[r + 1] > m0) map_offset_abs[m0] = first16[r];
}
__global__ void __launch_bounds__(256) widen_u8(const uint8_t* __restrict__ in, uint32_t n, uint16_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void __launch_bounds__(256) widen_u16(const uint16_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__device__ __forceinline__ uint32_t symbol_code(uint8_t c) {
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        case 'N': return VGL_EUKA_SYM_N;
        case '-': return VGL_EUKA_SYM_GAP;
        case 'W': case 'M': case 'K': case 'R': case 'Y': case 'B': case 'D': case 'H': case 'V': return VGL_EUKA_SYM_RARE;
        case 'S': return VGL_EUKA_SYM_SOFTCLIP;
        default: return VGL_EUKA_SYM_OTHER;
    }
}

// euka / soibean batches (vgl_euka_batch) from the reconstructed strings: one thread per lane.  Column record m =
// quality byte at ALIGNED COLUMN m | graph symbol << 8 | read symbol << 12; mapping records = node ids (euka) or, for
// soibean, one record per EDIT: size << 20 | node of the mapping with the same index (NO_NODE for the surplus entries
// when the two counts differ, src/getLCAfromGAM.h:156-160).
__global__ void __launch_bounds__(128)
sym_pack_kernel(ReconRawDev raw, ReconWorkDev W, uint32_t n_groups, int soibean, uint16_t* __restrict__ col, uint32_t* __restrict__ map,
                uint16_t* __restrict__ o_n_cols, uint16_t* __restrict__ o_n_maps, uint16_t* __restrict__ o_seq_len,
                uint8_t* __restrict__ o_mapq, uint8_t* __restrict__ o_flags, uint32_t* __restrict__ o_read_index) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = slot >> 5, lane = slot & 31;
    if (g >= n_groups) return;
    const uint32_t cb0 = W.grp_cb[g], ncb = W.grp_cb[g + 1] - cb0, mb0 = W.grp_mb[g], nmb = W.grp_mb[g + 1] - mb0;
    uint4* gc = reinterpret_cast<uint4*>(col) + (size_t)cb0 * 32 + lane;  // column block k (8 records): gc[k * 32]
    uint4* gm = reinterpret_cast<uint4*>(map) + (size_t)mb0 * 32 + lane;  // mapping block k (4 records): gm[k * 32]
    const bool live = slot < raw.n_reads && W.order[slot] != kReconNoRead && W.n_close[W.order[slot]] != 0xffffu;
    const uint32_t r = live ? W.order[slot] : 0u;
    const uint32_t L = live ? W.n_cols[r] : 0u, n_rec = live ? W.n_extra[r] : 0u;
    const uint32_t m0 = raw.map_off[r], M = raw.map_off[r + 1] - m0;
    const uint32_t q0 = raw.qual_off[r], Lq = raw.qual_off[r + 1] - q0;
    const uint8_t* gs = W.gstr + raw.str_off[r];
    const uint8_t* rs = W.rstr + raw.str_off[r];
    for (uint32_t cb = 0; cb < ncb; ++cb) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t m = cb * 8 + j;
            if (m >= L) continue;
            const uint32_t qb = m < Lq ? raw.qual[q0 + m] : 0u;  // quality by ALIGNED COLUMN (src/readGAM_Euka.h:264,286)
            const uint32_t rec = qb | (symbol_code(gs[m]) << 8) | (symbol_code(rs[m]) << 12);
            w[j >> 1] |= rec << (16 * (j & 1));
        }
        gc[(size_t)cb * 32] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const uint32_t e_first = (live && M) ? raw.edit_off[m0] : 0u;
    for (uint32_t mb = 0; mb < nmb; ++mb) {
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t i = mb * 4 + j;
            if (i >= n_rec) continue;
            if (soibean) {
                const bool none = n_rec != M && i >= M;
                v[j] = ((uint32_t)W.sizes[e_first + i] << 20) | (none ? VGL_SOIBEAN_NO_NODE : (raw.map_node[m0 + i] & 0x7fffffffu));
            } else {
                v[j] = raw.map_node[m0 + i] & 0x7fffffffu;
            }
        }
        gm[(size_t)mb * 32] = make_uint4(v[0], v[1], v[2], v[3]);
    }
    o_n_cols[slot] = (uint16_t)L;
    o_n_maps[slot] = (uint16_t)n_rec;
    o_seq_len[slot] = live ? raw.seq_len[r] : (uint16_t)0;
    o_mapq[slot] = live ? raw.mapq[r] : (uint8_t)0;
    o_flags[slot] = (live && M) ? (uint8_t)(raw.map_node[m0] >> 31) : (uint8_t)0;
    o_read_index[slot] = live ? raw.read_index[r] : 0u;
}

// Wire form -> 32-bit column records (include/vgan_gpu.h, vgl_hc_wire).  One thread per lane, walking from the read's last
// column: an OPEN mark takes the next node id of the lane's node stream, every owned column down to the mapping's first
// gets it.  Group g of the chunk owns 32-bit blocks [2 (c16[g] - c16[0]), 2 (c16[g+1] - c16[0])): twice its 16-bit blocks.
// kNarrow: the columns arrive as one byte (quality | base << 6) plus four flag bits each (col8 / flag4 / n_owned of vgl_hc_wire)
// instead of 16-bit records; the 16-bit record is put together here and the rest of the walk is the same.
template <bool kNarrow>
__global__ void __launch_bounds__(128)
hc_expand_kernel(uint32_t n_groups, const uint32_t* __restrict__ grp_c16, const uint32_t* __restrict__ grp_n16,
                 const uint16_t* __restrict__ col16, const uint8_t* __restrict__ col8, const uint32_t* __restrict__ flag4,
                 const uint16_t* __restrict__ n_owned, const uint16_t* __restrict__ node16, const uint16_t* __restrict__ n_cols,
                 uint32_t* __restrict__ col32, uint32_t* __restrict__ grp_c32) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = slot >> 5, lane = slot & 31;
    if (g >= n_groups) return;
    const uint32_t c0 = grp_c16[0], cb0 = grp_c16[g], ncb = grp_c16[g + 1] - cb0;
    if (lane == 0) {
        grp_c32[g] = 2u * (cb0 - c0);
        if (g + 1 == n_groups) grp_c32[n_groups] = 2u * (grp_c16[n_groups] - c0);
    }
    const uint4* src = reinterpret_cast<const uint4*>(col16) + (size_t)cb0 * 32 + lane;          // block k: src[k * 32]
    const uint4* nsrc = reinterpret_cast<const uint4*>(node16) + (size_t)grp_n16[g] * 32 + lane;  // node block k: nsrc[k * 32]
    uint4* dst = reinterpret_cast<uint4*>(col32) + (size_t)2 * (cb0 - c0) * 32 + lane;            // 32-bit block k: dst[k * 32]
    const uint32_t L = n_cols[slot];
    uint32_t n_own = 0;
    if constexpr (kNarrow) n_own = n_owned[slot];
    const uint2* src8 = reinterpret_cast<const uint2*>(col8) + (size_t)cb0 * 32 + lane;  // narrow block k: src8[k * 32], flag4 likewise
    const uint32_t* srcf = flag4 + (size_t)cb0 * 32 + lane;
    uint32_t n_used = 0, cur = VGL_HC_NO_NODE;
    uint4 nodes = make_uint4(0, 0, 0, 0);
    for (int cb = (int)ncb - 1; cb >= 0; --cb) {
        uint32_t w[4] = {0, 0, 0, 0};
        uint2 a = make_uint2(0, 0);
        uint32_t fl = 0;
        if constexpr (kNarrow) {
            a = src8[(size_t)cb * 32];
            fl = srcf[(size_t)cb * 32];
        } else {
            const uint4 v = src[(size_t)cb * 32];
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
        uint32_t out[8];
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            const uint32_t c = (uint32_t)cb * 8 + j;
            uint32_t rec;
            if constexpr (kNarrow) {
                const uint32_t byte = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xffu;
                const uint32_t nib = (fl >> (4 * j)) & 0xfu;
                static_assert(VGL_HC_COL_MISMATCH == 1u && VGL_HC_COL_VALID == 8u << 9 && VGL_HC_COL_OPEN == 2u << 12 && VGL_HC_COL_CLOSE == 4u << 12,
                              "flag nibble: bit 0 MISMATCH, 1 OPEN, 2 CLOSE, 3 VALID");
                rec = ((byte & 0x3fu) << VGL_HC_COL_QUAL_SHIFT) | (nib & 1u) | ((byte >> 6) << VGL_HC_COL_BASE_SHIFT) | ((nib & 8u) << 9) |
                      ((nib & 2u) << 12) | ((nib & 4u) << 12) | (c < n_own ? VGL_HC_WIRE_OWNED : 0u);
            } else {
                rec = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
            }
            if (c >= L) { out[j] = 0; continue; }
            if (rec & VGL_HC_COL_OPEN) {
                if ((n_used & 7u) == 0) nodes = nsrc[(size_t)(n_used >> 3) * 32];
                const uint32_t k = n_used & 7u;
                const uint32_t word = k < 2 ? nodes.x : k < 4 ? nodes.y : k < 6 ? nodes.z : nodes.w;
                cur = (word >> (16 * (k & 1))) & 0xffffu;
                ++n_used;
            }
            out[j] = (rec & 0x7fffu) | (((rec & VGL_HC_WIRE_OWNED) ? cur : VGL_HC_NO_NODE) << VGL_HC_COL_NODE_SHIFT);
        }
        dst[(size_t)(2 * cb + 1) * 32] = make_uint4(out[4], out[5], out[6], out[7]);
        dst[(size_t)(2 * cb) * 32] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// duplicate removal (reference: Dup_Remover::remove_duplicates_internal, src/rmdup.cpp:20-39,68-110)
// The reference marks read j when some EARLIER read i that is not paired has the same start coordinates (node id and offset of
// the first mapping, strand ignored) -- an O(n^2) scan.  Here: an open-addressing table keyed by the coordinates keeps, per
// key, the smallest index of an unpaired read (atomicMin); read j is a duplicate when that index is below j.
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr unsigned long long kDedupEmpty = 0xffffffffffffffffull;
__device__ __forceinline__ uint32_t dedup_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}
__global__ void __launch_bounds__(256)
dedup_insert_kernel(uint32_t n, const unsigned long long* __restrict__ key, const uint8_t* __restrict__ paired, uint32_t mask,
                    unsigned long long* __restrict__ tab_key, uint32_t* __restrict__ tab_first) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const unsigned long long k = key[r];
    uint32_t slot = dedup_hash(k) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(tab_key + slot, kDedupEmpty, k);
        if (old == kDedupEmpty || old == k) {
            if (!(paired && paired[r])) atomicMin(tab_first + slot, r);   // a paired read marks nothing (include/vgan_host.h: vgh_dedup)
            return;
        }
        slot = (slot + 1) & mask;
    }
}
__global__ void __launch_bounds__(256)
dedup_mark_kernel(uint32_t n, const unsigned long long* __restrict__ key, uint32_t mask, const unsigned long long* __restrict__ tab_key,
                  const uint32_t* __restrict__ tab_first, uint8_t* __restrict__ is_dup, unsigned int* __restrict__ n_dup) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t dup = 0;
    if (r < n) {
        const unsigned long long k = key[r];
        uint32_t slot = dedup_hash(k) & mask;
        while (tab_key[slot] != k) slot = (slot + 1) & mask;   // inserted by the first kernel
        dup = tab_first[slot] < r ? 1u : 0u;
        is_dup[r] = (uint8_t)dup;
    }
    const uint32_t votes = __ballot_sync(0xffffffffu, dup != 0);
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(n_dup, (unsigned int)__popc(votes));
}
}  // namespace

cudaError_t dedup_launch(uint32_t n, const unsigned long long* key, const uint8_t* paired, uint32_t table_size, unsigned long long* tab_key,
                         uint32_t* tab_first, uint8_t* is_dup, unsigned int* n_dup, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    dedup_insert_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, key, paired, table_size - 1, tab_key, tab_first);
    dedup_mark_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, key, table_size - 1, tab_key, tab_first, is_dup, n_dup);
    return cudaGetLastError();
}

cudaError_t hc_launch_expand(uint32_t n_groups, const uint32_t* grp_c16, const uint32_t* grp_n16, const uint16_t* col16,
                             const uint8_t* col8, const uint32_t* flag4, const uint16_t* n_owned, const uint16_t* node16,
                             const uint16_t* n_cols, uint32_t* col32, uint32_t* grp_c32, cudaStream_t s) {
    if (n_groups == 0) return cudaSuccess;
    const unsigned grid = (n_groups * 32 + 127) / 128;
    if (col8) hc_expand_kernel<true><<<grid, 128, 0, s>>>(n_groups, grp_c16, grp_n16, col16, col8, flag4, n_owned, node16, n_cols, col32, grp_c32);
    else hc_expand_kernel<false><<<grid, 128, 0, s>>>(n_groups, grp_c16, grp_n16, col16, col8, flag4, n_owned, node16, n_cols, col32, grp_c32);
    return cudaGetLastError();
}

cudaError_t hc_launch_recon(const ReconRawDev& raw, const ReconGraphDev& G, const ReconWorkDev& W, int mode, int* err_flag, cudaStream_t s) {
    if (raw.n_reads == 0) return cudaSuccess;
    hc_recon_kernel<<<(raw.n_reads + kReconReadsPerCta - 1) / kReconReadsPerCta, kReconWarps * 32, 0, s>>>(raw, G, W, mode, err_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    hc_scan4096<<<W.n_slots, 1024, 0, s>>>(W.hist, W.start, W);
    hc_order_kernel<<<(raw.n_reads + 255) / 256, 256, 0, s>>>(raw.n_reads, W.n_cols, raw.mapq, W.start, W.cursor, W.order, W);
    const uint32_t n_groups = W.n_lanes / 32;
    hc_groups_kernel<<<(n_groups + 255) / 256, 256, 0, s>>>(raw.n_reads, n_groups, W.order, W.n_cols, W.n_extra,
                                                            mode == kReconHaplocart ? (uint32_t)VGL_HC_COL_BLK : (uint32_t)VGL_EUKA_COL_BLK, W.grp_cb, W.grp_mb);
    hc_scan_groups<<<1, 1024, 0, s>>>(n_groups, W.grp_cb, W.grp_mb);
    return cudaGetLastError();
}

cudaError_t hc_launch_pack(const ReconRawDev& raw, const ReconWorkDev& W, uint32_t col_cap_blks, uint32_t map_cap_blks, uint32_t* col,
                           uint32_t* xmap, uint16_t* n_cols, uint16_t* n_close, uint16_t* n_extra, uint8_t* mapq, int* err_flag,
                           cudaStream_t s) {
    if (raw.n_reads == 0) return cudaSuccess;
    const uint32_t n_groups = W.n_lanes / 32;
    hc_pack_kernel<<<(n_groups * 32 + 127) / 128, 128, 0, s>>>(raw, W, n_groups, col_cap_blks, map_cap_blks, col, xmap, n_cols, n_close,
                                                               n_extra, mapq, err_flag);
    return cudaGetLastError();
}

cudaError_t raw_launch_offsets_from_counts(const uint8_t* counts, uint32_t n, uint32_t base, uint32_t* tile_scratch, uint32_t* out,
                                           cudaStream_t s) {
    const uint32_t n_tiles = n / kScanTile + 1;  // entry n (the total) lives in tile n / kScanTile
    scan_u8_tile_sums<<<n_tiles, 256, 0, s>>>(counts, n, tile_scratch);
    scan_tiles<<<1, 1024, 0, s>>>(n_tiles, base, tile_scratch);
    scan_u8_final<<<n_tiles, 256, 0, s>>>(counts, n, tile_scratch, out);
    return cudaGetLastError();
}

cudaError_t raw_launch_widen_node16(const uint16_t* in, uint32_t n, uint32_t* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    widen_node16<<<(n + 255) / 256, 256, 0, s>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t raw_launch_first_offsets(const uint32_t* map_off, const uint16_t* first16, uint32_t n, uint32_t* map_offset_abs, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    set_first_offsets<<<(n + 255) / 256, 256, 0, s>>>(map_off, first16, n, map_offset_abs);
    return cudaGetLastError();
}

cudaError_t raw_launch_widen_u8(const uint8_t* in, uint32_t n, uint16_t* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    widen_u8<<<(n + 255) / 256, 256, 0, s>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t raw_launch_widen_u16(const uint16_t* in, uint32_t n, uint32_t* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    widen_u16<<<(n + 255) / 256, 256, 0, s>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t sym_launch_pack(const ReconRawDev& raw, const ReconWorkDev& W, int soibean, uint16_t* col, uint32_t* map, uint16_t* n_cols,
                            uint16_t* n_maps, uint16_t* seq_len, uint8_t* mapq, uint8_t* flags, uint32_t* read_index, cudaStream_t s) {
    if (raw.n_reads == 0) return cudaSuccess;
    const uint32_t n_groups = (raw.n_reads + 31) / 32;
    sym_pack_kernel<<<(n_groups * 32 + 127) / 128, 128, 0, s>>>(raw, W, n_groups, soibean, col, map, n_cols, n_maps, seq_len, mapq, flags,
                                                               read_index);
    return cudaGetLastError();
}

}  // namespace vgl
