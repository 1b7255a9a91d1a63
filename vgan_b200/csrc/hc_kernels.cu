// HaploCart read x path likelihood kernels for sm_100a (B200).
//
// The reference evaluates, for every read, every mapping i and every one of the P paths,
//     ll[p] += pathsgo[node_i][p] ? A_i : B_i                      (src/process_mapping.cpp:54-88)
// where A_i (sum of per-base log mixtures, :59-77) and B_i (sum of log sequencing-error
// probabilities, :4-24) do not depend on p.  Regrouping the same sum by node,
//     final_vec[p] = sum_i A_i + sum_n [pathsgo[n][p] == 0] * Dn,   Dn = sum_{i at node n} (B_i - A_i)
// is exact algebra (no approximation) and turns the work into
//   K1  hc_read_kernel   : stream the read batch once from HBM (TMA bulk copies into a 3-stage
//                          shared-memory ring), one warp per read, compute every mapping's (A_i, B_i)
//                          and scatter B_i - A_i into a per-CTA fp64 node vector in shared memory;
//   K2  hc_path_kernel   : one pass over the bit-packed node x path membership (7.7 MB for the
//                          HaploCart graph) with the node vector staged by TMA bulk copy;
//   K3  hc_posterior_kernel : argmax, log-sum-exp, posterior of the call and its clades.
// "Supported" is the baseline so that the few best paths (the only ones that matter for the
// posterior) are sums of few terms and carry no cancellation.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stddef.h>
#include <stdint.h>

#include "vgl_internal.h"

namespace vgl {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// ---- read kernel geometry ------------------------------------------------------------------
constexpr int kReadBlock = 768;     // 24 warps: one read per thread, one 32-lane group per warp at a time
constexpr int kMapChunk = 96;       // mappings finalised per pass (per-thread product / quality lists live in local memory)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// ---- explicit shared-space accessors (32-bit shared addresses keep ptxas on LDS / ATOMS instead of generic LD) ----
// Tables are constant for the kernel's lifetime, so the loads are non-volatile and may be scheduled freely.
__device__ __forceinline__ uint32_t lds_u16_c(uint32_t a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64_c(uint32_t a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64_c(uint32_t a) { double2 v; asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void atoms_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// streaming global loads of the batch: every byte is used exactly once, keep it out of L1
__device__ __forceinline__ uint32_t ldg_stream_u16(const uint16_t* p) { uint16_t v; asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldg_l1_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }

constexpr uint32_t kOffFPair = offsetof(HcSmallTables, f_pair);
constexpr uint32_t kOffLogErr = offsetof(HcSmallTables, log_err);
constexpr uint32_t kOffLogTab = offsetof(HcSmallTables, log_tab);
constexpr uint32_t kOffPCorrect = offsetof(HcSmallTables, p_correct);
constexpr uint32_t kOffBg = offsetof(HcSmallTables, bg);

// log(x) for x > 0 (x == 0 gives -inf) to ~1e-16: 8-bit table on the mantissa + degree-5 log1p, straight log1p
// within 2^-9 of 1 (exactly 0 at x == 1).  Table at shared byte address `tab`.
__device__ __forceinline__ double fast_log_s(double x, uint32_t tab) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = lds_v2f64_c(tab + (uint32_t)((hi >> 12) & 0xff) * 16u);
    const double d = x - 1.0;
    const bool near = fabs(d) < 0.001953125;
    const double r = near ? d : fma(m, t.x, -1.0);
    const double base = near ? 0.0 : fma((double)e, 0.6931471805599453094, t.y);
    double p = fma(r, 0.2, -0.25);
    p = fma(r, p, 0.33333333333333333333);
    p = fma(r, p, -0.5);
    p = fma(r, p, 1.0);
    const double v = fma(r, p, base);
    return hi < 0x00100000 ? -CUDART_INF : v;  // zero; products are rescaled long before they go denormal
}

// ------------------------------------------------------------------------------------------
// K1: one read per thread, one 32-lane group per warp at a time.
//
// The batch is stored lane-interleaved in 16-byte blocks (include/vgan_gpu.h), so the 32 threads of a warp read
// 512 contiguous bytes per step: perfectly coalesced streaming loads, each byte used once, no staging needed.
// Per read the thread walks its columns from the LAST to the first:
//   * the running sum of log(err) over the columns seen so far is exactly the quality sum the reference takes
//     for the mapping that starts at the current column (src/update_likelihood.cpp:40-44 reads |algnseq|
//     qualities from position_in_read; the `pos` bytes past the end count as quality 0);
//   * the running product of the per-base mixtures (src/process_mapping.cpp:59-77) is the mapping's
//     exp(A_i); one log per mapping instead of one per base.
// When the walk reaches the first column of mapping i both are parked in per-thread lists (local memory,
// lane-interleaved by the hardware); a second, branch-free pass over the mappings takes the log, forms B_i
// and scatters B_i - A_i into the CTA's node vector with a shared-memory fp64 atomic.
//
// Shared memory: [ node_acc f64[n_rows] | HcSmallTables | node_class u16[n_rows] | ac table double2[100][n_class] ]
// The ac table holds {1 - pcm, pcm * K} (src/process_mapping.cpp:33-42) for every (mapping quality, table class)
// pair, so opening a mapping costs two shared loads and no arithmetic.
// ------------------------------------------------------------------------------------------
struct ReadKernelShared {
    uint32_t acc, f, bg, logtab, node_class, ac;  // shared byte addresses
    int n_rows, n_class;
    double le0;
};

// kLatch: some read of the group holds a real quality byte >= 90 (it switches mappings to the background error
// probability, src/update_likelihood.cpp:42); the common case compiles the test away.
// kSingle: no read of the group has more than kMapChunk mappings, so one pass covers it and the per-column range
// checks of the multi-pass walk disappear.
template <bool kLatch, bool kSingle>
__device__ __forceinline__ void hc_process_group(const HcBatchDev& b, const ReadKernelShared& S, uint32_t g, int lane, double& a_sum,
                                                 bool& bad_node) {
    const double scale = -0.23025850929940456840;  // -ln(10)/10: log(get_p_seq_error(Q)) = Q * scale for Q > 2
    const uint32_t cb0 = __ldg(b.grp_col_blk + g), mb0 = __ldg(b.grp_map_blk + g);
    const int ncb = (int)(__ldg(b.grp_col_blk + g + 1) - cb0);  // 8-column blocks of this group
    const int nmb = (int)(__ldg(b.grp_map_blk + g + 1) - mb0);  // 4-mapping blocks
    const uint4* gc = reinterpret_cast<const uint4*>(b.col) + (size_t)cb0 * 32 + lane;  // block k of this lane: gc[k * 32]
    const uint4* gm = reinterpret_cast<const uint4*>(b.map) + (size_t)mb0 * 32 + lane;
    const uint32_t* gm32 = reinterpret_cast<const uint32_t*>(gm);  // record i of this lane: gm32[(i >> 2) * 128 + (i & 3)]
    const uint32_t r = g * 32 + lane;
    const int Lp = (int)__ldg(b.n_cols + r);
    const int M = (int)__ldg(b.n_maps + r);
    const int f90 = kLatch ? (int)__ldg(b.first_q90 + r) : 0xFFFF;
    const uint32_t ac_row = S.ac + 16u * (uint32_t)(S.n_class + 1) * (uint32_t)__ldg(b.mapq + r);
    const int maxm = nmb * VGL_HC_MAP_BLK;

    double plist[kMapChunk];    // exp(A_i) of the mappings of the current pass
    uint32_t slist[kMapChunk];  // quality sums from the mapping's first column on: (sum of q for q > 2) | (count of q <= 2) << 20
    // Quality sums over the columns walked so far (from the end of the read).  The single-pass walk visits every record of
    // the group's blocks; records past a read's own columns are zero and would count as quality-0 columns, so the sum
    // starts that much lower (mod 2^32) instead of testing every column against the read's length.
    uint32_t suf = kSingle ? 0u - ((uint32_t)(ncb * VGL_HC_COL_BLK - Lp) << 20) : 0u;
    int c_done = Lp;  // columns >= c_done have been walked
    for (int hi_u = maxm; hi_u > 0; hi_u -= kMapChunk) {  // one iteration when kSingle
        const int lo_u = kSingle ? 0 : max(hi_u - kMapChunk, 0);
        const int hi = min(M, hi_u), lo = kSingle ? 0 : min(M, lo_u);
        // ---- pass 1: columns, descending ----
        int i = hi - 1;  // current mapping
        int pos = -1;    // its first column
        uint32_t and_m = 0x1ffu, or_m = 0u;  // f-table index = (record & and_m) | or_m
        double aa = 1.0, cc = 0.0, prod = 1.0, a_ext = 0.0;
        // mapping records are read one mapping ahead of need
        const uint32_t* prec = gm32 + (size_t)(max(i, 0) >> 2) * 128 + (max(i, 0) & 3);
        uint32_t rec_next = i >= lo ? ldg_stream_u32(prec) : 0u;
        auto open_mapping = [&]() {
            if (i < lo) { pos = -1; return; }
            const uint32_t rec = rec_next;
            if (i > lo) {
                prec -= (i & 3) ? 1 : 125;  // previous record of this lane: same block, or the last one of the block below
                rec_next = ldg_stream_u32(prec);
            }
            pos = (int)(rec >> VGL_HC_MAP_NODE_BITS);
            const uint32_t node = rec & VGL_HC_MAP_MAX_NODE;
            // class 0xFFFF: no pangenome_map entry -- the reference would terminate; flag it and use the neutral entry
            const uint32_t cls = node < (uint32_t)S.n_rows ? lds_u16_c(S.node_class + 2u * node) : 0xFFFFu;
            bad_node |= cls == 0xFFFFu;
            const double2 ac = lds_v2f64_c(ac_row + 16u * min(cls, (uint32_t)S.n_class));
            aa = ac.x;
            cc = ac.y;
            if (kLatch) {
                // the background error mode is on for this mapping once a real quality byte >= 90 lies before the end of
                // its quality window [pos, pos + Lp)  (src/update_likelihood.cpp:40-44)
                const bool latched = f90 < pos + Lp;
                and_m = latched ? 0x100u : 0x1ffu;
                or_m = latched ? 0x200u : 0u;
            }
        };
        open_mapping();
        while (i >= lo && pos >= c_done) {  // mappings that own no column (they start at or past the end)
            plist[i - lo] = 1.0;
            slist[i - lo] = kSingle ? 0u : suf;
            --i;
            open_mapping();
        }
        int c_stop = 0;
        if (!kSingle && lo > 0 && lo < M) c_stop = (int)(ldg_stream_u32(gm32 + (size_t)(lo >> 2) * 128 + (lo & 3)) >> VGL_HC_MAP_NODE_BITS);
        const bool walking = hi > lo;
        auto column = [&](int c, uint32_t rec) {
            if (!kSingle && !(c < c_done && c >= c_stop && walking)) return;
            const uint32_t q = rec & 0xffu;
            suf += ((int)(int8_t)q > 2) ? q : (1u << 20);
            if (rec & VGL_HC_COL_VALID) {
                const double f = lds_f64_c(S.f + 8u * (kLatch ? ((rec & and_m) | or_m) : (rec & 0x1ffu)));
                const double bgv = lds_f64_c(S.bg + 8u * VGL_HC_COL_BASE(rec));
                prod *= fma(cc, f, aa * bgv);
            }
            while (c == pos) {
                if (a_ext != 0.0) {  // the product was rescaled (dozens of very unlikely bases in one mapping)
                    const uint32_t node = ldg_stream_u32(gm32 + (size_t)(i >> 2) * 128 + (i & 3)) & VGL_HC_MAP_MAX_NODE;
                    if (node < (uint32_t)S.n_rows) atoms_add_f64(S.acc + 8u * node, -a_ext);
                    a_sum += a_ext;
                    a_ext = 0.0;
                }
                plist[i - lo] = prod;
                slist[i - lo] = suf;
                prod = 1.0;
                --i;
                open_mapping();
            }
        };
        // column records come in blocks of 8, fetched two blocks ahead of the walk
        uint4 cb_cur = ncb > 0 ? ldg_stream_v4(gc + (size_t)(ncb - 1) * 32) : make_uint4(0, 0, 0, 0);
        uint4 cb_nxt = ncb > 1 ? ldg_stream_v4(gc + (size_t)(ncb - 2) * 32) : make_uint4(0, 0, 0, 0);
        for (int cb = ncb - 1; cb >= 0; --cb) {
            const uint4 cb_nxt2 = cb > 1 ? ldg_stream_v4(gc + (size_t)(cb - 2) * 32) : make_uint4(0, 0, 0, 0);
            const int c8 = cb * VGL_HC_COL_BLK;
            column(c8 + 7, cb_cur.w >> 16);
            column(c8 + 6, cb_cur.w & 0xffffu);
            column(c8 + 5, cb_cur.z >> 16);
            column(c8 + 4, cb_cur.z & 0xffffu);
            column(c8 + 3, cb_cur.y >> 16);
            column(c8 + 2, cb_cur.y & 0xffffu);
            column(c8 + 1, cb_cur.x >> 16);
            column(c8 + 0, cb_cur.x & 0xffffu);
            if (prod < 1e-100) {  // keep the product normal (eight more factors cannot take it below 1e-200)
                a_ext += fast_log_s(prod, S.logtab);
                prod = 1.0;
            }
            cb_cur = cb_nxt;
            cb_nxt = cb_nxt2;
        }
        if (!kSingle && walking) c_done = c_stop;
        // ---- pass 2: mappings, four at a time so that the list loads of a block are in flight together ----
        for (int mbk = lo_u >> 2; mbk < (hi_u >> 2); ++mbk) {
            const uint4 rb = ldg_stream_v4(gm + (size_t)mbk * 32);
            const uint32_t recs[4] = {rb.x, rb.y, rb.z, rb.w};
            const int k0 = mbk * 4 - lo_u;
            double pr[4];
            uint32_t sv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool on = mbk * 4 + e < hi;
                pr[e] = on ? plist[k0 + e] : 1.0;
                sv[e] = on ? slist[k0 + e] : 0u;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (mbk * 4 + e < hi) {
                    const uint32_t rm = recs[e];
                    const uint32_t node = rm & VGL_HC_MAP_MAX_NODE;
                    const double A = fast_log_s(pr[e], S.logtab);
                    // log(err) over [pos, pos + Lp): the stored columns from pos on plus `pos` bytes past the end that
                    // count as quality 0  
                    const double B = fma((double)(sv[e] & 0xfffffu), scale, (double)((sv[e] >> 20) + (rm >> VGL_HC_MAP_NODE_BITS)) * S.le0);
                    if (node < (uint32_t)S.n_rows) atoms_add_f64(S.acc + 8u * node, B - A);
                    a_sum += A;
                }
            }
        }
        if (kSingle) break;
    }
}

__global__ void __launch_bounds__(kReadBlock, 1)
hc_read_kernel(HcBatchDev b, HcTablesDev t, double* __restrict__ cta_node_acc, double* __restrict__ cta_a_total,
               int* __restrict__ err_flag, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    size_t off = 0;
    double* s_acc = reinterpret_cast<double*>(smem_raw);
    off += ((size_t)t.n_rows * sizeof(double) + 15) & ~(size_t)15;
    HcSmallTables* s_tab = reinterpret_cast<HcSmallTables*>(smem_raw + off);
    off += (sizeof(HcSmallTables) + 15) & ~(size_t)15;
    uint16_t* s_class = reinterpret_cast<uint16_t*>(smem_raw + off);
    off += ((size_t)t.n_rows * sizeof(uint16_t) + 15) & ~(size_t)15;
    double2* s_ac = reinterpret_cast<double2*>(smem_raw + off);  // [100][n_class + 1]
    __shared__ double s_warp_a[kReadBlock / 32];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    for (int n = tid; n < t.n_rows; n += kReadBlock) {
        s_acc[n] = 0.0;
        s_class[n] = t.node_class[n];
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(t.small);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_tab);
        for (int k = tid; k < (int)(sizeof(HcSmallTables) / 4); k += kReadBlock) dst[k] = src[k];
    }
    __syncthreads();
    {
        // {1 - pcm, pcm * K} for every (mapping quality, class): p_correctly_mapped = (1 - p_incorrect[mq]) * mappability
        // (src/process_mapping.cpp:41); consensus-FASTA mode uses (1 - bep) * K alone (:73)
        const bool fasta = s_tab->is_fasta != 0;
        const int row = t.n_class + 1;  // entry n_class of every row is the neutral {1, 0} used for unknown nodes
        for (int k = tid; k < 100 * row; k += kReadBlock) {
            const int mq = k / row, cls = k - mq * row;
            if (cls == t.n_class) { s_ac[k] = make_double2(1.0, 0.0); continue; }
            const double2 mk = t.class_tab[cls];  // {mappability, match_prob}
            const double pcm = s_tab->p_correct[mq] * mk.x;
            s_ac[k] = make_double2(fasta ? 0.0 : 1.0 - pcm, (fasta ? s_tab->one_minus_bep : pcm) * mk.y);
        }
    }
    __syncthreads();

    ReadKernelShared S;
    S.acc = smem_u32(s_acc);
    S.f = smem_u32(s_tab) + kOffFPair;
    S.bg = smem_u32(s_tab) + kOffBg;
    S.logtab = smem_u32(s_tab) + kOffLogTab;
    S.node_class = smem_u32(s_class);
    S.ac = smem_u32(s_ac);
    S.n_rows = t.n_rows;
    S.n_class = t.n_class;
    S.le0 = s_tab->log_err[0];

    double a_sum = 0.0;
    bool bad_node = false;
    const uint32_t warps_total = gridDim.x * (kReadBlock / 32);
    for (uint32_t g = blockIdx.x * (kReadBlock / 32) + warp; g < b.n_groups; g += warps_total) {
        const bool latch = __any_sync(kFull, __ldg(b.first_q90 + g * 32 + lane) != 0xFFFFu);
        const uint32_t n_map_blk = __ldg(b.grp_map_blk + g + 1) - __ldg(b.grp_map_blk + g);
        if (n_map_blk == 0) continue;  // no read of this group has a mapping
        const bool single = n_map_blk * VGL_HC_MAP_BLK <= (uint32_t)kMapChunk;
        if (single && !latch) hc_process_group<false, true>(b, S, g, lane, a_sum, bad_node);
        else if (single) hc_process_group<true, true>(b, S, g, lane, a_sum, bad_node);
        else hc_process_group<true, false>(b, S, g, lane, a_sum, bad_node);
    }

    if (__any_sync(kFull, bad_node) && lane == 0) atomicOr(err_flag, VGL_ERR_BAD_NODE);
    a_sum = warp_sum_f64(a_sum);
    if (lane == 0) s_warp_a[warp] = a_sum;
    __syncthreads();
    double* my_acc = cta_node_acc + (size_t)blockIdx.x * (size_t)t.n_rows;
    for (int n = tid; n < t.n_rows; n += kReadBlock) {
        const double v = s_acc[n];
        if (v != 0.0) my_acc[n] += v;
    }
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kReadBlock / 32; ++w) tot += s_warp_a[w];
        cta_a_total[blockIdx.x] += tot;
    }
}

// ------------------------------------------------------------------------------------------
// K2a: node_acc[n] = sum over CTAs (fixed order, Neumaier-compensated within each of 8 slices, then the slices)
// ------------------------------------------------------------------------------------------
constexpr int kRedNodes = 32;   // nodes per CTA (one per lane)
constexpr int kRedSlices = 8;   // warps per CTA, each sums every 8th per-CTA vector

__device__ __forceinline__ void neumaier_add(double& s, double& comp, double v) {
    const double tsum = s + v;
    comp += (fabs(s) >= fabs(v)) ? ((s - tsum) + v) : ((v - tsum) + s);
    s = tsum;
}

__global__ void __launch_bounds__(kRedNodes * kRedSlices)
hc_reduce_ctas_kernel(const double* __restrict__ cta_node_acc, const double* __restrict__ cta_a_total, int n_ctas,
                      int n_rows, int n_rows_pad, double* __restrict__ node_acc, double* __restrict__ a_total) {
    __shared__ double s_part[kRedSlices][kRedNodes];
    __shared__ double s_comp[kRedSlices][kRedNodes];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int n = blockIdx.x * kRedNodes + lane;
    double s = 0.0, comp = 0.0;
    if (n < n_rows)
        for (int c = slice; c < n_ctas; c += kRedSlices) neumaier_add(s, comp, __ldg(cta_node_acc + (size_t)c * n_rows + n));
    s_part[slice][lane] = s;
    s_comp[slice][lane] = comp;
    __syncthreads();
    if (slice == 0 && n < n_rows_pad) {
        double ts = 0.0, tc = 0.0;
#pragma unroll
        for (int k = 0; k < kRedSlices; ++k) {
            neumaier_add(ts, tc, s_part[k][lane]);
            tc += s_comp[k][lane];
        }
        node_acc[n] = ts + tc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double ts = 0.0, tc = 0.0;
        for (int c = 0; c < n_ctas; ++c) neumaier_add(ts, tc, cta_a_total[c]);
        a_total[0] = ts + tc;
    }
}

// ------------------------------------------------------------------------------------------
// K2b: chunk_ll[chunk][p] = sum over the chunk's nodes n with pathsgo[n][p] == 0 of node_acc[n]
// ------------------------------------------------------------------------------------------
constexpr int kPathTile = 128;       // paths per CTA (one per thread)
constexpr int kWordsPerChunk = 10;   // 320 nodes per CTA

__global__ void __launch_bounds__(kPathTile)
hc_path_kernel(HcTablesDev t, const double* __restrict__ node_acc, double* __restrict__ chunk_ll) {
    __shared__ __align__(16) double s_d[kWordsPerChunk * 32];
    __shared__ __align__(8) unsigned long long s_bar;
    const int chunk = blockIdx.y;
    const int w0 = chunk * kWordsPerChunk;
    const int nw = min(kWordsPerChunk, t.n_words - w0);
    const int p = blockIdx.x * kPathTile + threadIdx.x;

    // Stage this chunk of the node vector with one TMA bulk copy (cp.async.bulk -> UBLKCP).
    const uint32_t bar = smem_u32(&s_bar);
    const uint32_t bytes = (uint32_t)nw * 32u * 8u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(s_d), node_acc + (size_t)w0 * 32, bytes, bar);
    }
    // Issue the membership loads while the bulk copy is in flight.
    uint32_t words[kWordsPerChunk];
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k)
        words[k] = (k < nw && p < t.n_paths) ? __ldg(t.node_bits + (size_t)(w0 + k) * t.n_paths + p) : 0xffffffffu;
    mbar_wait(bar, 0);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k) {
        const uint32_t unsup = ~words[k];
        if (k < nw) {
#pragma unroll
            for (int bit = 0; bit < 32; ++bit) {
                const double d = s_d[k * 32 + bit];
                acc += ((unsup >> bit) & 1u) ? d : 0.0;
            }
        }
    }
    if (p < t.n_paths) chunk_ll[(size_t)chunk * t.n_paths + p] = acc;
}

__global__ void hc_path_sum_kernel(const double* __restrict__ chunk_ll, int n_chunks, int n_paths,
                                   const double* __restrict__ a_total, double* __restrict__ ll) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += chunk_ll[(size_t)c * n_paths + p];
    ll[p] = a_total[0] + s;
}

__global__ void hc_sum_parts_kernel(const double* __restrict__ parts, int n_parts, int n_paths, double* __restrict__ ll) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    double s = 0.0;
    for (int k = 0; k < n_parts; ++k) s += parts[(size_t)k * n_paths + p];
    ll[p] = s;
}

// ------------------------------------------------------------------------------------------
// K3: argmax + log-softmax posterior + clade support, one CTA.  exp(ll[p] - max) is evaluated once
// per path into shared memory; every clade is then a masked sum over its descendant bitmap.
// ------------------------------------------------------------------------------------------
constexpr int kPostBlock = 1024;

__global__ void __launch_bounds__(kPostBlock)
hc_posterior_kernel(const double* __restrict__ ll, int n_paths, int row_words, const uint32_t* __restrict__ clade_bits,
                    const int32_t* __restrict__ anc_off, const int32_t* __restrict__ anc_clade, double* __restrict__ post,
                    int post_cap, double* __restrict__ out_scalars, int32_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) double s_e[];  // [n_paths] exp(ll - max)
    __shared__ double s_val[kPostBlock / 32];
    __shared__ int s_idx[kPostBlock / 32];
    __shared__ double s_max;
    __shared__ int s_arg;
    __shared__ double s_sum;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // argmax, first max wins (std::max_element, src/HaploCart.cpp:423)
    double best = -CUDART_INF;
    int arg = 0x7fffffff;
    for (int p = tid; p < n_paths; p += kPostBlock) {
        const double v = ll[p];
        if (arg == 0x7fffffff || v > best) { best = v; arg = p; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(kFull, best, d);
        const int oa = __shfl_xor_sync(kFull, arg, d);
        if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = arg; }
    __syncthreads();
    if (tid == 0) {
        double bb = s_val[0];
        int ba = s_idx[0];
        for (int w = 1; w < kPostBlock / 32; ++w) {
            const double ob = s_val[w];
            const int oa = s_idx[w];
            if (oa != 0x7fffffff && (ba == 0x7fffffff || ob > bb || (ob == bb && oa < ba))) { bb = ob; ba = oa; }
        }
        s_max = bb;
        s_arg = ba;
    }
    __syncthreads();
    const double mx = s_max;
    const int call = s_arg;

    // total = log sum_p exp(ll[p])   (sum_log_likelihoods, src/get_posterior.cpp:78-85)
    double se = 0.0;
    for (int p = tid; p < n_paths; p += kPostBlock) {
        const double e = exp(ll[p] - mx);
        s_e[p] = e;
        se += e;
    }
    se = warp_sum_f64(se);
    __syncthreads();
    if (lane == 0) s_val[warp] = se;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kPostBlock / 32; ++w) tot += s_val[w];
        s_sum = tot;
        out_scalars[0] = mx + log(tot);
        out_scalars[1] = mx;
        out_idx[0] = call;
        if (post_cap > 0) post[0] = 1.0 / tot;  // exp(ll[call] - total), src/get_posterior.cpp:103-106
    }
    __syncthreads();
    const double tot = s_sum;

    // clade support: one warp per listed ancestor clade (src/get_posterior.cpp:109-123)
    int n_anc = 0;
    if (anc_off != nullptr) n_anc = anc_off[call + 1] - anc_off[call];
    if (n_anc > post_cap - 1) n_anc = max(post_cap - 1, 0);
    for (int k = warp; k < n_anc; k += kPostBlock / 32) {
        const int clade = anc_clade[anc_off[call] + k];
        const uint32_t* bits = clade_bits + (size_t)clade * row_words;
        double s = 0.0;
        for (int w = lane; w < row_words; w += 32) {
            uint32_t word = bits[w];
            const int pbase = w * 32;
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                if (pbase + bit < n_paths) s += s_e[pbase + bit];
            }
        }
        s = warp_sum_f64(s);
        if (lane == 0) post[1 + k] = s / tot;
    }
    if (tid == 0) out_idx[1] = 1 + n_anc;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
static size_t read_smem_bytes(int n_rows, int n_class) {
    size_t s = 0;
    s += ((size_t)n_rows * sizeof(double) + 15) & ~(size_t)15;
    s += (sizeof(HcSmallTables) + 15) & ~(size_t)15;
    s += ((size_t)n_rows * sizeof(uint16_t) + 15) & ~(size_t)15;
    s += (size_t)100 * (n_class + 1) * sizeof(double2);
    return (s + 127) & ~(size_t)127;
}

HcReadLaunch hc_read_launch_config(int n_rows, int n_class, int sm_count, int max_smem_optin) {
    HcReadLaunch c;
    c.grid = sm_count;
    c.block = kReadBlock;
    c.smem = read_smem_bytes(n_rows, n_class);
    if (c.smem + 256 > (size_t)max_smem_optin) c.grid = 0;  // caller reports VGL_ERR_INVALID (tables too large for this layout)
    return c;
}

cudaError_t hc_read_configure(size_t smem) {
    return cudaFuncSetAttribute(hc_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t hc_launch_read_kernel(const HcReadLaunch& cfg, const HcBatchDev& batch, const HcTablesDev& tab,
                                  double* cta_node_acc, double* cta_a_total, int* err_flag,
                                  unsigned long long* stats, cudaStream_t s) {
    hc_read_kernel<<<cfg.grid, cfg.block, cfg.smem, s>>>(batch, tab, cta_node_acc, cta_a_total, err_flag, stats);
    return cudaGetLastError();
}

cudaError_t hc_launch_reduce_ctas(const double* cta_node_acc, const double* cta_a_total, int n_ctas, int n_rows,
                                  int n_rows_pad, double* node_acc, double* a_total, cudaStream_t s) {
    hc_reduce_ctas_kernel<<<(n_rows_pad + kRedNodes - 1) / kRedNodes, kRedNodes * kRedSlices, 0, s>>>(
        cta_node_acc, cta_a_total, n_ctas, n_rows, n_rows_pad, node_acc, a_total);
    return cudaGetLastError();
}

int hc_path_chunks(int n_words) { return (n_words + kWordsPerChunk - 1) / kWordsPerChunk; }

cudaError_t hc_launch_path_kernel(const HcTablesDev& tab, const double* node_acc, double* chunk_ll, cudaStream_t s) {
    dim3 grid((tab.n_paths + kPathTile - 1) / kPathTile, hc_path_chunks(tab.n_words));
    hc_path_kernel<<<grid, kPathTile, 0, s>>>(tab, node_acc, chunk_ll);
    return cudaGetLastError();
}

cudaError_t hc_launch_path_sum(const double* chunk_ll, int n_chunks, int n_paths, const double* a_total, double* ll,
                               cudaStream_t s) {
    const int block = 128;
    hc_path_sum_kernel<<<(n_paths + block - 1) / block, block, 0, s>>>(chunk_ll, n_chunks, n_paths, a_total, ll);
    return cudaGetLastError();
}

cudaError_t hc_launch_sum_parts(const double* parts, int n_parts, int n_paths, double* ll, cudaStream_t s) {
    const int block = 128;
    hc_sum_parts_kernel<<<(n_paths + block - 1) / block, block, 0, s>>>(parts, n_parts, n_paths, ll);
    return cudaGetLastError();
}

size_t hc_posterior_smem(int n_paths) { return (size_t)n_paths * sizeof(double); }

cudaError_t hc_posterior_configure(size_t smem) {
    return cudaFuncSetAttribute(hc_posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t hc_launch_posterior(const double* ll, int n_paths, int row_words, const uint32_t* clade_bits,
                                const int32_t* anc_off, const int32_t* anc_clade, double* post, int post_cap,
                                double* out_scalars, int32_t* out_idx, cudaStream_t s) {
    hc_posterior_kernel<<<1, kPostBlock, hc_posterior_smem(n_paths), s>>>(ll, n_paths, row_words, clade_bits, anc_off,
                                                                          anc_clade, post, post_cap, out_scalars, out_idx);
    return cudaGetLastError();
}

}  // namespace vgl
