// HaploCart read x path likelihood kernels for sm_100a (B200).
//
// The reference evaluates, for every read, every mapping i and every one of the P paths,
//     ll[p] += pathsgo[node_i][p] ? A_i : B_i                      (src/process_mapping.cpp:54-88)
// where A_i (sum of per-base log mixtures, :59-77) and B_i (sum of log sequencing-error
// probabilities, :4-24) do not depend on p.  Regrouping the same sum by node,
//     final_vec[p] = sum_i A_i + sum_n [pathsgo[n][p] == 0] * Dn,   Dn = sum_{i at node n} (B_i - A_i)
// is exact algebra (no approximation) and turns the work into
//   K1  hc_read_kernel   : stream the read batch once from HBM, one read per thread, compute every
//                          mapping's (A_i, B_i) and add B_i - A_i into one of VGL_HC_ACC_COPIES fp64 node
//                          vectors (global memory, L2 resident) with fire-and-forget reductions;
//   K2  hc_path_kernel   : one pass over the bit-packed node x path membership (7.7 MB for the
//                          HaploCart graph) with the node vectors staged by TMA bulk copies;
//   K3  hc_posterior_kernel : argmax, log-sum-exp, posterior of the call and its clades;
//   hc_peer_push / hc_peer_sum : the multi-GPU exchange of the per-path vector over NVLink peer memory.
// "Supported" is the baseline so that the few best paths (the only ones that matter for the
// posterior) are sums of few terms and carry no cancellation.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stddef.h>
#include <stdint.h>

#include <type_traits>

#include "vgl_internal.h"

namespace vgl {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// ---- read kernel geometry ------------------------------------------------------------------
#ifndef VGL_HC_READ_BLOCK
#define VGL_HC_READ_BLOCK 768
#endif
constexpr int kReadBlock = VGL_HC_READ_BLOCK;  // 24 warps: one read per thread, one 32-lane group per warp at a time

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
// fire-and-forget fp64 add performed by the L2 atomic unit (SASS: RED.E.ADD.F64).  Shared memory has no native fp64 add:
// red.shared.add.f64 is a compare-and-swap loop that costs ~37 shared wavefronts per warp instruction.
__device__ __forceinline__ void red_add_f64(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
// streaming global loads of the batch: every byte is used exactly once, keep it out of L1
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }

constexpr uint32_t kOffFPair = offsetof(HcSmallTables, f_col);
constexpr uint32_t kOffLogTab = offsetof(HcSmallTables, log_tab);
constexpr uint32_t kOffBg = offsetof(HcSmallTables, bg);
constexpr uint32_t kOffPCorrect = offsetof(HcSmallTables, p_correct);

// log(x) for x > 0 (x == 0 gives -inf): 8-bit table on the mantissa + degree-5 log1p of the remainder (|r| < 2^-9).
// Absolute error ~1e-16 (one ulp of ln 2 when the exponent is non-zero), which is what a SUM of logs needs; the result
// for x within 2^-9 of 1 is not relatively accurate and does not have to be.  Table at shared byte address `tab`.
__device__ __forceinline__ double fast_log_s(double x, uint32_t tab) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = lds_v2f64_c(tab + (uint32_t)((hi >> 12) & 0xff) * 16u);
    const double r = fma(m, t.x, -1.0);
    const double base = fma((double)e, 0.6931471805599453094, t.y);
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
// Every global load of the walk is issued by the whole warp at a fixed point of the loop, well ahead of its use:
// a load issued from divergent code would make every lane wait for the most recent one of its neighbours
// (scoreboards are per warp), i.e. a full HBM round trip per mapping.  That is why the column record carries
// everything the walk needs (node id, OPEN / CLOSE marks) and no per-mapping array is read here.
// Per read the thread walks its columns from the LAST to the first:
//   * the running sum of log(err) over the columns seen so far is exactly the quality sum the reference takes
//     for the mapping that starts at the current column (src/update_likelihood.cpp:40-44 reads |algnseq|
//     qualities from position_in_read; the `pos` bytes past the end count as quality 0);
//   * the running product of the per-base mixtures (src/process_mapping.cpp:59-77) is the mapping's
//     exp(A_i); one log per mapping instead of one per base;
//   * a second running product over ALL valid columns of the read gives sum_i A_i with one log per read, so a mapping's
//     own A_i is needed only where it feeds D[n]: not at all for a node every path goes through.
// At a CLOSE mark (first column of a mapping) whose node some path avoids, the pair is pushed onto the warp's queue in
// shared memory (ballot + popc give every closing lane its slot, one 16-byte store each).  After every block of four
// columns, while the queue holds 32 entries or more, the warp drains them, one entry per lane: log of the product, B_i from the quality sums, and
// B_i - A_i added to the CTA's node vector (global memory, L2 resident) with a fire-and-forget fp64 reduction.
// Entries are self-contained, so the queue simply carries over from one group to the next and the finalisation
// always runs on full warps.
// (Per-thread lists in local memory, the obvious alternative, cost ~9 KB of L2 traffic per read.)
//
// Shared memory: [ HcSmallTables | node_class u16[n_rows + 1] | ac table double2[100][n_class + 1]
//                  | close queues uint4[warps][kQueueSlots], every ring on its own boundary ]
// The ac table holds {1 - pcm, pcm * K} (src/process_mapping.cpp:33-42) for every (mapping quality, table class)
// pair, so opening a mapping costs one shared load and no arithmetic.
// ------------------------------------------------------------------------------------------
constexpr int kAccCopies = VGL_HC_ACC_COPIES;  // node vectors the CTAs add into (CTA b uses copy b % kAccCopies)
// Columns walked between two looks at the queue.  A drain leaves < 32 entries and every column adds <= 32, so the ring needs
// 32 * (1 + kDrainEvery) slots; with more than one column per look the columns in between form one basic block and their
// independent work overlaps.
#ifndef VGL_HC_DRAIN_EVERY
#define VGL_HC_DRAIN_EVERY 4
#endif
constexpr int kDrainEvery = VGL_HC_DRAIN_EVERY;
static_assert(kDrainEvery == 1 || kDrainEvery == 2 || kDrainEvery == 4, "a look at the queue every 1, 2 or 4 columns");
constexpr int kQueueSlots = kDrainEvery == 1 ? 64 : kDrainEvery == 2 ? 128 : 256;  // per warp, a power of two

struct ReadKernelShared {
    uint32_t f, bg, logtab, node_class, ac, queue, p_correct;  // shared byte addresses (queue: this warp's ring)
    double* acc;                                                // this CTA's node vector (global, lives in L2)
    const double2* class_tab;                                   // [n_class + 1] {mappability, K}; last entry neutral
    int n_rows, n_class, ac_stride;  // ac_stride: bytes per mapping-quality row of the mode-0 table
    bool fasta;
    double le0, one_minus_bep;
};

// What a lane needs to know about its read of a group; loaded one group ahead of use.
struct GroupHdr {
    const uint4* gc;  // column block k of this lane: gc[k * 32]
    int ncb;          // 4-column blocks of the group
    int n_cols;       // this read's own columns
    uint32_t mapq;
    int n_extra;      // this read's empty mappings
};
__device__ __forceinline__ GroupHdr hc_load_hdr(const HcBatchDev& b, uint32_t g, int lane) {
    GroupHdr h;
    h.gc = nullptr;
    h.ncb = h.n_cols = h.n_extra = 0;
    h.mapq = 0;
    if (g < b.n_groups) {
        const uint32_t cb0 = __ldg(b.grp_col_blk + g);
        h.ncb = (int)(__ldg(b.grp_col_blk + g + 1) - cb0);
        h.gc = reinterpret_cast<const uint4*>(b.col) + (size_t)cb0 * 32 + lane;
        const uint32_t r = g * 32 + lane;
        h.n_cols = (int)__ldg(b.n_cols + r);
        h.mapq = __ldg(b.mapq + r);
        h.n_extra = (int)__ldg(b.n_extra + r);
    }
    return h;
}
__device__ __forceinline__ uint4 hc_fetch(const GroupHdr& h, int cb) {
    return cb >= 0 ? ldg_stream_v4(h.gc + (size_t)cb * 32) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// Finalise one queued mapping: entry = {exp(A) lo, hi, quality sums, node | first column << 17}.
// Only mappings whose D[n] is read are queued: a node every path goes through (ALLPATHS) never shows up in the sum
// over unsupported nodes, and the sum of all A_i comes from the read-level product (see `pall` below).
__device__ __forceinline__ void hc_finalise(const ReadKernelShared& S, const uint4& e) {
    const double scale = -0.23025850929940456840;  // -ln(10)/10: log(get_p_seq_error(Q)) = Q * scale for Q > 2
    const uint32_t node = e.w & VGL_HC_NO_NODE, pos = (e.w >> 17) & 0xfffu;
    const double A = fast_log_s(__hiloint2double((int)e.y, (int)e.x), S.logtab);
    if (node < (uint32_t)S.n_rows) {
        // log(err) over [pos, pos + n): the stored columns from pos on plus `pos` bytes past the end that count as
        // quality 0  
        const double B = fma((double)(e.z & 0xfffffu), scale, (double)((e.z >> 20) + pos) * S.le0);
        red_add_f64(S.acc + node, B - A);
    }
}

// One group.  On entry v0, v1 hold its last two column blocks (the walk is descending); on exit they hold those of
// the warp's next group `gn`, whose header is loaded here as well: neither the header nor the first blocks of a
// group are ever waited for.  q_head / q_tail: the warp's close queue (uniform across the warp).
template <int kAcMode>
__device__ __forceinline__ void hc_process_group(const HcBatchDev& b, const ReadKernelShared& S, uint32_t g, uint32_t gn, const GroupHdr& H,
                                                 GroupHdr& Hn, uint4& v0, uint4& v1, uint32_t& q_head, uint32_t& q_tail, int lane,
                                                 double& a_sum, uint32_t& bad_cls, bool& bad_node) {
    const double scale = -0.23025850929940456840;
    Hn = hc_load_hdr(b, gn, lane);
    const int ncb = H.ncb;
    // kAcMode 0: S.ac is the {1 - pcm, pcm * K} table [100][n_class + 1] in shared memory -> row of this read's mapping quality
    // kAcMode 1: S.ac is the class table {mappability, K} [n_class + 1] in shared memory (too many classes for mode 0)
    // kAcMode 2: the class table stays in global memory (S.class_tab)
    const uint32_t ac_row = kAcMode == 0 ? S.ac + (uint32_t)S.ac_stride * H.mapq : S.ac;
    const double pc = kAcMode == 0 ? 0.0 : lds_f64_c(S.p_correct + 8u * H.mapq);  // 1 - p_incorrect_map[mapq]
    double mk_m = 0.0, mk_k = 0.0;                                                 // modes 1, 2: {mappability, K} of the open mapping
    const uint32_t lt_mask = (1u << lane) - 1u;
    int x = H.n_extra - 1;  // empty mappings (no column of their own): rare, handled on the side
    bool rare = x >= 0;     // ... as is a rescaled product
    // Quality sums over the columns walked so far (from the end of the read): (sum of q for q > 2) | (count of q <= 2) << 20.
    // The walk visits every record of the group's blocks; records past a read's own columns are zero and would count as
    // quality-0 columns, so the sum starts that much lower (mod 2^32) instead of testing every column against the
    // read's length.
    uint32_t suf = 0u - ((uint32_t)(ncb * VGL_HC_COL_BLK - H.n_cols) << 20);
    double aa = 1.0, cc = 0.0, prod = 1.0, a_ext = 0.0;
    // product of EVERY valid column's mixture of this read: sum_i A_i = log(pall), one log per read instead of one per
    // mapping; only the mappings whose node some path avoids need their own A_i (for D[n])
    double pall = 1.0;
    int n_scale = 0;  // times pall was multiplied by 2^400
    uint32_t open_cls = 0;  // node-table entry of the mapping being walked

    auto extra_rec = [&](int i) -> uint32_t {
        const uint32_t* gx = b.map + ((size_t)__ldg(b.grp_map_blk + g) * 32 + lane) * VGL_HC_MAP_BLK;
        return __ldg(gx + (size_t)(i >> 2) * 128 + (i & 3));
    };
    auto add_extra = [&](uint32_t xn, double Bx) {
        if (xn < (uint32_t)S.n_rows && (lds_u16_c(S.node_class + 2u * xn) & VGL_HC_CLASS_MASK) != VGL_HC_CLASS_NONE) red_add_f64(S.acc + xn, Bx);
        else bad_node = true;
    };

    // kRare: some lane of the warp has a rescaled product or empty mappings pending; the common instantiation carries
    // neither test.  Predicated loads / stores (inline PTX) instead of branches: at every column SOME lane opens or
    // closes a mapping, so a branch would be taken by the warp anyway and only add reconvergence overhead.
    auto column = [&](auto rare_tag, int c, uint32_t rec, double f, double bgv, uint32_t cls_raw) {
        constexpr bool kRare = decltype(rare_tag)::value;
        const uint32_t q = VGL_HC_COL_QUAL(rec);   // canonical: 0 stands for every quality whose error probability is 0.25
        suf += q ? q : (1u << 20);
        const uint32_t open = rec & VGL_HC_COL_OPEN;
        // table entry: class id | ALLPATHS bit.  Class VGL_HC_CLASS_NONE: no pangenome_map entry -- the reference would
        // terminate; flag it and use the neutral entry
        const uint32_t cls = cls_raw & VGL_HC_CLASS_MASK;
        const uint32_t fetch = open;  // (fetching only when the class differs from the previous mapping's saves L1 wavefronts but costs as many issue slots)
        open_cls = open ? cls_raw : open_cls;
        bad_cls = max(bad_cls, cls);  // cls_raw is 0 where no mapping opens (predicated load)
        if (kAcMode == 0) {
            asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0,%1}, [%2];\n\t}"
                : "+d"(aa), "+d"(cc)
                : "r"(ac_row + 16u * min(cls, (uint32_t)S.n_class)), "r"(fetch));
        } else {
            if (kAcMode == 1)
                asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0,%1}, [%2];\n\t}"
                    : "+d"(mk_m), "+d"(mk_k)
                    : "r"(ac_row + 16u * min(cls, (uint32_t)S.n_class)), "r"(fetch));
            else
                asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.v2.f64 {%0,%1}, [%2];\n\t}"
                    : "+d"(mk_m), "+d"(mk_k)
                    : "l"(S.class_tab + min(cls, (uint32_t)S.n_class)), "r"(fetch));
            // p_correctly_mapped = (1 - p_incorrect[mq]) * mappability (src/process_mapping.cpp:41); consensus-FASTA mode
            // uses (1 - bep) * K alone (:73)
            const double pcm = pc * mk_m;
            aa = S.fasta ? 0.0 : 1.0 - pcm;
            cc = (S.fasta ? S.one_minus_bep : pcm) * mk_k;
        }
        const double fac0 = fma(cc, f, aa * bgv);
        const double fac = (rec & VGL_HC_COL_VALID) ? fac0 : 1.0;
        prod *= fac;
        pall *= fac;
        const uint32_t close = rec & VGL_HC_COL_CLOSE;
        if (kRare) {
            if (close && rare) {
                const uint32_t node = VGL_HC_COL_NODE(rec);
                if (a_ext != 0.0) {  // the product was rescaled (dozens of very unlikely bases in one mapping)
                    if (node < (uint32_t)S.n_rows) red_add_f64(S.acc + node, -a_ext);
                    a_ext = 0.0;
                }
                // empty mappings at or past this column: A = 0, B as for a mapping that starts here (past the last
                // column: no stored quality at all)
                const double B = fma((double)(suf & 0xfffffu), scale, (double)((suf >> 20) + (uint32_t)c) * S.le0);
                while (x >= 0) {
                    const uint32_t xr = extra_rec(x);
                    const uint32_t xp = xr >> VGL_HC_MAP_NODE_BITS;
                    if (xp < (uint32_t)c) break;
                    add_extra(xr & VGL_HC_MAP_MAX_NODE, xp == (uint32_t)c ? B : (double)xp * S.le0);
                    --x;
                }
                rare = x >= 0;
            }
        }
        // push {exp(A), quality sums, node | column} unless every path goes through the node; a CLOSE on a column nobody
        // owns only served the empty mappings
        static_assert(VGL_HC_CLASS_ALLPATHS == 2 * VGL_HC_COL_CLOSE, "ALLPATHS >> 1 lines up with CLOSE");
        const uint32_t push = ((rec & ~(open_cls >> 1) & VGL_HC_COL_CLOSE) != 0u && rec < (VGL_HC_NO_NODE << VGL_HC_COL_NODE_SHIFT)) ? 1u : 0u;
        const uint32_t m = __ballot_sync(kFull, push);
        const uint32_t slot = q_tail + __popc(m & lt_mask);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"r"(S.queue | ((16u * slot) & (16u * (kQueueSlots - 1)))),
                     "r"((uint32_t)__double2loint(prod)), "r"((uint32_t)__double2hiint(prod)), "r"(suf),
                     "r"(VGL_HC_COL_NODE(rec) | ((uint32_t)c << 17)), "r"(push)
                     : "memory");
        prod = close ? 1.0 : prod;
        q_tail += __popc(m);
    };
    auto drain = [&]() {
#pragma unroll 1
        while (q_tail - q_head >= 32u) {
            __syncwarp();
            const uint4 e = lds_v4(S.queue | ((16u * (q_head + lane)) & (16u * (kQueueSlots - 1))));
            __syncwarp();
            q_head += 32u;
            hc_finalise(S, e);
        }
    };
    auto block = [&](int cb, const uint4& v) {
        const uint32_t rec[4] = {v.x, v.y, v.z, v.w};
        // the table lookups of the four columns do not depend on the walk's state: issue them together
        double f[4], bgv[4];
        uint32_t cls[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[j] = lds_f64_c(S.f + 8u * (rec[j] & 0xfffu));
            bgv[j] = lds_f64_c(S.bg + 8u * VGL_HC_COL_BASE(rec[j]));
            // the class is needed where a mapping opens only: predicated, so that the other lanes cost no bank conflicts
            uint32_t c32 = 0;  // (a 16-bit load may target a 32-bit register: zero-extended, no pack / unpack around it)
            asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u16 %0, [%1];\n\t}"
                : "+r"(c32)
                : "r"(S.node_class + 2u * min(VGL_HC_COL_NODE(rec[j]), (uint32_t)S.n_rows)), "r"(rec[j] & VGL_HC_COL_OPEN));  // entry n_rows: CLASS_NONE
            cls[j] = c32;
        }
        const int c4 = cb * VGL_HC_COL_BLK;
        if (__any_sync(kFull, rare)) {
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                column(std::true_type{}, c4 + j, rec[j], f[j], bgv[j], cls[j]);
                if (j % kDrainEvery == 0) drain();
            }
        } else {
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                column(std::false_type{}, c4 + j, rec[j], f[j], bgv[j], cls[j]);
                if (j % kDrainEvery == 0) drain();
            }
        }
        // keep the products normal (four more factors cannot take them out of range).  The read-level product falls below
        // any threshold several times per read (the mismatch columns of Q2 cost ~1e-4 each): it is rescaled by 2^400 without
        // a branch and the count is undone in the log at the end of the group
        const bool low = pall < 0x1p-400;
        const double up = pall * 0x1p400;
        pall = low ? up : pall;
        n_scale += low ? 1 : 0;
        if (prod < 1e-100) {  // one mapping's own product: dozens of very unlikely bases in one mapping, rare
            a_ext += fast_log_s(prod, S.logtab);
            prod = 1.0;
            rare = true;
        }
    };
    // two blocks in flight from ONE load site: a block is waited for a whole block (hundreds of issue slots) after its
    // load, and no other outstanding load shares its scoreboard
    for (int cb = ncb - 1; cb >= 0; --cb) {
        const uint4 v = v0;
        v0 = v1;
        v1 = hc_fetch(H, cb - 2);
        block(cb, v);
    }
    // the next group's last blocks are requested now; they arrive while it is being set up
    v0 = hc_fetch(Hn, Hn.ncb - 1);
    v1 = hc_fetch(Hn, Hn.ncb - 2);
    a_sum += fma((double)n_scale, -400.0 * 0.6931471805599453094, fast_log_s(pall, S.logtab));
    while (x >= 0) {  // empty mappings that start past the last column
        const uint32_t xr = extra_rec(x);
        add_extra(xr & VGL_HC_MAP_MAX_NODE, (double)(xr >> VGL_HC_MAP_NODE_BITS) * S.le0);
        --x;
    }
}

template <int kAcMode>
__global__ void __launch_bounds__(kReadBlock, 1)
hc_read_kernel(HcBatchDev b, HcTablesDev t, double* __restrict__ cta_node_acc, double* __restrict__ cta_a_total,
               int* __restrict__ err_flag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double s_warp_a[kReadBlock / 32];
    const HcImageLayout L = t.layout;
    HcSmallTables* s_tab = reinterpret_cast<HcSmallTables*>(smem_raw + L.off_small);
    uint16_t* s_class = reinterpret_cast<uint16_t*>(smem_raw + L.off_class);
    double2* s_ac = reinterpret_cast<double2*>(smem_raw + L.off_ac);  // mode 0: [100][row]; mode 1: [n_class + 1]
    // [warps][kQueueSlots], every warp's ring on its own 1 KB boundary of the shared window (slot address = base | offset)
    constexpr uint32_t kRing = kQueueSlots * 16u;
    const uint32_t q_base = (smem_u32(smem_raw) + L.bytes + kRing - 1u) & ~(kRing - 1u);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    // the tables arrive as one TMA bulk copy of the prepared image (cp.async.bulk -> UBLKCP)
    {
        const uint32_t bar = smem_u32(&s_bar);
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, L.bytes);
            bulk_g2s(smem_u32(smem_raw), t.image, L.bytes, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
    }

    ReadKernelShared S;
    S.acc = cta_node_acc + (size_t)(blockIdx.x % kAccCopies) * (size_t)(t.n_words * 32);
    S.f = smem_u32(s_tab) + kOffFPair;
    S.bg = smem_u32(s_tab) + kOffBg;
    S.logtab = smem_u32(s_tab) + kOffLogTab;
    S.node_class = smem_u32(s_class);
    S.ac = smem_u32(s_ac);
    S.p_correct = smem_u32(s_tab) + kOffPCorrect;
    S.class_tab = t.class_tab;
    S.fasta = s_tab->is_fasta != 0;
    S.one_minus_bep = s_tab->one_minus_bep;
    S.queue = q_base + (uint32_t)warp * kRing;
    S.n_rows = t.n_rows;
    S.n_class = t.n_class;
    S.ac_stride = (int)L.ac_row_bytes;
    S.le0 = s_tab->log_err[0];

    double a_sum = 0.0;
    bool bad_node = false;
    uint32_t bad_cls = 0;  // largest class seen at an OPEN mark (CLASS_NONE: a node without pangenome_map entry)
    uint32_t q_head = 0, q_tail = 0;
    const uint32_t warps_total = gridDim.x * (kReadBlock / 32);
    uint32_t g = blockIdx.x * (kReadBlock / 32) + warp;
    GroupHdr H = hc_load_hdr(b, g, lane), Hn;
    uint4 v0 = hc_fetch(H, H.ncb - 1), v1 = hc_fetch(H, H.ncb - 2);
    while (g < b.n_groups) {
        const uint32_t gn = g + warps_total;
        hc_process_group<kAcMode>(b, S, g, gn, H, Hn, v0, v1, q_head, q_tail, lane, a_sum, bad_cls, bad_node);
        H = Hn;
        g = gn;
    }
    __syncwarp();
    if ((uint32_t)lane < q_tail - q_head) hc_finalise(S, lds_v4(S.queue | ((16u * (q_head + lane)) & (16u * (kQueueSlots - 1)))));

    if (__any_sync(kFull, bad_node || bad_cls == VGL_HC_CLASS_NONE) && lane == 0) atomicOr(err_flag, VGL_ERR_BAD_NODE);
    a_sum = warp_sum_f64(a_sum);
    if (lane == 0) s_warp_a[warp] = a_sum;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kReadBlock / 32; ++w) tot += s_warp_a[w];
        red_add_f64(cta_a_total + blockIdx.x % kAccCopies, tot);
    }
}

// ------------------------------------------------------------------------------------------
// K1, log-term mode.  Same batch, same walk (one read per thread, columns from the last to the first), but the
// per-column work is ONE shared-memory load and two fp64 adds: the term
//     log((1 - pcm) * bg[base] + pcm * K * (1 - eps(quality, mismatch)))                 src/process_mapping.cpp:59-77
// of every (mapping quality, node class, quality, mismatch, base) combination comes from tables staged in shared memory
// (rows of the mapping qualities present in the batch, qualities up to the batch's largest), indexed by bits 0-8 of the
// column record.  A mapping's A_i is then a plain sum, so a mapping is finalised in line at its CLOSE mark -- no product,
// no log, no rescaling, no queue: B_i from the integer quality sums, B_i - A_i added to the node vector with a
// fire-and-forget fp64 reduction.  Needs: no quality above 63 in the batch (so nothing is latched either) and tables
// that fit (hc_logterm_plan); everything else goes through hc_read_kernel.
//
// Quality sums.  The window of a mapping that starts at column c holds the stored qualities of columns c .. n - 1 plus c
// bytes past the end that count as quality 0.  With nz = the number of columns >= c whose canonical quality
// is not 0, the entries that cost log 0.25 number (n - c - nz) + c = n - nz, whatever c is:
//     B_i = (sum of q over columns >= c) * (-ln 10 / 10) + (n - nz) * log 0.25.
// Records past a read's own columns are zero and change neither sum.
//
// Mapping qualities.  The batch's groups come in segments, one mapping quality each (vgl_hc_batch_meta); the CTAs are dealt
// out to the segments in proportion to their groups and a CTA stages the tables of its segment's mapping quality only --
// (n_class + 1) x q_rows x 64 bytes, whatever the number of mapping qualities in the batch.
//
// Shared memory: [ node table u32[n_rows + 1] | pair tables double[n_class + 1][q_rows * 8] ]
// node table entry: byte offset of the class's table, bit 31 = every path goes through the node.
// ------------------------------------------------------------------------------------------
#ifndef VGL_HC_LT_BLOCK
#define VGL_HC_LT_BLOCK 768
#endif
constexpr int kLtBlock = VGL_HC_LT_BLOCK;

struct LtShared {
    uint32_t node_tab, tables;           // shared byte addresses
    double* acc;
    int n_rows;
    uint32_t none_entry;                 // node-table entry of the neutral class (unknown node, no pangenome_map entry)
    double le0, scale8;                  // log 0.25;  (-ln 10 / 10) / 8
};
// node table entry (u32): byte offset of the class's table | kLtAllPaths when every path goes through the node
constexpr uint32_t kLtAllPaths = 0x80000000u;

__device__ __forceinline__ uint32_t lds_u32_c(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// fire-and-forget fp64 add into base[index] where `on` is non-zero (one 64-bit multiply-add forms the address; that of the
// other lanes is never dereferenced)
__device__ __forceinline__ void red_add_f64_if(uint32_t on, uint64_t base, uint32_t index, double v) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\tsetp.ne.u32 p, %3, 0;\n\tmad.wide.u32 a, %1, 8, %0;\n\t@p red.global.add.f64 [a], %2;\n\t}" ::"l"(base),
                 "r"(index), "d"(v), "r"(on)
                 : "memory");
}

// Per-thread state of the walk through one read.
struct LtWalk {
    uint32_t sufq8;   // 8 * (sum of the canonical qualities of the columns walked so far)
    uint32_t nz;      // columns walked so far whose canonical quality is not 0
    double a_m;       // sum of the terms of the mapping being walked
    double a_all;     // ... of every valid column of the read: sum_i A_i
    uint32_t pbase;   // shared byte address of the table of the mapping being walked (its node's class)
    uint32_t allp;    // its node-table entry (bit 31: every path goes through the node)
};

// One block of four columns, no lane of the warp with empty mappings pending: straight-line code, nothing but predication.
__device__ __forceinline__ void lt_block_fast(const LtShared& S, uint64_t acc, const uint32_t (&rec)[4], const uint32_t (&ent)[4], uint32_t n_cols,
                                              LtWalk& w) {
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        const uint32_t r = rec[j];
        // OPEN (last column of a mapping): its class's table
        const bool open = (r & VGL_HC_COL_OPEN) != 0u;
        w.pbase = open ? S.tables + (ent[j] & ~kLtAllPaths) : w.pbase;
        w.allp = open ? ent[j] : w.allp;
        // the column's term; columns that take no part (N, non-ACGT, no owner) add nothing
        double t = 0.0;
        asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.f64 %0, [%1];\n\t}"
            : "+d"(t)
            : "r"(w.pbase + ((r & 0x1ffu) << 3)), "r"(r & VGL_HC_COL_VALID));
        w.a_m += t;
        w.a_all += t;
        const uint32_t q8 = r & (0xffu << VGL_HC_COL_QUAL_SHIFT);
        w.sufq8 += q8;
        w.nz += q8 ? 1u : 0u;
        // CLOSE (first column of a mapping): finalise it, unless every path goes through its node (its D[n] is never read);
        // a CLOSE on a column nobody owns carries the node field VGL_HC_NO_NODE, which is no row of the table
        static_assert((kLtAllPaths >> 17) == VGL_HC_COL_CLOSE, "the all-paths bit, shifted, lines up with CLOSE");
        const uint32_t node = VGL_HC_COL_NODE(r);
        const uint32_t close = r & VGL_HC_COL_CLOSE;
        const uint32_t push = ((r & ~(w.allp >> 17) & VGL_HC_COL_CLOSE) != 0u && node < (uint32_t)S.n_rows) ? 1u : 0u;
        const double B = fma((double)w.sufq8, S.scale8, (double)(n_cols - w.nz) * S.le0);
        red_add_f64_if(push, acc, node, B - w.a_m);
        w.a_m = close ? 0.0 : w.a_m;
    }
}

__device__ __forceinline__ void lt_process_group(const HcBatchDev& b, const LtShared& S, uint32_t g, uint32_t gn, const GroupHdr& H,
                                                 GroupHdr& Hn, uint4& v0, uint4& v1, int lane, double& a_sum, uint32_t& worst_entry,
                                                 bool& bad_node) {
    Hn = hc_load_hdr(b, gn, lane);
    const int ncb = H.ncb;
    int x = H.n_extra - 1;   // empty mappings (no column of their own): rare, handled on the side
    bool rare = x >= 0;
    LtWalk w;
    w.sufq8 = 0; w.nz = 0; w.a_m = 0.0; w.a_all = 0.0;
    w.pbase = S.tables + S.none_entry;
    w.allp = S.none_entry;

    auto extra_rec = [&](int i) -> uint32_t {
        const uint32_t* gx = b.map + ((size_t)__ldg(b.grp_map_blk + g) * 32 + lane) * VGL_HC_MAP_BLK;
        return __ldg(gx + (size_t)(i >> 2) * 128 + (i & 3));
    };
    auto add_extra = [&](uint32_t xn, double Bx) {
        if (xn < (uint32_t)S.n_rows && (lds_u32_c(S.node_tab + 4u * xn) & ~kLtAllPaths) != S.none_entry) red_add_f64(S.acc + xn, Bx);
        else bad_node = true;
    };
    // the same block with empty mappings pending on some lane: they receive B of the column they start at
    auto block_slow = [&](int cb, const uint32_t (&rec)[4], const uint32_t (&ent)[4]) {
#pragma unroll
        for (int j = 3; j >= 0; --j) {   // (unrolled: a dynamic index would put the records into local memory)
            const uint32_t r = rec[j];
            const bool open = (r & VGL_HC_COL_OPEN) != 0u;
            w.pbase = open ? S.tables + (ent[j] & ~kLtAllPaths) : w.pbase;
            w.allp = open ? ent[j] : w.allp;
            double t = 0.0;
            if (r & VGL_HC_COL_VALID) t = lds_f64_c(w.pbase + ((r & 0x1ffu) << 3));
            w.a_m += t;
            w.a_all += t;
            const uint32_t q8 = r & (0xffu << VGL_HC_COL_QUAL_SHIFT);
            w.sufq8 += q8;
            w.nz += q8 ? 1u : 0u;
            const uint32_t node = VGL_HC_COL_NODE(r);
            const uint32_t close = r & VGL_HC_COL_CLOSE;
            const double B = fma((double)w.sufq8, S.scale8, (double)((uint32_t)H.n_cols - w.nz) * S.le0);
            if (close && rare) {
                const int c = cb * VGL_HC_COL_BLK + j;
                while (x >= 0) {
                    const uint32_t xr = extra_rec(x);
                    const uint32_t xp = xr >> VGL_HC_MAP_NODE_BITS;
                    if (xp < (uint32_t)c) break;
                    add_extra(xr & VGL_HC_MAP_MAX_NODE, xp == (uint32_t)c ? B : (double)xp * S.le0);
                    --x;
                }
                rare = x >= 0;
            }
            if (close != 0u && (w.allp & kLtAllPaths) == 0u && node < (uint32_t)S.n_rows) red_add_f64(S.acc + node, B - w.a_m);
            w.a_m = close ? 0.0 : w.a_m;
        }
    };
    // The node-table lookups do not depend on the walk's state: the four of a block are issued together, each only where a
    // mapping opens (predicated: the other lanes cost no bank conflicts), and ONE BLOCK AHEAD of their use (the next block's
    // records are in registers already), so that their latency never shows.
    auto lookup = [&](const uint4& v, uint32_t (&ent)[4]) {
        const uint32_t rec[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t e32 = 0;
            asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
                : "+r"(e32)
                : "r"(S.node_tab + 4u * min(VGL_HC_COL_NODE(rec[j]), (uint32_t)S.n_rows)), "r"(rec[j] & VGL_HC_COL_OPEN));  // entry n_rows: neutral
            ent[j] = e32;
        }
    };
    const uint64_t acc = (uint64_t)S.acc;
    auto step = [&](int cb, const uint4& v, const uint32_t (&ent)[4]) {
        const uint32_t rec[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) worst_entry = max(worst_entry, ent[j] & ~kLtAllPaths);  // (0 where no mapping opens; the neutral class has the largest offset)
        if (__any_sync(kFull, rare)) block_slow(cb, rec, ent);
        else lt_block_fast(S, acc, rec, ent, (uint32_t)H.n_cols, w);
    };
    // two blocks per trip with the two register sets swapping roles (no copies): va / ea = the block being walked, vb / eb the
    // one after it, whose records are in flight or in registers and whose node-table entries are looked up meanwhile
    uint4 va = v0, vb = v1;
    uint32_t ea[4], eb[4];
    lookup(va, ea);
    int cb = ncb - 1;
    for (; cb >= 1; cb -= 2) {
        const uint4 vc = hc_fetch(H, cb - 2);
        lookup(vb, eb);
        step(cb, va, ea);
        va = hc_fetch(H, cb - 3);
        lookup(vc, ea);
        step(cb - 1, vb, eb);
        vb = va;   // (cb - 3)
        va = vc;   // (cb - 2)
    }
    if (cb == 0) step(0, va, ea);
    v0 = hc_fetch(Hn, Hn.ncb - 1);
    v1 = hc_fetch(Hn, Hn.ncb - 2);
    a_sum += w.a_all;
    while (x >= 0) {  // empty mappings that start past the last column
        const uint32_t xr = extra_rec(x);
        add_extra(xr & VGL_HC_MAP_MAX_NODE, (double)(xr >> VGL_HC_MAP_NODE_BITS) * S.le0);
        --x;
    }
}

__global__ void __launch_bounds__(kLtBlock, 1)
hc_read_logterm_kernel(HcBatchDev b, HcTablesDev t, HcLogTermLaunch plan, double* __restrict__ cta_node_acc, double* __restrict__ cta_a_total,
                       int* __restrict__ err_flag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double s_warp_a[kLtBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t node_bytes = (uint32_t)((((size_t)t.n_rows + 1) * sizeof(uint32_t) + 15) & ~(size_t)15);
    const uint32_t cls_bytes = (uint32_t)((((size_t)t.n_rows + 1) * sizeof(uint16_t) + 15) & ~(size_t)15);
    const uint32_t pair_bytes = (uint32_t)plan.q_rows * 64u;                       // q_rows * 8 entries of 8 bytes
    const uint32_t row_bytes = pair_bytes * (uint32_t)(t.n_class + 1);
    uint32_t* s_node = reinterpret_cast<uint32_t*>(smem_raw);
    unsigned char* s_tab = smem_raw + node_bytes;
    // the u16 class ids arrive in the upper half of the node table's space and are widened in place, front to back
    uint16_t* s_cls = reinterpret_cast<uint16_t*>(smem_raw + node_bytes - cls_bytes);
    // this CTA's segment: the groups of one mapping quality
    int seg = 0;
    while (seg + 1 < plan.n_segs && blockIdx.x >= plan.cta_begin[seg + 1]) ++seg;
    const uint32_t seg_mapq = plan.seg_mapq[seg];
    const uint32_t g_begin = plan.seg_begin[seg], g_end = plan.seg_begin[seg + 1];
    const uint32_t cta_in_seg = blockIdx.x - plan.cta_begin[seg], ctas_of_seg = plan.cta_begin[seg + 1] - plan.cta_begin[seg];

    // staging: the node classes (from the general kernel's image) and one table per (present mapping quality, class),
    // all as TMA bulk copies on one mbarrier
    {
        const uint32_t bar = smem_u32(&s_bar);
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, cls_bytes + row_bytes);
            bulk_g2s(smem_u32(s_cls), t.image + t.layout.off_class, cls_bytes, bar);
        }
        __syncthreads();
        for (int cls = tid; cls <= t.n_class; cls += kLtBlock) {
            const double* src = t.logterm + ((size_t)seg_mapq * (size_t)(t.n_class + 1) + (size_t)cls) * kLogTermEntries;
            bulk_g2s(smem_u32(s_tab) + (uint32_t)cls * pair_bytes, src, pair_bytes, bar);
        }
        mbar_wait(bar, 0);
        __syncthreads();
        // class id -> byte offset of the class's table | kLtAllPaths.  In place: entry n (4 bytes at 4n) overwrites class ids
        // that sit at or before 2n + (node_bytes - cls_bytes) ... so go in rounds of kLtBlock entries, reading a round before
        // writing it, from the front (a written entry never reaches past the ids still to be read: 4n + 4 <= base + 2n
        // for every n in the rounds that follow once base >= 2 (n + 2), which holds from the first round on because
        // base = node_bytes - cls_bytes >= 2 (n_rows + 1) - 16)
        for (int n0 = 0; n0 <= t.n_rows; n0 += kLtBlock) {
            const int n = n0 + tid;
            uint32_t e = 0;
            if (n <= t.n_rows) e = s_cls[n];
            __syncthreads();
            if (n <= t.n_rows) {
                const uint32_t cls = min(e & VGL_HC_CLASS_MASK, (uint32_t)t.n_class);
                s_node[n] = cls * pair_bytes | ((e & VGL_HC_CLASS_ALLPATHS) ? kLtAllPaths : 0u);
            }
            __syncthreads();
        }
    }

    LtShared S;
    S.node_tab = smem_u32(s_node);
    S.tables = smem_u32(s_tab);
    S.acc = cta_node_acc + (size_t)(blockIdx.x % kAccCopies) * (size_t)(t.n_words * 32);
    S.n_rows = t.n_rows;
    S.none_entry = (uint32_t)t.n_class * pair_bytes;
    S.le0 = -1.3862943611198906188;               // log(0.25) = log(get_p_seq_error(Q <= 2)), src/miscfunc.h:186
    S.scale8 = -0.23025850929940456840 / 8.0;     // log(get_p_seq_error(Q)) = Q * (-ln 10 / 10) for Q > 2

    double a_sum = 0.0;
    bool bad_node = false;
    uint32_t worst_entry = 0;
    const uint32_t warps_of_seg = ctas_of_seg * (kLtBlock / 32);
    uint32_t g = g_begin + cta_in_seg * (kLtBlock / 32) + warp;
    GroupHdr H = hc_load_hdr(b, g < g_end ? g : b.n_groups, lane), Hn;
    uint4 v0 = hc_fetch(H, H.ncb - 1), v1 = hc_fetch(H, H.ncb - 2);
    bool wrong_mapq = false;
    while (g < g_end) {
        const uint32_t gn = g + warps_of_seg;
        // (a lane with columns whose mapping quality is not the segment's: the producer's segments are wrong)
        wrong_mapq = wrong_mapq || (H.n_cols != 0 && H.mapq != seg_mapq);
        lt_process_group(b, S, g, gn < g_end ? gn : b.n_groups, H, Hn, v0, v1, lane, a_sum, worst_entry, bad_node);
        H = Hn;
        g = gn;
    }
    if (__any_sync(kFull, bad_node || worst_entry == S.none_entry) && lane == 0) atomicOr(err_flag, VGL_ERR_BAD_NODE);
    if (__any_sync(kFull, wrong_mapq) && lane == 0) atomicOr(err_flag, 0x08000000);   // decoded by vgl_hc_finalize
    a_sum = warp_sum_f64(a_sum);
    if (lane == 0) s_warp_a[warp] = a_sum;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kLtBlock / 32; ++w) tot += s_warp_a[w];
        red_add_f64(cta_a_total + blockIdx.x % kAccCopies, tot);
    }
}

// ------------------------------------------------------------------------------------------
// K2a: D[n] = sum of the kAccCopies node vectors the read kernel's CTAs added into, in copy order (the same order whichever
//      CTA added what), and the sum of the A totals -- once per sample, so that the 41 path tiles of K2b stage 2.5 KB each
//      instead of every copy of their node slice.  node_sum of a sample: [n_words * 32] doubles, then the A total.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hc_collapse_kernel(const double* __restrict__ node_acc, size_t acc_stride, int n_pad, double* __restrict__ node_sum, size_t sum_stride) {
    const double* acc = node_acc + (size_t)blockIdx.y * acc_stride;
    double* out = node_sum + (size_t)blockIdx.y * sum_stride;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n_pad) {
        double d = 0.0;
#pragma unroll 8
        for (int c = 0; c < kAccCopies; ++c) d += __ldcg(acc + (size_t)c * n_pad + n);
        out[n] = d;
    }
    if (n == 0) {
        const double* a_total = acc + (size_t)kAccCopies * n_pad;
        double a = 0.0;
        for (int c = 0; c < kAccCopies; ++c) a += __ldcg(a_total + c);
        out[n_pad] = a;
    }
}

// ------------------------------------------------------------------------------------------
// K2b: chunk_ll[chunk][p] = sum over the chunk's nodes n with pathsgo[n][p] == 0 of D[n]
// ------------------------------------------------------------------------------------------
struct PeerBufs {
    double* base[VGL_HC_MAX_PEERS];
};
// The multi-GPU exchange fused into K2 / K3: world <= 1 switches it off.  Exchange buffer of a rank (vgl_hc_peer_export):
// parts[2][world][n_pad] doubles, then flags[2][world] u64.  Generation = epoch & 1; a flag counts the path tiles its rank
// has delivered into that generation since the buffers were created, so the value to wait for is uses * tiles with
// uses = (epoch + 1) / 2.
struct PeerFuse {
    PeerBufs bufs;
    int rank, world, n_pad;
    unsigned long long epoch;
    unsigned long long timeout_ns;
};

constexpr int kPathTile = 128;       // paths per CTA (one per thread)
constexpr int kWordsPerChunk = 10;   // 320 nodes per CTA

// The masked sum over a chunk's nodes goes four nodes at a time: s_t[g][m] holds the sum of D[4g + b] over the bits b of m (added in the
// order b = 0..3), so a path costs one 8-byte lookup and one add per nibble of its membership word instead of two selects and an add per
// bit.  A row is 128 bytes: the lanes of a warp read different entries of the same row without bank conflicts.  The table is the same
// for every path, so identically supported paths still get bit-identical sums.
__device__ __forceinline__ void path_build_nibble_table(const double* __restrict__ s_d, double (*s_t)[16], int nw) {
    for (int e = threadIdx.x; e < nw * 8 * 16; e += kPathTile) {
        const int g = e >> 4, m = e & 15;
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v += ((m >> b) & 1) ? s_d[4 * g + b] : 0.0;
        s_t[g][m] = v;
    }
}
__device__ __forceinline__ double path_masked_sum(const double (*s_t)[16], const uint32_t (&words)[kWordsPerChunk], int nw) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k) {
        const uint32_t unsup = ~words[k];
        if (k < nw) {
#pragma unroll
            for (int nib = 0; nib < 8; ++nib) acc += s_t[k * 8 + nib][(unsup >> (4 * nib)) & 15u];
        }
    }
    return acc;
}

__global__ void __launch_bounds__(kPathTile)
hc_path_kernel(HcTablesDev t, const double* __restrict__ node_sum, double* __restrict__ chunk_ll,
               double* __restrict__ ll, unsigned int* __restrict__ tile_done, PeerFuse pf) {
    __shared__ __align__(16) double s_d[kWordsPerChunk * 32];
    __shared__ __align__(128) double s_t[kWordsPerChunk * 8][16];
    __shared__ __align__(8) unsigned long long s_bar;
    const int chunk = blockIdx.y;
    const int w0 = chunk * kWordsPerChunk;
    const int nw = min(kWordsPerChunk, t.n_words - w0);
    const int p = blockIdx.x * kPathTile + threadIdx.x;

    // Stage this chunk of the node vector with a TMA bulk copy (cp.async.bulk -> UBLKCP).
    const uint32_t bar = smem_u32(&s_bar);
    const uint32_t bytes = (uint32_t)nw * 32u * 8u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(&s_d[0]), node_sum + (size_t)w0 * 32, bytes, bar);
    }
    // Issue the membership loads while the bulk copy is in flight.
    uint32_t words[kWordsPerChunk];
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k)
        words[k] = (k < nw && p < t.n_paths) ? __ldg(t.node_bits + (size_t)(w0 + k) * t.n_paths + p) : 0xffffffffu;
    mbar_wait(bar, 0);
    path_build_nibble_table(s_d, s_t, nw);
    __syncthreads();
    const double acc = path_masked_sum(s_t, words, nw);
    if (p < t.n_paths) chunk_ll[(size_t)chunk * t.n_paths + p] = acc;
    // K2c, run by whichever CTA of this path tile finishes last: ll[p] = sum of the A totals + sum over chunks, both in a
    // fixed order that is the same for every p (who does the sum changes from run to run, the sum does not)
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(tile_done + blockIdx.x, 1u) == gridDim.y - 1;
        if (s_last) tile_done[blockIdx.x] = 0;  // re-armed for the next sample
    }
    __syncthreads();
    if (s_last && p < t.n_paths) {
        __threadfence();
        const double a = __ldcg(node_sum + (size_t)t.n_words * 32);   // sum of the A totals (K2a)
        double sum = 0.0;
        for (unsigned c = 0; c < gridDim.y; ++c) sum += __ldcg(chunk_ll + (size_t)c * t.n_paths + p);
        ll[p] = a + sum;
        // multi-GPU: this tile of the rank's vector goes straight into slot `rank` of EVERY rank's exchange buffer (peer
        // memory over NVLink) -- the exchange is the epilogue of this kernel, no separate push launch
        if (pf.world > 1) {
            const int gen = (int)(pf.epoch & 1ull);
            for (int q = 0; q < pf.world; ++q) pf.bufs.base[q][((size_t)gen * pf.world + pf.rank) * (size_t)pf.n_pad + p] = a + sum;
        }
    }
    if (pf.world > 1 && s_last) {
        __threadfence_system();
        __syncthreads();   // (s_last is uniform across the CTA)
        if ((int)threadIdx.x < pf.world) {   // one release per peer: "rank has delivered one more tile"
            const int gen = (int)(pf.epoch & 1ull);
            unsigned long long* flag = reinterpret_cast<unsigned long long*>(pf.bufs.base[threadIdx.x] + (size_t)2 * pf.world * pf.n_pad) +
                                       (size_t)gen * pf.world + pf.rank;
            asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(flag), "l"(1ull) : "memory");
        }
    }
}

// K2 for S samples at once (vgl_hc_finalize_samples): the CTA of (path tile, node chunk) fetches its membership words ONCE and
// walks the samples of its slice (blockIdx.z) with them -- the 7.7 MB bit matrix is read once per slice of samples, not once per
// sample.  Sample s keeps its node vectors at node_acc + s * acc_stride (layout of the single-sample buffer: [copies][n_words * 32]
// then the A totals), its chunk partials at chunk_ll + s * chunks * P, its vector at ll + s * P.  Same fixed summation orders as
// hc_path_kernel, so a sample's result does not depend on which call computed it.
__global__ void __launch_bounds__(kPathTile)
hc_path_samples_kernel(HcTablesDev t, const double* __restrict__ node_sum, size_t sum_stride, double* __restrict__ chunk_ll,
                       double* __restrict__ ll, unsigned int* __restrict__ tile_done, int n_samples, int samples_per_slice) {
    __shared__ __align__(16) double s_d[kWordsPerChunk * 32];
    __shared__ __align__(128) double s_t[kWordsPerChunk * 8][16];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ bool s_last;
    const int chunk = blockIdx.y;
    const int w0 = chunk * kWordsPerChunk;
    const int nw = min(kWordsPerChunk, t.n_words - w0);
    const int p = blockIdx.x * kPathTile + threadIdx.x;
    const uint32_t bar = smem_u32(&s_bar);
    const uint32_t bytes = (uint32_t)nw * 32u * 8u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t words[kWordsPerChunk];
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k)
        words[k] = (k < nw && p < t.n_paths) ? __ldg(t.node_bits + (size_t)(w0 + k) * t.n_paths + p) : 0xffffffffu;
    __syncthreads();
    const int s0 = blockIdx.z * samples_per_slice, s1 = min(n_samples, s0 + samples_per_slice);
    uint32_t phase = 0;
    for (int sm = s0; sm < s1; ++sm) {
        const double* sum_s = node_sum + (size_t)sm * sum_stride;
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(&s_d[0]), sum_s + (size_t)w0 * 32, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        path_build_nibble_table(s_d, s_t, nw);
        __syncthreads();
        const double acc = path_masked_sum(s_t, words, nw);
        double* chunk_s = chunk_ll + (size_t)sm * gridDim.y * t.n_paths;
        if (p < t.n_paths) chunk_s[(size_t)chunk * t.n_paths + p] = acc;
        __threadfence();
        __syncthreads();   // (also: everybody is done with s_d / s_t before the next sample's copy lands)
        unsigned int* done = tile_done + (size_t)sm * gridDim.x + blockIdx.x;
        if (threadIdx.x == 0) {
            s_last = atomicAdd(done, 1u) == gridDim.y - 1;
            if (s_last) *done = 0;
        }
        __syncthreads();
        if (s_last && p < t.n_paths) {
            __threadfence();
            const double a = __ldcg(sum_s + (size_t)t.n_words * 32);
            double sum = 0.0;
            for (unsigned c = 0; c < gridDim.y; ++c) sum += __ldcg(chunk_s + (size_t)c * t.n_paths + p);
            ll[(size_t)sm * t.n_paths + p] = a + sum;
        }
    }
}

__global__ void hc_sum_parts_kernel(const double* __restrict__ parts, int n_parts, int n_paths, double* __restrict__ ll) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    double s = 0.0;
    for (int k = 0; k < n_parts; ++k) s += parts[(size_t)k * n_paths + p];
    ll[p] = s;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU exchange over NVLink peer memory (no library collective).
// Buffer of every rank: parts[2][world][n_pad] doubles, then flags[2][world] u64; generation = epoch & 1.
//   push: CTA q copies this rank's ll[P] into slot `rank` of rank q's buffer (peer stores), then publishes `epoch`
//         in q's flag with a system-scope release;
//   sum : every CTA waits (system-scope acquire) until the `world` flags of the local buffer hold `epoch`, then adds
//         the slots in rank order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hc_peer_push_kernel(PeerBufs peers, int rank, int world, int n_paths, int n_pad, unsigned long long epoch, int n_tiles,
                    const double* __restrict__ ll) {
    const int q = blockIdx.x;
    const int gen = (int)(epoch & 1ull);
    double* slot = peers.base[q] + ((size_t)gen * world + rank) * (size_t)n_pad;
    for (int p = threadIdx.x; p < n_paths; p += blockDim.x) slot[p] = ll[p];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.base[q] + (size_t)2 * world * n_pad) + (size_t)gen * world + rank;
        __threadfence_system();
        // (the flag counts delivered path tiles, as in the fused exchange: a whole vector is n_tiles of them)
        asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(flag), "l"((unsigned long long)n_tiles) : "memory");
    }
}

__global__ void __launch_bounds__(256)
hc_peer_sum_kernel(const double* __restrict__ local, int world, int n_paths, int n_pad, unsigned long long epoch, int n_tiles,
                   double* __restrict__ ll, int* __restrict__ err_flag) {
    const int gen = (int)(epoch & 1ull);
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(local + (size_t)2 * world * n_pad) + (size_t)gen * world;
    int gone = 0;
    if ((int)threadIdx.x < world) {
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= ((epoch + 1ull) / 2ull) * (unsigned long long)n_tiles) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 120000000000ull) {  // 120 s: a peer is gone; report instead of hanging the GPU
                atomicOr(err_flag, 0x40000000);
                gone = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    if (__syncthreads_or(gone)) return;   // never add up slots that were not delivered
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const double* parts = local + (size_t)gen * world * (size_t)n_pad;
    double s = 0.0;
    for (int k = 0; k < world; ++k) s += __ldcg(parts + (size_t)k * n_pad + p);  // written by peers: bypass L1
    ll[p] = s;
}

// ------------------------------------------------------------------------------------------
// K3: argmax + log-softmax posterior + clade support, one CTA.  exp(ll[p] - max) is evaluated once
// per path into shared memory; every clade is then a masked sum over its descendant bitmap.
// ------------------------------------------------------------------------------------------
constexpr int kPostBlock = 1024;

__global__ void __launch_bounds__(kPostBlock)
hc_posterior_kernel(const double* ll, int n_paths, int row_words, const uint32_t* __restrict__ clade_bits,
                    const int32_t* __restrict__ anc_off, const int32_t* __restrict__ anc_clade, double* __restrict__ post,
                    int post_cap, double* __restrict__ out_scalars, int32_t* __restrict__ out_idx, const int* __restrict__ err_flag,
                    size_t ll_stride, size_t result_stride, size_t err_stride, double* ll_rw, PeerFuse pf, int n_tiles) {
    // one CTA per sample (blockIdx.x): sample s has its vector at ll + s * ll_stride, its result block at + s * result_stride
    // doubles and its error word at err_flag + s * err_stride ints (all strides 0 for the single-sample call)
    ll += (size_t)blockIdx.x * ll_stride;
    post += (size_t)blockIdx.x * result_stride;
    out_scalars += (size_t)blockIdx.x * result_stride;
    out_idx += (size_t)blockIdx.x * result_stride * 2;
    err_flag += (size_t)blockIdx.x * err_stride;
    extern __shared__ __align__(16) double s_e[];  // [n_paths] exp(ll - max)
    __shared__ double s_val[kPostBlock / 32];
    __shared__ int s_idx[kPostBlock / 32];
    __shared__ double s_max;
    __shared__ int s_arg;
    __shared__ double s_sum;
    __shared__ int s_peer_err;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // multi-GPU: the other half of the fused exchange.  Wait until every rank has delivered all its path tiles into the local
    // buffer's current generation, then ll = the ranks' vectors added IN RANK ORDER (bit-identical on every rank, and exact
    // ties between identically supported paths survive).
    if (pf.world > 1) {
        if (tid == 0) s_peer_err = 0;
        __syncthreads();
        const int gen = (int)(pf.epoch & 1ull);
        const double* local = pf.bufs.base[pf.rank];
        if (tid < pf.world) {
            const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(local + (size_t)2 * pf.world * pf.n_pad) + (size_t)gen * pf.world + tid;
            const unsigned long long want = ((pf.epoch + 1ull) / 2ull) * (unsigned long long)n_tiles;
            unsigned long long t0, t1, v;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
                if (v >= want) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > pf.timeout_ns) {  // a peer is gone: report instead of hanging the GPU, and leave ll alone
                    s_peer_err = 1;
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (s_peer_err) {
            if (tid == 0) out_idx[2] = 0x40000000;
            return;
        }
        const double* parts = local + (size_t)gen * pf.world * (size_t)pf.n_pad;
        for (int p = tid; p < n_paths; p += kPostBlock) {
            double sum = 0.0;
            for (int k = 0; k < pf.world; ++k) sum += __ldcg(parts + (size_t)k * pf.n_pad + p);  // written by peers: bypass L1
            ll_rw[p] = sum;
        }
        __threadfence_block();
        __syncthreads();
    }

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
        // lane l looks at bit l of every word (one uniform load per word, independent iterations): a clade near the
        // root holds thousands of paths, walking its set bits one by one would serialise the whole kernel
        double s = 0.0;
#pragma unroll 8
        for (int w = 0; w < row_words; ++w) {
            const uint32_t word = __ldg(bits + w);
            const int p = w * 32 + lane;
            const double e = p < n_paths ? s_e[p] : 0.0;
            s += ((word >> lane) & 1u) ? e : 0.0;
        }
        s = warp_sum_f64(s);
        if (lane == 0) post[1 + k] = s / tot;
    }
    if (tid == 0) {
        out_idx[1] = 1 + n_anc;
        out_idx[2] = *err_flag;  // travels with the results: one device-to-host copy per sample
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
static size_t read_smem_bytes(int n_rows, int n_class, int ac_mode) {
    size_t s = (size_t)hc_image_layout(n_rows, n_class, ac_mode).bytes + kQueueSlots * sizeof(uint4);  // + alignment slack of the rings
    s += (size_t)(kReadBlock / 32) * kQueueSlots * sizeof(uint4);
    return (s + 127) & ~(size_t)127;
}

HcReadLaunch hc_read_launch_config(int n_rows, int n_class, int sm_count, int max_smem_optin, int force_mode) {
    HcReadLaunch c;
    c.grid = sm_count;
    c.block = kReadBlock;
    // the fastest table layout that fits: per-(mapq, class) constants, else the class table, else the class table in global memory
    for (c.ac_mode = force_mode >= 0 ? force_mode : 0; c.ac_mode < 3; ++c.ac_mode) {
        c.smem = read_smem_bytes(n_rows, n_class, c.ac_mode);
        if (c.smem + 256 <= (size_t)max_smem_optin) return c;
        if (force_mode >= 0) break;
    }
    c.grid = 0;  // caller reports VGL_ERR_INVALID (node table too large for this layout)
    return c;
}

cudaError_t hc_read_configure(const HcReadLaunch& cfg) {
    const void* fn = cfg.ac_mode == 0 ? (const void*)hc_read_kernel<0> : cfg.ac_mode == 1 ? (const void*)hc_read_kernel<1> : (const void*)hc_read_kernel<2>;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
}

cudaError_t hc_launch_read_kernel(const HcReadLaunch& cfg, const HcBatchDev& batch, const HcTablesDev& tab,
                                  double* cta_node_acc, double* cta_a_total, int* err_flag,
                                  cudaStream_t s) {
    if (cfg.ac_mode == 0) hc_read_kernel<0><<<cfg.grid, cfg.block, cfg.smem, s>>>(batch, tab, cta_node_acc, cta_a_total, err_flag);
    else if (cfg.ac_mode == 1) hc_read_kernel<1><<<cfg.grid, cfg.block, cfg.smem, s>>>(batch, tab, cta_node_acc, cta_a_total, err_flag);
    else hc_read_kernel<2><<<cfg.grid, cfg.block, cfg.smem, s>>>(batch, tab, cta_node_acc, cta_a_total, err_flag);
    return cudaGetLastError();
}

size_t hc_logterm_plan(const HcBatchMeta& meta, uint32_t n_groups, int n_ctas, int n_rows, int n_class, bool have_tables, int max_smem_optin,
                       HcLogTermLaunch& plan) {
    if (!have_tables || !meta.valid || meta.qual_max > 63u || meta.n_segs == 0 || n_ctas < (int)meta.n_segs) return 0;
    if (meta.seg_end[meta.n_segs - 1] != n_groups) return 0;   // the segments must cover the batch
    plan.n_segs = (int)meta.n_segs;
    plan.q_rows = (int)meta.qual_max + 1;
    const size_t node_bytes = (((size_t)n_rows + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t pair_bytes = (size_t)plan.q_rows * 64;
    // + slack behind the last table: a record whose quality exceeds the meta's qual_max (a wrong meta, or a chunk whose packing
    // was abandoned) indexes up to 512 entries into a table of q_rows * 8 -- garbage in, but never past the shared-memory window
    const size_t slack = (size_t)(64 - plan.q_rows) * 64;
    const size_t smem = node_bytes + (size_t)(n_class + 1) * pair_bytes + slack;
    if (smem + 1024 > (size_t)max_smem_optin || smem >= (1u << 20)) return 0;  // (one mbarrier counts the staged bytes)
    // CTAs in proportion to the segments' groups (largest remainder), at least one each
    uint32_t begin = 0;
    plan.seg_begin[0] = 0;
    int given = 0;
    double frac[VGL_HC_MAX_SEGS];
    int share[VGL_HC_MAX_SEGS];
    for (int k = 0; k < plan.n_segs; ++k) {
        if (meta.seg_end[k] < begin || meta.seg_mapq[k] >= 100) return 0;
        plan.seg_mapq[k] = meta.seg_mapq[k];
        plan.seg_begin[k + 1] = meta.seg_end[k];
        const double want = n_groups ? (double)(meta.seg_end[k] - begin) * (double)(n_ctas - plan.n_segs) / (double)n_groups : 0.0;
        share[k] = 1 + (int)want;
        frac[k] = want - (double)(int)want;
        given += share[k];
        begin = meta.seg_end[k];
    }
    while (given < n_ctas) {
        int best = 0;
        for (int k = 1; k < plan.n_segs; ++k)
            if (frac[k] > frac[best]) best = k;
        share[best] += 1;
        frac[best] = -1.0;
        given += 1;
    }
    plan.cta_begin[0] = 0;
    for (int k = 0; k < plan.n_segs; ++k) plan.cta_begin[k + 1] = plan.cta_begin[k] + (uint32_t)share[k];
    return (smem + 127) & ~(size_t)127;
}

cudaError_t hc_logterm_configure(int max_smem_optin) {
    return cudaFuncSetAttribute(hc_read_logterm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin - 1024);
}

cudaError_t hc_launch_read_logterm_kernel(int grid, size_t smem, const HcLogTermLaunch& plan, const HcBatchDev& batch, const HcTablesDev& tab,
                                          double* cta_node_acc, double* cta_a_total, int* err_flag, cudaStream_t s) {
    hc_read_logterm_kernel<<<grid, kLtBlock, smem, s>>>(batch, tab, plan, cta_node_acc, cta_a_total, err_flag);
    return cudaGetLastError();
}

int hc_path_chunks(int n_words) { return (n_words + kWordsPerChunk - 1) / kWordsPerChunk; }

static PeerFuse no_peers() {
    PeerFuse pf;
    memset(&pf, 0, sizeof(pf));
    return pf;
}
static PeerFuse make_fuse(double* const* peer_bufs, int rank, int world, int n_pad, unsigned long long epoch, double timeout_s) {
    PeerFuse pf = no_peers();
    for (int k = 0; k < VGL_HC_MAX_PEERS; ++k) pf.bufs.base[k] = k < world ? peer_bufs[k] : nullptr;
    pf.rank = rank; pf.world = world; pf.n_pad = n_pad; pf.epoch = epoch;
    pf.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    return pf;
}

static cudaError_t launch_collapse(const HcTablesDev& tab, int n_samples, const double* node_acc, size_t acc_stride, double* node_sum,
                                   size_t sum_stride, cudaStream_t s) {
    const int n_pad = tab.n_words * 32;
    hc_collapse_kernel<<<dim3((n_pad + 255) / 256, n_samples), 256, 0, s>>>(node_acc, acc_stride, n_pad, node_sum, sum_stride);
    return cudaGetLastError();
}
size_t hc_node_sum_doubles(int n_words) { return (((size_t)n_words * 32 + 1) + 31) & ~(size_t)31; }   // per sample: D[n], then the A total

cudaError_t hc_launch_path_kernel(const HcTablesDev& tab, const double* node_acc, double* node_sum, double* chunk_ll, double* ll,
                                  unsigned int* tile_done, cudaStream_t s) {
    if (cudaError_t e = launch_collapse(tab, 1, node_acc, 0, node_sum, 0, s)) return e;
    dim3 grid((tab.n_paths + kPathTile - 1) / kPathTile, hc_path_chunks(tab.n_words));
    hc_path_kernel<<<grid, kPathTile, 0, s>>>(tab, node_sum, chunk_ll, ll, tile_done, no_peers());
    return cudaGetLastError();
}

cudaError_t hc_launch_path_kernel_fused(const HcTablesDev& tab, const double* node_acc, double* node_sum, double* chunk_ll, double* ll,
                                        unsigned int* tile_done, double* const* peer_bufs, int rank, int world, int n_pad,
                                        unsigned long long epoch, cudaStream_t s) {
    if (cudaError_t e = launch_collapse(tab, 1, node_acc, 0, node_sum, 0, s)) return e;
    dim3 grid((tab.n_paths + kPathTile - 1) / kPathTile, hc_path_chunks(tab.n_words));
    hc_path_kernel<<<grid, kPathTile, 0, s>>>(tab, node_sum, chunk_ll, ll, tile_done, make_fuse(peer_bufs, rank, world, n_pad, epoch, 0));
    return cudaGetLastError();
}
int hc_path_tiles(int n_paths) { return (n_paths + kPathTile - 1) / kPathTile; }

cudaError_t hc_launch_path_samples_kernel(const HcTablesDev& tab, int n_samples, const double* node_acc, size_t acc_stride, double* node_sum,
                                          double* chunk_ll, double* ll, unsigned int* tile_done, cudaStream_t s) {
    const size_t sum_stride = hc_node_sum_doubles(tab.n_words);
    if (cudaError_t e = launch_collapse(tab, n_samples, node_acc, acc_stride, node_sum, sum_stride, s)) return e;
    // slices of samples along z: enough CTAs to fill the GPU, few enough that the bit matrix is re-read only a handful of times
    const int slices = n_samples >= 32 ? 4 : n_samples >= 8 ? 2 : 1;
    const int per_slice = (n_samples + slices - 1) / slices;
    dim3 grid((tab.n_paths + kPathTile - 1) / kPathTile, hc_path_chunks(tab.n_words), (n_samples + per_slice - 1) / per_slice);
    hc_path_samples_kernel<<<grid, kPathTile, 0, s>>>(tab, node_sum, sum_stride, chunk_ll, ll, tile_done, n_samples, per_slice);
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
                                double* out_scalars, int32_t* out_idx, const int* err_flag, cudaStream_t s) {
    hc_posterior_kernel<<<1, kPostBlock, hc_posterior_smem(n_paths), s>>>(ll, n_paths, row_words, clade_bits, anc_off,
                                                                          anc_clade, post, post_cap, out_scalars, out_idx, err_flag, 0, 0, 0,
                                                                          nullptr, no_peers(), 0);
    return cudaGetLastError();
}

cudaError_t hc_launch_posterior_fused(double* ll, int n_paths, int row_words, const uint32_t* clade_bits, const int32_t* anc_off,
                                      const int32_t* anc_clade, double* post, int post_cap, double* out_scalars, int32_t* out_idx,
                                      const int* err_flag, double* const* peer_bufs, int rank, int world, int n_pad,
                                      unsigned long long epoch, double timeout_s, cudaStream_t s) {
    hc_posterior_kernel<<<1, kPostBlock, hc_posterior_smem(n_paths), s>>>(ll, n_paths, row_words, clade_bits, anc_off, anc_clade, post, post_cap,
                                                                          out_scalars, out_idx, err_flag, 0, 0, 0, ll,
                                                                          make_fuse(peer_bufs, rank, world, n_pad, epoch, timeout_s),
                                                                          hc_path_tiles(n_paths));
    return cudaGetLastError();
}

cudaError_t hc_launch_posterior_samples(int n_samples, const double* ll, int n_paths, int row_words, const uint32_t* clade_bits,
                                        const int32_t* anc_off, const int32_t* anc_clade, double* results, int post_cap, size_t result_stride,
                                        const int* err_flag, size_t err_stride, cudaStream_t s) {
    hc_posterior_kernel<<<n_samples, kPostBlock, hc_posterior_smem(n_paths), s>>>(
        ll, n_paths, row_words, clade_bits, anc_off, anc_clade, results + 4, post_cap, results, reinterpret_cast<int32_t*>(results + 2), err_flag,
        (size_t)n_paths, result_stride, err_stride, nullptr, no_peers(), 0);
    return cudaGetLastError();
}

cudaError_t hc_launch_peer_exchange(double* const* peer_bufs, int rank, int world, int n_paths, int n_pad,
                                    unsigned long long epoch, int n_tiles, double* ll, int* err_flag, cudaStream_t s) {
    PeerBufs pb;
    for (int k = 0; k < VGL_HC_MAX_PEERS; ++k) pb.base[k] = k < world ? peer_bufs[k] : nullptr;
    hc_peer_push_kernel<<<world, 256, 0, s>>>(pb, rank, world, n_paths, n_pad, epoch, n_tiles, ll);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    hc_peer_sum_kernel<<<(n_paths + 255) / 256, 256, 0, s>>>(peer_bufs[rank], world, n_paths, n_pad, epoch, n_tiles, ll, err_flag);
    return cudaGetLastError();
}

}  // namespace vgl
