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
constexpr int kConsumerWarps = 16;
constexpr int kReadBlock = (kConsumerWarps + 1) * 32;  // + one producer warp
constexpr int kTileReads = 32;                         // reads per staged tile
constexpr int kStages = 3;
constexpr int kTileColCap = 4096;                      // column records per staged tile
constexpr int kTileMapCap = 2048;                      // mapping records per staged tile
constexpr int kFastCols = 128;                         // fast path: the read is one 128-column chunk ...
constexpr int kFastMaps = 64;                          // ... and at most two rounds of 32 mappings

// byte layout of one stage (every sub-buffer 16-byte aligned; +slack for the 16-byte alignment of bulk copies)
constexpr int kStHdr = 0;                                       // StageHeader
constexpr int kStColOff = 64;                                   // (32 + 1) u32, up to 3 before / 3 after
constexpr int kStMapOff = kStColOff + 176;
constexpr int kStNq = kStMapOff + 176;                          // 32 u16
constexpr int kStMq = kStNq + 96;                               // 32 u8
constexpr int kStCol = kStMq + 64;                              // kTileColCap u16
constexpr int kStMap = kStCol + kTileColCap * 2 + 32;           // kTileMapCap u32
constexpr int kStageBytes = kStMap + kTileMapCap * 4 + 32;
static_assert(kStageBytes % 16 == 0, "stage size must keep 16-byte alignment");

struct StageHeader {
    uint32_t r0, n_reads, direct;  // direct != 0: column / mapping records were too many to stage; read them from global
    uint32_t sh_col_off, sh_map_off, sh_nq, sh_mq, sh_col, sh_map;  // element shift of each staged array
    uint32_t c0, m0;               // first column / mapping record of the tile
};

// per consumer warp scratch
struct WarpBuf {
    double xs[kFastCols];       // per-column mixture value x (1.0 where the column does not count)
    uint32_t pq[kFastCols];     // inclusive prefix over columns of (sum of q for q > 2) | (count of q <= 2) << 16
    double2 ac[kFastMaps];      // per mapping {1 - pcm, pcm * K}
    uint16_t pad16[kFastMaps];  // (spare)
    uint32_t mask[4];           // bit c set: a mapping with at least one column starts at column c
    uint32_t pad[4];
};
static_assert(sizeof(WarpBuf) % 16 == 0, "WarpBuf must keep 16-byte alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the producer lane waits for a stage to drain with a sleep between probes so it does not compete for issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
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
// 16-byte aligned window around [base + first, base + first + nbytes): returns its size, sets the source and the shift
__device__ __forceinline__ uint32_t aligned_window(const void* base, size_t first, size_t nbytes, const void*& src,
                                                   uint32_t& shift) {
    const uintptr_t a = (uintptr_t)base + first;
    const uintptr_t a0 = a & ~(uintptr_t)15;
    shift = (uint32_t)(a - a0);
    src = (const void*)a0;
    return (uint32_t)(((a + nbytes + 15) & ~(uintptr_t)15) - a0);
}

__device__ __forceinline__ int warp_incl_scan_i32(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ double warp_incl_scan_f64(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// ---- explicit shared-space accessors (32-bit shared addresses; keeps ptxas on LDS/STS/ATOMS instead of generic LD/ST) ----
// `_c` variants read tables that are constant for the kernel's lifetime (non-volatile: may be scheduled freely);
// the others touch staged tiles / warp scratch that are rewritten, so they stay ordered.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4u32(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16_c(uint32_t a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64_c(uint32_t a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64_c(uint32_t a) { double2 v; asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4u32(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ void atoms_or_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_add_f64(uint32_t a, double v) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// byte offsets inside HcSmallTables
constexpr uint32_t kOffFPair = offsetof(HcSmallTables, f_pair);
constexpr uint32_t kOffLogErr = offsetof(HcSmallTables, log_err);
constexpr uint32_t kOffLogTab = offsetof(HcSmallTables, log_tab);
constexpr uint32_t kOffPCorrect = offsetof(HcSmallTables, p_correct);
constexpr uint32_t kOffBg = offsetof(HcSmallTables, bg);

// log(x) for x > 0 (x == 0 gives -inf) with the table in shared memory at byte address `tab`.
__device__ __forceinline__ double fast_log_s(double x, uint32_t tab) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = lds_v2f64_c(tab + (uint32_t)((hi >> 12) & 0xff) * 16u);
    const double d = x - 1.0;
    const bool near = fabs(d) < 0.001953125;  // 2^-9: straight log1p keeps the relative accuracy near 1
    const double r = near ? d : fma(m, t.x, -1.0);
    const double base = near ? 0.0 : fma((double)e, 0.6931471805599453094, t.y);
    double p = fma(r, 0.2, -0.25);
    p = fma(r, p, 0.33333333333333333333);
    p = fma(r, p, -0.5);
    p = fma(r, p, 1.0);
    const double v = fma(r, p, base);
    return hi < 0x00100000 ? -CUDART_INF : v;  // zero (a denormal product cannot occur: factors are renormalised every 16)
}

// log(x) for x > 0 to ~1e-16 absolute + 1e-16 relative: 8-bit table on the mantissa, degree-5 log1p.
// Exactly 0 at x == 1.  Denormals / zero go through the library log.
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    if (hi < 0x00100000) return log(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = tab[(hi >> 12) & 0xff];
    double r = fma(m, t.x, -1.0);
    double base = fma((double)e, 0.6931471805599453094, t.y);
    const double d = x - 1.0;
    if (fabs(d) < 0.001953125) {  // 2^-9: straight log1p keeps the relative accuracy near 1
        r = d;
        base = 0.0;
    }
    double p = fma(r, 0.2, -0.25);
    p = fma(r, p, 0.33333333333333333333);
    p = fma(r, p, -0.5);
    p = fma(r, p, 1.0);
    return fma(r, p, base);
}

struct ReadSmem {
    double* acc;            // [n_rows]
    const HcSmallTables* tab;
    const double2* cls_cache;  // [kClassCache]
    const uint16_t* node_class;
};

// {1 - pcm, pcm * K} of a mapping at `node` (src/process_mapping.cpp:33-42); false if the node has no table entry
__device__ __forceinline__ bool mapping_scalars(const ReadSmem& S, const HcTablesDev& t, int node, double pc, double& aa,
                                                double& cc) {
    if (node >= t.n_rows) return false;
    const uint32_t cls = S.node_class[node];
    if (cls == 0xFFFFu) return false;
    const double2 mk = cls < (uint32_t)kClassCache ? S.cls_cache[cls] : t.class_tab[cls];  // {mappability, match_prob}
    if (S.tab->is_fasta) {
        aa = 0.0;
        cc = S.tab->one_minus_bep * mk.y;
    } else {
        const double pcm = pc * mk.x;
        aa = 1.0 - pcm;
        cc = pcm * mk.y;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// Generic per-read path: any length, any number of mappings, zero-length mappings anywhere, the
// background-error latch.  Lane per mapping with an inline column loop, records read from `col` /
// `map` wherever they live (shared or global).  Warp-collective.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void hc_generic_read(const uint16_t* __restrict__ col, const uint32_t* __restrict__ map, int Lp,
                                             int M, int Lq, int mq, const ReadSmem& S, const HcTablesDev& t,
                                             double& a_sum, int* err_flag, int lane) {
    const HcSmallTables* T = S.tab;
    const double le0 = T->log_err[0];
    double qtot = 0.0;
    int f90 = 0x7fffffff;  // first real quality byte >= 90: latches the background error mode (src/update_likelihood.cpp:42)
    for (int c = lane; c < Lp; c += 32) {
        const uint32_t qb = col[c] & 0xffu;
        qtot += T->log_err[qb];
        if (c < Lq && (int)(int8_t)qb >= 90) f90 = min(f90, c);
    }
    qtot = warp_sum_f64(qtot);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) f90 = min(f90, __shfl_xor_sync(kFull, f90, d));
    const double pc = T->p_correct[mq];
    int carry_len = 0;
    double carry_seg = 0.0;
    for (int base = 0; base < M; base += 32) {
        const int i = base + lane;
        const bool active = i < M;
        const uint32_t rec = active ? map[i] : 0u;
        const int node = (int)(rec & VGL_HC_MAP_MAX_NODE);
        const int len = (int)(rec >> VGL_HC_MAP_NODE_BITS);
        const int incl = warp_incl_scan_i32(len, lane);
        const int pos = min(carry_len + incl - len, Lp);  // position_in_read (src/update_likelihood.cpp:37,45)
        const int lg = min(len, Lp - pos);
        carry_len += __shfl_sync(kFull, incl, 31);
        double aa = 0.0, cc = 0.0;
        bool ok = false;
        if (active) {
            ok = mapping_scalars(S, t, node, pc, aa, cc);
            if (!ok) atomicOr(err_flag, VGL_ERR_BAD_NODE);
        }
        const bool latched = T->is_fasta || (f90 < pos + Lp);
        double A = 0.0, seg = 0.0, prod = 1.0;
        for (int j = 0; j < lg; ++j) {
            const uint32_t cr = col[pos + j];
            const uint32_t qb = cr & 0xffu;
            seg += T->log_err[qb];
            if (cr & VGL_HC_COL_VALID) {
                const bool match = (cr & VGL_HC_COL_MATCH) != 0;
                double f;
                if (latched) f = match ? T->f_match_bep : T->f_mis_bep;
                else f = T->f_pair[(match ? 0u : 256u) + qb];
                prod *= fma(cc, f, aa * T->bg[VGL_HC_COL_BASE(cr)]);
                if ((j & 15) == 15) {
                    A += fast_log(prod, T->log_tab);
                    prod = 1.0;
                }
            }
        }
        A += fast_log(prod, T->log_tab);
        // log(err) over [pos, pos + Lp): stored columns from pos on, plus `pos` bytes past the end that count
        // as quality 0 (src/update_likelihood.cpp:40-44)
        const double seg_incl = warp_incl_scan_f64(seg, lane);
        const double before = carry_seg + seg_incl - seg;
        carry_seg += __shfl_sync(kFull, seg_incl, 31);
        if (ok) {
            const double B = (qtot - before) + (double)pos * le0;
            atomicAdd(&S.acc[node], B - A);
            a_sum += A;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fast per-read path: Lp <= 128 columns, M <= 64 mappings, no quality byte >= 90, zero-length
// mappings only at the end, at most kClassCache table classes.  Returns false (before any side effect)
// when the read needs the generic path.  Warp-collective.  Everything it touches is addressed as shared
// memory: `col` / `map` are byte addresses of the read's records inside the staged tile.
// ------------------------------------------------------------------------------------------
struct FastCtx {
    uint32_t acc;        // node accumulators (f64)
    uint32_t tab;        // HcSmallTables
    uint32_t cls;        // class cache (double2)
    uint32_t node_class; // u16 per node
    uint32_t wb;         // this warp's WarpBuf
    int n_rows;
    bool fasta;
    double one_minus_bep;
};
constexpr uint32_t kWbXs = offsetof(WarpBuf, xs);
constexpr uint32_t kWbPq = offsetof(WarpBuf, pq);
constexpr uint32_t kWbAc = offsetof(WarpBuf, ac);
constexpr uint32_t kWbMask = offsetof(WarpBuf, mask);

__device__ __forceinline__ bool hc_fast_read(uint32_t col, uint32_t map, int Lp, int M, int Lq, int mq, const FastCtx& F,
                                             double& a_sum, bool& bad_node, int lane) {
    // (1) quality pre-pass, blocked layout: this lane owns columns 4*lane .. 4*lane + 3
    uint32_t qpack = 0;
    {
        const int cb = 4 * lane;
        const uint32_t a = col + 2u * (uint32_t)cb;
        const uint32_t r0 = cb + 0 < Lp ? lds_u16(a + 0) : 0u;
        const uint32_t r1 = cb + 1 < Lp ? lds_u16(a + 2) : 0u;
        const uint32_t r2 = cb + 2 < Lp ? lds_u16(a + 4) : 0u;
        const uint32_t r3 = cb + 3 < Lp ? lds_u16(a + 6) : 0u;
        qpack = (r0 & 0xffu) | ((r1 & 0xffu) << 8) | ((r2 & 0xffu) << 16) | (r3 << 24);
    }
    const int nin = max(0, min(4, Lp - 4 * lane));
    const int nq = max(0, min(4, Lq - 4 * lane));
    const uint32_t inb = nin >= 4 ? 0xffffffffu : ((1u << (8 * nin)) - 1u);
    const uint32_t qin = nq >= 4 ? 0xffffffffu : ((1u << (8 * nq)) - 1u);
    if (__any_sync(kFull, (__vcmpges4(qpack, 0x5a5a5a5au) & qin) != 0u)) return false;  // a real byte >= 90
    const uint32_t gt = __vcmpgts4(qpack, 0x02020202u) & inb;  // 0xff where (signed) quality > 2
    // per column: quality if > 2, else 1 << 16 (a stored column whose log(err) is log(0.25))
    const uint32_t low = ~gt & inb;
    const uint32_t qg = qpack & gt;
    const uint32_t t0 = (qg & 0xffu) + ((low & 1u) << 16);
    const uint32_t t1 = t0 + ((qg >> 8) & 0xffu) + ((low & 0x100u) << 8);
    const uint32_t t2 = t1 + ((qg >> 16) & 0xffu) + (low & 0x10000u);
    const uint32_t t3 = t2 + (qg >> 24) + ((low & 0x1000000u) >> 8);
    const uint32_t incl = (uint32_t)warp_incl_scan_i32((int)t3, lane);
    const uint32_t excl = incl - t3;
    const uint32_t qtot = __shfl_sync(kFull, incl, 31);
    sts_v4u32(F.wb + kWbPq + 16u * (uint32_t)lane, make_uint4(excl + t0, excl + t1, excl + t2, excl + t3));
    if (lane < 4) sts_u32(F.wb + kWbMask + 4u * (uint32_t)lane, 0u);
    __syncwarp();

    // (2) per-mapping scalars, two rounds of 32 lanes
    const double pc = lds_f64_c(F.tab + kOffPCorrect + 8u * (uint32_t)mq);
    int node_[2], pos_[2], lg_[2];
    bool ok_[2];
    int carry = 0;
    bool irregular = false;
    uint32_t nz0 = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
        const bool active = i < M;
        const uint32_t rec = active ? lds_u32(map + 4u * (uint32_t)i) : 0u;
        const int node = (int)(rec & VGL_HC_MAP_MAX_NODE);
        const int len = (int)(rec >> VGL_HC_MAP_NODE_BITS);
        const int inc = warp_incl_scan_i32(len, lane);
        const int pos = min(carry + inc - len, Lp);
        const int lg = min(len, Lp - pos);
        carry += __shfl_sync(kFull, inc, 31);
        const uint32_t nz = __ballot_sync(kFull, lg > 0);
        irregular |= (nz & (nz + 1u)) != 0u;  // columns must belong to a prefix of the mappings
        if (s == 0) nz0 = nz;
        else irregular |= (nz0 != kFull && nz != 0u);
        bool ok = !active;
        if (active) {
            const uint32_t cls = node < F.n_rows ? lds_u16_c(F.node_class + 2u * (uint32_t)node) : 0xFFFFu;
            ok = cls != 0xFFFFu;
            const double2 mk = lds_v2f64_c(F.cls + 16u * (ok ? cls : 0u));  // {mappability, match_prob}
            const double pcm = pc * mk.x;                                   // p_correctly_mapped, src/process_mapping.cpp:41
            const double aa = F.fasta ? 0.0 : 1.0 - pcm;
            const double cc = (F.fasta ? F.one_minus_bep : pcm) * mk.y;
            sts_v2f64(F.wb + kWbAc + 16u * (uint32_t)i, aa, cc);
            if (lg > 0) atoms_or_u32(F.wb + kWbMask + 4u * (uint32_t)(pos >> 5), 1u << (pos & 31));
        }
        node_[s] = node; pos_[s] = pos; lg_[s] = lg; ok_[s] = ok;
    }
    if (irregular) return false;
    bad_node |= !ok_[0] | !ok_[1];
    __syncwarp();

    // (3) columns, strided layout: lane handles columns lane, lane + 32, ...
    {
        const uint4 m4 = lds_v4u32(F.wb + kWbMask);
        const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
        int before = -1;
        const uint32_t lemask = 0xffffffffu >> (31 - lane);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            if (32 * it < Lp) {
                const int c = 32 * it + lane;
                const uint32_t rec = c < Lp ? lds_u16(col + 2u * (uint32_t)c) : 0u;
                const int rank = max(before + __popc(mw[it] & lemask), 0);
                const double2 ac = lds_v2f64(F.wb + kWbAc + 16u * (uint32_t)rank);
                const double f = lds_f64_c(F.tab + kOffFPair + 8u * (((rec & VGL_HC_COL_MATCH) ? 0u : 256u) + (rec & 0xffu)));
                const double bgv = lds_f64_c(F.tab + kOffBg + 8u * VGL_HC_COL_BASE(rec));
                const double x = fma(ac.y, f, ac.x * bgv);
                sts_f64(F.wb + kWbXs + 8u * (uint32_t)c, (rec & VGL_HC_COL_VALID) ? x : 1.0);
            }
            before += __popc(mw[it]);
        }
    }
    __syncwarp();

    // (4) per-mapping product -> A_i, quality prefix -> B_i, scatter
    const double le0 = lds_f64_c(F.tab + kOffLogErr);
    const double scale = -0.23025850929940456840;  // -ln(10)/10: log(get_p_seq_error(Q)) = Q * scale for Q > 2
    const bool any_long = __any_sync(kFull, lg_[0] > 8 || lg_[1] > 8);
    double a_read = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
        if (i < M) {
            const int pos = pos_[s], lg = lg_[s];
            const uint32_t xa = F.wb + kWbXs + 8u * (uint32_t)pos;
            double prod = 1.0, A = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < lg) prod *= lds_f64(xa + 8u * j);
            if (any_long) {
                for (int j = 8; j < lg; ++j) {
                    prod *= lds_f64(xa + 8u * (uint32_t)j);
                    if ((j & 15) == 15) {
                        A += fast_log_s(prod, F.tab + kOffLogTab);
                        prod = 1.0;
                    }
                }
            }
            A += fast_log_s(prod, F.tab + kOffLogTab);
            const uint32_t pre = pos > 0 ? lds_u32(F.wb + kWbPq + 4u * (uint32_t)(pos - 1)) : 0u;
            const int sq = (int)(qtot & 0xffffu) - (int)(pre & 0xffffu);
            const int nl = (int)(qtot >> 16) - (int)(pre >> 16) + pos;  // + `pos` bytes past the end (quality 0)
            const double B = fma((double)sq, scale, (double)nl * le0);
            if (ok_[s]) {
                atoms_add_f64(F.acc + 8u * (uint32_t)node_[s], B - A);
                a_read += A;
            }
        }
    }
    a_sum += a_read;
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
// Shared memory: [ node_acc f64[n_rows] | HcSmallTables | class cache | node_class u16[n_rows] | WarpBuf x 16 |
//                  stages x 3 | mbarriers ]
__global__ void __launch_bounds__(kReadBlock, 1)
hc_read_kernel(HcBatchDev b, HcTablesDev t, double* __restrict__ cta_node_acc, double* __restrict__ cta_a_total,
               int* __restrict__ err_flag, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    size_t off = 0;
    double* s_acc = reinterpret_cast<double*>(smem_raw);
    off += ((size_t)t.n_rows * sizeof(double) + 15) & ~(size_t)15;
    HcSmallTables* s_tab = reinterpret_cast<HcSmallTables*>(smem_raw + off);
    off += (sizeof(HcSmallTables) + 15) & ~(size_t)15;
    double2* s_cls = reinterpret_cast<double2*>(smem_raw + off);
    off += kClassCache * sizeof(double2);
    uint16_t* s_class = reinterpret_cast<uint16_t*>(smem_raw + off);
    off += ((size_t)t.n_rows * sizeof(uint16_t) + 15) & ~(size_t)15;
    WarpBuf* s_wb = reinterpret_cast<WarpBuf*>(smem_raw + off);
    off += kConsumerWarps * sizeof(WarpBuf);
    unsigned char* s_stage = smem_raw + off;
    off += (size_t)kStages * kStageBytes;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw + off);  // full[kStages], empty[kStages]
    __shared__ double s_warp_a[kConsumerWarps];
    __shared__ unsigned int s_generic;

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
        if (tid < kClassCache) s_cls[tid] = t.class_tab[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&s_bar[s]), 1);                         // full: the producer's expect_tx arrival
            mbar_init(smem_u32(&s_bar[kStages + s]), kConsumerWarps);  // empty: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_generic = 0;
    }
    __syncthreads();

    const uint32_t n_tiles = (b.n_reads + kTileReads - 1) / kTileReads;

    if (warp == kConsumerWarps) {
        // ===== producer warp: one elected lane stages tiles with TMA bulk copies =====
        if (lane == 0) {
            uint32_t k = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
                const uint32_t st = k % kStages;
                const uint32_t phase = (k / kStages) & 1u;
                unsigned char* sb = s_stage + (size_t)st * kStageBytes;
                const uint32_t full = smem_u32(&s_bar[st]);
                mbar_wait_backoff(smem_u32(&s_bar[kStages + st]), phase ^ 1u);
                const uint32_t r0 = tile * kTileReads;
                const uint32_t nr = min((uint32_t)kTileReads, b.n_reads - r0);
                const uint32_t c0 = __ldg(b.col_off + r0), c1 = __ldg(b.col_off + r0 + nr);
                const uint32_t m0 = __ldg(b.map_off + r0), m1 = __ldg(b.map_off + r0 + nr);
                StageHeader h;
                h.r0 = r0; h.n_reads = nr; h.c0 = c0; h.m0 = m0;
                h.direct = (c1 - c0 > (uint32_t)kTileColCap || m1 - m0 > (uint32_t)kTileMapCap) ? 1u : 0u;
                const void *p_co, *p_mo, *p_nq, *p_mq, *p_col = nullptr, *p_map = nullptr;
                uint32_t sh;
                const uint32_t n_co = aligned_window(b.col_off, (size_t)r0 * 4, (size_t)(nr + 1) * 4, p_co, sh);
                h.sh_col_off = sh / 4;
                const uint32_t n_mo = aligned_window(b.map_off, (size_t)r0 * 4, (size_t)(nr + 1) * 4, p_mo, sh);
                h.sh_map_off = sh / 4;
                const uint32_t n_nq = aligned_window(b.n_qual, (size_t)r0 * 2, (size_t)nr * 2, p_nq, sh);
                h.sh_nq = sh / 2;
                const uint32_t n_mq = aligned_window(b.mapq, (size_t)r0, (size_t)nr, p_mq, sh);
                h.sh_mq = sh;
                uint32_t n_col = 0, n_map = 0;
                h.sh_col = h.sh_map = 0;
                if (!h.direct) {
                    if (c1 > c0) { n_col = aligned_window(b.col, (size_t)c0 * 2, (size_t)(c1 - c0) * 2, p_col, sh); h.sh_col = sh / 2; }
                    if (m1 > m0) { n_map = aligned_window(b.map, (size_t)m0 * 4, (size_t)(m1 - m0) * 4, p_map, sh); h.sh_map = sh / 4; }
                }
                *reinterpret_cast<StageHeader*>(sb + kStHdr) = h;
                mbar_expect_tx(full, n_co + n_mo + n_nq + n_mq + n_col + n_map);
                bulk_g2s(smem_u32(sb + kStColOff), p_co, n_co, full);
                bulk_g2s(smem_u32(sb + kStMapOff), p_mo, n_mo, full);
                bulk_g2s(smem_u32(sb + kStNq), p_nq, n_nq, full);
                bulk_g2s(smem_u32(sb + kStMq), p_mq, n_mq, full);
                if (n_col) bulk_g2s(smem_u32(sb + kStCol), p_col, n_col, full);
                if (n_map) bulk_g2s(smem_u32(sb + kStMap), p_map, n_map, full);
            }
        }
        __syncwarp();
    } else {
        // ===== consumer warps: one read at a time =====
        ReadSmem S{s_acc, s_tab, s_cls, s_class};
        FastCtx F;
        F.acc = smem_u32(s_acc); F.tab = smem_u32(s_tab); F.cls = smem_u32(s_cls); F.node_class = smem_u32(s_class);
        F.wb = smem_u32(&s_wb[warp]);
        F.n_rows = t.n_rows; F.fasta = s_tab->is_fasta != 0; F.one_minus_bep = s_tab->one_minus_bep;
        const bool classes_cached = t.n_class <= kClassCache;
        const uint32_t stage_base = smem_u32(s_stage), bar_base = smem_u32(s_bar);
        bool bad_node = false;
        double a_sum = 0.0;
        unsigned int n_generic = 0;
        uint32_t k = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const uint32_t st = k % kStages;
            const uint32_t phase = (k / kStages) & 1u;
            const unsigned char* sb = s_stage + (size_t)st * kStageBytes;
            const uint32_t sa = stage_base + st * (uint32_t)kStageBytes;
            mbar_wait(bar_base + 8u * st, phase);
            // StageHeader fields (explicit shared loads)
            const uint32_t h_n_reads = lds_u32(sa + kStHdr + offsetof(StageHeader, n_reads));
            const uint32_t h_direct = lds_u32(sa + kStHdr + offsetof(StageHeader, direct));
            const uint32_t a_co = sa + kStColOff + 4u * lds_u32(sa + kStHdr + offsetof(StageHeader, sh_col_off));
            const uint32_t a_mo = sa + kStMapOff + 4u * lds_u32(sa + kStHdr + offsetof(StageHeader, sh_map_off));
            const uint32_t a_nq = sa + kStNq + 2u * lds_u32(sa + kStHdr + offsetof(StageHeader, sh_nq));
            const uint32_t a_mq = sa + kStMq + lds_u32(sa + kStHdr + offsetof(StageHeader, sh_mq));
            const uint32_t h_sh_col = lds_u32(sa + kStHdr + offsetof(StageHeader, sh_col));
            const uint32_t h_sh_map = lds_u32(sa + kStHdr + offsetof(StageHeader, sh_map));
            const uint32_t h_c0 = lds_u32(sa + kStHdr + offsetof(StageHeader, c0));
            const uint32_t h_m0 = lds_u32(sa + kStHdr + offsetof(StageHeader, m0));
            const uint32_t a_col = sa + kStCol + 2u * h_sh_col;  // record c of the tile is at a_col + 2 * (c - c0)
            const uint32_t a_map = sa + kStMap + 4u * h_sh_map;
            for (uint32_t ri = warp; ri < h_n_reads; ri += kConsumerWarps) {
                const uint32_t c_lo = lds_u32(a_co + 4u * ri), m_lo = lds_u32(a_mo + 4u * ri);
                const int Lp = (int)(lds_u32(a_co + 4u * ri + 4u) - c_lo);
                const int M = (int)(lds_u32(a_mo + 4u * ri + 4u) - m_lo);
                const int Lq = (int)lds_u16(a_nq + 2u * ri);
                const int mq = (int)lds_u8(a_mq + ri);
                bool done = false;
                if (!h_direct && classes_cached && Lp <= kFastCols && M <= kFastMaps && M > 0)
                    done = hc_fast_read(a_col + 2u * (c_lo - h_c0), a_map + 4u * (m_lo - h_m0), Lp, M, Lq, mq, F, a_sum, bad_node, lane);
                if (!done && M > 0) {
                    if (h_direct) {
                        hc_generic_read(b.col + c_lo, b.map + m_lo, Lp, M, Lq, mq, S, t, a_sum, err_flag, lane);
                    } else {
                        const uint16_t* t_col = reinterpret_cast<const uint16_t*>(sb + kStCol) + h_sh_col;
                        const uint32_t* t_map = reinterpret_cast<const uint32_t*>(sb + kStMap) + h_sh_map;
                        hc_generic_read(t_col + (c_lo - h_c0), t_map + (m_lo - h_m0), Lp, M, Lq, mq, S, t, a_sum, err_flag, lane);
                    }
                    n_generic += 1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8u * (kStages + st));
        }
        if (__any_sync(kFull, bad_node) && lane == 0) atomicOr(err_flag, VGL_ERR_BAD_NODE);
        a_sum = warp_sum_f64(a_sum);
        if (lane == 0) {
            s_warp_a[warp] = a_sum;
            if (n_generic) atomicAdd(&s_generic, n_generic);
        }
    }
    __syncthreads();
    double* my_acc = cta_node_acc + (size_t)blockIdx.x * (size_t)t.n_rows;
    for (int n = tid; n < t.n_rows; n += kReadBlock) {
        const double v = s_acc[n];
        if (v != 0.0) my_acc[n] += v;
    }
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kConsumerWarps; ++w) tot += s_warp_a[w];
        cta_a_total[blockIdx.x] += tot;
        if (s_generic) atomicAdd(stats, (unsigned long long)s_generic);
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
static size_t read_smem_bytes(int n_rows) {
    size_t s = 0;
    s += ((size_t)n_rows * sizeof(double) + 15) & ~(size_t)15;
    s += (sizeof(HcSmallTables) + 15) & ~(size_t)15;
    s += kClassCache * sizeof(double2);
    s += ((size_t)n_rows * sizeof(uint16_t) + 15) & ~(size_t)15;
    s += kConsumerWarps * sizeof(WarpBuf);
    s += (size_t)kStages * kStageBytes;
    s += 2 * kStages * sizeof(unsigned long long);
    return (s + 127) & ~(size_t)127;
}

HcReadLaunch hc_read_launch_config(int n_rows, int sm_count, int max_smem_optin) {
    HcReadLaunch c;
    c.grid = sm_count;
    c.block = kReadBlock;
    c.smem = read_smem_bytes(n_rows);
    if (c.smem + 256 > (size_t)max_smem_optin) c.grid = 0;  // caller reports VGL_ERR_INVALID (table too large for this layout)
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
