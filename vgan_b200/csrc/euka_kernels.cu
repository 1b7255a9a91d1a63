// euka per-read kernel for sm_100a: the body of the per-alignment lambda of readGAM3
// (reference src/readGAM_Euka.h:103-549) with one read per thread over the warp-striped batch.
//
// Per aligned column the reference takes log() of a 4-term marginal over the post-damage base and adds it to the
// in-clade log-likelihood (:312-394); here the marginal is multiplied into a running product and one log is taken
// per read (rescaled long before it can underflow); the constant-per-column cases (N, gap, rare IUPAC code, soft
// clip, model 2) are table look-ups.  euka scores each read against exactly one clade.
#include <cuda_runtime.h>

#include <algorithm>
#include <math_constants.h>
#include <stdint.h>

#include "vgl_internal.h"

namespace vgl {

namespace {

__device__ __forceinline__ uint4 ldg_stream_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr int kEukaBlock = 128;

__global__ void __launch_bounds__(kEukaBlock)
euka_read_kernel(EukaBatchDev b, EukaTablesDev t, EukaOutDev o) {
    __shared__ EukaSmallTables s_tab;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(t.small);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
        for (int k = threadIdx.x; k < (int)(sizeof(EukaSmallTables) / 4); k += kEukaBlock) dst[k] = src[k];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * (kEukaBlock / 32) + (threadIdx.x >> 5);
    if (g >= b.n_groups) return;
    const uint32_t r = g * 32 + lane;
    const int nc = (int)__ldg(b.n_cols + r);
    const int nm = (int)__ldg(b.n_maps + r);
    if (nm == 0) return;  // unused lane
    const uint32_t cb0 = __ldg(b.grp_col_blk + g), mb0 = __ldg(b.grp_map_blk + g);
    const uint4* gc = reinterpret_cast<const uint4*>(b.col) + (size_t)cb0 * 32 + lane;
    const uint4* gm = reinterpret_cast<const uint4*>(b.map) + (size_t)mb0 * 32 + lane;
    const int Lseq = (int)__ldg(b.seq_len + r);
    const int mq = (int)__ldg(b.mapq + r);
    const bool isrev = (__ldg(b.flags + r) & 1) != 0;
    const uint32_t ridx = __ldg(b.read_index + r);

    // clade of the read: the last bin (all clades, file order) that contains its first node (:115-140),
    // looked up in the node -> clade table the library built from the bins
    const uint32_t first_node = ldg_stream_v4(gm).x;
    const int c_n = first_node <= (uint32_t)t.max_node ? (int)__ldg(t.node_clade + first_node) : 0;
    const double d = __ldg(t.clade_dist + c_n);  // pair_dist
    const double one_minus_d = 1.0 - d;
    const double log_rare_in = log(one_minus_d * 0.001);

    double in_acc = 0.0, not_acc = 0.0, prod = 1.0;
    int n = isrev ? Lseq - 1 : 0;  // coordinate on the fragment, from its 5' end
    int softclips = 0;
    const int ncb = (nc + 7) >> 3;
    for (int cb = 0; cb < ncb; ++cb) {
        const uint4 blk = ldg_stream_v4(gc + (size_t)cb * 32);
        const uint32_t w[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int m = cb * 8 + k;
            if (m < nc) {
                const uint32_t rec = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
                const uint32_t q = rec & 0xffu;
                const int G = (int)((rec >> 8) & 15u), R = (int)(rec >> 12);
                if (G == VGL_EUKA_SYM_N || R == VGL_EUKA_SYM_N) {
                    const double v = s_tab.base_freq_log[R];
                    in_acc += v;
                    not_acc += v;
                } else if (G == VGL_EUKA_SYM_GAP || R == VGL_EUKA_SYM_GAP) {
                    in_acc += s_tab.log_gap_in;
                    not_acc += s_tab.log_gap_not;
                } else if (G == VGL_EUKA_SYM_RARE || R == VGL_EUKA_SYM_RARE) {
                    in_acc += log_rare_in;
                    not_acc += s_tab.log_rare_not;
                } else if (G == VGL_EUKA_SYM_SOFTCLIP || R == VGL_EUKA_SYM_SOFTCLIP) {
                    ++softclips;
                    in_acc += (softclips % 3 == 0) ? s_tab.log_one_minus_e[q] : s_tab.log_e_third[q];
                    not_acc += s_tab.log_quarter;
                } else {
                    // model 1: mutation since the clade's sequences (pair_dist, ts/tv), damage, sequencing error (:312-394)
                    const int i5 = min(n, t.k5 - 1), i3 = min(Lseq - 1 - n, t.k3 - 1);
                    const double* __restrict__ D = t.deam + ((size_t)i5 * t.k3 + i3) * 16;
                    double pre[4];
#pragma unroll
                    for (int bpo = 0; bpo < 4; ++bpo)
                        pre[bpo] = (bpo == G) ? one_minus_d : (G < 4 ? d * s_tab.tt[G][bpo] : 0.0);
                    const double e1 = s_tab.one_minus_e[q], e3 = s_tab.e_third[q];
                    double s = 0.0;
#pragma unroll
                    for (int bpd = 0; bpd < 4; ++bpd) {
                        double post = 0.0;
#pragma unroll
                        for (int bpo = 0; bpo < 4; ++bpo) post = fma(pre[bpo], __ldg(D + bpo * 4 + bpd), post);
                        s = fma(post, (bpd == R) ? e1 : e3, s);
                    }
                    prod *= s;
                    if (prod < 1e-280) {
                        in_acc += log(prod);
                        prod = 1.0;
                    }
                    not_acc += (G == R) ? s_tab.log_match2 : s_tab.log_mis2;  // model 2 (:420-443)
                }
                if (R != VGL_EUKA_SYM_GAP) n += isrev ? -1 : 1;
            }
        }
    }
    const double in_lik = in_acc + log(prod);
    const double not_lik = not_acc;
    // clade_like = map_q * exp(in - oplusInitnatl(in, not))  (:485-492; oplusInitnatl(0, y) = y)
    const double hi = fmax(in_lik, not_lik), lo = fmin(in_lik, not_lik);
    const double lse = (in_lik == 0.0) ? not_lik : hi + log1p(exp(lo - hi));
    const double map_q = 1.0 - pow(10.0, ((-1 * mq) * 0.1));
    const double like = map_q * exp(in_lik - lse);
    const bool pass = (in_lik - not_lik > 1.0) && ((uint32_t)mq > t.minimum_mq);
    if (ridx < o.n_out) {
        if (o.clade) o.clade[ridx] = c_n;
        if (o.in_lik) o.in_lik[ridx] = in_lik;
        if (o.not_lik) o.not_lik[ridx] = not_lik;
        if (o.like) o.like[ridx] = like;
        if (o.pass) o.pass[ridx] = pass ? 1 : 0;
    }
    // Clade::clade_like gets one entry per scored read (:491); the abundance MCMC only ever uses sum_k log(frac * clade_like[k])
    // per clade (src/MCMC.cpp:1190-1193, where `1/334` is an integer division that removes the clade_not_like term)
    atomicAdd(o.like_n + c_n, 1ull);
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(o.like_logsum + c_n), "d"(log(like)) : "memory");
    if (pass) {  // count the read and spread 1/M over the bins its nodes fall into (:504-549)
        atomicAdd(o.clade_count + c_n, 1ull);
        const double share = 1.0 / (double)nm;
        const int b0 = __ldg(t.bin_off + c_n), b1 = __ldg(t.bin_off + c_n + 1);
        for (int i = 0; i < nm; ++i) {
            const uint4 mb = ldg_stream_v4(gm + (size_t)(i >> 2) * 32);
            const int e = i & 3;
            const int node = (int)(e == 0 ? mb.x : e == 1 ? mb.y : e == 2 ? mb.z : mb.w);
            for (int j = b0; j < b1; ++j)
                if (node >= __ldg(t.bin_lo + j) && node <= __ldg(t.bin_hi + j)) atomicAdd(o.bin_cov + j, share);
        }
    }
}

}  // namespace

cudaError_t euka_launch_read_kernel(const EukaBatchDev& b, const EukaTablesDev& t, const EukaOutDev& o, cudaStream_t s) {
    if (b.n_groups == 0) return cudaSuccess;
    const int warps = kEukaBlock / 32;
    euka_read_kernel<<<(b.n_groups + warps - 1) / warps, kEukaBlock, 0, s>>>(b, t, o);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// soibean: per-read x per-path log-likelihoods (reference src/getLCAfromGAM.h:102-604)
// ------------------------------------------------------------------------------------------
namespace {

// K5a: one read per thread.  For edit i the reference evaluates, for every path, either the "supported" or the
// "unsupported" sum over the slice of aligned columns; both are path independent, so they are computed once here.
__global__ void __launch_bounds__(kEukaBlock)
soibean_edit_kernel(EukaBatchDev b, SoibeanTablesDev t, double* __restrict__ sup, double* __restrict__ uns, uint32_t* __restrict__ mcode,
                    int* __restrict__ err_flag) {
    __shared__ EukaSmallTables s_tab;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(t.small);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
        for (int k = threadIdx.x; k < (int)(sizeof(EukaSmallTables) / 4); k += kEukaBlock) dst[k] = src[k];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * (kEukaBlock / 32) + (threadIdx.x >> 5);
    if (g >= b.n_groups) return;
    const uint32_t r = g * 32 + lane;
    const int nc = (int)__ldg(b.n_cols + r);
    const int ne = (int)__ldg(b.n_maps + r);  // edits
    if (ne == 0) return;
    const uint32_t cb0 = __ldg(b.grp_col_blk + g), mb0 = __ldg(b.grp_map_blk + g);
    const uint16_t* colbase = b.col + ((size_t)cb0 * 32 + lane) * 8;
    const size_t mapbase = ((size_t)mb0 * 32 + lane) * 4;
    const bool rev = (__ldg(b.flags + r) & 1) != 0;
    auto col_at = [&](int c) -> uint32_t { return (uint32_t)__ldg(colbase + (size_t)(c >> 3) * 256 + (c & 7)); };
    const double log_quarter = s_tab.log_quarter, log_gap = log(0.02), log_cap = log(0.9999999);
    int baseIX = rev ? (int)__ldg(b.seq_len + r) - 1 : 0;
    for (int i = 0; i < ne; ++i) {
        const size_t mi = mapbase + (size_t)(i >> 2) * 128 + (i & 3);
        const uint32_t rec = __ldg(b.map + mi);
        const int size = (int)(rec >> 20);
        int start = baseIX;
        if (rev) start = (baseIX - size - 1 >= 0) ? (baseIX - size - 1) : 0;
        if (start < 0 || start > nc) {  // std::string::substr would throw inside the reference
            atomicOr(err_flag, 1);
            return;
        }
        const int len = min(size, nc - start);
        double s_sup = 0.0, s_uns = 0.0;
        int bor = baseIX;  // baseOnRead
        bool need_damage = false;
        // for the MCMC kernel: how many regular columns of the edit pair A/A, C/C, G/G, T/T (7 bits each); bit 31: some regular
        // column is not such a match (or a count overflowed) and the MCMC kernel has to walk the columns itself
        uint32_t pairs = 0;
        for (int s = 0; s < len; ++s) {
            const uint32_t cr = col_at(start + s);
            const uint32_t q = col_at(s) & 0xffu;  // a.quality()[s]: indexed by the position INSIDE the slice
            const int G = (int)((cr >> 8) & 15u), R = (int)(cr >> 12);
            if (G == VGL_EUKA_SYM_N || R == VGL_EUKA_SYM_N) { s_sup += log_quarter; s_uns += log_quarter; }
            else if (G == VGL_EUKA_SYM_SOFTCLIP || R == VGL_EUKA_SYM_SOFTCLIP) { s_sup += s_tab.log_e_third[q]; s_uns += s_tab.log_e_third[q]; }
            else if (G == VGL_EUKA_SYM_GAP || R == VGL_EUKA_SYM_GAP) { s_sup += log_gap; s_uns += log_gap; }
            else {
                need_damage = true;
                if (G == R && G < 4 && ((pairs >> (7 * G)) & 127u) != 127u) pairs += 1u << (7 * G);
                else pairs |= 0x80000000u;
                if (baseIX >= 0 && baseIX < nc) {
                    const int i5 = min(baseIX, t.k5 - 1), i3 = min(nc - baseIX - 1, t.k3 - 1);
                    const double* __restrict__ D = t.deam + ((size_t)i5 * t.k3 + i3) * 4;  // row sums of the damage matrix
                    const double e1 = s_tab.one_minus_e[q], e3 = s_tab.e_third[q];
                    double tot = 0.0;
#pragma unroll
                    for (int bpo = 0; bpo < 4; ++bpo) tot = fma((bpo == G) ? e1 : e3, __ldg(D + bpo), tot);
                    // min(log(tot), log(0.9999999)) (src/getLCAfromGAM.h:351-353): the log is only needed below the cap
                    s_sup += tot >= 0.9999999 ? log_cap : fmin(log(tot), log_cap);
                }
                s_uns += ((bor < 0 ? -bor : bor) % t.penalty == 0) ? s_tab.log_one_minus_e[q] : s_tab.log_e_third[q];
            }
            if (R != VGL_EUKA_SYM_GAP) bor += rev ? -1 : 1;
        }
        // the damage table is indexed [|graph_seq|][baseIX]: outside it the reference reads out of range
        if (need_damage && (rec & 0xFFFFFu) != VGL_SOIBEAN_NO_NODE && (baseIX < 0 || baseIX >= nc)) atomicOr(err_flag, 1);
        sup[mi] = s_sup;
        uns[mi] = s_uns;
        mcode[mi] = pairs;
        baseIX = rev ? start : baseIX + size;
    }
}

// K5b: one warp per read.  pathMap[p] = sum over edits of (path p goes through the edit's node ? sup : uns)
//                              = sum of sup  +  sum over edits of !member(edit, p) * (uns - sup)
// ("supported" is the baseline: a well supported path is a sum of small terms and carries no cancellation),
// so a lane keeps the running sums of "its" paths (p = 32 k + lane, bit `lane` of membership word k) in registers and
// every edit costs one uniform load per membership word instead of a pass over the edit list per 32 paths.
// kSoibeanWordsPerPass membership words (32 paths each) are covered per pass over the edit list.
template <int kSoibeanWordsPerPass>
__global__ void __launch_bounds__(kEukaBlock)
soibean_path_kernel(EukaBatchDev b, SoibeanTablesDev t, const double* __restrict__ sup, const double* __restrict__ uns, uint32_t n_out,
                    double* __restrict__ out, uint8_t* __restrict__ used, int* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (kEukaBlock / 32) + (threadIdx.x >> 5);  // read lane index
    if (r >= b.n_groups * 32u) return;
    const int ne = (int)__ldg(b.n_maps + r);
    if (ne == 0) return;
    const uint32_t ridx = __ldg(b.read_index + r);
    if (ridx >= n_out) return;
    const uint32_t g = r >> 5, l = r & 31;
    const size_t mapbase = ((size_t)__ldg(b.grp_map_blk + g) * 32 + l) * 4;
    const int n_words = (t.n_paths + 31) >> 5;
    for (int w0 = 0; w0 < n_words; w0 += kSoibeanWordsPerPass) {
        double acc[kSoibeanWordsPerPass];
#pragma unroll
        for (int k = 0; k < kSoibeanWordsPerPass; ++k) acc[k] = 0.0;
        double base = 0.0;
        for (int i = 0; i < ne; ++i) {
            const size_t mi = mapbase + (size_t)(i >> 2) * 128 + (i & 3);
            const uint32_t node = __ldg(b.map + mi) & 0xFFFFFu;
            const double u = __ldg(uns + mi);
            const int row = (int)node - t.min_id;
            if (node == VGL_SOIBEAN_NO_NODE || row < 0 || row >= t.n_rows) {
                // surplus entries get no support on any path (src/getLCAfromGAM.h:156-160); a node outside the table: nodepaths.at() would throw
                if (node != VGL_SOIBEAN_NO_NODE && lane == 0 && w0 == 0) atomicOr(err_flag, 2);
                base += u;
                continue;
            }
            const double sp = __ldg(sup + mi);
            base += sp;
            const double delta = u - sp;
            const uint32_t* bits = t.member_bits + (size_t)row * t.row_words + w0;
#pragma unroll
            for (int k = 0; k < kSoibeanWordsPerPass; ++k)
                if (w0 + k < n_words && !((__ldg(bits + k) >> lane) & 1u)) acc[k] += delta;
        }
#pragma unroll
        for (int k = 0; k < kSoibeanWordsPerPass; ++k) {
            const int p = (w0 + k) * 32 + lane;
            if (w0 + k < n_words && p < t.n_paths) out[(size_t)ridx * t.n_paths + p] = base + acc[k];
        }
    }
    if (lane == 0 && used) used[ridx] = 1;
}

// ------------------------------------------------------------------------------------------
// soibean MCMC: log-likelihood of proposals (reference: read loop of MCMC::run_tree_proportion, src/MCMC.cpp:738-993,
// computeBaseLogLike, src/MCMC.h:111-296, calculateLogWeightedAverage, src/MCMC.h:299-316).
//
// The reference keeps, per read, path, edit and base, a BaseInfo {read base, graph base, pathSupport, logLikelihood}
// (detailMap, src/getLCAfromGAM.h:192-530) and re-walks it for every proposal.  Everything in it is a function of the
// batch that is already resident in HBM: pathSupport is true exactly for the "regular" columns (neither base N, 'S' or
// '-') of an edit whose node the path goes through, there logLikelihood is the column's share of the edit's "supported"
// sum; every other entry's logLikelihood is its share of the "supported" (N / S / gap: the same on both sides) or
// "unsupported" sum.  computeBaseLogLike adds to a supported regular column a term that depends only on (read base,
// graph base, t, con): 25 values per (source, child | parent), tabulated in shared memory when the CTA starts.  So
//   ll(path) = sum over edits of  member(edit, path) ? sup[edit] + sum over its regular columns of H[read][graph] : uns[edit]
// with sup / uns as K5a left them.  One read per thread, one proposal per blockIdx.y; the per-read contributions are
// reduced in a fixed order (CTA partials, then soibean_mcmc_sum_kernel).
// ------------------------------------------------------------------------------------------

struct McmcFreqs { double f[7]; };  // params.freqs['A','C','G','T','R','Y','M']

__device__ __forceinline__ double mcmc_oplus(double x, double y) { return x > y ? x + log1p(exp(y - x)) : y + log1p(exp(x - y)); }
__device__ __forceinline__ double mcmc_oplus_init(double x, double y) { return x == 0.0 ? y : mcmc_oplus(x, y); }  // libgab oplusInitnatl

// The HKY part of computeBaseLogLike for read base class rb and graph base class gb (0-3 = ACGT, 4 = anything else).
__device__ double mcmc_hky_log_marg(int rb, int gb, const McmcFreqs& F, double t, double con) {
    const double kappa = 0.0;  // `double kappa = 1/22;` is an integer division (src/MCMC.h:66)
    const double mu = F.f[6];
    double lm = -CUDART_INF;
    for (int bpo = 0; bpo < 4; ++bpo) {
        const double frb = F.f[bpo];
        const bool purine = (bpo == 0 || bpo == 2);
        const double fam = purine ? F.f[4] : F.f[5];
        const double A = 1 + fam * (kappa - 1);
        const double jut1 = frb + frb * ((1 / fam) - 1) * exp(-(mu * t));
        double p;
        if (bpo == gb) {
            p = jut1 + ((fam - frb) / fam) * exp(-(mu * t * A));
        } else if (gb < 4 && (bpo ^ gb) == 2) {  // A <-> G, C <-> T
            const double jut11 = (frb / fam) * exp(-(mu * t * A));
            p = jut1 > jut11 ? jut1 - jut11 : jut11 - jut1;
        } else {
            p = frb * (1 - exp(-(mu * t)));
        }
        if (p < 1e-8) p = 1e-8;
        lm = mcmc_oplus_init(lm, log(p) + (bpo == rb ? log(1 - con) : log(con / 3)));
    }
    if (lm > 1e-8) lm = log(0.999999999);
    return lm;
}

// H[q][y][child at t2 | parent at t1][read class * 5 + graph class] for every proposal q and source y: 50 values per source
__global__ void __launch_bounds__(128)
soibean_mcmc_tables_kernel(const vgl_soibean_source* __restrict__ sources, int n_entries, McmcFreqs F, double con, double* __restrict__ H) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_entries) return;
    const vgl_soibean_source src = sources[k / 50];
    const int w = (k / 25) & 1, pair = k % 25;
    double tt = src.dist;
    if (tt == 0.0) tt = 0.00001;  // only at the root, which has no branch (src/MCMC.cpp:757-759)
    const double t1 = src.pos_branch * tt;
    H[k] = mcmc_hky_log_marg(pair / 5, pair % 5, F, w ? t1 : tt - t1, con);
}

// One read per thread, a tile of `tile` proposals per blockIdx.y: the read's records are fetched from HBM / L2 once per
// tile and come from L1 for the other proposals.  Dynamic shared memory: H tables [tile][n_sources][50] doubles, the tile's
// sources, [tile][warps] partial sums.  out[n_proposals] doubles followed by one double that is set non-zero on error
// is NOT written here: partials[q][blockIdx.x] go to soibean_mcmc_sum_kernel.
__global__ void __launch_bounds__(kEukaBlock)
soibean_mcmc_kernel(EukaBatchDev b, SoibeanTablesDev t, const double* __restrict__ sup, const double* __restrict__ uns,
                    const uint32_t* __restrict__ mcode, const vgl_soibean_source* __restrict__ sources, const double* __restrict__ H, int n_proposals, int n_sources,
                    int tile, uint32_t n_out, double* __restrict__ per_read, double* __restrict__ partials, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char mcmc_smem[];
    const int q0 = blockIdx.y * tile;
    const int nq = min(tile, n_proposals - q0);
    double* s_H = reinterpret_cast<double*>(mcmc_smem);                                         // [tile][n_sources][50]
    vgl_soibean_source* s_src = reinterpret_cast<vgl_soibean_source*>(s_H + (size_t)tile * n_sources * 50);  // [tile][n_sources]
    double* s_red = reinterpret_cast<double*>(s_src + (size_t)tile * n_sources);               // [tile][warps]
    for (int k = threadIdx.x; k < nq * n_sources * 50; k += kEukaBlock) s_H[k] = H[(size_t)q0 * n_sources * 50 + k];
    for (int k = threadIdx.x; k < nq * n_sources; k += kEukaBlock) s_src[k] = sources[(size_t)q0 * n_sources + k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x * (kEukaBlock / 32) + warp;
    bool bad = false;
    uint32_t r = 0, ridx = 0xffffffffu;
    int ne = 0, nc = 0, seq_len = 0;
    bool rev = false;
    const uint16_t* colbase = nullptr;
    size_t mapbase = 0;
    if (g < b.n_groups) {
        r = g * 32 + lane;
        ne = (int)__ldg(b.n_maps + r);
        if (ne > 0) {
            nc = (int)__ldg(b.n_cols + r);
            colbase = b.col + ((size_t)__ldg(b.grp_col_blk + g) * 32 + lane) * 8;
            mapbase = ((size_t)__ldg(b.grp_map_blk + g) * 32 + lane) * 4;
            rev = (__ldg(b.flags + r) & 1) != 0;
            seq_len = (int)__ldg(b.seq_len + r);
            if (per_read) ridx = __ldg(b.read_index + r);
        }
    }
    auto col_at = [&](int c) -> uint32_t { return (uint32_t)__ldg(colbase + (size_t)(c >> 3) * 256 + (c & 7)); };
    for (int q = 0; q < nq; ++q) {
        double contribution = 0.0;
        if (ne > 0) {
            double inter = -CUDART_INF;
            for (int y = 0; y < n_sources; ++y) {
                const vgl_soibean_source& S = s_src[q * n_sources + y];
                const int pc = S.child, pp = S.parent;
                const double* __restrict__ Hc = s_H + ((size_t)q * n_sources + y) * 50;
                const double* __restrict__ Hp = Hc + 25;
                double llc = 0.0, llp = 0.0;
                int baseIX = rev ? seq_len - 1 : 0;
                for (int i = 0; i < ne; ++i) {
                    const size_t mi = mapbase + (size_t)(i >> 2) * 128 + (i & 3);
                    const uint32_t rec = __ldg(b.map + mi);
                    const int size = (int)(rec >> 20);
                    int start = baseIX;
                    if (rev) start = (baseIX - size - 1 >= 0) ? (baseIX - size - 1) : 0;
                    if (start < 0 || start > nc) { bad = true; break; }  // K5a has reported it already
                    const int len = min(size, nc - start);
                    const uint32_t node = rec & 0xFFFFFu;
                    const int row = (int)node - t.min_id;
                    bool in_c = false, in_p = false;
                    if (node != VGL_SOIBEAN_NO_NODE && row >= 0 && row < t.n_rows) {
                        const uint32_t* bits = t.member_bits + (size_t)row * t.row_words;
                        in_c = (__ldg(bits + (pc >> 5)) >> (pc & 31)) & 1u;
                        in_p = (__ldg(bits + (pp >> 5)) >> (pp & 31)) & 1u;
                    }
                    const double sp = __ldg(sup + mi), u = __ldg(uns + mi);
                    double hc = 0.0, hp = 0.0;
                    const uint32_t pairs = (in_c || in_p) ? __ldg(mcode + mi) : 0u;
                    if (!(pairs >> 31)) {  // every regular column pairs a base with itself: four counts, no column walk
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const double n = (double)((pairs >> (7 * k)) & 127u);
                            hc = fma(n, Hc[6 * k], hc);
                            hp = fma(n, Hp[6 * k], hp);
                        }
                    } else {
                        for (int s = 0; s < len; ++s) {
                            const uint32_t cr = col_at(start + s);
                            const int G = (int)((cr >> 8) & 15u), R = (int)(cr >> 12);
                            const bool special = G == VGL_EUKA_SYM_N || R == VGL_EUKA_SYM_N || G == VGL_EUKA_SYM_SOFTCLIP ||
                                                 R == VGL_EUKA_SYM_SOFTCLIP || G == VGL_EUKA_SYM_GAP || R == VGL_EUKA_SYM_GAP;
                            if (!special) {
                                const int pair = min(R, 4) * 5 + min(G, 4);
                                hc += Hc[pair];
                                hp += Hp[pair];
                            }
                        }
                    }
                    llc += in_c ? sp + hc : u;
                    llp += in_p ? sp + hp : u;
                    baseIX = rev ? start : baseIX + size;
                }
                // the reference throws when a running sum is positive, nan or infinite (src/MCMC.cpp:841-856,936-953)
                if (llc > 0.0 || isnan(llc) || isinf(llc) || llp > 0.0 || isnan(llp) || isinf(llp)) bad = true;
                const double pb = S.pos_branch;
                if (n_sources == 1) {  // calculateLogWeightedAverage(readLogLike, pos_branch, readLogLikeP, 1 - pos_branch)
                    const double a = llc + log(pb), c = llp + log(1 - pb);
                    const double mx = fmax(a, c);
                    const double lse = mx + log(exp(a - mx) + exp(c - mx));
                    const double lws = log(pb + (1 - pb));
                    contribution = isinf(lws) ? -CUDART_INF : lse - lws;
                } else {
                    const double inter2 = mcmc_oplus(log(pb) + llc, log(1 - pb) + llp);
                    inter = mcmc_oplus_init(inter, inter2 + log(S.proportion));
                }
            }
            if (n_sources != 1) contribution = inter;
            if (ridx < n_out) per_read[(size_t)(q0 + q) * n_out + ridx] = contribution;
        }
        // fixed-order reduction: lanes, then (below) warps
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contribution += __shfl_xor_sync(0xffffffffu, contribution, d);
        if (lane == 0) s_red[q * (kEukaBlock / 32) + warp] = contribution;
    }
    if (bad) atomicOr(err_flag, 4);
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += kEukaBlock) {
        double tot = 0.0;
        for (int w = 0; w < kEukaBlock / 32; ++w) tot += s_red[q * (kEukaBlock / 32) + w];
        partials[(size_t)(q0 + q) * gridDim.x + blockIdx.x] = tot;
    }
}

// one warp per proposal: the CTA partials in a fixed order.  out[n_proposals] is followed by one double that carries the
// error flag, so that results and status come back in one copy.
__global__ void __launch_bounds__(32) soibean_mcmc_sum_kernel(const double* __restrict__ partials, int n_blocks, int n_proposals,
                                                              double* __restrict__ out, const int* __restrict__ err_flag) {
    const double* p = partials + (size_t)blockIdx.x * n_blocks;
    double s = 0.0;
    for (int k = threadIdx.x; k < n_blocks; k += 32) s += p[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) {
        out[blockIdx.x] = s;
        if (blockIdx.x == 0) out[n_proposals] = (double)*err_flag;  // set by the read kernels, which have all finished
    }
}

}  // namespace

int soibean_mcmc_blocks(const EukaBatchDev& b) { return (int)((b.n_groups + kEukaBlock / 32 - 1) / (kEukaBlock / 32)); }

cudaError_t soibean_mcmc_launch(const EukaBatchDev& b, const SoibeanTablesDev& t, const double* sup, const double* uns, const uint32_t* mcode,
                                const vgl_soibean_source* sources, int n_proposals, int n_sources, const double* freqs7, double con,
                                uint32_t n_out, double* per_read, double* H, double* partials, double* loglike, int* err_flag,
                                cudaStream_t s) {
    const int nb = soibean_mcmc_blocks(b);
    if (nb == 0 || n_proposals == 0) return cudaSuccess;
    McmcFreqs F;
    for (int k = 0; k < 7; ++k) F.f[k] = freqs7[k];
    const int n_entries = n_proposals * n_sources * 50;
    soibean_mcmc_tables_kernel<<<(n_entries + 127) / 128, 128, 0, s>>>(sources, n_entries, F, con, H);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // proposals per CTA: as many as fit 40 KB of tables (16 at most); with few reads, fewer so that the grid still fills the GPU
    const size_t per_q = (size_t)n_sources * (50 * sizeof(double) + sizeof(vgl_soibean_source)) + (kEukaBlock / 32) * sizeof(double);
    int tile = (int)std::min<size_t>(16, std::max<size_t>(1, (40 * 1024) / per_q));
    tile = std::min(tile, n_proposals);
    while (tile > 1 && (long long)nb * ((n_proposals + tile - 1) / tile) < 4 * 148) tile = (tile + 1) / 2;
    const int tiles = (n_proposals + tile - 1) / tile;
    for (int t0 = 0; t0 < tiles; t0 += 32768) {  // grid.y is limited to 65535
        const int nt = std::min(32768, tiles - t0);
        const int q0 = t0 * tile, nq = std::min(n_proposals - q0, nt * tile);
        soibean_mcmc_kernel<<<dim3((unsigned)nb, (unsigned)nt), kEukaBlock, per_q * tile, s>>>(
            b, t, sup, uns, mcode, sources + (size_t)q0 * n_sources, H + (size_t)q0 * n_sources * 50, nq, n_sources, tile, n_out,
            per_read ? per_read + (size_t)q0 * n_out : nullptr, partials + (size_t)q0 * nb, err_flag);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    soibean_mcmc_sum_kernel<<<(unsigned)n_proposals, 32, 0, s>>>(partials, nb, n_proposals, loglike, err_flag);
    return cudaGetLastError();
}

cudaError_t soibean_launch(const EukaBatchDev& b, const SoibeanTablesDev& t, double* sup, double* uns, uint32_t* mcode, uint32_t n_out,
                           double* out, uint8_t* used, int* err_flag, cudaStream_t s) {
    if (b.n_groups == 0) return cudaSuccess;
    const int warps = kEukaBlock / 32;
    soibean_edit_kernel<<<(b.n_groups + warps - 1) / warps, kEukaBlock, 0, s>>>(b, t, sup, uns, mcode, err_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t lanes = b.n_groups * 32u;
    const int n_words = (t.n_paths + 31) / 32;
    const unsigned grid = (lanes + warps - 1) / warps;
    if (n_words <= 1) soibean_path_kernel<1><<<grid, kEukaBlock, 0, s>>>(b, t, sup, uns, n_out, out, used, err_flag);
    else if (n_words <= 4) soibean_path_kernel<4><<<grid, kEukaBlock, 0, s>>>(b, t, sup, uns, n_out, out, used, err_flag);
    else if (n_words <= 8) soibean_path_kernel<8><<<grid, kEukaBlock, 0, s>>>(b, t, sup, uns, n_out, out, used, err_flag);
    else soibean_path_kernel<16><<<grid, kEukaBlock, 0, s>>>(b, t, sup, uns, n_out, out, used, err_flag);
    return cudaGetLastError();
}

}  // namespace vgl
