// HaploCart read x path likelihood kernels for sm_100a (B200).
//
// The reference evaluates, for every read, every mapping i and every one of the P paths,
//     ll[p] += pathsgo[node_i][p] ? A_i : B_i                      (src/process_mapping.cpp:54-88)
// where A_i (sum of per-base log mixtures, :59-77) and B_i (sum of log sequencing-error
// probabilities, :4-24) do not depend on p.  Regrouping the same sum by node,
//     final_vec[p] = sum_i A_i + sum_n [pathsgo[n][p] == 0] * Dn,   Dn = sum_{i at node n} (B_i - A_i)
// is exact algebra (no approximation) and turns the work into
//   K1  hc_read_kernel   : stream the read batch once from HBM, one warp per read, compute every
//                          mapping's (A_i, B_i) and scatter B_i - A_i into a per-CTA node vector held
//                          in shared memory (fp64), sum A_i in registers;
//   K2  hc_path_kernel   : one pass over the bit-packed node x path membership (7.7 MB for the
//                          HaploCart graph) with the node vector staged by TMA bulk copy;
//   K3  hc_posterior_kernel : argmax, log-sum-exp, posterior of the call and its clades.
// "Supported" is used as the baseline so that the few best paths (the only ones that matter for
// the posterior) are sums of few terms and carry no cancellation.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "vgl_internal.h"

namespace vgl {

namespace {

constexpr int kReadBlock = 512;            // threads per CTA of the read kernel
constexpr int kWarpsPerCta = kReadBlock / 32;
constexpr unsigned kFull = 0xffffffffu;

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

// ------------------------------------------------------------------------------------------
// K1: one warp per read.
// ------------------------------------------------------------------------------------------
// Shared memory: [ double node_acc[n_rows] | HcSmallTables | uint16 node_class[n_rows] ]
__global__ void __launch_bounds__(kReadBlock, 1)
hc_read_kernel(HcBatchDev b, HcTablesDev t, double* __restrict__ cta_node_acc, double* __restrict__ cta_a_total,
               int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_acc = reinterpret_cast<double*>(smem_raw);
    HcSmallTables* s_tab = reinterpret_cast<HcSmallTables*>(s_acc + t.n_rows);
    uint16_t* s_class = reinterpret_cast<uint16_t*>(s_tab + 1);
    __shared__ double s_warp_a[kWarpsPerCta];

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

    const bool is_fasta = s_tab->is_fasta != 0;
    const double le0 = s_tab->log_err[0];
    double a_sum = 0.0;  // per-lane sum of A_i

    const uint32_t warps_total = gridDim.x * kWarpsPerCta;
    for (uint32_t r = blockIdx.x * kWarpsPerCta + warp; r < b.n_reads; r += warps_total) {
        const uint32_t c0 = b.col_off[r];
        const int Lp = (int)(b.col_off[r + 1] - c0);
        const uint32_t m0 = b.map_off[r];
        const int M = (int)(b.map_off[r + 1] - m0);
        const int Lq = (int)b.n_qual[r];
        const int mq = (int)b.mapq[r];
        const uint8_t* __restrict__ g = b.gcol + c0;
        const uint8_t* __restrict__ rb = b.rcol + c0;
        const uint8_t* __restrict__ q = b.qual + c0;

        // quality pass: total of log(err) over the stored columns, and the first real byte >= 90
        // (src/update_likelihood.cpp:40-44: that byte latches the background error mode).
        double qtot = 0.0;
        int f90 = 0x7fffffff;
        for (int c = lane; c < Lp; c += 32) {
            const uint8_t qb = q[c];
            qtot += s_tab->log_err[qb];
            if (c < Lq && (int)(int8_t)qb >= 90) f90 = min(f90, c);
        }
        qtot = warp_sum_f64(qtot);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) f90 = min(f90, __shfl_xor_sync(kFull, f90, d));

        const double pc = s_tab->p_correct[mq];
        int carry_len = 0;       // sum of sizes of earlier rounds
        double carry_seg = 0.0;  // sum of log(err) over the columns owned by earlier rounds

        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            const bool active = i < M;
            const int len = active ? (int)b.map_len[m0 + i] : 0;
            const int node = active ? b.map_node[m0 + i] : 0;
            const int incl = warp_incl_scan_i32(len, lane);
            // position_in_read advances by |read_seq.substr(pos, size)| (src/update_likelihood.cpp:37,45)
            const int pos = min(carry_len + incl - len, Lp);
            const int lg = min(len, Lp - pos);
            carry_len += __shfl_sync(kFull, incl, 31);

            double aa = 0.0, cc = 0.0;
            bool ok = false;
            if (active) {
                if (node >= 0 && node < t.n_rows) {
                    const uint16_t cls = s_class[node];
                    if (cls != 0xFFFFu) {
                        ok = true;
                        const double2 mk = t.class_tab[cls];  // {mappability, match_prob}
                        if (is_fasta) {
                            aa = 0.0;
                            cc = s_tab->one_minus_bep * mk.y;
                        } else {
                            const double pcm = pc * mk.x;  // p_correctly_mapped, src/process_mapping.cpp:41
                            aa = 1.0 - pcm;
                            cc = pcm * mk.y;
                        }
                    }
                }
                if (!ok) atomicOr(err_flag, VGL_ERR_BAD_NODE);
            }
            const bool latched = is_fasta || (f90 < pos + Lp);

            double A = 0.0, seg = 0.0;
            for (int j = 0; j < lg; ++j) {
                const uint8_t gb = g[pos + j];
                const uint8_t bb = rb[j];  // whole aligned read string, indexed from 0 (src/update_likelihood.cpp:46)
                const uint8_t qb = q[pos + j];
                seg += s_tab->log_err[qb];
                const int cg = s_tab->base_code[gb];
                const int cb = s_tab->base_code[bb];
                if (cg < 4 && cb < 4) {
                    const bool match = gb == bb;
                    double f;
                    if (latched) f = match ? s_tab->f_match_bep : s_tab->f_mis_bep;
                    else f = match ? s_tab->f_match[qb] : s_tab->f_mis[qb];
                    const double x = fma(cc, f, aa * s_tab->bg[cb]);
                    A += log(x);
                }
            }
            // log(err) summed over [pos, pos + Lp): stored columns from pos on, plus `pos` bytes past
            // the end that count as quality 0 (src/update_likelihood.cpp:40-44).
            const double seg_incl = warp_incl_scan_f64(seg, lane);
            const double before = carry_seg + seg_incl - seg;
            carry_seg += __shfl_sync(kFull, seg_incl, 31);
            if (ok) {
                const double B = (qtot - before) + (double)pos * le0;
                atomicAdd(&s_acc[node], B - A);
                a_sum += A;
            }
        }
    }

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
        for (int w = 0; w < kWarpsPerCta; ++w) tot += s_warp_a[w];
        cta_a_total[blockIdx.x] += tot;
    }
}

// ------------------------------------------------------------------------------------------
// K2a: node_acc[n] = sum over CTAs (fixed order, Neumaier-compensated)
// ------------------------------------------------------------------------------------------
__global__ void hc_reduce_ctas_kernel(const double* __restrict__ cta_node_acc, const double* __restrict__ cta_a_total,
                                      int n_ctas, int n_rows, int n_rows_pad, double* __restrict__ node_acc,
                                      double* __restrict__ a_total) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n_rows_pad) {
        double s = 0.0, comp = 0.0;
        if (n < n_rows) {
            for (int c = 0; c < n_ctas; ++c) {
                const double v = cta_node_acc[(size_t)c * n_rows + n];
                const double tsum = s + v;
                comp += (fabs(s) >= fabs(v)) ? ((s - tsum) + v) : ((v - tsum) + s);
                s = tsum;
            }
        }
        node_acc[n] = s + comp;
    }
    if (n == 0) {
        double s = 0.0, comp = 0.0;
        for (int c = 0; c < n_ctas; ++c) {
            const double v = cta_a_total[c];
            const double tsum = s + v;
            comp += (fabs(s) >= fabs(v)) ? ((s - tsum) + v) : ((v - tsum) + s);
            s = tsum;
        }
        a_total[0] = s + comp;
    }
}

// ------------------------------------------------------------------------------------------
// K2b: chunk_ll[chunk][p] = sum over the chunk's nodes n with pathsgo[n][p] == 0 of node_acc[n]
// ------------------------------------------------------------------------------------------
constexpr int kPathTile = 128;       // paths per CTA (one per thread)
constexpr int kWordsPerChunk = 10;   // 320 nodes per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_d)),
            "l"(node_acc + (size_t)w0 * 32), "r"(bytes), "r"(bar)
            : "memory");
    }
    // Issue the membership loads while the bulk copy is in flight.
    uint32_t words[kWordsPerChunk];
#pragma unroll
    for (int k = 0; k < kWordsPerChunk; ++k)
        words[k] = (k < nw && p < t.n_paths) ? __ldg(t.node_bits + (size_t)(w0 + k) * t.n_paths + p) : 0xffffffffu;
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(0u)
                : "memory");
        }
    }
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
// K3: argmax + log-softmax posterior + clade support, one CTA.
// ------------------------------------------------------------------------------------------
constexpr int kPostBlock = 1024;

__global__ void __launch_bounds__(kPostBlock)
hc_posterior_kernel(const double* __restrict__ ll, int n_paths, int row_words, const uint32_t* __restrict__ clade_bits,
                    const int32_t* __restrict__ anc_off, const int32_t* __restrict__ anc_clade, double* __restrict__ post,
                    int post_cap, double* __restrict__ out_scalars, int32_t* __restrict__ out_idx) {
    __shared__ double s_val[kPostBlock / 32];
    __shared__ int s_idx[kPostBlock / 32];
    __shared__ double s_max;
    __shared__ int s_arg;
    __shared__ double s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // argmax, first max wins (std::max_element, src/HaploCart.cpp:423)
    double best = -CUDART_INF;
    int arg = 0x7fffffff;
    for (int p = tid; p < n_paths; p += kPostBlock) {
        const double v = ll[p];
        if (v > best || (v == best && p < arg) || arg == 0x7fffffff) {
            if (arg == 0x7fffffff || v > best || (v == best && p < arg)) { best = v; arg = p; }
        }
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
    for (int p = tid; p < n_paths; p += kPostBlock) se += exp(ll[p] - mx);
    se = warp_sum_f64(se);
    __syncthreads();
    if (lane == 0) s_val[warp] = se;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kPostBlock / 32; ++w) tot += s_val[w];
        s_total = mx + log(tot);
        out_scalars[0] = s_total;
        out_scalars[1] = mx;
        out_idx[0] = call;
        if (post_cap > 0) post[0] = exp(mx - s_total);  // posterior of the call, src/get_posterior.cpp:103-106
    }
    __syncthreads();
    const double total = s_total;

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
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int p = w * 32 + bit;
                if (p < n_paths) s += exp(ll[p] - mx);
            }
        }
        s = warp_sum_f64(s);
        if (lane == 0) post[1 + k] = exp(mx + log(s) - total);
    }
    if (tid == 0) out_idx[1] = 1 + n_anc;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
static size_t read_smem_bytes(int n_rows) {
    size_t s = (size_t)n_rows * sizeof(double) + sizeof(HcSmallTables) + (size_t)n_rows * sizeof(uint16_t);
    return (s + 15) & ~(size_t)15;
}

HcReadLaunch hc_read_launch_config(int n_rows, int sm_count, int max_smem_optin) {
    HcReadLaunch c;
    c.grid = sm_count;
    c.block = kReadBlock;
    c.smem = read_smem_bytes(n_rows);
    if (c.smem > (size_t)max_smem_optin) c.grid = 0;  // caller reports VGL_ERR_INVALID (table too large for this layout)
    return c;
}

cudaError_t hc_read_configure(size_t smem) {
    return cudaFuncSetAttribute(hc_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t hc_launch_read_kernel(const HcReadLaunch& cfg, const HcBatchDev& batch, const HcTablesDev& tab,
                                  double* cta_node_acc, double* cta_a_total, int* err_flag, cudaStream_t s) {
    hc_read_kernel<<<cfg.grid, cfg.block, cfg.smem, s>>>(batch, tab, cta_node_acc, cta_a_total, err_flag);
    return cudaGetLastError();
}

cudaError_t hc_launch_reduce_ctas(const double* cta_node_acc, const double* cta_a_total, int n_ctas, int n_rows,
                                  int n_rows_pad, double* node_acc, double* a_total, cudaStream_t s) {
    const int block = 256;
    hc_reduce_ctas_kernel<<<(n_rows_pad + block - 1) / block, block, 0, s>>>(cta_node_acc, cta_a_total, n_ctas, n_rows,
                                                                            n_rows_pad, node_acc, a_total);
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
    const int block = 256;
    hc_path_sum_kernel<<<(n_paths + block - 1) / block, block, 0, s>>>(chunk_ll, n_chunks, n_paths, a_total, ll);
    return cudaGetLastError();
}

cudaError_t hc_launch_sum_parts(const double* parts, int n_parts, int n_paths, double* ll, cudaStream_t s) {
    const int block = 256;
    hc_sum_parts_kernel<<<(n_paths + block - 1) / block, block, 0, s>>>(parts, n_parts, n_paths, ll);
    return cudaGetLastError();
}

cudaError_t hc_launch_posterior(const double* ll, int n_paths, int row_words, const uint32_t* clade_bits,
                                const int32_t* anc_off, const int32_t* anc_clade, double* post, int post_cap,
                                double* out_scalars, int32_t* out_idx, cudaStream_t s) {
    hc_posterior_kernel<<<1, kPostBlock, 0, s>>>(ll, n_paths, row_words, clade_bits, anc_off, anc_clade, post, post_cap,
                                                 out_scalars, out_idx);
    return cudaGetLastError();
}

}  // namespace vgl
