// C-ABI layer of libvgan_gpu.so (include/vgan_gpu.h).  Owns device memory, streams and the table
// preprocessing; the kernels live in hc_kernels.cu.  No CPU fallback exists anywhere in this
// library: without a CUDA device vgl_create() fails with VGL_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <utility>
#include <vector>

#include "vgl_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define VGL_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(VGL_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n) {
        if (n <= cap) return VGL_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e != cudaSuccess) return fail(VGL_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
        cap = want;
        return VGL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct vgl_ctx {
    int device = 0;
    int sm_count = 0;
    // duplicate removal (vgl_dedup): keys, hash table, mask -- kept between calls (allocation costs more than the two kernels)
    DevBuf<unsigned long long> dd_key, dd_tab_key;
    DevBuf<uint32_t> dd_tab_first;
    DevBuf<uint8_t> dd_paired, dd_dup;
    DevBuf<unsigned int> dd_n;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;

    // HaploCart tables
    bool have_tables = false;
    bool have_tree = false;
    vgl::HcTablesDev tab{};
    int row_words = 0;
    int n_rows_pad = 0;
    DevBuf<double2> d_class_tab;
    DevBuf<double> d_logterm;   // log-term tables [100][n_class + 1][512] (vgl_internal.h)
    DevBuf<unsigned char> d_image;  // the read kernel's shared-memory image (vgl::HcImageLayout)
    DevBuf<uint32_t> d_node_bits;
    vgl::HcReadLaunch read_cfg{};

    // accumulators
    // one allocation, one memset per sample: node vectors [copies][n_rows_pad] | A totals [copies] | error flag
    DevBuf<double> d_cta_node_acc, d_node_sum, d_chunk_ll, d_ll;   // d_node_sum: the copies collapsed (K2a), hc_node_sum_doubles()
    DevBuf<unsigned int> d_tile_done;  // per path tile: CTAs of the path kernel that have finished
    double* d_a_total = nullptr;
    // Accumulator set the accumulate entry points add into: the single-sample buffer above, or sample `cur_sample` of the
    // multi-sample buffer (vgl_hc_samples_begin / vgl_hc_sample_select).  One set = [copies][n_rows_pad] | A totals | error word.
    double* cur_acc = nullptr;
    double* cur_a_total = nullptr;
    int* cur_err = nullptr;
    DevBuf<double> m_acc, m_node_sum, m_chunk_ll, m_ll, m_result;   // [S] sets / [S] collapsed vectors / [S][chunks][P] / [S][P] / [S][4 + cap]
    DevBuf<unsigned int> m_tile_done;                   // [S][tiles]
    double* m_h_result = nullptr;                       // page-locked mirror of m_result
    size_t m_h_result_cap = 0;
    int m_samples = 0, cur_sample = -1;
    // (a multiple of 32 doubles: every sample's set starts on a 256-byte boundary, as the bulk copies of the path kernel need)
    size_t acc_set_doubles() const { return ((size_t)VGL_HC_ACC_COPIES * (size_t)n_rows_pad + VGL_HC_ACC_COPIES + 1 + 31) & ~(size_t)31; }
    int* d_err = nullptr;
    int n_chunks = 0;
    bool ll_valid = false;

    // tree
    int n_clades = 0;
    int max_anc = 0;
    DevBuf<uint32_t> d_clade_bits;
    DevBuf<int32_t> d_anc_off, d_anc_clade;

    // posterior outputs
    // device-side reconstruction (recon_kernels.cu): node sequences, staged raw alignments, scratch, the packed batch
    bool have_graph = false;
    uint32_t g_min_id = 0, g_n_nodes = 0;
    DevBuf<uint32_t> g_seq_off;
    DevBuf<uint8_t> g_seq;
    struct RawDev {
        DevBuf<uint32_t> qual_off, str_off, map_off, map_node, map_offset, edit_off, edit_seq_off;
        DevBuf<uint16_t> edit_from, edit_to;
        DevBuf<uint8_t> qual, mapq, edit_seq;
        DevBuf<uint8_t> gstr, rstr;
        DevBuf<uint16_t> sizes, n_cols, n_close, n_extra;
        DevBuf<uint32_t> hist3, order, grp_cb, grp_mb;   // hist3: hist | start | cursor
        DevBuf<uint32_t> col, xmap;
        DevBuf<uint16_t> o_n_cols, o_n_close, o_n_extra;
        DevBuf<uint8_t> o_mapq;
        uint32_t col_cap_blks = 0, map_cap_blks = 0;
        // euka / soibean batches built on the device
        DevBuf<uint16_t> seq_len, s_col, s_n_cols, s_n_maps, s_seq_len;
        DevBuf<uint32_t> read_index, s_map, s_read_index;
        DevBuf<uint8_t> s_mapq, s_flags;
        DevBuf<int> s_err;
        // narrow arrays as shipped + scratch for the prefix sums
        DevBuf<uint16_t> off16, n16, fo16;
        DevBuf<uint8_t> ne8, sl8, f8, t8;
        DevBuf<uint32_t> tiles;
        cudaEvent_t copied = nullptr, consumed = nullptr;  // H2D of the chunk done / its kernels done
        bool used = false;
    } rawdev[2];

    // multi-GPU exchange through peer memory
    double* peer_local = nullptr;             // this rank's buffer (cudaMalloc)
    double* peer_bufs[VGL_HC_MAX_PEERS] = {};  // every rank's buffer as mapped here
    int peer_rank = -1, peer_world = 0, peer_pad = 0;
    unsigned long long peer_epoch = 0;

    // one block {double total, max; int32 argmax, n_post, err, pad; double post[]} and its pinned host mirror
    // result blocks {total, ll[argmax]}, {argmax, n_post, err}, post[cap]: TWO of them (device + page-locked mirror + event), so that
    // a sample's posterior kernel and copy can be in flight while the next sample is already being accumulated
    // (vgl_hc_finalize_begin / _end); slot k starts at k * (4 + h_result_cap) doubles
    DevBuf<double> d_result;
    double* h_result = nullptr;
    size_t h_result_cap = 0;
    cudaEvent_t res_ev[2] = {nullptr, nullptr};
    uint32_t res_seq = 0;
    bool res_pending[2] = {false, false};

    // staging for host batches: two slots so that the H2D copy of chunk k+1 overlaps the read kernel on chunk k
    struct Slot {
        DevBuf<uint32_t> grp_col_blk, grp_map_blk, col, map;
        DevBuf<uint16_t> n_cols, n_close, n_extra;
        DevBuf<uint8_t> mapq;
        DevBuf<uint32_t> grp_c16, grp_n16;  // wire form of the chunk (vgl_hc_wire)
        DevBuf<uint16_t> col16, node16, n_owned;
        DevBuf<uint8_t> col8;  // narrow columns of the wire form
        DevBuf<uint32_t> flag4;
        cudaEvent_t copied = nullptr, consumed = nullptr;
        bool used = false;
    } slot[2];
    cudaStream_t copy_stream = nullptr;
    uint32_t chunk_groups = 4096;  // 131072 reads per H2D chunk
    bool peer_fused_pending = false;  // vgl_hc_exchange_paths ran K2 with the exchange epilogue: the next posterior kernel completes it
    double peer_timeout_s = 120.0;    // a peer that never delivers raises VGL_ERR_PEER after this long (VGL_PEER_TIMEOUT_S)
    bool logterm_off = false;      // VGL_HC_TABLE_MODE pins the general kernel (tests)
    uint64_t logterm_launches = 0; // read-kernel launches that took the log-term kernel
    uint32_t raw_seq = 0;          // chunks submitted through vgl_hc_accumulate_raw so far (selects the staging set)

    // read-kernel timing
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;

    // euka
    bool have_euka = false;
    vgl::EukaTablesDev etab{};
    DevBuf<uint16_t> e_node_clade;
    DevBuf<double> e_dist, e_deam, e_bin_cov, e_in, e_not, e_like;
    DevBuf<int32_t> e_bin_off, e_bin_lo, e_bin_hi, e_clade;
    DevBuf<vgl::EukaSmallTables> e_small;
    DevBuf<unsigned long long> e_count, e_like_n;
    DevBuf<double> e_like_logsum;
    DevBuf<uint8_t> e_pass;
    // soibean
    bool have_soibean = false;
    vgl::SoibeanTablesDev stab{};
    DevBuf<uint32_t> s_member;
    DevBuf<double> s_deam, s_sup, s_uns, s_out;
    DevBuf<uint32_t> s_mcode;
    DevBuf<vgl::EukaSmallTables> s_small;
    DevBuf<uint8_t> s_used;
    DevBuf<int> s_err;
    // soibean MCMC: the batch of the last soibean run stays resident (staging area + s_sup / s_uns)
    bool s_last_valid = false;
    vgl::EukaBatchDev s_last{};
    size_t s_last_nmap = 0;                      // edit records of s_last
    bool e_last_valid = false;                   // euka: the batch the last vgl_euka_run / _run_raw scored is still in HBM
    vgl::EukaBatchDev e_last{};
    cudaEvent_t sib_ev0 = nullptr, sib_ev1 = nullptr;  // around the last euka read kernel / soibean kernel pair
    bool sib_timed = false;
    DevBuf<vgl_soibean_source> m_src;
    DevBuf<double> m_partials, m_loglike, m_per_read, m_H;
    unsigned char* m_pinned = nullptr;  // proposal in, log-likelihoods + error flag out
    size_t m_pinned_cap = 0;
    struct EukaStage {
        DevBuf<uint32_t> grp_col_blk, grp_map_blk, map, read_index;
        DevBuf<uint16_t> col, n_cols, n_maps, seq_len;
        DevBuf<uint8_t> mapq, flags;
    } estage;
};

namespace {

int use_device(vgl_ctx* ctx) {
    if (!ctx) return fail(VGL_ERR_INVALID, "null context");
    VGL_CUDA(cudaSetDevice(ctx->device));
    return VGL_OK;
}

int upload_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return VGL_OK;
    VGL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return VGL_OK;
}

vgl::HcBatchMeta to_meta(const vgl_hc_batch_meta& m) {
    vgl::HcBatchMeta r;
    r.valid = m.valid != 0;
    if (r.valid) {
        r.qual_max = m.qual_max;
        r.n_segs = std::min<uint32_t>(m.n_segs, VGL_HC_MAX_SEGS);
        for (uint32_t k = 0; k < r.n_segs; ++k) { r.seg_mapq[k] = m.seg_mapq[k]; r.seg_end[k] = m.seg_end[k]; }
    }
    return r;
}

// The segments of groups [g0, g1) of a batch, renumbered from g0 (chunked launches).
vgl::HcBatchMeta clip_meta(const vgl::HcBatchMeta& m, uint32_t g0, uint32_t g1) {
    vgl::HcBatchMeta r;
    r.valid = m.valid;
    r.qual_max = m.qual_max;
    uint32_t begin = 0;
    for (uint32_t k = 0; k < m.n_segs; ++k) {
        const uint32_t end = m.seg_end[k], lo = std::max(begin, g0), hi = std::min(end, g1);
        if (hi > lo) {
            r.seg_mapq[r.n_segs] = m.seg_mapq[k];
            r.seg_end[r.n_segs] = hi - g0;
            r.n_segs += 1;
        }
        begin = end;
    }
    return r;
}

int launch_read(vgl_ctx* ctx, const vgl::HcBatchDev& b, const vgl::HcBatchMeta& meta) {
    if (b.n_groups == 0) return VGL_OK;
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t a, c;
        VGL_CUDA(cudaEventCreate(&a));
        VGL_CUDA(cudaEventCreate(&c));
        ctx->ev_pool.push_back(std::make_pair(a, c));
    }
    auto& ev = ctx->ev_pool[ctx->ev_used++];
    // the log-term kernel where the batch allows it (no quality above 63, tables of the pairs present fit), else the general one
    vgl::HcLogTermLaunch plan;
    const size_t lt_smem = ctx->logterm_off ? 0 : vgl::hc_logterm_plan(meta, b.n_groups, ctx->sm_count, ctx->tab.n_rows, ctx->tab.n_class,
                                                                      ctx->tab.logterm != nullptr, ctx->max_smem_optin, plan);
    VGL_CUDA(cudaEventRecord(ev.first, ctx->stream));
    if (lt_smem) {
        VGL_CUDA(vgl::hc_launch_read_logterm_kernel(ctx->sm_count, lt_smem, plan, b, ctx->tab, ctx->cur_acc, ctx->cur_a_total, ctx->cur_err,
                                                    ctx->stream));
        ctx->logterm_launches += 1;
    } else {
        VGL_CUDA(vgl::hc_launch_read_kernel(ctx->read_cfg, b, ctx->tab, ctx->cur_acc, ctx->cur_a_total,
                                            ctx->cur_err, ctx->stream));
    }
    VGL_CUDA(cudaEventRecord(ev.second, ctx->stream));
    ctx->launches += 1;
    ctx->ll_valid = false;
    return VGL_OK;
}

void peer_detach(vgl_ctx* ctx) {
    for (int k = 0; k < VGL_HC_MAX_PEERS; ++k) {
        if (ctx->peer_bufs[k] && ctx->peer_bufs[k] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_bufs[k]);
        ctx->peer_bufs[k] = nullptr;
    }
    if (ctx->peer_local) cudaFree(ctx->peer_local);
    ctx->peer_local = nullptr;
    ctx->peer_rank = -1;
    ctx->peer_world = 0;
}

int reserve_result(vgl_ctx* ctx, size_t n_post) {
    for (int k = 0; k < 2; ++k)
        if (!ctx->res_ev[k]) VGL_CUDA(cudaEventCreateWithFlags(&ctx->res_ev[k], cudaEventDisableTiming));
    if (n_post <= ctx->h_result_cap) return VGL_OK;
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));   // (a result in flight still points into the old blocks)
    ctx->res_pending[0] = ctx->res_pending[1] = false;
    if (int rc = ctx->d_result.reserve(2 * (4 + n_post))) return rc;
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    ctx->h_result = nullptr;
    ctx->h_result_cap = 0;
    VGL_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_result), 2 * (4 + n_post) * sizeof(double)));
    ctx->h_result_cap = n_post;
    return VGL_OK;
}

int sibling_events(vgl_ctx* ctx) {
    if (!ctx->sib_ev0) VGL_CUDA(cudaEventCreate(&ctx->sib_ev0));
    if (!ctx->sib_ev1) VGL_CUDA(cudaEventCreate(&ctx->sib_ev1));
    return VGL_OK;
}
int euka_run_on_device(vgl_ctx* ctx, const vgl::EukaBatchDev& d, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik,
                       double* like, uint8_t* pass);
int soibean_run_on_device(vgl_ctx* ctx, const vgl::EukaBatchDev& d, size_t n_map_rec, uint32_t n_out, double* path_ll, uint8_t* used);
int pack_raw_symbols(vgl_ctx* ctx, const vgl_hc_raw* raw, int soibean, vgl::EukaBatchDev& d, size_t& n_map_rec);
int stage_symbol_batch(vgl_ctx* ctx, const vgl_euka_batch* b, vgl::EukaBatchDev& d);

int check_batch(const vgl_hc_batch* b) {
    if (!b) return fail(VGL_ERR_INVALID, "null batch");
    if (b->n_groups == 0) return VGL_OK;
    if ((uint64_t)b->n_groups * VGL_HC_LANES < b->n_reads) return fail(VGL_ERR_INVALID, "n_groups too small for n_reads");
    if (!b->grp_col_blk || !b->grp_map_blk || !b->n_cols || !b->n_close || !b->n_extra || !b->mapq)
        return fail(VGL_ERR_INVALID, "batch header arrays missing");
    if (b->n_col_blks && !b->col) return fail(VGL_ERR_INVALID, "batch column records missing");
    if (b->n_map_blks && !b->map) return fail(VGL_ERR_INVALID, "batch extra records missing");
    return VGL_OK;
}

}  // namespace

extern "C" {

int vgl_abi_version(void) { return VGL_ABI_VERSION; }
const char* vgl_last_error(void) { return g_err; }

int vgl_create(int device_id, vgl_ctx** out) {
    if (!out) return fail(VGL_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(VGL_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_id < 0 || device_id >= n) return fail(VGL_ERR_INVALID, "device %d out of range (have %d)", device_id, n);
    vgl_ctx* ctx = new (std::nothrow) vgl_ctx();
    if (!ctx) return fail(VGL_ERR_INVALID, "out of host memory");
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
        delete ctx;
        return fail(VGL_ERR_CUDA, "cannot select device %d", device_id);
    }
    if (prop.major < 10) {
        delete ctx;
        return fail(VGL_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a only", device_id, prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(VGL_ERR_CUDA, "cudaStreamCreate failed");
    }
    ctx->stream = ctx->own_stream;
    bool ok = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& sl : ctx->slot)
        ok = ok && cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        vgl_destroy(ctx);
        return fail(VGL_ERR_CUDA, "cannot create the copy stream / events");
    }
    *out = ctx;
    return VGL_OK;
}

void vgl_destroy(vgl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->dd_key.release(); ctx->dd_tab_key.release(); ctx->dd_tab_first.release(); ctx->dd_paired.release(); ctx->dd_dup.release(); ctx->dd_n.release();
    ctx->d_class_tab.release(); ctx->d_image.release(); ctx->d_node_bits.release(); ctx->d_logterm.release();
    ctx->d_cta_node_acc.release();
    ctx->d_node_sum.release(); ctx->m_node_sum.release();
    ctx->d_chunk_ll.release(); ctx->d_ll.release(); ctx->d_tile_done.release();
    ctx->m_acc.release(); ctx->m_chunk_ll.release(); ctx->m_ll.release(); ctx->m_result.release(); ctx->m_tile_done.release();
    if (ctx->m_h_result) cudaFreeHost(ctx->m_h_result);
    ctx->d_clade_bits.release(); ctx->d_anc_off.release(); ctx->d_anc_clade.release();
    ctx->d_result.release();
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    for (auto& ev : ctx->res_ev) if (ev) cudaEventDestroy(ev);
    peer_detach(ctx);
    ctx->g_seq_off.release(); ctx->g_seq.release();
    for (vgl_ctx::RawDev& R : ctx->rawdev) {
        R.qual_off.release(); R.str_off.release(); R.map_off.release(); R.map_node.release(); R.map_offset.release(); R.edit_off.release();
        R.edit_seq_off.release(); R.edit_from.release(); R.edit_to.release(); R.qual.release(); R.mapq.release(); R.edit_seq.release();
        R.gstr.release(); R.rstr.release(); R.sizes.release(); R.n_cols.release(); R.n_close.release(); R.n_extra.release();
        R.hist3.release(); R.order.release(); R.grp_cb.release(); R.grp_mb.release(); R.col.release(); R.xmap.release();
        R.o_n_cols.release(); R.o_n_close.release(); R.o_n_extra.release(); R.o_mapq.release();
        R.seq_len.release(); R.s_col.release(); R.s_n_cols.release(); R.s_n_maps.release(); R.s_seq_len.release();
        R.off16.release(); R.n16.release(); R.fo16.release(); R.ne8.release(); R.sl8.release(); R.f8.release(); R.t8.release(); R.tiles.release();
        R.read_index.release(); R.s_map.release(); R.s_read_index.release(); R.s_mapq.release(); R.s_flags.release(); R.s_err.release();
        if (R.copied) cudaEventDestroy(R.copied);
        if (R.consumed) cudaEventDestroy(R.consumed);
    }
    for (auto& sl : ctx->slot) {
        sl.grp_col_blk.release(); sl.grp_map_blk.release(); sl.map.release(); sl.col.release(); sl.n_cols.release();
        sl.n_close.release(); sl.n_extra.release(); sl.mapq.release();
        sl.grp_c16.release(); sl.grp_n16.release(); sl.col16.release(); sl.node16.release(); sl.col8.release(); sl.flag4.release(); sl.n_owned.release();
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.consumed) cudaEventDestroy(sl.consumed);
    }
    ctx->e_node_clade.release(); ctx->e_dist.release(); ctx->e_deam.release(); ctx->e_bin_cov.release(); ctx->e_in.release();
    ctx->e_not.release(); ctx->e_like.release(); ctx->e_bin_off.release(); ctx->e_bin_lo.release(); ctx->e_bin_hi.release();
    ctx->e_clade.release(); ctx->e_small.release(); ctx->e_count.release(); ctx->e_pass.release();
    ctx->e_like_n.release(); ctx->e_like_logsum.release();
    ctx->s_member.release(); ctx->s_deam.release(); ctx->s_sup.release(); ctx->s_uns.release(); ctx->s_out.release(); ctx->s_mcode.release();
    ctx->m_src.release(); ctx->m_partials.release(); ctx->m_loglike.release(); ctx->m_per_read.release(); ctx->m_H.release();
    if (ctx->m_pinned) cudaFreeHost(ctx->m_pinned);
    ctx->s_small.release(); ctx->s_used.release(); ctx->s_err.release();
    ctx->estage.grp_col_blk.release(); ctx->estage.grp_map_blk.release(); ctx->estage.map.release(); ctx->estage.read_index.release();
    ctx->estage.col.release(); ctx->estage.n_cols.release(); ctx->estage.n_maps.release(); ctx->estage.seq_len.release();
    ctx->estage.mapq.release(); ctx->estage.flags.release();
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& ev : ctx->ev_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int vgl_set_stream(vgl_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(VGL_ERR_INVALID, "null context");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return VGL_OK;
}

int vgl_synchronize(vgl_ctx* ctx) {
    if (int rc = use_device(ctx)) return rc;
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    return VGL_OK;
}

int vgl_host_alloc(void** out, size_t bytes) {
    if (!out) return fail(VGL_ERR_INVALID, "null out pointer");
    VGL_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return VGL_OK;
}
int vgl_host_free(void* p) {
    if (p) VGL_CUDA(cudaFreeHost(p));
    return VGL_OK;
}
int vgl_device_alloc(vgl_ctx* ctx, void** out, size_t bytes) {
    if (int rc = use_device(ctx)) return rc;
    if (!out) return fail(VGL_ERR_INVALID, "null out pointer");
    VGL_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return VGL_OK;
}
int vgl_device_free(vgl_ctx* ctx, void* p) {
    if (int rc = use_device(ctx)) return rc;
    if (p) VGL_CUDA(cudaFree(p));
    return VGL_OK;
}
int vgl_memcpy_h2d(vgl_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (int rc = use_device(ctx)) return rc;
    VGL_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    return VGL_OK;
}
int vgl_memcpy_d2h(vgl_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (int rc = use_device(ctx)) return rc;
    VGL_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    return VGL_OK;
}

uint64_t vgl_launch_count(const vgl_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t vgl_hc_logterm_launches(const vgl_ctx* ctx) { return ctx ? ctx->logterm_launches : 0; }

uint64_t vgl_hc_batch_bytes(const vgl_hc_batch* b) {
    if (!b) return 0;
    return (uint64_t)(b->n_groups + 1) * 8 + (b->n_col_blks + b->n_map_blks) * 512 +
           (uint64_t)b->n_groups * VGL_HC_LANES * 7;
}

// -------------------------------------------------------------------------------------------
// tables
// -------------------------------------------------------------------------------------------
int vgl_hc_set_tables(vgl_ctx* ctx, const vgl_hc_tables* T) {
    if (int rc = use_device(ctx)) return rc;
    if (!T) return fail(VGL_ERR_INVALID, "null tables");
    if (T->n_rows <= 0 || T->n_paths <= 0) return fail(VGL_ERR_INVALID, "empty tables (%d rows, %d paths)", T->n_rows, T->n_paths);
    if (T->row_words < (T->n_paths + 31) / 32) return fail(VGL_ERR_INVALID, "row_words %d too small for %d paths", T->row_words, T->n_paths);
    if (!T->support_bits || !T->node_mappability || !T->node_match_prob || !T->node_valid)
        return fail(VGL_ERR_INVALID, "table arrays missing");
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));

    const int N = T->n_rows, P = T->n_paths;
    const int n_words = (N + 31) / 32;

    // (mappability, match_prob) dictionary: one class per distinct pair, so the read kernel can keep a
    // 2-byte class id per node in shared memory instead of 16 bytes per node.
    std::vector<uint16_t> node_class((size_t)N, (uint16_t)VGL_HC_CLASS_NONE);
    std::vector<double2> class_tab;
    {
        std::map<std::pair<double, double>, int> dict;
        for (int n = 0; n < N; ++n) {
            if (!T->node_valid[n]) continue;
            const std::pair<double, double> key(T->node_mappability[n], T->node_match_prob[n]);
            auto it = dict.find(key);
            int id;
            if (it == dict.end()) {
                id = (int)class_tab.size();
                if (id >= (int)VGL_HC_CLASS_NONE) return fail(VGL_ERR_INVALID, "more than 32767 distinct (mappability, match_prob) pairs");
                dict.insert(std::make_pair(key, id));
                class_tab.push_back(make_double2(key.first, key.second));
            } else {
                id = it->second;
            }
            node_class[(size_t)n] = (uint16_t)id;
            // a node every path goes through never appears in any path's sum over unsupported nodes: mark it, the read
            // kernel then skips the scatter for its mappings
            const uint32_t* row = T->support_bits + (size_t)n * (size_t)T->row_words;
            bool all = true;
            for (int p = 0; p < P && all; p += 32) {
                const uint32_t want = P - p >= 32 ? 0xffffffffu : (1u << (P - p)) - 1u;
                all = (row[p >> 5] & want) == want;
            }
            if (all) node_class[(size_t)n] |= (uint16_t)VGL_HC_CLASS_ALLPATHS;
        }
        if (class_tab.empty()) class_tab.push_back(make_double2(0.0, 0.0));
    }

    // Small tables, with the reference's exact double-precision expressions.
    vgl::HcSmallTables small;
    memset(&small, 0, sizeof(small));
    const double bep = T->background_error_prob;
    small.f_match_bep = 1 - bep;
    small.f_mis_bep = 1 - (1 - bep);
    small.one_minus_bep = 1 - bep;
    small.is_fasta = T->is_consensus_fasta ? 1 : 0;
    small.log_err_scale = -std::log(10.0) / 10.0;
    // per-column factor 1 - eps, by the low 12 bits of the column record: mismatch | base << 1 | quality << 3 | latched << 11.
    // The quality field is canonical (0 for Q <= 2 and for bytes that are negative as signed char, i.e. wherever the error
    // probability is 0.25: qscore_vec[0..2] = 0.25, src/miscfunc.h:198-212; a negative byte would index qscore_vec out of
    // range in the reference -- defined as 0.25).
    for (int idx = 0; idx < 4096; ++idx) {
        const int mm = idx & 1, Q = (idx >> VGL_HC_COL_QUAL_SHIFT) & 0xff, latched = (idx >> 11) & 1;
        // get_p_no_seq_error_mapping (src/get_p_obs_base.cpp:3-27): eps = qscore[Q] for a match, 1 - qscore[Q] for a mismatch;
        // the background error probability instead of qscore[Q] once a quality byte >= 90 was seen, and for every base in
        // consensus-FASTA mode (src/HaploCart.cpp:397-401).  get_p_obs_base: match * (1 - epsilon), src/get_p_obs_base.cpp:67.
        const double qs = (Q >= 0 && Q < 100) ? T->qscore[Q] : 0.25;
        const double eps = (latched || small.is_fasta) ? (mm ? 1 - bep : bep) : (mm ? 1 - qs : qs);
        small.f_col[idx] = 1 - eps;
    }
    for (int b = 0; b < 256; ++b) {
        const int Q = (int)(int8_t)(uint8_t)b;  // the reference reads quality bytes as (signed) char
        const double perr = (Q > 2) ? std::pow(10, ((-1 * Q) * 0.1)) : 0.25;  // get_p_seq_error, src/miscfunc.h:180-187
        small.log_err[b] = std::log(perr);
        // fast_log table: rc ~ 1 / c_k and -log(rc) for the bucket centre c_k = 1 + (k + 0.5) / 256
        const double rc = 1.0 / (1.0 + (b + 0.5) / 256.0);
        small.log_tab[b] = make_double2(rc, (double)(-logl((long double)rc)));
    }
    for (int q = 0; q < 100; ++q) small.p_correct[q] = 1 - T->p_incorrect_map[q];
    for (int k = 0; k < 4; ++k) small.bg[k] = T->bg_freq[k];

    // Membership matrix re-laid as node-words x paths: word (w, p) holds pathsgo[p] of nodes 32w..32w+31.
    // Bits of nodes >= n_rows are set ("supported": contributes nothing).
    std::vector<uint32_t> node_bits((size_t)n_words * (size_t)P, 0u);
    for (int n = 0; n < n_words * 32; ++n) {
        uint32_t* dst = node_bits.data() + (size_t)(n >> 5) * (size_t)P;
        const uint32_t bit = 1u << (n & 31);
        if (n >= N) {
            for (int p = 0; p < P; ++p) dst[p] |= bit;
            continue;
        }
        const uint32_t* row = T->support_bits + (size_t)n * (size_t)T->row_words;
        for (int w = 0; w < (P + 31) / 32; ++w) {
            uint32_t word = row[w];
            const int pbase = w * 32;
            while (word) {
                const int k = __builtin_ctz(word);
                word &= word - 1;
                if (pbase + k < P) dst[pbase + k] |= bit;
            }
        }
    }

    const int n_class = (int)class_tab.size();
    class_tab.push_back(make_double2(0.0, 0.0));  // neutral entry for nodes without a class
    int force_mode = -1;  // VGL_HC_TABLE_MODE=0|1|2 pins the table layout of the read kernel (tests)
    if (const char* e = getenv("VGL_HC_TABLE_MODE")) force_mode = atoi(e);
    ctx->read_cfg = vgl::hc_read_launch_config(N, n_class, ctx->sm_count, ctx->max_smem_optin, force_mode);
    if (ctx->read_cfg.grid == 0)
        return fail(VGL_ERR_INVALID, "%d node rows need %zu bytes of shared memory per CTA (limit %d)", N, ctx->read_cfg.smem,
                    ctx->max_smem_optin);
    VGL_CUDA(vgl::hc_read_configure(ctx->read_cfg));

    // the read kernel's shared-memory image
    const vgl::HcImageLayout L = vgl::hc_image_layout(N, n_class, ctx->read_cfg.ac_mode);
    std::vector<unsigned char> image((size_t)L.bytes, 0);
    memcpy(image.data() + L.off_small, &small, sizeof(small));
    {
        uint16_t* cls = reinterpret_cast<uint16_t*>(image.data() + L.off_class);
        memcpy(cls, node_class.data(), (size_t)N * sizeof(uint16_t));
        cls[N] = (uint16_t)VGL_HC_CLASS_NONE;  // "no such node"
        double2* ac = reinterpret_cast<double2*>(image.data() + L.off_ac);
        if (ctx->read_cfg.ac_mode == 0) {
            // {1 - pcm, pcm * K} for every (mapping quality, class): p_correctly_mapped = (1 - p_incorrect[mq]) * mappability
            // (src/process_mapping.cpp:41); consensus-FASTA mode uses (1 - bep) * K alone (:73)
            const size_t stride = L.ac_row_bytes / 16;
            for (int mq = 0; mq < 100; ++mq) {
                for (int c = 0; c < n_class; ++c) {
                    const double pcm = small.p_correct[mq] * class_tab[(size_t)c].x;
                    ac[(size_t)mq * stride + c] = make_double2(small.is_fasta ? 0.0 : 1.0 - pcm,
                                                               (small.is_fasta ? small.one_minus_bep : pcm) * class_tab[(size_t)c].y);
                }
                ac[(size_t)mq * stride + n_class] = make_double2(1.0, 0.0);  // neutral entry for unknown nodes
            }
        } else if (ctx->read_cfg.ac_mode == 1) {
            memcpy(ac, class_tab.data(), class_tab.size() * sizeof(double2));
        }
    }
    if (int rc = ctx->d_class_tab.reserve(class_tab.size())) return rc;
    if (int rc = ctx->d_image.reserve(image.size())) return rc;
    if (int rc = ctx->d_node_bits.reserve(node_bits.size())) return rc;
    VGL_CUDA(cudaMemcpy(ctx->d_class_tab.p, class_tab.data(), class_tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->d_image.p, image.data(), image.size(), cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->d_node_bits.p, node_bits.data(), node_bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));

    ctx->tab.n_rows = N;
    ctx->tab.n_paths = P;
    ctx->tab.n_words = n_words;
    ctx->tab.n_class = n_class;
    ctx->tab.class_tab = ctx->d_class_tab.p;
    ctx->tab.image = ctx->d_image.p;
    ctx->tab.layout = L;
    ctx->tab.node_bits = ctx->d_node_bits.p;
    // Log-term tables for the fastest read kernel: for every (mapping quality, class) the term of every (quality <= 63,
    // mismatch, base) combination, computed exactly as the general kernel forms the factor (fma(cc, f, aa * bg)) and then
    // log() of it.  Built when the class count is small enough for some launch to be able to use them.
    ctx->tab.logterm = nullptr;
    ctx->logterm_off = force_mode >= 0;
    if (n_class + 1 <= 256) {
        const size_t per_mq = (size_t)(n_class + 1) * vgl::kLogTermEntries;
        std::vector<double> lt(100 * per_mq);
        for (int mq = 0; mq < 100; ++mq) {
            for (int c = 0; c <= n_class; ++c) {
                double aa = 1.0, cc = 0.0;  // the neutral entry of unknown nodes (they raise VGL_ERR_BAD_NODE anyway)
                if (c < n_class) {
                    const double pcm = small.p_correct[mq] * class_tab[(size_t)c].x;
                    aa = small.is_fasta ? 0.0 : 1.0 - pcm;
                    cc = (small.is_fasta ? small.one_minus_bep : pcm) * class_tab[(size_t)c].y;
                }
                double* row = lt.data() + (size_t)mq * per_mq + (size_t)c * vgl::kLogTermEntries;
                for (int idx = 0; idx < vgl::kLogTermEntries; ++idx) {
                    const double f = small.f_col[idx];  // quality <= 63, not latched: the low 9 bits are the whole index
                    const double bgv = small.bg[(idx >> VGL_HC_COL_BASE_SHIFT) & 3];
                    row[idx] = std::log(std::fma(cc, f, aa * bgv));
                }
            }
        }
        if (int rc = ctx->d_logterm.reserve(lt.size())) return rc;
        VGL_CUDA(cudaMemcpy(ctx->d_logterm.p, lt.data(), lt.size() * sizeof(double), cudaMemcpyHostToDevice));
        ctx->tab.logterm = ctx->d_logterm.p;
        VGL_CUDA(vgl::hc_logterm_configure(ctx->max_smem_optin));
    }
    ctx->row_words = (P + 31) / 32;
    ctx->n_rows_pad = n_words * 32;
    ctx->n_chunks = vgl::hc_path_chunks(n_words);

    if (int rc = ctx->d_cta_node_acc.reserve((size_t)VGL_HC_ACC_COPIES * (size_t)ctx->n_rows_pad + VGL_HC_ACC_COPIES + 1)) return rc;
    ctx->d_a_total = ctx->d_cta_node_acc.p + (size_t)VGL_HC_ACC_COPIES * (size_t)ctx->n_rows_pad;
    ctx->d_err = reinterpret_cast<int*>(ctx->d_a_total + VGL_HC_ACC_COPIES);
    ctx->cur_acc = ctx->d_cta_node_acc.p; ctx->cur_a_total = ctx->d_a_total; ctx->cur_err = ctx->d_err;
    ctx->m_samples = 0; ctx->cur_sample = -1;
    if (int rc = ctx->d_node_sum.reserve(vgl::hc_node_sum_doubles(ctx->tab.n_words))) return rc;
    if (int rc = ctx->d_chunk_ll.reserve((size_t)ctx->n_chunks * (size_t)P)) return rc;
    if (int rc = ctx->d_ll.reserve((size_t)P)) return rc;
    if (int rc = ctx->d_tile_done.reserve((size_t)vgl::hc_path_tiles(P))) return rc;
    VGL_CUDA(cudaMemset(ctx->d_tile_done.p, 0, (size_t)vgl::hc_path_tiles(P) * sizeof(unsigned int)));
    if (int rc = reserve_result(ctx, 64)) return rc;
    if (vgl::hc_posterior_smem(P) > 48 * 1024) VGL_CUDA(vgl::hc_posterior_configure(vgl::hc_posterior_smem(P)));
    ctx->have_tables = true;
    ctx->have_tree = false;
    return vgl_hc_reset(ctx);
}

int vgl_hc_set_tree(vgl_ctx* ctx, int32_t n_clades, const uint32_t* clade_bits, const int32_t* anc_off,
                    const int32_t* anc_clade) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (n_clades < 0 || !anc_off || (n_clades > 0 && (!clade_bits || !anc_clade))) return fail(VGL_ERR_INVALID, "tree arrays missing");
    const int P = ctx->tab.n_paths;
    int max_anc = 0;
    if (anc_off[0] != 0) return fail(VGL_ERR_INVALID, "anc_off[0] must be 0");
    for (int p = 0; p < P; ++p) {
        const int k = anc_off[p + 1] - anc_off[p];
        if (k < 0) return fail(VGL_ERR_INVALID, "anc_off not monotone at path %d", p);
        max_anc = std::max(max_anc, k);
    }
    const int n_anc = anc_off[P];
    for (int k = 0; k < n_anc; ++k)
        if (anc_clade[k] < 0 || anc_clade[k] >= n_clades) return fail(VGL_ERR_INVALID, "anc_clade[%d] = %d out of range", k, anc_clade[k]);
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (int rc = ctx->d_clade_bits.reserve((size_t)std::max(n_clades, 1) * (size_t)ctx->row_words)) return rc;
    if (int rc = ctx->d_anc_off.reserve((size_t)P + 1)) return rc;
    if (int rc = ctx->d_anc_clade.reserve((size_t)std::max(n_anc, 1))) return rc;
    if (int rc = reserve_result(ctx, (size_t)max_anc + 1)) return rc;
    if (n_clades)
        VGL_CUDA(cudaMemcpy(ctx->d_clade_bits.p, clade_bits, (size_t)n_clades * ctx->row_words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->d_anc_off.p, anc_off, ((size_t)P + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (n_anc) VGL_CUDA(cudaMemcpy(ctx->d_anc_clade.p, anc_clade, (size_t)n_anc * sizeof(int32_t), cudaMemcpyHostToDevice));
    ctx->n_clades = n_clades;
    ctx->max_anc = max_anc;
    ctx->have_tree = true;
    return VGL_OK;
}

// -------------------------------------------------------------------------------------------
// hot path
// -------------------------------------------------------------------------------------------
int vgl_hc_reset(vgl_ctx* ctx) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    cudaStream_t s = ctx->stream;
    VGL_CUDA(cudaMemsetAsync(ctx->d_cta_node_acc.p, 0, ((size_t)VGL_HC_ACC_COPIES * ctx->n_rows_pad + VGL_HC_ACC_COPIES + 1) * sizeof(double), s));
    ctx->cur_acc = ctx->d_cta_node_acc.p; ctx->cur_a_total = ctx->d_a_total; ctx->cur_err = ctx->d_err;   // single-sample mode
    ctx->peer_fused_pending = false;
    ctx->m_samples = 0; ctx->cur_sample = -1;
    ctx->ev_used = 0;
    ctx->ll_valid = false;
    return VGL_OK;
}

int vgl_hc_accumulate_device(vgl_ctx* ctx, const vgl_hc_batch* b) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (int rc = check_batch(b)) return rc;
    vgl::HcBatchDev d{b->n_groups, b->grp_col_blk, b->grp_map_blk, b->col, b->map, b->n_cols, b->n_close, b->n_extra, b->mapq};
    return launch_read(ctx, d, to_meta(b->meta));
}

int vgl_hc_accumulate(vgl_ctx* ctx, const vgl_hc_batch* b) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (int rc = check_batch(b)) return rc;
    if (b->n_groups == 0) return VGL_OK;
    // The batch is cut into chunks of groups; the H2D copy of a chunk runs on the copy stream into one of two
    // staging slots while the read kernel works on the other.  Row offsets stay absolute: the kernel is given
    // record bases shifted back by the chunk's first row.
    const uint32_t G = b->n_groups;
    uint32_t n_chunk = 0;
    for (uint32_t g0 = 0; g0 < G; g0 += ctx->chunk_groups, ++n_chunk) {
        const uint32_t g1 = std::min(G, g0 + ctx->chunk_groups);
        const uint32_t ng = g1 - g0;
        const size_t nl = (size_t)ng * VGL_HC_LANES;
        const uint64_t c0 = b->grp_col_blk[g0], c1 = b->grp_col_blk[g1];
        const uint64_t m0 = b->grp_map_blk[g0], m1 = b->grp_map_blk[g1];
        if (c1 < c0 || m1 < m0 || c1 > b->n_col_blks || m1 > b->n_map_blks) return fail(VGL_ERR_INVALID, "batch block offsets are not monotone");
        const size_t ncol = (size_t)(c1 - c0) * VGL_HC_LANES * VGL_HC_COL_BLK, nmap = (size_t)(m1 - m0) * VGL_HC_LANES * VGL_HC_MAP_BLK;
        vgl_ctx::Slot& sl = ctx->slot[n_chunk & 1];
        if (sl.used) VGL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, sl.consumed, 0));  // the kernel that last read this slot
        if (ng + 1 > sl.grp_col_blk.cap || ncol > sl.col.cap || nmap > sl.map.cap) VGL_CUDA(cudaStreamSynchronize(ctx->stream));
        if (int rc = sl.grp_col_blk.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.grp_map_blk.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.n_cols.reserve(nl)) return rc;
        if (int rc = sl.n_close.reserve(nl)) return rc;
        if (int rc = sl.n_extra.reserve(nl)) return rc;
        if (int rc = sl.mapq.reserve(nl)) return rc;
        if (int rc = sl.col.reserve(ncol)) return rc;
        if (int rc = sl.map.reserve(nmap)) return rc;
        cudaStream_t cs = ctx->copy_stream;
        if (int rc = upload_async(sl.grp_col_blk.p, b->grp_col_blk + g0, ((size_t)ng + 1) * 4, cs)) return rc;
        if (int rc = upload_async(sl.grp_map_blk.p, b->grp_map_blk + g0, ((size_t)ng + 1) * 4, cs)) return rc;
        if (int rc = upload_async(sl.n_cols.p, b->n_cols + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.n_close.p, b->n_close + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.n_extra.p, b->n_extra + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.mapq.p, b->mapq + (size_t)g0 * VGL_HC_LANES, nl, cs)) return rc;
        if (int rc = upload_async(sl.col.p, b->col + (size_t)c0 * VGL_HC_LANES * VGL_HC_COL_BLK, ncol * 4, cs)) return rc;
        if (int rc = upload_async(sl.map.p, b->map + (size_t)m0 * VGL_HC_LANES * VGL_HC_MAP_BLK, nmap * 4, cs)) return rc;
        VGL_CUDA(cudaEventRecord(sl.copied, cs));
        VGL_CUDA(cudaStreamWaitEvent(ctx->stream, sl.copied, 0));
        vgl::HcBatchDev d;
        d.n_groups = ng;
        d.grp_col_blk = sl.grp_col_blk.p;
        d.grp_map_blk = sl.grp_map_blk.p;
        // (uintptr_t arithmetic: the shifted bases are only ever dereferenced at rows >= c0 / m0, i.e. inside the slot)
        d.col = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(sl.col.p) - (uintptr_t)c0 * 512);
        d.map = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(sl.map.p) - (uintptr_t)m0 * 512);
        d.n_cols = sl.n_cols.p;
        d.n_close = sl.n_close.p;
        d.n_extra = sl.n_extra.p;
        d.mapq = sl.mapq.p;
        if (int rc = launch_read(ctx, d, clip_meta(to_meta(b->meta), g0, g1))) return rc;
        VGL_CUDA(cudaEventRecord(sl.consumed, ctx->stream));
        sl.used = true;
    }
    // The caller may reuse its host buffers as soon as this returns.
    VGL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return VGL_OK;
}

uint64_t vgl_hc_raw_bytes(const vgl_hc_raw* r) {
    if (!r) return 0;
    const uint64_t seq_bytes = (r->reserved & VGL_HC_RAW_NARROW_ONLY) ? r->edit_seq_off[1] - r->edit_seq_off[0] : r->edit_seq_off[r->n_edits];
    const uint64_t per_read = (uint64_t)(r->n_reads + 1) * 12 + r->n_reads + r->qual_off[r->n_reads] + seq_bytes;
    const uint64_t lens = (uint64_t)r->n_edits * (r->edit_from8 ? 2 : 4);  // from_length + to_length per edit
    const uint64_t nodes = (uint64_t)r->n_maps * (r->map_node16 ? 2 : 4);
    if (r->first_offset16 && r->map_n_edits8 && r->edit_slen8)  // one offset per read, edit count 1 per mapping, sequence length 1 per edit
        return per_read + nodes + (uint64_t)r->n_reads * 2 + (uint64_t)r->n_maps + (uint64_t)r->n_edits + lens;
    if (r->map_offset16)  // narrow: offset 2 + edit count 1 per mapping; sequence length 1 per edit
        return per_read + nodes + (uint64_t)r->n_maps * 3 + (uint64_t)r->n_edits + lens;
    return per_read + nodes + (uint64_t)r->n_maps * 4 + (uint64_t)(r->n_maps + 1) * 4 + (uint64_t)(r->n_edits + 1) * 4 + lens;
}

int vgl_hc_set_graph(vgl_ctx* ctx, int64_t min_id, int64_t n_nodes, const int64_t* seq_off, const char* seq) {
    if (int rc = use_device(ctx)) return rc;
    // (HaploCart's column record holds node ids <= VGL_HC_MAX_NODE: checked per mapping by the reconstruction kernel)
    if (min_id < 0 || n_nodes <= 0 || !seq_off || !seq || min_id + n_nodes > 0x7fffffffll || seq_off[n_nodes] > 0xffffffffll)
        return fail(VGL_ERR_INVALID, "bad node table");
    std::vector<uint32_t> off((size_t)n_nodes + 1);
    for (int64_t k = 0; k <= n_nodes; ++k) off[(size_t)k] = (uint32_t)seq_off[k];
    if (int rc = ctx->g_seq_off.reserve(off.size())) return rc;
    if (int rc = ctx->g_seq.reserve((size_t)seq_off[n_nodes] + 1)) return rc;
    VGL_CUDA(cudaMemcpy(ctx->g_seq_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->g_seq.p, seq, (size_t)seq_off[n_nodes], cudaMemcpyHostToDevice));
    ctx->g_min_id = (uint32_t)min_id;
    ctx->g_n_nodes = (uint32_t)n_nodes;
    ctx->have_graph = true;
    return VGL_OK;
}

// Lanes of a device-built batch: the reads of one mapping quality take one run of groups (a "slot"; its last group filled up with
// unused lanes), so that no group mixes two mapping qualities -- the segments of vgl_hc_batch_meta, which let the log-term read
// kernel run.  With more than VGL_HC_MAX_SEGS distinct values: one slot, groups by shape only.
struct SlotPlan {
    uint32_t n_slots = 1;
    uint32_t lane_base[VGL_HC_MAX_SEGS] = {0};
    uint8_t slot_of_mapq[128] = {0};
    size_t G = 0;                 // groups of the batch
    vgl::HcBatchMeta segs;        // n_segs / seg_mapq / seg_end (valid and qual_max are the producer's business)
};
static SlotPlan plan_slots(const uint32_t mq_count[128], size_t n) {
    SlotPlan P;
    int n_distinct = 0;
    for (int q = 0; q < 128; ++q) n_distinct += mq_count[q] ? 1 : 0;
    P.G = (n + 31) / 32;
    if (n_distinct >= 1 && n_distinct <= VGL_HC_MAX_SEGS) {
        P.n_slots = 0;
        P.G = 0;
        for (int q = 0; q < 128; ++q) {
            if (!mq_count[q]) continue;
            P.slot_of_mapq[q] = (uint8_t)P.n_slots;
            P.lane_base[P.n_slots] = (uint32_t)(P.G * 32);
            P.G += (mq_count[q] + 31) / 32;
            P.segs.seg_mapq[P.n_slots] = (uint8_t)q;
            P.segs.seg_end[P.n_slots] = (uint32_t)P.G;
            ++P.n_slots;
        }
        P.segs.n_segs = P.n_slots;
    }
    return P;
}

// Where reads [ra, rb) of a raw batch sit in its arrays (a narrow-only view is one chunk: its two wide offset arrays are
// stand-ins that hold only their end points, VGL_HC_RAW_NARROW_ONLY).
struct RawSpan {
    size_t n, ma, nm, ea, ne, qa, nq, sa, nseq, ca, ncap;
};
static RawSpan raw_span(const vgl_hc_raw* raw, size_t ra, size_t rb) {
    const bool narrow_only = (raw->reserved & VGL_HC_RAW_NARROW_ONLY) != 0;
    RawSpan sp;
    sp.n = rb - ra;
    sp.ma = raw->map_off[ra];
    sp.nm = raw->map_off[rb] - sp.ma;
    sp.ea = narrow_only ? raw->edit_off[0] : raw->edit_off[sp.ma];
    sp.ne = (narrow_only ? raw->edit_off[1] : raw->edit_off[sp.ma + sp.nm]) - sp.ea;
    sp.qa = raw->qual_off[ra];
    sp.nq = raw->qual_off[rb] - sp.qa;
    sp.sa = narrow_only ? raw->edit_seq_off[0] : raw->edit_seq_off[sp.ea];
    sp.nseq = (narrow_only ? raw->edit_seq_off[1] : raw->edit_seq_off[sp.ea + sp.ne]) - sp.sa;
    sp.ca = raw->str_off[ra];
    sp.ncap = raw->str_off[rb] - sp.ca;
    return sp;
}

// Device buffers for the input arrays of one span (wide and narrow forms).
static int reserve_raw_inputs(vgl_ctx::RawDev& R, const RawSpan& sp) {
    const size_t n = sp.n, nm = sp.nm, ne = sp.ne;
#define RSV(buf, cnt) if (int rc = (buf).reserve(std::max<size_t>((cnt), 1))) return rc
    RSV(R.qual_off, n + 1); RSV(R.str_off, n + 1); RSV(R.map_off, n + 1); RSV(R.mapq, n); RSV(R.qual, sp.nq);
    RSV(R.map_node, nm); RSV(R.map_offset, nm); RSV(R.edit_off, nm + 1); RSV(R.edit_from, ne); RSV(R.edit_to, ne);
    RSV(R.edit_seq_off, ne + 1); RSV(R.edit_seq, sp.nseq);
    RSV(R.off16, nm); RSV(R.fo16, n); RSV(R.n16, nm); RSV(R.ne8, nm); RSV(R.sl8, ne); RSV(R.f8, ne); RSV(R.t8, ne); RSV(R.tiles, std::max(nm, ne) / 4096 + 2);
#undef RSV
    return VGL_OK;
}

// H2D of the span's input arrays on `copy_stream` (event R.copied), then on ctx->stream the widening of whatever travelled
// in narrow form: 16-bit node ids, one node offset per read, counts and lengths as bytes -> the 32-bit values and offsets the
// reconstruction kernel reads.  The slices keep their absolute indices.
static int stage_raw_inputs(vgl_ctx* ctx, const vgl_hc_raw* raw, size_t ra, const RawSpan& sp, vgl_ctx::RawDev& R, cudaStream_t copy_stream) {
    cudaStream_t s = ctx->stream;
    const size_t n = sp.n, ma = sp.ma, nm = sp.nm, ea = sp.ea, ne = sp.ne, qa = sp.qa, nq = sp.nq, sa = sp.sa, nseq = sp.nseq;
    if (!R.copied) VGL_CUDA(cudaEventCreateWithFlags(&R.copied, cudaEventDisableTiming));
#define H2D(dst, src, bytes) if ((size_t)(bytes) != 0) VGL_CUDA(cudaMemcpyAsync((dst), (src), (size_t)(bytes), cudaMemcpyHostToDevice, copy_stream))
    H2D(R.qual_off.p, raw->qual_off + ra, (n + 1) * 4); H2D(R.str_off.p, raw->str_off + ra, (n + 1) * 4);
    H2D(R.map_off.p, raw->map_off + ra, (n + 1) * 4); H2D(R.mapq.p, raw->mapq + ra, n); H2D(R.qual.p, raw->qual + qa, nq);
    const bool first_only = raw->first_offset16 != nullptr;  // one offset per read instead of one per mapping
    const bool narrow = (raw->map_offset16 || first_only) && raw->map_n_edits8 && raw->edit_slen8;
    const bool narrow_len = raw->edit_from8 && raw->edit_to8;
    if (raw->map_node16) {
        H2D(R.n16.p, raw->map_node16 + ma, nm * 2);
    } else {
        H2D(R.map_node.p, raw->map_node + ma, nm * 4);
    }
    if (narrow_len) {
        H2D(R.f8.p, raw->edit_from8 + ea, ne); H2D(R.t8.p, raw->edit_to8 + ea, ne);
    } else {
        H2D(R.edit_from.p, raw->edit_from + ea, ne * 2); H2D(R.edit_to.p, raw->edit_to + ea, ne * 2);
    }
    H2D(R.edit_seq.p, raw->edit_seq + sa, nseq);
    if (narrow) {
        if (first_only) {
            H2D(R.fo16.p, raw->first_offset16 + ra, n * 2);
        } else {
            H2D(R.off16.p, raw->map_offset16 + ma, nm * 2);
        }
        H2D(R.ne8.p, raw->map_n_edits8 + ma, nm); H2D(R.sl8.p, raw->edit_slen8 + ea, ne);
    } else {
        H2D(R.map_offset.p, raw->map_offset + ma, nm * 4); H2D(R.edit_off.p, raw->edit_off + ma, (nm + 1) * 4);
        H2D(R.edit_seq_off.p, raw->edit_seq_off + ea, (ne + 1) * 4);
    }
#undef H2D
    VGL_CUDA(cudaEventRecord(R.copied, copy_stream));
    if (copy_stream != s) VGL_CUDA(cudaStreamWaitEvent(s, R.copied, 0));
    if (raw->map_node16) {
        VGL_CUDA(vgl::raw_launch_widen_node16(R.n16.p, (uint32_t)nm, R.map_node.p, s));
        ctx->launches += 1;
    }
    if (narrow_len) {
        VGL_CUDA(vgl::raw_launch_widen_u8(R.f8.p, (uint32_t)ne, R.edit_from.p, s));
        VGL_CUDA(vgl::raw_launch_widen_u8(R.t8.p, (uint32_t)ne, R.edit_to.p, s));
        ctx->launches += 2;
    }
    if (narrow) {  // rebuild the 32-bit offsets (absolute indices: the prefix sums start at the chunk's first edit / byte)
        if (first_only) {
            VGL_CUDA(cudaMemsetAsync(R.map_offset.p, 0, nm * sizeof(uint32_t), s));
            VGL_CUDA(vgl::raw_launch_first_offsets(R.map_off.p, R.fo16.p, (uint32_t)n, R.map_offset.p - ma, s));
        } else {
            VGL_CUDA(vgl::raw_launch_widen_u16(R.off16.p, (uint32_t)nm, R.map_offset.p, s));
        }
        VGL_CUDA(vgl::raw_launch_offsets_from_counts(R.ne8.p, (uint32_t)nm, (uint32_t)ea, R.tiles.p, R.edit_off.p, s));
        VGL_CUDA(vgl::raw_launch_offsets_from_counts(R.sl8.p, (uint32_t)ne, (uint32_t)sa, R.tiles.p, R.edit_seq_off.p, s));
        ctx->launches += 7;
    }
    return VGL_OK;
}

// Reads [ra, rb) of `raw`: H2D of their slices on `copy_stream` into staging set `R`, then reconstruction and packing on
// ctx->stream; fills the read kernel's batch descriptor.  The slices keep their absolute indices (offsets into the
// whole batch): the device pointers are shifted back by each slice's first element and only ever dereferenced inside.
static int pack_raw_chunk(vgl_ctx* ctx, const vgl_hc_raw* raw, size_t ra, size_t rb, vgl_ctx::RawDev& R, cudaStream_t copy_stream,
                          vgl::HcBatchDev& d, vgl::HcBatchMeta* seg_out = nullptr) {
    cudaStream_t s = ctx->stream;
    const RawSpan sp = raw_span(raw, ra, rb);
    const size_t n = sp.n, ma = sp.ma, nm = sp.nm, ea = sp.ea, ne = sp.ne, qa = sp.qa, nq = sp.nq, sa = sp.sa, nseq = sp.nseq, ca = sp.ca, ncap = sp.ncap;
    uint32_t mq_count[128];
    memset(mq_count, 0, sizeof(mq_count));
    for (size_t r = ra; r < rb; ++r) mq_count[raw->mapq[r] & 127u] += 1;
    const SlotPlan plan = plan_slots(mq_count, n);
    const uint32_t n_slots = plan.n_slots;
    const size_t G = plan.G;
    if (seg_out) {
        seg_out->n_segs = plan.segs.n_segs;
        memcpy(seg_out->seg_mapq, plan.segs.seg_mapq, sizeof(plan.segs.seg_mapq));
        memcpy(seg_out->seg_end, plan.segs.seg_end, sizeof(plan.segs.seg_end));
    }
    // sorted groups: sum over groups of the longest lane <= (sum of all lanes) / 32 + the longest read (per slot)
    const size_t col_cap = ncap / 128 + (size_t)n_slots * 1024 + G + 1, map_cap = G + 4096;
    if (col_cap > 0x7fffffffull) return fail(VGL_ERR_INVALID, "batch too large; split it");
    if (!R.copied) VGL_CUDA(cudaEventCreateWithFlags(&R.copied, cudaEventDisableTiming));
    if (!R.consumed) VGL_CUDA(cudaEventCreateWithFlags(&R.consumed, cudaEventDisableTiming));
    if (R.used) VGL_CUDA(cudaStreamWaitEvent(copy_stream, R.consumed, 0));  // the kernels that last read this staging set
    // growing a buffer frees it: the kernels still reading it must have finished
    if (n + 1 > R.qual_off.cap || nq > R.qual.cap || nm > R.map_node.cap || ne > R.edit_from.cap || nseq > R.edit_seq.cap ||
        ncap > R.gstr.cap || col_cap * 128 > R.col.cap || map_cap * 128 > R.xmap.cap || G * 32 > R.order.cap || G * 32 > R.o_n_cols.cap ||
        3 * 4096 * (size_t)n_slots > R.hist3.cap || G + 1 > R.grp_cb.cap)
        VGL_CUDA(cudaStreamSynchronize(s));
    if (int rc = reserve_raw_inputs(R, sp)) return rc;
#define RSV(buf, cnt) if (int rc = (buf).reserve(std::max<size_t>((cnt), 1))) return rc
    RSV(R.gstr, ncap); RSV(R.rstr, ncap); RSV(R.sizes, ne); RSV(R.n_cols, n); RSV(R.n_close, n); RSV(R.n_extra, n);
    RSV(R.hist3, 3 * 4096 * (size_t)n_slots); RSV(R.order, G * 32); RSV(R.grp_cb, G + 1); RSV(R.grp_mb, G + 1);
    RSV(R.col, col_cap * 128); RSV(R.xmap, map_cap * 128);
    RSV(R.o_n_cols, G * 32); RSV(R.o_n_close, G * 32); RSV(R.o_n_extra, G * 32); RSV(R.o_mapq, G * 32);
#undef RSV
    R.col_cap_blks = (uint32_t)col_cap;
    R.map_cap_blks = (uint32_t)map_cap;
    d = vgl::HcBatchDev{(uint32_t)G, R.grp_cb.p, R.grp_mb.p, R.col.p, R.xmap.p, R.o_n_cols.p, R.o_n_close.p, R.o_n_extra.p, R.o_mapq.p};
    if (n == 0) return VGL_OK;
    if (int rc = stage_raw_inputs(ctx, raw, ra, sp, R, copy_stream)) return rc;
    VGL_CUDA(cudaMemsetAsync(R.hist3.p, 0, 3 * 4096 * (size_t)n_slots * sizeof(uint32_t), s));
    VGL_CUDA(cudaMemsetAsync(R.order.p, 0xff, G * 32 * sizeof(uint32_t), s));
    VGL_CUDA(cudaMemsetAsync(R.xmap.p, 0, map_cap * 512, s));
    const vgl::ReconGraphDev Gd{ctx->g_min_id, ctx->g_n_nodes, ctx->g_seq_off.p, ctx->g_seq.p};
    // per-read arrays are indexed by the read's position in the chunk; everything they point INTO by absolute index
    const vgl::ReconRawDev Rd{(uint32_t)n, R.qual_off.p, R.qual.p - qa, R.mapq.p, R.str_off.p, R.map_off.p, R.map_node.p - ma,
                              R.map_offset.p - ma, R.edit_off.p - ma, R.edit_from.p - ea, R.edit_to.p - ea, R.edit_seq_off.p - ea,
                              R.edit_seq.p - sa, nullptr, nullptr};
    vgl::ReconWorkDev Wd{R.gstr.p - ca, R.rstr.p - ca, R.sizes.p - ea, R.n_cols.p, R.n_close.p, R.n_extra.p, R.hist3.p,
                         R.hist3.p + 4096 * (size_t)n_slots, R.hist3.p + 8192 * (size_t)n_slots, R.order.p, R.grp_cb.p, R.grp_mb.p};
    Wd.n_slots = n_slots;
    Wd.n_lanes = (uint32_t)(G * 32);
    memcpy(Wd.lane_base, plan.lane_base, sizeof(plan.lane_base));
    memcpy(Wd.slot_of_mapq, plan.slot_of_mapq, sizeof(plan.slot_of_mapq));
    VGL_CUDA(vgl::hc_launch_recon(Rd, Gd, Wd, 0, ctx->cur_err, s));
    VGL_CUDA(vgl::hc_launch_pack(Rd, Wd, R.col_cap_blks, R.map_cap_blks, R.col.p, R.xmap.p, R.o_n_cols.p, R.o_n_close.p, R.o_n_extra.p,
                                 R.o_mapq.p, ctx->cur_err, s));
    ctx->launches += 6;
    return VGL_OK;
}

static int check_raw(vgl_ctx* ctx, const vgl_hc_raw* raw) {
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (!ctx->have_graph) return fail(VGL_ERR_STATE, "vgl_hc_set_graph must come first");
    if (!raw || !raw->qual_off || !raw->str_off || !raw->map_off || !raw->edit_off || !raw->edit_seq_off || !raw->mapq)
        return fail(VGL_ERR_INVALID, "raw alignment arrays missing");
    if (raw->reserved & VGL_HC_RAW_NARROW_ONLY) {
        if (!(raw->map_node16 || raw->map_node) || !raw->first_offset16 || !raw->map_n_edits8 || !raw->edit_slen8 || !raw->edit_from8 || !raw->edit_to8)
            return fail(VGL_ERR_INVALID, "a narrow-only raw batch needs the narrow arrays (node ids as 16 or 32 bits)");
        if ((size_t)raw->n_reads > (size_t)ctx->chunk_groups * VGL_HC_LANES) return fail(VGL_ERR_INVALID, "a narrow-only raw batch holds at most one chunk");
    } else if ((raw->n_maps && (!raw->map_node || !raw->map_offset)) || (raw->n_edits && (!raw->edit_from || !raw->edit_to))) {
        return fail(VGL_ERR_INVALID, "raw alignment arrays missing");
    }
    return VGL_OK;
}

int vgl_hc_pack_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, vgl_hc_batch* out) {
    if (int rc = use_device(ctx)) return rc;
    if (!out) return fail(VGL_ERR_INVALID, "null output");
    if (int rc = check_raw(ctx, raw)) return rc;
    vgl::HcBatchDev d;
    vgl::HcBatchMeta meta = to_meta(raw->meta);
    if (int rc = pack_raw_chunk(ctx, raw, 0, raw->n_reads, ctx->rawdev[0], ctx->stream, d, &meta)) return rc;
    memset(out, 0, sizeof(*out));
    out->meta = raw->meta;   // value ranges as the producer gave them; the segments are those of the device-built batch
    out->meta.n_segs = meta.n_segs;
    for (uint32_t k = 0; k < VGL_HC_MAX_SEGS; ++k) { out->meta.seg_mapq[k] = meta.seg_mapq[k]; out->meta.seg_end[k] = meta.seg_end[k]; }
    out->n_reads = raw->n_reads;
    out->n_groups = d.n_groups;
    out->grp_col_blk = d.grp_col_blk; out->grp_map_blk = d.grp_map_blk; out->col = d.col; out->map = d.map;
    out->n_cols = d.n_cols; out->n_close = d.n_close; out->n_extra = d.n_extra; out->mapq = d.mapq;
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (d.n_groups) {
        uint32_t tot[2] = {0, 0};
        VGL_CUDA(cudaMemcpy(&tot[0], d.grp_col_blk + d.n_groups, 4, cudaMemcpyDeviceToHost));
        VGL_CUDA(cudaMemcpy(&tot[1], d.grp_map_blk + d.n_groups, 4, cudaMemcpyDeviceToHost));
        out->n_col_blks = tot[0];
        out->n_map_blks = tot[1];
    }
    return VGL_OK;
}

int vgl_hc_accumulate_raw(vgl_ctx* ctx, const vgl_hc_raw* raw) {
    if (int rc = use_device(ctx)) return rc;
    if (int rc = check_raw(ctx, raw)) return rc;
    // Chunks of reads through two staging sets: the H2D copy of a chunk (copy stream) overlaps the reconstruction, packing
    // and read kernel of the chunk before it.
    // (the staging sets alternate across calls too: a caller that submits chunk-sized batches -- vgh_hc_feed -- gets the same overlap)
    const size_t n = raw->n_reads, step = (size_t)ctx->chunk_groups * VGL_HC_LANES;
    for (size_t ra = 0; ra < n; ra += step) {
        const size_t rb = std::min(n, ra + step);
        vgl_ctx::RawDev& R = ctx->rawdev[ctx->raw_seq++ & 1];
        vgl::HcBatchDev d;
        vgl::HcBatchMeta meta = to_meta(raw->meta);   // value ranges from the producer; the segments are this chunk's own
        if (int rc = pack_raw_chunk(ctx, raw, ra, rb, R, ctx->copy_stream, d, &meta)) return rc;
        if (d.n_groups)
            if (int rc = launch_read(ctx, d, meta)) return rc;
        VGL_CUDA(cudaEventRecord(R.consumed, ctx->stream));
        R.used = true;
    }
    // The caller may reuse its host buffers as soon as this returns (the kernels run on).
    VGL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return VGL_OK;
}

uint64_t vgl_hc_wire_bytes(const vgl_hc_wire* w) {
    if (!w) return 0;
    const bool narrow = w->col8 && w->flag4 && w->n_owned;
    const uint64_t cols = narrow ? w->n_col_blks * (256 + 128) + (uint64_t)w->n_groups * VGL_HC_LANES * 2 : w->n_col_blks * 512;
    return (uint64_t)(w->n_groups + 1) * 12 + cols + (w->n_node_blks + w->n_map_blks) * 512 + (uint64_t)w->n_groups * VGL_HC_LANES * 7;
}

int vgl_hc_accumulate_wire(vgl_ctx* ctx, const vgl_hc_wire* w) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (!w) return fail(VGL_ERR_INVALID, "null batch");
    if (w->n_groups == 0) return VGL_OK;
    if ((uint64_t)w->n_groups * VGL_HC_LANES < w->n_reads) return fail(VGL_ERR_INVALID, "n_groups too small for n_reads");
    if (!w->grp_col_blk || !w->grp_node_blk || !w->grp_map_blk || !w->n_cols || !w->n_close || !w->n_extra || !w->mapq ||
        (w->n_node_blks && !w->node) || (w->n_map_blks && !w->map))
        return fail(VGL_ERR_INVALID, "wire batch arrays missing");
    const bool narrow = w->col8 && w->flag4 && w->n_owned;  // 12-bit columns; otherwise the 16-bit records
    if (!narrow && (w->col8 || w->flag4 || w->n_owned)) return fail(VGL_ERR_INVALID, "col8, flag4 and n_owned come together");
    if (!narrow && w->n_col_blks && !w->col) return fail(VGL_ERR_INVALID, "wire batch arrays missing");
    // As vgl_hc_accumulate: chunks of groups, two staging slots, the H2D copy of a chunk overlaps the kernels of the one
    // before.  Each chunk is expanded into 32-bit records on the device before the read kernel runs on it.
    const uint32_t G = w->n_groups;
    uint32_t n_chunk = 0;
    for (uint32_t g0 = 0; g0 < G; g0 += ctx->chunk_groups, ++n_chunk) {
        const uint32_t g1 = std::min(G, g0 + ctx->chunk_groups);
        const uint32_t ng = g1 - g0;
        const size_t nl = (size_t)ng * VGL_HC_LANES;
        const uint64_t c0 = w->grp_col_blk[g0], c1 = w->grp_col_blk[g1];
        const uint64_t n0 = w->grp_node_blk[g0], n1 = w->grp_node_blk[g1];
        const uint64_t m0 = w->grp_map_blk[g0], m1 = w->grp_map_blk[g1];
        if (c1 < c0 || n1 < n0 || m1 < m0 || c1 > w->n_col_blks || n1 > w->n_node_blks || m1 > w->n_map_blks)
            return fail(VGL_ERR_INVALID, "batch block offsets are not monotone");
        const size_t ncol16 = (size_t)(c1 - c0) * 256, nnode16 = (size_t)(n1 - n0) * 256, nmap = (size_t)(m1 - m0) * 128;
        const size_t ncol32 = (size_t)(c1 - c0) * 2 * 128;
        vgl_ctx::Slot& sl = ctx->slot[n_chunk & 1];
        if (sl.used) VGL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, sl.consumed, 0));  // the kernels that last read this slot
        if (ng + 1 > sl.grp_c16.cap || (!narrow && ncol16 > sl.col16.cap) || (narrow && (ncol16 > sl.col8.cap || ncol16 / 8 > sl.flag4.cap || nl > sl.n_owned.cap)) ||
            nnode16 > sl.node16.cap || ncol32 > sl.col.cap || nmap > sl.map.cap || ng + 1 > sl.grp_col_blk.cap)
            VGL_CUDA(cudaStreamSynchronize(ctx->stream));
        if (int rc = sl.grp_c16.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.grp_n16.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.grp_col_blk.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.grp_map_blk.reserve((size_t)ng + 1)) return rc;
        if (int rc = sl.n_cols.reserve(nl)) return rc;
        if (int rc = sl.n_close.reserve(nl)) return rc;
        if (int rc = sl.n_extra.reserve(nl)) return rc;
        if (int rc = sl.mapq.reserve(nl)) return rc;
        if (narrow) {
            if (int rc = sl.col8.reserve(std::max<size_t>(ncol16, 1))) return rc;
            if (int rc = sl.flag4.reserve(std::max<size_t>(ncol16 / 8, 1))) return rc;
            if (int rc = sl.n_owned.reserve(nl)) return rc;
        } else {
            if (int rc = sl.col16.reserve(std::max<size_t>(ncol16, 1))) return rc;
        }
        if (int rc = sl.node16.reserve(std::max<size_t>(nnode16, 1))) return rc;
        if (int rc = sl.col.reserve(std::max<size_t>(ncol32, 1))) return rc;
        if (int rc = sl.map.reserve(std::max<size_t>(nmap, 1))) return rc;
        cudaStream_t cs = ctx->copy_stream;
        if (int rc = upload_async(sl.grp_c16.p, w->grp_col_blk + g0, ((size_t)ng + 1) * 4, cs)) return rc;
        if (int rc = upload_async(sl.grp_n16.p, w->grp_node_blk + g0, ((size_t)ng + 1) * 4, cs)) return rc;
        if (int rc = upload_async(sl.grp_map_blk.p, w->grp_map_blk + g0, ((size_t)ng + 1) * 4, cs)) return rc;
        if (int rc = upload_async(sl.n_cols.p, w->n_cols + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.n_close.p, w->n_close + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.n_extra.p, w->n_extra + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        if (int rc = upload_async(sl.mapq.p, w->mapq + (size_t)g0 * VGL_HC_LANES, nl, cs)) return rc;
        if (narrow) {
            if (int rc = upload_async(sl.col8.p, w->col8 + (size_t)c0 * 256, ncol16, cs)) return rc;
            if (int rc = upload_async(sl.flag4.p, w->flag4 + (size_t)c0 * 32, ncol16 / 8 * 4, cs)) return rc;
            if (int rc = upload_async(sl.n_owned.p, w->n_owned + (size_t)g0 * VGL_HC_LANES, nl * 2, cs)) return rc;
        } else {
            if (int rc = upload_async(sl.col16.p, w->col + (size_t)c0 * 256, ncol16 * 2, cs)) return rc;
        }
        if (int rc = upload_async(sl.node16.p, w->node + (size_t)n0 * 256, nnode16 * 2, cs)) return rc;
        if (int rc = upload_async(sl.map.p, w->map + (size_t)m0 * 128, nmap * 4, cs)) return rc;
        VGL_CUDA(cudaEventRecord(sl.copied, cs));
        VGL_CUDA(cudaStreamWaitEvent(ctx->stream, sl.copied, 0));
        // the 16-bit arrays keep their absolute block indices: bases shifted back by the chunk's first block (only ever
        // dereferenced at blocks >= c0 / n0 / m0, i.e. inside the slot)
        const uint16_t* col16 = narrow ? nullptr : reinterpret_cast<const uint16_t*>(reinterpret_cast<uintptr_t>(sl.col16.p) - (uintptr_t)c0 * 512);
        const uint8_t* col8 = narrow ? reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(sl.col8.p) - (uintptr_t)c0 * 256) : nullptr;
        const uint32_t* flag4 = narrow ? reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(sl.flag4.p) - (uintptr_t)c0 * 128) : nullptr;
        const uint16_t* node16 = reinterpret_cast<const uint16_t*>(reinterpret_cast<uintptr_t>(sl.node16.p) - (uintptr_t)n0 * 512);
        VGL_CUDA(vgl::hc_launch_expand(ng, sl.grp_c16.p, sl.grp_n16.p, col16, col8, flag4, narrow ? sl.n_owned.p : nullptr, node16, sl.n_cols.p,
                                       sl.col.p, sl.grp_col_blk.p, ctx->stream));
        ctx->launches += 1;
        vgl::HcBatchDev d;
        d.n_groups = ng;
        d.grp_col_blk = sl.grp_col_blk.p;  // chunk relative (written by the expansion)
        d.grp_map_blk = sl.grp_map_blk.p;
        d.col = sl.col.p;
        d.map = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(sl.map.p) - (uintptr_t)m0 * 512);
        d.n_cols = sl.n_cols.p;
        d.n_close = sl.n_close.p;
        d.n_extra = sl.n_extra.p;
        d.mapq = sl.mapq.p;
        if (int rc = launch_read(ctx, d, clip_meta(to_meta(w->meta), g0, g1))) return rc;
        VGL_CUDA(cudaEventRecord(sl.consumed, ctx->stream));
        sl.used = true;
    }
    // The caller may reuse its host buffers as soon as this returns.
    VGL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return VGL_OK;
}

// -------------------------------------------------------------------------------------------
// several samples per pass (src/HaploCart.cpp:303 loops over samples; a real sample is <= ~2 x 10^4 reads after duplicate
// removal, microseconds of GPU work, so the per-sample fixed costs -- the 7.7 MB bit matrix of K2, K3's launch, the result copy
// and the host synchronisation -- are paid once per pass instead)
// -------------------------------------------------------------------------------------------
// -------------------------------------------------------------------------------------------
// GAM wire decode on the device: framing from the host (vgh_gam_frame), messages parsed on the GPU straight
// into the raw-alignment arrays, then the same reconstruction + packing + read kernel as the raw route
// -------------------------------------------------------------------------------------------
namespace {
struct GamScratch {
    DevBuf<uint8_t> bytes, keep;
    DevBuf<unsigned long long> off;
    DevBuf<uint32_t> len, counts, small;   // small: totals[6] | mapq_hist[128] | qual_max
};
GamScratch g_gam[VGL_HC_MAX_PEERS];   // one per device (a context decodes on its own device)

// Decodes messages [c0, c1) into staging set R; returns the kept reads and their value ranges.
int gam_decode_chunk(vgl_ctx* ctx, const uint8_t* bytes, const uint64_t* msg_off, const uint32_t* msg_len, size_t c0, size_t c1,
                     vgl_ctx::RawDev& R, bool want_index, uint32_t totals[6], uint32_t mq_count[128], uint32_t* qual_max) {
    cudaStream_t s = ctx->stream;
    GamScratch& S = g_gam[ctx->device % VGL_HC_MAX_PEERS];
    const size_t n = c1 - c0;
    const uint64_t b0 = msg_off[c0], b1 = msg_off[c1 - 1] + msg_len[c1 - 1];
    VGL_CUDA(cudaStreamSynchronize(s));   // the scratch and the staging set are about to be rewritten
    if (int rc = S.bytes.reserve((size_t)(b1 - b0) + 16)) return rc;
    if (int rc = S.off.reserve(n)) return rc;
    if (int rc = S.len.reserve(n)) return rc;
    if (int rc = S.keep.reserve(n)) return rc;
    if (int rc = S.counts.reserve(6 * n)) return rc;
    if (int rc = S.small.reserve(6 + 128 + 1)) return rc;
    VGL_CUDA(cudaMemcpyAsync(S.bytes.p, bytes + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemcpyAsync(S.off.p, msg_off + c0, n * 8, cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemcpyAsync(S.len.p, msg_len + c0, n * 4, cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemsetAsync(S.small.p, 0, (6 + 128 + 1) * sizeof(uint32_t), s));
    VGL_CUDA(vgl::gam_launch_count(S.bytes.p, reinterpret_cast<const uint64_t*>(S.off.p), S.len.p, (uint32_t)n, b0, S.counts.p, S.keep.p,
                                   S.small.p + 6, S.small.p + 6 + 128, S.small.p, ctx->cur_err, s));
    ctx->launches += 2;
    uint32_t host_small[6 + 128 + 1];
    VGL_CUDA(cudaMemcpyAsync(host_small, S.small.p, sizeof(host_small), cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    memcpy(totals, host_small, 6 * sizeof(uint32_t));
    memcpy(mq_count, host_small + 6, 128 * sizeof(uint32_t));
    *qual_max = host_small[6 + 128];
    const size_t nr = totals[0], nm = totals[1], ne = totals[2], nq = totals[3], nseq = totals[4];
#define RSV(buf, cnt) if (int rc = (buf).reserve(std::max<size_t>((cnt), 1))) return rc
    RSV(R.qual_off, nr + 1); RSV(R.str_off, nr + 1); RSV(R.map_off, nr + 1); RSV(R.mapq, nr); RSV(R.qual, nq);
    RSV(R.map_node, nm); RSV(R.map_offset, nm); RSV(R.edit_off, nm + 1); RSV(R.edit_from, ne); RSV(R.edit_to, ne);
    RSV(R.edit_seq_off, ne + 1); RSV(R.edit_seq, nseq); RSV(R.seq_len, nr); RSV(R.read_index, nr);
#undef RSV
    vgl::GamOutDev o{R.qual_off.p, R.qual.p, R.mapq.p, R.str_off.p, R.map_off.p, R.map_node.p, R.map_offset.p, R.edit_off.p, R.edit_from.p,
                     R.edit_to.p, R.edit_seq_off.p, R.edit_seq.p, R.seq_len.p, want_index ? R.read_index.p : nullptr};
    VGL_CUDA(vgl::gam_launch_fill(S.bytes.p, reinterpret_cast<const uint64_t*>(S.off.p), S.len.p, (uint32_t)n, b0, S.counts.p, S.keep.p,
                                  S.small.p, o, ctx->cur_err, s));
    ctx->launches += 1;
    return VGL_OK;
}
}  // namespace

extern "C" int vgl_hc_accumulate_gam(vgl_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* msg_off, const uint32_t* msg_len,
                                     uint64_t n_msgs, uint64_t* n_reads_out) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (!ctx->have_graph) return fail(VGL_ERR_STATE, "vgl_hc_set_graph must come first");
    if (n_msgs && (!bytes || !msg_off || !msg_len)) return fail(VGL_ERR_INVALID, "null arrays");
    if (n_reads_out) *n_reads_out = 0;
    for (uint64_t k = 0; k < n_msgs; ++k)
        if (msg_off[k] + msg_len[k] > n_bytes || (k && msg_off[k] < msg_off[k - 1])) return fail(VGL_ERR_INVALID, "message %llu lies outside the byte buffer", (unsigned long long)k);
    cudaStream_t s = ctx->stream;
    const size_t step = (size_t)ctx->chunk_groups * VGL_HC_LANES;
    for (size_t c0 = 0; c0 < n_msgs; c0 += step) {
        const size_t c1 = std::min<size_t>(n_msgs, c0 + step);
        vgl_ctx::RawDev& R = ctx->rawdev[0];
        uint32_t totals[6], mq_count[128], qmax = 0;
        if (int rc = gam_decode_chunk(ctx, bytes, msg_off, msg_len, c0, c1, R, false, totals, mq_count, &qmax)) return rc;
        const size_t n = totals[0], ne = totals[2], ncap = totals[5];
        if (n_reads_out) *n_reads_out += n;
        if (n == 0) continue;
        const SlotPlan plan = plan_slots(mq_count, n);
        const size_t G = plan.G, col_cap = ncap / 128 + (size_t)plan.n_slots * 1024 + G + 1, map_cap = G + 4096;
#define RSV(buf, cnt) if (int rc = (buf).reserve(std::max<size_t>((cnt), 1))) return rc
        RSV(R.gstr, ncap); RSV(R.rstr, ncap); RSV(R.sizes, ne); RSV(R.n_cols, n); RSV(R.n_close, n); RSV(R.n_extra, n);
        RSV(R.hist3, 3 * 4096 * (size_t)plan.n_slots); RSV(R.order, G * 32); RSV(R.grp_cb, G + 1); RSV(R.grp_mb, G + 1);
        RSV(R.col, col_cap * 128); RSV(R.xmap, map_cap * 128);
        RSV(R.o_n_cols, G * 32); RSV(R.o_n_close, G * 32); RSV(R.o_n_extra, G * 32); RSV(R.o_mapq, G * 32);
#undef RSV
        R.col_cap_blks = (uint32_t)col_cap;
        R.map_cap_blks = (uint32_t)map_cap;
        VGL_CUDA(cudaMemsetAsync(R.hist3.p, 0, 3 * 4096 * (size_t)plan.n_slots * sizeof(uint32_t), s));
        VGL_CUDA(cudaMemsetAsync(R.order.p, 0xff, G * 32 * sizeof(uint32_t), s));
        VGL_CUDA(cudaMemsetAsync(R.xmap.p, 0, map_cap * 512, s));
        const vgl::ReconGraphDev Gd{ctx->g_min_id, ctx->g_n_nodes, ctx->g_seq_off.p, ctx->g_seq.p};
        const vgl::ReconRawDev Rd{(uint32_t)n, R.qual_off.p, R.qual.p, R.mapq.p, R.str_off.p, R.map_off.p, R.map_node.p, R.map_offset.p, R.edit_off.p,
                                  R.edit_from.p, R.edit_to.p, R.edit_seq_off.p, R.edit_seq.p, nullptr, nullptr};
        vgl::ReconWorkDev Wd{R.gstr.p, R.rstr.p, R.sizes.p, R.n_cols.p, R.n_close.p, R.n_extra.p, R.hist3.p,
                             R.hist3.p + 4096 * (size_t)plan.n_slots, R.hist3.p + 8192 * (size_t)plan.n_slots, R.order.p, R.grp_cb.p, R.grp_mb.p};
        Wd.n_slots = plan.n_slots;
        Wd.n_lanes = (uint32_t)(G * 32);
        memcpy(Wd.lane_base, plan.lane_base, sizeof(plan.lane_base));
        memcpy(Wd.slot_of_mapq, plan.slot_of_mapq, sizeof(plan.slot_of_mapq));
        VGL_CUDA(vgl::hc_launch_recon(Rd, Gd, Wd, 0, ctx->cur_err, s));
        VGL_CUDA(vgl::hc_launch_pack(Rd, Wd, R.col_cap_blks, R.map_cap_blks, R.col.p, R.xmap.p, R.o_n_cols.p, R.o_n_close.p, R.o_n_extra.p, R.o_mapq.p,
                                     ctx->cur_err, s));
        ctx->launches += 6;
        vgl::HcBatchDev d{(uint32_t)G, R.grp_cb.p, R.grp_mb.p, R.col.p, R.xmap.p, R.o_n_cols.p, R.o_n_close.p, R.o_n_extra.p, R.o_mapq.p};
        vgl::HcBatchMeta meta = plan.segs;
        meta.valid = true;
        meta.qual_max = qmax;
        if (int rc = launch_read(ctx, d, meta)) return rc;
        R.used = false;   // (synchronised at the next chunk's start: no event needed)
    }
    return VGL_OK;
}

// The decoded alignments of a (small) GAM, copied back in the flat layout, for tests: every array as the raw route sees it.
extern "C" int vgl_gam_decode(vgl_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint64_t n_msgs,
                              uint32_t* totals6, uint32_t* qual_off, uint8_t* qual, uint8_t* mapq, uint32_t* map_off, uint32_t* map_node,
                              uint32_t* map_offset, uint32_t* edit_off, uint16_t* edit_from, uint16_t* edit_to, uint32_t* edit_seq_off,
                              uint8_t* edit_seq, uint16_t* seq_len, uint32_t* read_index, uint32_t* str_off) {
    if (int rc = use_device(ctx)) return rc;
    if (!totals6) return fail(VGL_ERR_INVALID, "null totals");
    memset(totals6, 0, 6 * sizeof(uint32_t));
    if (n_msgs == 0) return VGL_OK;
    if (n_msgs > (uint64_t)ctx->chunk_groups * VGL_HC_LANES) return fail(VGL_ERR_INVALID, "vgl_gam_decode takes at most one chunk of messages");
    for (uint64_t k = 0; k < n_msgs; ++k)
        if (msg_off[k] + msg_len[k] > n_bytes) return fail(VGL_ERR_INVALID, "message %llu lies outside the byte buffer", (unsigned long long)k);
    vgl_ctx::RawDev& R = ctx->rawdev[0];
    uint32_t mq_count[128], qmax = 0;
    int zero = 0;
    // (standalone use: errors go to a scratch word of the staging set so that no sample is needed)
    if (int rc = R.s_err.reserve(1)) return rc;
    int* saved = ctx->cur_err;
    ctx->cur_err = R.s_err.p;
    cudaMemcpy(R.s_err.p, &zero, sizeof(int), cudaMemcpyHostToDevice);
    const int rc = gam_decode_chunk(ctx, bytes, msg_off, msg_len, 0, (size_t)n_msgs, R, true, totals6, mq_count, &qmax);
    ctx->cur_err = saved;
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    const size_t n = totals6[0], nm = totals6[1], ne = totals6[2], nq = totals6[3], ns = totals6[4];
    int err = 0;
#define D2H(dst, src, bytes_) if ((dst) != nullptr && (size_t)(bytes_) != 0) VGL_CUDA(cudaMemcpyAsync((dst), (src), (bytes_), cudaMemcpyDeviceToHost, s))
    D2H(qual_off, R.qual_off.p, (n + 1) * 4); D2H(qual, R.qual.p, nq); D2H(mapq, R.mapq.p, n); D2H(map_off, R.map_off.p, (n + 1) * 4);
    D2H(map_node, R.map_node.p, nm * 4); D2H(map_offset, R.map_offset.p, nm * 4); D2H(edit_off, R.edit_off.p, (nm + 1) * 4);
    D2H(edit_from, R.edit_from.p, ne * 2); D2H(edit_to, R.edit_to.p, ne * 2); D2H(edit_seq_off, R.edit_seq_off.p, (ne + 1) * 4);
    D2H(edit_seq, R.edit_seq.p, ns); D2H(seq_len, R.seq_len.p, n * 2); D2H(read_index, R.read_index.p, n * 4); D2H(str_off, R.str_off.p, (n + 1) * 4);
#undef D2H
    VGL_CUDA(cudaMemcpyAsync(&err, R.s_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    if (err) return fail(VGL_ERR_BAD_READ, "malformed vg.Alignment message (or values that do not fit the raw-alignment arrays)");
    return VGL_OK;
}

// -------------------------------------------------------------------------------------------
// duplicate removal on the device (the O(n^2) scan in front of the hot loop, src/HaploCart.cpp:390, src/rmdup.cpp:20-110)
// -------------------------------------------------------------------------------------------
int vgl_dedup(vgl_ctx* ctx, uint32_t n_reads, const uint64_t* start_key, const uint8_t* paired, uint8_t* is_dup, uint64_t* n_kept) {
    if (int rc = use_device(ctx)) return rc;
    if (n_reads && (!start_key || !is_dup)) return fail(VGL_ERR_INVALID, "null arrays");
    if (n_kept) *n_kept = 0;
    if (n_reads == 0) return VGL_OK;
    if (n_reads > 0x40000000u) return fail(VGL_ERR_INVALID, "too many reads for one call");
    cudaStream_t s = ctx->stream;
    uint32_t size = 1024;
    while (size < 2u * n_reads) size <<= 1;
    DevBuf<unsigned long long>&d_key = ctx->dd_key, &t_key = ctx->dd_tab_key;
    DevBuf<uint32_t>& t_first = ctx->dd_tab_first;
    DevBuf<uint8_t>&d_paired = ctx->dd_paired, &d_dup = ctx->dd_dup;
    DevBuf<unsigned int>& d_n = ctx->dd_n;
    int rc = VGL_OK;
    auto done = [&](int code) { return code; };
    if (n_reads > d_key.cap || size > t_key.cap) VGL_CUDA(cudaStreamSynchronize(s));   // growing a buffer frees it
    if ((rc = d_key.reserve(n_reads)) || (rc = t_key.reserve(size)) || (rc = t_first.reserve(size)) || (rc = d_dup.reserve(n_reads)) ||
        (rc = d_n.reserve(1)) || (paired && (rc = d_paired.reserve(n_reads))))
        return done(rc);
#define DD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(VGL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    DD(cudaMemcpyAsync(d_key.p, start_key, (size_t)n_reads * 8, cudaMemcpyHostToDevice, s));
    if (paired) DD(cudaMemcpyAsync(d_paired.p, paired, n_reads, cudaMemcpyHostToDevice, s));
    DD(cudaMemsetAsync(t_key.p, 0xff, (size_t)size * 8, s));
    DD(cudaMemsetAsync(t_first.p, 0xff, (size_t)size * 4, s));
    DD(cudaMemsetAsync(d_n.p, 0, sizeof(unsigned int), s));
    DD(vgl::dedup_launch(n_reads, d_key.p, paired ? d_paired.p : nullptr, size, t_key.p, t_first.p, d_dup.p, d_n.p, s));
    ctx->launches += 2;
    unsigned int n_dup = 0;
    DD(cudaMemcpyAsync(is_dup, d_dup.p, n_reads, cudaMemcpyDeviceToHost, s));
    DD(cudaMemcpyAsync(&n_dup, d_n.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    DD(cudaStreamSynchronize(s));
#undef DD
    if (n_kept) *n_kept = (uint64_t)n_reads - n_dup;
    return done(VGL_OK);
}

int vgl_hc_samples_begin(vgl_ctx* ctx, int32_t n_samples) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (n_samples < 1 || n_samples > 4096) return fail(VGL_ERR_INVALID, "1 <= n_samples <= 4096");
    cudaStream_t s = ctx->stream;
    const size_t set = ctx->acc_set_doubles(), P = (size_t)ctx->tab.n_paths, tiles = (size_t)vgl::hc_path_tiles((int)P);
    const size_t cap = (size_t)(ctx->have_tree ? ctx->max_anc + 1 : 1), blk = 4 + cap;
    if ((size_t)n_samples * set > ctx->m_acc.cap || (size_t)n_samples * P > ctx->m_ll.cap) VGL_CUDA(cudaStreamSynchronize(s));
    if (int rc = ctx->m_acc.reserve((size_t)n_samples * set)) return rc;
    if (int rc = ctx->m_node_sum.reserve((size_t)n_samples * vgl::hc_node_sum_doubles(ctx->tab.n_words))) return rc;
    if (int rc = ctx->m_chunk_ll.reserve((size_t)n_samples * (size_t)ctx->n_chunks * P)) return rc;
    if (int rc = ctx->m_ll.reserve((size_t)n_samples * P)) return rc;
    if (int rc = ctx->m_result.reserve((size_t)n_samples * blk)) return rc;
    const size_t old_tiles = ctx->m_tile_done.cap;
    if (int rc = ctx->m_tile_done.reserve((size_t)n_samples * tiles)) return rc;
    if (ctx->m_tile_done.cap != old_tiles) VGL_CUDA(cudaMemsetAsync(ctx->m_tile_done.p, 0, ctx->m_tile_done.cap * sizeof(unsigned int), s));
    if ((size_t)n_samples * blk > ctx->m_h_result_cap) {
        if (ctx->m_h_result) cudaFreeHost(ctx->m_h_result);
        ctx->m_h_result = nullptr;
        ctx->m_h_result_cap = 0;
        VGL_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->m_h_result), (size_t)n_samples * blk * sizeof(double)));
        ctx->m_h_result_cap = (size_t)n_samples * blk;
    }
    VGL_CUDA(cudaMemsetAsync(ctx->m_acc.p, 0, (size_t)n_samples * set * sizeof(double), s));
    ctx->m_samples = n_samples;
    ctx->ev_used = 0;
    ctx->ll_valid = false;
    return vgl_hc_sample_select(ctx, 0);
}

int vgl_hc_sample_select(vgl_ctx* ctx, int32_t sample) {
    if (!ctx) return fail(VGL_ERR_INVALID, "null context");
    if (ctx->m_samples < 1) return fail(VGL_ERR_STATE, "vgl_hc_samples_begin must come first");
    if (sample < 0 || sample >= ctx->m_samples) return fail(VGL_ERR_INVALID, "sample %d outside [0, %d)", sample, ctx->m_samples);
    const size_t set = ctx->acc_set_doubles();
    ctx->cur_acc = ctx->m_acc.p + (size_t)sample * set;
    ctx->cur_a_total = ctx->cur_acc + (size_t)VGL_HC_ACC_COPIES * (size_t)ctx->n_rows_pad;
    ctx->cur_err = reinterpret_cast<int*>(ctx->cur_a_total + VGL_HC_ACC_COPIES);
    ctx->cur_sample = sample;
    return VGL_OK;
}

int vgl_hc_finalize_samples(vgl_ctx* ctx, double* ll_out, int32_t* argmax_out, double* total_out, double* post_out, int32_t post_cap,
                            int32_t* n_post_out, int32_t* status_out) {
    if (int rc = use_device(ctx)) return rc;
    if (ctx->m_samples < 1) return fail(VGL_ERR_STATE, "vgl_hc_samples_begin must come first");
    cudaStream_t s = ctx->stream;
    const int S = ctx->m_samples, P = ctx->tab.n_paths;
    const size_t set = ctx->acc_set_doubles();
    const int cap = ctx->have_tree ? ctx->max_anc + 1 : 1;
    const size_t blk = 4 + (size_t)cap;
    VGL_CUDA(vgl::hc_launch_path_samples_kernel(ctx->tab, S, ctx->m_acc.p, set, ctx->m_node_sum.p, ctx->m_chunk_ll.p, ctx->m_ll.p, ctx->m_tile_done.p, s));
    const int* err0 = reinterpret_cast<const int*>(ctx->m_acc.p + (size_t)VGL_HC_ACC_COPIES * (size_t)ctx->n_rows_pad + VGL_HC_ACC_COPIES);
    VGL_CUDA(vgl::hc_launch_posterior_samples(S, ctx->m_ll.p, P, ctx->row_words, ctx->have_tree ? ctx->d_clade_bits.p : nullptr,
                                              ctx->have_tree ? ctx->d_anc_off.p : nullptr, ctx->have_tree ? ctx->d_anc_clade.p : nullptr,
                                              ctx->m_result.p, cap, blk, err0, set * 2, s));
    ctx->launches += 3;
    VGL_CUDA(cudaMemcpyAsync(ctx->m_h_result, ctx->m_result.p, (size_t)S * blk * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ll_out) VGL_CUDA(cudaMemcpyAsync(ll_out, ctx->m_ll.p, (size_t)S * (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    int worst = VGL_OK;
    for (int k = 0; k < S; ++k) {
        const double* r = ctx->m_h_result + (size_t)k * blk;
        const int32_t* idx = reinterpret_cast<const int32_t*>(r + 2);
        const int err = idx[2];
        int st = VGL_OK;
        if (err & 0x04000000) st = VGL_ERR_INVALID;   // (device-built batch out of room: nothing else the flags say counts)
        else if (err & 0x30000000) st = (err & 0x20000000) ? VGL_ERR_BAD_READ : VGL_ERR_BAD_NODE;
        else if (err & VGL_ERR_BAD_NODE) st = VGL_ERR_BAD_NODE;
        else if (err & 0x08000000) st = VGL_ERR_INVALID;
        if (status_out) status_out[k] = st;
        if (st != VGL_OK && worst == VGL_OK) worst = st;
        if (argmax_out) argmax_out[k] = idx[0];
        if (total_out) total_out[k] = r[0];
        const int n_post = idx[1];
        if (n_post_out) n_post_out[k] = n_post;
        if (post_out)
            for (int j = 0; j < std::min(n_post, (int)post_cap); ++j) post_out[(size_t)k * (size_t)post_cap + j] = r[4 + j];
    }
    // back to single-sample mode
    ctx->m_samples = 0;
    ctx->cur_sample = -1;
    ctx->cur_acc = ctx->d_cta_node_acc.p; ctx->cur_a_total = ctx->d_a_total; ctx->cur_err = ctx->d_err;
    if (worst != VGL_OK && !status_out) return fail(worst, "a sample of the pass holds a read the reference could not process (malformed alignment or unknown node)");
    return VGL_OK;
}

int vgl_hc_paths(vgl_ctx* ctx, double* ll_dev) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    cudaStream_t s = ctx->stream;
    VGL_CUDA(vgl::hc_launch_path_kernel(ctx->tab, ctx->d_cta_node_acc.p, ctx->d_node_sum.p, ctx->d_chunk_ll.p, ctx->d_ll.p, ctx->d_tile_done.p, s));
    ctx->launches += 2;
    ctx->ll_valid = true;
    if (ll_dev) VGL_CUDA(cudaMemcpyAsync(ll_dev, ctx->d_ll.p, (size_t)ctx->tab.n_paths * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return VGL_OK;
}

int vgl_hc_set_paths_from_parts(vgl_ctx* ctx, const double* parts_dev, int32_t n_parts) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (!parts_dev || n_parts <= 0) return fail(VGL_ERR_INVALID, "no parts");
    VGL_CUDA(vgl::hc_launch_sum_parts(parts_dev, n_parts, ctx->tab.n_paths, ctx->d_ll.p, ctx->stream));
    ctx->launches += 1;
    ctx->ll_valid = true;
    return VGL_OK;
}

int vgl_hc_peer_export(vgl_ctx* ctx, int32_t world, unsigned char* handle_out) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (world < 1 || world > VGL_HC_MAX_PEERS || !handle_out) return fail(VGL_ERR_INVALID, "world size %d outside [1, %d]", world, VGL_HC_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) <= VGL_PEER_HANDLE_BYTES, "IPC handle does not fit");
    peer_detach(ctx);
    ctx->peer_pad = (ctx->tab.n_paths + 31) & ~31;
    const size_t bytes = (size_t)2 * world * ctx->peer_pad * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
    VGL_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->peer_local), bytes));
    VGL_CUDA(cudaMemset(ctx->peer_local, 0, bytes));
    cudaIpcMemHandle_t h;
    VGL_CUDA(cudaIpcGetMemHandle(&h, ctx->peer_local));
    memset(handle_out, 0, VGL_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    ctx->peer_world = world;
    ctx->peer_epoch = 0;
    ctx->peer_fused_pending = false;
    if (const char* e = getenv("VGL_PEER_TIMEOUT_S")) { const double v = atof(e); if (v > 0) ctx->peer_timeout_s = v; }
    return VGL_OK;
}

int vgl_hc_peer_attach(vgl_ctx* ctx, int32_t rank, int32_t world, const unsigned char* handles) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->peer_local || world != ctx->peer_world) return fail(VGL_ERR_STATE, "vgl_hc_peer_export(world = %d) must come first", world);
    if (rank < 0 || rank >= world || !handles) return fail(VGL_ERR_INVALID, "bad rank %d of %d", rank, world);
    for (int k = 0; k < world; ++k) {
        if (k == rank) { ctx->peer_bufs[k] = ctx->peer_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)k * VGL_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        VGL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_bufs[k] = static_cast<double*>(p);
    }
    ctx->peer_rank = rank;
    return VGL_OK;
}

int vgl_hc_exchange_paths(vgl_ctx* ctx) {
    if (int rc = use_device(ctx)) return rc;
    if (ctx->peer_rank < 0) return fail(VGL_ERR_STATE, "vgl_hc_peer_attach must come first");
    if (!ctx->ll_valid) {
        // the usual case: the collapse has not run yet -- run it with the exchange as its epilogue (every path tile goes to
        // all peers as soon as its last CTA has it); vgl_hc_finalize's posterior kernel acquires and sums in rank order
        ctx->peer_epoch += 1;
        VGL_CUDA(vgl::hc_launch_path_kernel_fused(ctx->tab, ctx->d_cta_node_acc.p, ctx->d_node_sum.p, ctx->d_chunk_ll.p, ctx->d_ll.p,
                                                  ctx->d_tile_done.p, ctx->peer_bufs, ctx->peer_rank, ctx->peer_world, ctx->peer_pad,
                                                  ctx->peer_epoch, ctx->stream));
        ctx->launches += 2;
        ctx->ll_valid = true;
        ctx->peer_fused_pending = true;
        return VGL_OK;
    }
    // this rank's vector is already on the device (vgl_hc_paths was called): the two-kernel exchange
    ctx->peer_epoch += 1;
    const int gen = (int)(ctx->peer_epoch & 1ull);
    (void)gen;
    VGL_CUDA(vgl::hc_launch_peer_exchange(ctx->peer_bufs, ctx->peer_rank, ctx->peer_world, ctx->tab.n_paths, ctx->peer_pad,
                                          ctx->peer_epoch, vgl::hc_path_tiles(ctx->tab.n_paths), ctx->d_ll.p, ctx->d_err, ctx->stream));
    ctx->launches += 2;
    return VGL_OK;
}

}  // extern "C"

namespace {
// The posterior kernel and the copy of its result block into slot `slot`, enqueued; the slot's event marks the copy.
int finalize_enqueue(vgl_ctx* ctx, int slot) {
    if (!ctx->have_tables) return fail(VGL_ERR_STATE, "vgl_hc_set_tables must come first");
    if (!ctx->ll_valid)
        if (int rc = vgl_hc_paths(ctx, nullptr)) return rc;
    cudaStream_t s = ctx->stream;
    const int P = ctx->tab.n_paths;
    const int cap = (int)std::min<size_t>(ctx->h_result_cap, (size_t)(ctx->have_tree ? ctx->max_anc + 1 : 1));
    const size_t at = (size_t)slot * (4 + ctx->h_result_cap);
    double* d_scalars = ctx->d_result.p + at;
    int32_t* d_idx = reinterpret_cast<int32_t*>(d_scalars + 2);
    double* d_post = d_scalars + 4;
    if (ctx->peer_fused_pending) {
        VGL_CUDA(vgl::hc_launch_posterior_fused(ctx->d_ll.p, P, ctx->row_words, ctx->have_tree ? ctx->d_clade_bits.p : nullptr,
                                                ctx->have_tree ? ctx->d_anc_off.p : nullptr, ctx->have_tree ? ctx->d_anc_clade.p : nullptr,
                                                d_post, cap, d_scalars, d_idx, ctx->d_err, ctx->peer_bufs, ctx->peer_rank, ctx->peer_world,
                                                ctx->peer_pad, ctx->peer_epoch, ctx->peer_timeout_s, s));
        ctx->peer_fused_pending = false;
    } else {
        VGL_CUDA(vgl::hc_launch_posterior(ctx->d_ll.p, P, ctx->row_words, ctx->have_tree ? ctx->d_clade_bits.p : nullptr,
                                          ctx->have_tree ? ctx->d_anc_off.p : nullptr,
                                          ctx->have_tree ? ctx->d_anc_clade.p : nullptr, d_post, cap, d_scalars, d_idx,
                                          ctx->d_err, s));
    }
    ctx->launches += 1;
    VGL_CUDA(cudaMemcpyAsync(ctx->h_result + at, d_scalars, (4 + (size_t)cap) * sizeof(double), cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaEventRecord(ctx->res_ev[slot], s));
    ctx->res_pending[slot] = true;
    return VGL_OK;
}

// Waits for slot `slot` and decodes it.
int finalize_collect(vgl_ctx* ctx, int slot, int32_t* argmax_out, double* total_out, double* post_out, int32_t post_cap, int32_t* n_post_out) {
    if (!ctx->res_pending[slot]) return fail(VGL_ERR_STATE, "no result pending in this slot (vgl_hc_finalize_begin comes first)");
    VGL_CUDA(cudaEventSynchronize(ctx->res_ev[slot]));
    ctx->res_pending[slot] = false;
    const double* scalars = ctx->h_result + (size_t)slot * (4 + ctx->h_result_cap);
    const int32_t* idx = reinterpret_cast<const int32_t*>(scalars + 2);
    const double* post = scalars + 4;
    const int err = idx[2];
    if (err & 0x40000000)
        return fail(VGL_ERR_PEER, "multi-GPU exchange: a peer rank did not deliver its vector in time (VGL_PEER_TIMEOUT_S, default 120 s)");
    // (first: the chunk's batch was never written, so whatever the read kernel made of it means nothing)
    if (err & 0x04000000)
        return fail(VGL_ERR_INVALID, "device-side packing: a chunk holds more empty mappings than the device-built batch reserves "
                                     "(groups + 4096 blocks of extra records); pack this sample on the host (vgh_hc_pack + vgl_hc_accumulate)");
    if (err & 0x20000000)
        return fail(VGL_ERR_BAD_READ, "device-side reconstruction rejected a read (malformed alignment: lengths, offsets, mapping quality or edit sizes out of range)");
    if (err & 0x10000000)
        return fail(VGL_ERR_BAD_NODE, "device-side reconstruction: a mapping names a node outside the graph");
    if (err & VGL_ERR_BAD_NODE)
        return fail(VGL_ERR_BAD_NODE, "a mapping names a node without a pangenome_map entry (the reference would terminate in map::at)");
    if (err & 0x08000000)
        return fail(VGL_ERR_INVALID, "a batch's vgl_hc_batch_meta is wrong: a segment holds a read of another mapping quality");
    if (argmax_out) *argmax_out = idx[0];
    if (total_out) *total_out = scalars[0];
    const int n_post = idx[1];
    if (n_post_out) *n_post_out = n_post;
    if (post_out) {
        if (post_cap < n_post) return fail(VGL_ERR_INVALID, "post_cap %d < %d posteriors", post_cap, n_post);
        for (int k = 0; k < n_post; ++k) post_out[k] = post[k];
    }
    return VGL_OK;
}
}  // namespace

extern "C" {

int vgl_hc_finalize(vgl_ctx* ctx, double* ll_out, int32_t* argmax_out, double* total_out, double* post_out,
                    int32_t post_cap, int32_t* n_post_out) {
    if (int rc = use_device(ctx)) return rc;
    const int slot = (int)(ctx->res_seq++ & 1u);
    if (ctx->res_pending[slot]) return fail(VGL_ERR_STATE, "two results are in flight already (vgl_hc_finalize_end first)");
    if (int rc = finalize_enqueue(ctx, slot)) return rc;
    if (ll_out) {
        VGL_CUDA(cudaMemcpyAsync(ll_out, ctx->d_ll.p, (size_t)ctx->tab.n_paths * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return finalize_collect(ctx, slot, argmax_out, total_out, post_out, post_cap, n_post_out);
}

int vgl_hc_finalize_begin(vgl_ctx* ctx, int32_t* ticket_out) {
    if (int rc = use_device(ctx)) return rc;
    if (!ticket_out) return fail(VGL_ERR_INVALID, "null ticket");
    const int slot = (int)(ctx->res_seq & 1u);
    if (ctx->res_pending[slot]) return fail(VGL_ERR_STATE, "two results are in flight already (vgl_hc_finalize_end first)");
    if (int rc = finalize_enqueue(ctx, slot)) return rc;
    ctx->res_seq += 1;
    *ticket_out = slot;
    return VGL_OK;
}

int vgl_hc_finalize_end(vgl_ctx* ctx, int32_t ticket, int32_t* argmax_out, double* total_out, double* post_out, int32_t post_cap,
                        int32_t* n_post_out) {
    if (int rc = use_device(ctx)) return rc;
    if (ticket < 0 || ticket > 1) return fail(VGL_ERR_INVALID, "bad ticket");
    return finalize_collect(ctx, ticket, argmax_out, total_out, post_out, post_cap, n_post_out);
}

int vgl_hc_read_kernel_ms(vgl_ctx* ctx, float* ms_out, int32_t* launches_out) {
    if (int rc = use_device(ctx)) return rc;
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    float tot = 0.f;
    for (size_t k = 0; k < ctx->ev_used; ++k) {
        float ms = 0.f;
        VGL_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[k].first, ctx->ev_pool[k].second));
        tot += ms;
    }
    if (ms_out) *ms_out = tot;
    if (launches_out) *launches_out = (int32_t)ctx->ev_used;
    return VGL_OK;
}

// -------------------------------------------------------------------------------------------
// euka
// -------------------------------------------------------------------------------------------
int vgl_euka_set_tables(vgl_ctx* ctx, const vgl_euka_tables* T) {
    if (int rc = use_device(ctx)) return rc;
    if (!T || T->n_clades < 1 || T->n_clades > 65535 || !T->clade_dist || !T->bin_off || !T->deam || T->k5 < 1 || T->k3 < 1 || T->max_node < 0)
        return fail(VGL_ERR_INVALID, "bad euka tables");
    if (T->n_bins != T->bin_off[T->n_clades] || (T->n_bins && (!T->bin_lo || !T->bin_hi))) return fail(VGL_ERR_INVALID, "bad euka bins");
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    // node -> clade: the LAST bin over all clades (file order) that contains the node wins; nodes in no bin go to
    // clade 0, the reference's initial value of c_n (src/readGAM_Euka.h:115-140)
    std::vector<uint16_t> node_clade((size_t)T->max_node + 1, 0);
    for (int c = 0; c < T->n_clades; ++c)
        for (int j = T->bin_off[c]; j < T->bin_off[c + 1]; ++j) {
            const int64_t lo = std::max<int64_t>(T->bin_lo[j], 0), hi = std::min<int64_t>(T->bin_hi[j], T->max_node);
            for (int64_t n = lo; n <= hi; ++n) node_clade[(size_t)n] = (uint16_t)c;
        }
    vgl::EukaSmallTables small;
    memset(&small, 0, sizeof(small));
    for (int b = 0; b < 256; ++b) {
        const int Q = (int)(int8_t)(uint8_t)b;
        const double qs = (Q >= 0 && Q < 100) ? T->qscore[Q] : 0.25;
        small.one_minus_e[b] = 1 - qs;
        small.e_third[b] = qs / 3;
        small.log_one_minus_e[b] = std::log(1 - qs);
        small.log_e_third[b] = std::log(qs / 3);
    }
    for (int k = 0; k < 5; ++k) small.base_freq_log[k] = T->base_freq_log[k];  // symbol codes 0..4 = A C G T N
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) small.tt[x][y] = T->tt_ratio[x][y];
    small.log_gap_in = std::log(0.002);   // src/readGAM_Euka.h:244-249
    small.log_gap_not = std::log(0.2);
    small.log_rare_not = std::log(0.001); // :252-257
    small.log_quarter = std::log(0.25);   // :278
    small.log_match2 = std::log(1 - 0.25536);  // :423
    small.log_mis2 = std::log(0.25536);        // :440
    const size_t nb = (size_t)std::max(T->n_bins, 1);
    if (int rc = ctx->e_node_clade.reserve(node_clade.size())) return rc;
    if (int rc = ctx->e_dist.reserve((size_t)T->n_clades)) return rc;
    if (int rc = ctx->e_bin_off.reserve((size_t)T->n_clades + 1)) return rc;
    if (int rc = ctx->e_bin_lo.reserve(nb)) return rc;
    if (int rc = ctx->e_bin_hi.reserve(nb)) return rc;
    if (int rc = ctx->e_deam.reserve((size_t)T->k5 * T->k3 * 16)) return rc;
    if (int rc = ctx->e_small.reserve(1)) return rc;
    if (int rc = ctx->e_count.reserve((size_t)T->n_clades)) return rc;
    if (int rc = ctx->e_like_n.reserve((size_t)T->n_clades)) return rc;
    if (int rc = ctx->e_like_logsum.reserve((size_t)T->n_clades)) return rc;
    if (int rc = ctx->e_bin_cov.reserve(nb)) return rc;
    VGL_CUDA(cudaMemcpy(ctx->e_node_clade.p, node_clade.data(), node_clade.size() * 2, cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->e_dist.p, T->clade_dist, (size_t)T->n_clades * 8, cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->e_bin_off.p, T->bin_off, ((size_t)T->n_clades + 1) * 4, cudaMemcpyHostToDevice));
    if (T->n_bins) {
        VGL_CUDA(cudaMemcpy(ctx->e_bin_lo.p, T->bin_lo, (size_t)T->n_bins * 4, cudaMemcpyHostToDevice));
        VGL_CUDA(cudaMemcpy(ctx->e_bin_hi.p, T->bin_hi, (size_t)T->n_bins * 4, cudaMemcpyHostToDevice));
    }
    VGL_CUDA(cudaMemcpy(ctx->e_deam.p, T->deam, (size_t)T->k5 * T->k3 * 16 * 8, cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemcpy(ctx->e_small.p, &small, sizeof(small), cudaMemcpyHostToDevice));
    vgl::EukaTablesDev& d = ctx->etab;
    d.n_clades = T->n_clades; d.n_bins = T->n_bins; d.max_node = T->max_node; d.k5 = T->k5; d.k3 = T->k3;
    d.minimum_mq = T->minimum_mq;
    d.node_clade = ctx->e_node_clade.p; d.clade_dist = ctx->e_dist.p; d.bin_off = ctx->e_bin_off.p;
    d.bin_lo = ctx->e_bin_lo.p; d.bin_hi = ctx->e_bin_hi.p; d.deam = ctx->e_deam.p; d.small = ctx->e_small.p;
    ctx->have_euka = true;
    return vgl_euka_reset(ctx);
}

int vgl_euka_reset(vgl_ctx* ctx) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka) return fail(VGL_ERR_STATE, "vgl_euka_set_tables must come first");
    VGL_CUDA(cudaMemsetAsync(ctx->e_count.p, 0, (size_t)ctx->etab.n_clades * sizeof(unsigned long long), ctx->stream));
    VGL_CUDA(cudaMemsetAsync(ctx->e_bin_cov.p, 0, (size_t)std::max(ctx->etab.n_bins, 1) * sizeof(double), ctx->stream));
    VGL_CUDA(cudaMemsetAsync(ctx->e_like_n.p, 0, (size_t)ctx->etab.n_clades * sizeof(unsigned long long), ctx->stream));
    VGL_CUDA(cudaMemsetAsync(ctx->e_like_logsum.p, 0, (size_t)ctx->etab.n_clades * sizeof(double), ctx->stream));
    return VGL_OK;
}

uint64_t vgl_euka_batch_bytes(const vgl_euka_batch* b) {
    if (!b) return 0;
    return (uint64_t)(b->n_groups + 1) * 8 + (b->n_col_blks + b->n_map_blks) * 512 + (uint64_t)b->n_groups * VGL_HC_LANES * 12;
}

int vgl_euka_run(vgl_ctx* ctx, const vgl_euka_batch* b, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik,
                 double* like, uint8_t* pass) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka) return fail(VGL_ERR_STATE, "vgl_euka_set_tables must come first");
    if (!b) return fail(VGL_ERR_INVALID, "null batch");
    if (b->n_groups == 0) return VGL_OK;
    if (!b->grp_col_blk || !b->grp_map_blk || !b->col || !b->map || !b->n_cols || !b->n_maps || !b->seq_len || !b->mapq || !b->flags ||
        !b->read_index)
        return fail(VGL_ERR_INVALID, "euka batch arrays missing");
    cudaStream_t s = ctx->stream;
    VGL_CUDA(cudaStreamSynchronize(s));  // single staging area
    vgl::EukaBatchDev d;
    if (int rc = stage_symbol_batch(ctx, b, d)) return rc;
    return euka_run_on_device(ctx, d, n_out, clade, in_lik, not_lik, like, pass);
}

int vgl_euka_run_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik,
                     double* like, uint8_t* pass) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka) return fail(VGL_ERR_STATE, "vgl_euka_set_tables must come first");
    vgl::EukaBatchDev d;
    size_t n_map_rec = 0;
    if (int rc = pack_raw_symbols(ctx, raw, /*soibean=*/0, d, n_map_rec)) return rc;
    if (d.n_groups == 0) return VGL_OK;
    return euka_run_on_device(ctx, d, n_out, clade, in_lik, not_lik, like, pass);
}

int vgl_euka_rerun(vgl_ctx* ctx, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik, double* like, uint8_t* pass) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka || !ctx->e_last_valid) return fail(VGL_ERR_STATE, "vgl_euka_run / vgl_euka_run_raw must come first (its batch is what is scored again)");
    return euka_run_on_device(ctx, ctx->e_last, n_out, clade, in_lik, not_lik, like, pass);
}

int vgl_sibling_kernel_ms(vgl_ctx* ctx, float* ms_out) {
    if (int rc = use_device(ctx)) return rc;
    if (!ms_out) return fail(VGL_ERR_INVALID, "null output");
    *ms_out = 0.0f;
    if (!ctx->sib_timed) return VGL_OK;
    VGL_CUDA(cudaEventSynchronize(ctx->sib_ev1));
    VGL_CUDA(cudaEventElapsedTime(ms_out, ctx->sib_ev0, ctx->sib_ev1));
    return VGL_OK;
}

int vgl_euka_like_sums(vgl_ctx* ctx, int64_t* like_n, double* like_logsum) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka) return fail(VGL_ERR_STATE, "vgl_euka_set_tables must come first");
    cudaStream_t s = ctx->stream;
    if (like_n) VGL_CUDA(cudaMemcpyAsync(like_n, ctx->e_like_n.p, (size_t)ctx->etab.n_clades * 8, cudaMemcpyDeviceToHost, s));
    if (like_logsum) VGL_CUDA(cudaMemcpyAsync(like_logsum, ctx->e_like_logsum.p, (size_t)ctx->etab.n_clades * 8, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    return VGL_OK;
}

double vgl_euka_proposal_loglike(int32_t n_listed, const int32_t* clade_ids, const double* frac, const int64_t* like_n, const double* like_logsum) {
    double ll = 0.0;
    for (int32_t j = 0; j < n_listed; ++j) {
        const int32_t c = clade_ids[j];
        if (like_n[c] > 0) ll += (double)like_n[c] * std::log(frac[j]) + like_logsum[c];
    }
    return ll;
}

int vgl_euka_results(vgl_ctx* ctx, int64_t* clade_count, double* bin_cov) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_euka) return fail(VGL_ERR_STATE, "vgl_euka_set_tables must come first");
    cudaStream_t s = ctx->stream;
    if (clade_count) VGL_CUDA(cudaMemcpyAsync(clade_count, ctx->e_count.p, (size_t)ctx->etab.n_clades * 8, cudaMemcpyDeviceToHost, s));
    if (bin_cov && ctx->etab.n_bins) VGL_CUDA(cudaMemcpyAsync(bin_cov, ctx->e_bin_cov.p, (size_t)ctx->etab.n_bins * 8, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    return VGL_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------
// soibean
// -------------------------------------------------------------------------------------------
namespace {
int stage_symbol_batch(vgl_ctx* ctx, const vgl_euka_batch* b, vgl::EukaBatchDev& d) {
    cudaStream_t s = ctx->stream;
    ctx->s_last_valid = false;  // the staging area is about to be overwritten
    ctx->e_last_valid = false;
    const size_t G = b->n_groups, nl = G * VGL_HC_LANES;
    const size_t ncol = (size_t)b->n_col_blks * VGL_HC_LANES * VGL_EUKA_COL_BLK, nmap = (size_t)b->n_map_blks * VGL_HC_LANES * VGL_HC_MAP_BLK;
    vgl_ctx::EukaStage& st = ctx->estage;
    if (int rc = st.grp_col_blk.reserve(G + 1)) return rc;
    if (int rc = st.grp_map_blk.reserve(G + 1)) return rc;
    if (int rc = st.col.reserve(ncol)) return rc;
    if (int rc = st.map.reserve(nmap)) return rc;
    if (int rc = st.n_cols.reserve(nl)) return rc;
    if (int rc = st.n_maps.reserve(nl)) return rc;
    if (int rc = st.seq_len.reserve(nl)) return rc;
    if (int rc = st.mapq.reserve(nl)) return rc;
    if (int rc = st.flags.reserve(nl)) return rc;
    if (int rc = st.read_index.reserve(nl)) return rc;
    if (int rc = upload_async(st.grp_col_blk.p, b->grp_col_blk, (G + 1) * 4, s)) return rc;
    if (int rc = upload_async(st.grp_map_blk.p, b->grp_map_blk, (G + 1) * 4, s)) return rc;
    if (int rc = upload_async(st.col.p, b->col, ncol * 2, s)) return rc;
    if (int rc = upload_async(st.map.p, b->map, nmap * 4, s)) return rc;
    if (int rc = upload_async(st.n_cols.p, b->n_cols, nl * 2, s)) return rc;
    if (int rc = upload_async(st.n_maps.p, b->n_maps, nl * 2, s)) return rc;
    if (int rc = upload_async(st.seq_len.p, b->seq_len, nl * 2, s)) return rc;
    if (int rc = upload_async(st.mapq.p, b->mapq, nl, s)) return rc;
    if (int rc = upload_async(st.flags.p, b->flags, nl, s)) return rc;
    if (int rc = upload_async(st.read_index.p, b->read_index, nl * 4, s)) return rc;
    d = vgl::EukaBatchDev{b->n_groups, st.grp_col_blk.p, st.grp_map_blk.p, st.col.p, st.map.p, st.n_cols.p, st.n_maps.p,
                          st.seq_len.p, st.mapq.p, st.flags.p, st.read_index.p};
    return VGL_OK;
}
}  // namespace

extern "C" {

int vgl_soibean_set_tables(vgl_ctx* ctx, const vgl_soibean_tables* T) {
    if (int rc = use_device(ctx)) return rc;
    if (!T || T->n_rows < 1 || T->n_paths < 1 || T->row_words < (T->n_paths + 31) / 32 || !T->member_bits || !T->deam || T->k5 < 1 ||
        T->k3 < 1 || T->penalty < 1)
        return fail(VGL_ERR_INVALID, "bad soibean tables");
    VGL_CUDA(cudaStreamSynchronize(ctx->stream));
    vgl::EukaSmallTables small;
    memset(&small, 0, sizeof(small));
    for (int b = 0; b < 256; ++b) {
        const int Q = (int)(int8_t)(uint8_t)b;
        const double qs = (Q >= 0 && Q < 100) ? T->qscore[Q] : 0.25;
        small.one_minus_e[b] = 1 - qs;
        small.e_third[b] = qs / 3;
        small.log_one_minus_e[b] = std::log(1 - qs);
        small.log_e_third[b] = std::log(qs / 3);
    }
    small.log_quarter = std::log(0.25);
    if (int rc = ctx->s_member.reserve((size_t)T->n_rows * T->row_words)) return rc;
    if (int rc = ctx->s_deam.reserve((size_t)T->k5 * T->k3 * 4)) return rc;
    if (int rc = ctx->s_small.reserve(1)) return rc;
    if (int rc = ctx->s_err.reserve(1)) return rc;
    VGL_CUDA(cudaMemcpy(ctx->s_member.p, T->member_bits, (size_t)T->n_rows * T->row_words * 4, cudaMemcpyHostToDevice));
    {
        // soibean marginalises over ALL four post-damage bases (src/getLCAfromGAM.h:340-349): per original base only the
        // row sum of the damage matrix matters
        std::vector<double> rowsum((size_t)T->k5 * T->k3 * 4);
        for (size_t m = 0; m < (size_t)T->k5 * T->k3; ++m)
            for (int bpo = 0; bpo < 4; ++bpo) {
                double s = 0.0;
                for (int bpd = 0; bpd < 4; ++bpd) s += T->deam[m * 16 + (size_t)bpo * 4 + bpd];
                rowsum[m * 4 + bpo] = s;
            }
        VGL_CUDA(cudaMemcpy(ctx->s_deam.p, rowsum.data(), rowsum.size() * 8, cudaMemcpyHostToDevice));
    }
    VGL_CUDA(cudaMemcpy(ctx->s_small.p, &small, sizeof(small), cudaMemcpyHostToDevice));
    VGL_CUDA(cudaMemset(ctx->s_err.p, 0, sizeof(int)));
    vgl::SoibeanTablesDev& d = ctx->stab;
    d.n_rows = T->n_rows; d.min_id = T->min_id; d.n_paths = T->n_paths; d.row_words = T->row_words; d.k5 = T->k5; d.k3 = T->k3;
    d.penalty = T->penalty; d.member_bits = ctx->s_member.p; d.deam = ctx->s_deam.p; d.small = ctx->s_small.p;
    ctx->have_soibean = true;
    return VGL_OK;
}

int vgl_soibean_run(vgl_ctx* ctx, const vgl_soibean_batch* b, uint32_t n_out, double* path_ll, uint8_t* used) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_soibean) return fail(VGL_ERR_STATE, "vgl_soibean_set_tables must come first");
    if (!b) return fail(VGL_ERR_INVALID, "null batch");
    if (b->n_groups == 0) return VGL_OK;
    if (!b->grp_col_blk || !b->grp_map_blk || !b->col || !b->map || !b->n_cols || !b->n_maps || !b->seq_len || !b->mapq || !b->flags ||
        !b->read_index)
        return fail(VGL_ERR_INVALID, "soibean batch arrays missing");
    cudaStream_t s = ctx->stream;
    VGL_CUDA(cudaStreamSynchronize(s));  // single staging area
    vgl::EukaBatchDev d;
    if (int rc = stage_symbol_batch(ctx, b, d)) return rc;
    return soibean_run_on_device(ctx, d, (size_t)b->n_map_blks * VGL_HC_LANES * VGL_HC_MAP_BLK, n_out, path_ll, used);
}

int vgl_soibean_rerun(vgl_ctx* ctx, uint32_t n_out, double* path_ll, uint8_t* used) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_soibean || !ctx->s_last_valid) return fail(VGL_ERR_STATE, "vgl_soibean_run / vgl_soibean_run_raw must come first (its batch is what is scored again)");
    const vgl::EukaBatchDev d = ctx->s_last;
    return soibean_run_on_device(ctx, d, ctx->s_last_nmap, n_out, path_ll, used);
}

int vgl_soibean_mcmc_loglike(vgl_ctx* ctx, int32_t n_proposals, int32_t n_sources, const vgl_soibean_source* sources,
                             const double* freqs7, double con, uint32_t n_out, double* per_read, double* loglike) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_soibean || !ctx->s_last_valid) return fail(VGL_ERR_STATE, "vgl_soibean_run / vgl_soibean_run_raw must come first (its batch is what the proposals are scored on)");
    if (n_proposals < 0 || n_sources < 1 || n_sources > VGL_SOIBEAN_MAX_SOURCES || !sources || !freqs7 || !loglike)
        return fail(VGL_ERR_INVALID, "bad proposal arguments (1 <= n_sources <= %d)", VGL_SOIBEAN_MAX_SOURCES);
    if (n_proposals == 0) return VGL_OK;
    const size_t ns = (size_t)n_proposals * (size_t)n_sources;
    for (size_t k = 0; k < ns; ++k)
        if (sources[k].child < 0 || sources[k].child >= ctx->stab.n_paths || sources[k].parent < 0 || sources[k].parent >= ctx->stab.n_paths)
            return fail(VGL_ERR_INVALID, "proposal %zu names a path outside [0, %d)", k / (size_t)n_sources, ctx->stab.n_paths);
    cudaStream_t s = ctx->stream;
    VGL_CUDA(cudaStreamSynchronize(s));  // the staging blocks are reused
    const int nb = vgl::soibean_mcmc_blocks(ctx->s_last);
    if (int rc = ctx->m_src.reserve(ns)) return rc;
    if (int rc = ctx->m_H.reserve(ns * 50)) return rc;
    if (int rc = ctx->m_partials.reserve((size_t)n_proposals * std::max(nb, 1))) return rc;
    if (int rc = ctx->m_loglike.reserve((size_t)n_proposals + 1)) return rc;
    if (per_read)
        if (int rc = ctx->m_per_read.reserve((size_t)n_proposals * std::max<size_t>(n_out, 1))) return rc;
    // one page-locked block carries the proposal in and the log-likelihoods + error flag out: two async copies, one sync
    const size_t in_bytes = ns * sizeof(vgl_soibean_source), out_bytes = ((size_t)n_proposals + 1) * sizeof(double);
    if (in_bytes + out_bytes > ctx->m_pinned_cap) {
        if (ctx->m_pinned) cudaFreeHost(ctx->m_pinned);
        ctx->m_pinned = nullptr;
        ctx->m_pinned_cap = 0;
        const size_t want = 2 * (in_bytes + out_bytes) + 256;
        VGL_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->m_pinned), want));
        ctx->m_pinned_cap = want;
    }
    double* h_out = reinterpret_cast<double*>(ctx->m_pinned);
    memcpy(ctx->m_pinned + out_bytes, sources, in_bytes);
    VGL_CUDA(cudaMemcpyAsync(ctx->m_src.p, ctx->m_pinned + out_bytes, in_bytes, cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemsetAsync(ctx->s_err.p, 0, sizeof(int), s));
    if (per_read) VGL_CUDA(cudaMemsetAsync(ctx->m_per_read.p, 0, (size_t)n_proposals * n_out * 8, s));
    VGL_CUDA(vgl::soibean_mcmc_launch(ctx->s_last, ctx->stab, ctx->s_sup.p, ctx->s_uns.p, ctx->s_mcode.p, ctx->m_src.p, n_proposals, n_sources, freqs7, con,
                                      n_out, per_read ? ctx->m_per_read.p : nullptr, ctx->m_H.p, ctx->m_partials.p, ctx->m_loglike.p,
                                      ctx->s_err.p, s));
    ctx->launches += 3;
    VGL_CUDA(cudaMemcpyAsync(h_out, ctx->m_loglike.p, out_bytes, cudaMemcpyDeviceToHost, s));
    if (per_read) VGL_CUDA(cudaMemcpyAsync(per_read, ctx->m_per_read.p, (size_t)n_proposals * n_out * 8, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    memcpy(loglike, h_out, (size_t)n_proposals * sizeof(double));
    if (h_out[n_proposals] != 0.0) return fail(VGL_ERR_BAD_READ, "a read's log-likelihood under a proposal is positive, nan or infinite (the reference throws)");
    return VGL_OK;
}

int vgl_soibean_run_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, uint32_t n_out, double* path_ll, uint8_t* used) {
    if (int rc = use_device(ctx)) return rc;
    if (!ctx->have_soibean) return fail(VGL_ERR_STATE, "vgl_soibean_set_tables must come first");
    vgl::EukaBatchDev d;
    size_t n_map_rec = 0;
    if (int rc = pack_raw_symbols(ctx, raw, /*soibean=*/1, d, n_map_rec)) return rc;
    if (d.n_groups == 0) return VGL_OK;
    return soibean_run_on_device(ctx, d, n_map_rec, n_out, path_ll, used);
}

}  // extern "C"

namespace {
int soibean_run_on_device(vgl_ctx* ctx, const vgl::EukaBatchDev& d, size_t n_map_rec, uint32_t n_out, double* path_ll, uint8_t* used) {
    cudaStream_t s = ctx->stream;
    const size_t nmap = std::max<size_t>(n_map_rec, 1);
    const size_t P = (size_t)ctx->stab.n_paths, no = std::max<size_t>(n_out, 1);
    if (int rc = ctx->s_sup.reserve(nmap)) return rc;
    if (int rc = ctx->s_uns.reserve(nmap)) return rc;
    if (int rc = ctx->s_mcode.reserve(nmap)) return rc;
    if (int rc = ctx->s_out.reserve(no * P)) return rc;
    if (int rc = ctx->s_used.reserve(no)) return rc;
    VGL_CUDA(cudaMemsetAsync(ctx->s_out.p, 0, no * P * 8, s));
    VGL_CUDA(cudaMemsetAsync(ctx->s_used.p, 0, no, s));
    VGL_CUDA(cudaMemsetAsync(ctx->s_err.p, 0, sizeof(int), s));
    if (int rc = sibling_events(ctx)) return rc;
    VGL_CUDA(cudaEventRecord(ctx->sib_ev0, s));
    VGL_CUDA(vgl::soibean_launch(d, ctx->stab, ctx->s_sup.p, ctx->s_uns.p, ctx->s_mcode.p, n_out, ctx->s_out.p, ctx->s_used.p, ctx->s_err.p, s));
    VGL_CUDA(cudaEventRecord(ctx->sib_ev1, s));
    ctx->sib_timed = true;
    ctx->launches += 2;
    int err = 0;
    VGL_CUDA(cudaMemcpyAsync(&err, ctx->s_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (path_ll) VGL_CUDA(cudaMemcpyAsync(path_ll, ctx->s_out.p, (size_t)n_out * P * 8, cudaMemcpyDeviceToHost, s));
    if (used) VGL_CUDA(cudaMemcpyAsync(used, ctx->s_used.p, (size_t)n_out, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    if (err & 2) return fail(VGL_ERR_BAD_NODE, "an edit names a node outside the membership table (the reference would throw from vector::at)");
    if (err & 1) return fail(VGL_ERR_BAD_READ, "an alignment slices past its aligned strings or the damage table (the reference would throw / read out of range)");
    ctx->s_last = d;  // resident for vgl_soibean_mcmc_loglike / vgl_soibean_rerun
    ctx->s_last_nmap = n_map_rec;
    ctx->s_last_valid = true;
    return VGL_OK;
}

}  // namespace

namespace {

int euka_run_on_device(vgl_ctx* ctx, const vgl::EukaBatchDev& d, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik,
                       double* like, uint8_t* pass) {
    cudaStream_t s = ctx->stream;
    const size_t no = std::max<size_t>(n_out, 1);
    if (int rc = ctx->e_clade.reserve(no)) return rc;
    if (int rc = ctx->e_in.reserve(no)) return rc;
    if (int rc = ctx->e_not.reserve(no)) return rc;
    if (int rc = ctx->e_like.reserve(no)) return rc;
    if (int rc = ctx->e_pass.reserve(no)) return rc;
    VGL_CUDA(cudaMemsetAsync(ctx->e_clade.p, 0xff, no * 4, s));  // reads that were not in the batch keep clade -1
    VGL_CUDA(cudaMemsetAsync(ctx->e_in.p, 0, no * 8, s));
    VGL_CUDA(cudaMemsetAsync(ctx->e_not.p, 0, no * 8, s));
    VGL_CUDA(cudaMemsetAsync(ctx->e_like.p, 0, no * 8, s));
    VGL_CUDA(cudaMemsetAsync(ctx->e_pass.p, 0, no, s));
    vgl::EukaOutDev o{n_out, ctx->e_clade.p, ctx->e_in.p, ctx->e_not.p, ctx->e_like.p, ctx->e_pass.p, ctx->e_count.p, ctx->e_bin_cov.p,
                      ctx->e_like_n.p, ctx->e_like_logsum.p};
    if (int rc = sibling_events(ctx)) return rc;
    VGL_CUDA(cudaEventRecord(ctx->sib_ev0, s));
    VGL_CUDA(vgl::euka_launch_read_kernel(d, ctx->etab, o, s));
    VGL_CUDA(cudaEventRecord(ctx->sib_ev1, s));
    ctx->sib_timed = true;
    ctx->launches += 1;
    if (clade) VGL_CUDA(cudaMemcpyAsync(clade, ctx->e_clade.p, (size_t)n_out * 4, cudaMemcpyDeviceToHost, s));
    if (in_lik) VGL_CUDA(cudaMemcpyAsync(in_lik, ctx->e_in.p, (size_t)n_out * 8, cudaMemcpyDeviceToHost, s));
    if (not_lik) VGL_CUDA(cudaMemcpyAsync(not_lik, ctx->e_not.p, (size_t)n_out * 8, cudaMemcpyDeviceToHost, s));
    if (like) VGL_CUDA(cudaMemcpyAsync(like, ctx->e_like.p, (size_t)n_out * 8, cudaMemcpyDeviceToHost, s));
    if (pass) VGL_CUDA(cudaMemcpyAsync(pass, ctx->e_pass.p, (size_t)n_out, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    ctx->e_last = d;  // resident for vgl_euka_rerun
    ctx->e_last_valid = true;
    return VGL_OK;
}

// euka / soibean: the whole raw batch in one piece -- H2D, reconstruction (mode 1 / 2), group sizes read back, packing.
int pack_raw_symbols(vgl_ctx* ctx, const vgl_hc_raw* raw, int soibean, vgl::EukaBatchDev& d, size_t& n_map_rec) {
    ctx->s_last_valid = false;  // the device-built batch is about to be overwritten
    ctx->e_last_valid = false;
    if (!ctx->have_graph) return fail(VGL_ERR_STATE, "vgl_hc_set_graph must come first");
    if (!raw || !raw->qual_off || !raw->str_off || !raw->map_off || !raw->edit_off || !raw->edit_seq_off || !raw->mapq || !raw->seq_len || !raw->read_index)
        return fail(VGL_ERR_INVALID, "raw alignment arrays missing (euka / soibean need seq_len and read_index: vgh_euka_raw_build)");
    if (raw->reserved & VGL_HC_RAW_NARROW_ONLY) {
        if (!(raw->map_node16 || raw->map_node) || !raw->first_offset16 || !raw->map_n_edits8 || !raw->edit_slen8 || !raw->edit_from8 || !raw->edit_to8)
            return fail(VGL_ERR_INVALID, "a narrow-only raw batch needs the narrow arrays (node ids as 16 or 32 bits)");
    } else if ((raw->n_maps && (!raw->map_node || !raw->map_offset)) || (raw->n_edits && (!raw->edit_from || !raw->edit_to))) {
        return fail(VGL_ERR_INVALID, "raw alignment arrays missing (euka / soibean need seq_len and read_index: vgh_euka_raw_build)");
    }
    vgl_ctx::RawDev& R = ctx->rawdev[0];
    cudaStream_t s = ctx->stream;
    VGL_CUDA(cudaStreamSynchronize(s));  // single staging area
    const RawSpan sp = raw_span(raw, 0, raw->n_reads);
    const size_t n = sp.n, ne = sp.ne, ncap = sp.ncap, G = (n + 31) / 32;
    d = vgl::EukaBatchDev{};
    n_map_rec = 0;
    if (n == 0) return VGL_OK;
    if (int rc = reserve_raw_inputs(R, sp)) return rc;
#define RSV(buf, cnt) if (int rc = (buf).reserve(std::max<size_t>((cnt), 1))) return rc
    RSV(R.seq_len, n); RSV(R.read_index, n);
    RSV(R.gstr, ncap); RSV(R.rstr, ncap); RSV(R.sizes, ne); RSV(R.n_cols, n); RSV(R.n_close, n); RSV(R.n_extra, n);
    RSV(R.hist3, 3 * 4096); RSV(R.order, G * 32); RSV(R.grp_cb, G + 1); RSV(R.grp_mb, G + 1); RSV(R.s_err, 1);
    RSV(R.s_n_cols, G * 32); RSV(R.s_n_maps, G * 32); RSV(R.s_seq_len, G * 32); RSV(R.s_mapq, G * 32); RSV(R.s_flags, G * 32);
    RSV(R.s_read_index, G * 32);
    // the input arrays (wide, or narrow and widened on the device like HaploCart's raw route), then the two per-read arrays
    if (int rc = stage_raw_inputs(ctx, raw, 0, sp, R, s)) return rc;
    VGL_CUDA(cudaMemcpyAsync(R.seq_len.p, raw->seq_len, n * 2, cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemcpyAsync(R.read_index.p, raw->read_index, n * 4, cudaMemcpyHostToDevice, s));
    VGL_CUDA(cudaMemsetAsync(R.hist3.p, 0, 3 * 4096 * sizeof(uint32_t), s));
    VGL_CUDA(cudaMemsetAsync(R.s_err.p, 0, sizeof(int), s));
    const vgl::ReconGraphDev Gd{ctx->g_min_id, ctx->g_n_nodes, ctx->g_seq_off.p, ctx->g_seq.p};
    // per-read arrays are indexed by the read's position; everything they point INTO by absolute index (as pack_raw_chunk)
    const vgl::ReconRawDev Rd{(uint32_t)n, R.qual_off.p, R.qual.p - sp.qa, R.mapq.p, R.str_off.p, R.map_off.p, R.map_node.p - sp.ma,
                              R.map_offset.p - sp.ma, R.edit_off.p - sp.ma, R.edit_from.p - sp.ea, R.edit_to.p - sp.ea, R.edit_seq_off.p - sp.ea,
                              R.edit_seq.p - sp.sa, R.seq_len.p, R.read_index.p};
    vgl::ReconWorkDev Wd{R.gstr.p - sp.ca, R.rstr.p - sp.ca, R.sizes.p - sp.ea, R.n_cols.p, R.n_close.p, R.n_extra.p, R.hist3.p, R.hist3.p + 4096,
                         R.hist3.p + 8192, R.order.p, R.grp_cb.p, R.grp_mb.p};
    Wd.n_slots = 1;   // euka / soibean: groups by shape only
    Wd.n_lanes = (uint32_t)(G * 32);
    memset(Wd.lane_base, 0, sizeof(Wd.lane_base));
    memset(Wd.slot_of_mapq, 0, sizeof(Wd.slot_of_mapq));
    VGL_CUDA(cudaMemsetAsync(R.order.p, 0xff, G * 32 * sizeof(uint32_t), s));
    VGL_CUDA(vgl::hc_launch_recon(Rd, Gd, Wd, soibean ? 2 : 1, R.s_err.p, s));
    // the group tables are final: size the record arrays from them
    uint32_t tot[2] = {0, 0};
    int err = 0;
    VGL_CUDA(cudaMemcpyAsync(&tot[0], R.grp_cb.p + G, 4, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaMemcpyAsync(&tot[1], R.grp_mb.p + G, 4, cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaMemcpyAsync(&err, R.s_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGL_CUDA(cudaStreamSynchronize(s));
    if (err & 0x10000000) return fail(VGL_ERR_BAD_NODE, "device-side reconstruction: a mapping names a node outside the graph (or beyond the edit record's node field)");
    if (err) return fail(VGL_ERR_BAD_READ, "device-side reconstruction rejected a read (lengths, offsets, fragment size outside [15, 1000], or the path spells more bases than the read has)");
    RSV(R.s_col, (size_t)tot[0] * VGL_HC_LANES * VGL_EUKA_COL_BLK);
    RSV(R.s_map, (size_t)tot[1] * VGL_HC_LANES * VGL_HC_MAP_BLK);
#undef RSV
    VGL_CUDA(vgl::sym_launch_pack(Rd, Wd, soibean, R.s_col.p, R.s_map.p, R.s_n_cols.p, R.s_n_maps.p, R.s_seq_len.p, R.s_mapq.p, R.s_flags.p,
                                  R.s_read_index.p, s));
    ctx->launches += 6;
    n_map_rec = (size_t)tot[1] * VGL_HC_LANES * VGL_HC_MAP_BLK;
    d = vgl::EukaBatchDev{(uint32_t)G, R.grp_cb.p, R.grp_mb.p, R.s_col.p, R.s_map.p, R.s_n_cols.p, R.s_n_maps.p, R.s_seq_len.p, R.s_mapq.p,
                          R.s_flags.p, R.s_read_index.p};
    return VGL_OK;
}

}  // namespace
