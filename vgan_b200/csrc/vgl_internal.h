// Internal declarations shared by the C-ABI layer (vgl_api.cu) and the kernels (hc_kernels.cu).
// Not installed; the public surface is include/vgan_gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vgan_gpu.h"

#ifndef VGL_HC_ACC_COPIES
#define VGL_HC_ACC_COPIES 32
#endif

#define VGL_HC_MAX_PEERS 16
// node table entry (u16): class id in bits 0-14 (VGL_HC_CLASS_NONE: no pangenome_map entry), bit 15: every path goes
// through the node
#define VGL_HC_CLASS_MASK 0x7FFFu
#define VGL_HC_CLASS_NONE 0x7FFFu
#define VGL_HC_CLASS_ALLPATHS 0x8000u

namespace vgl {

// Small per-process lookup tables, copied into shared memory by every CTA of the read kernel.
// Indexed by the RAW quality byte (0..255); bytes >= 128 are negative as `signed char`, which is
// how the reference sees them (src/update_likelihood.cpp:41).
struct HcSmallTables {
    double f_col[4096];    // 1 - eps, indexed by bits 0-11 of the column record (mismatch | base << 1 | quality << 3 | latched << 11;
                           // the base bits do not matter): match eps = qscore[Q], mismatch eps = 1 - qscore[Q]; latched (or
                           // consensus-FASTA mode): the same with the background error probability
    double log_err[256];   // log(get_p_seq_error(Q))           src/process_mapping.cpp:11
    double2 log_tab[256];  // fast_log table: {rc_k ~ 1/c_k, -log(rc_k)}, c_k = 1 + (k + 0.5)/256
    double p_correct[100]; // 1 - p_incorrect_map[mapq]         src/process_mapping.cpp:41
    double bg[4];          // background frequency of A,C,G,T   src/haplocart_functions.cpp:81-97
    double f_match_bep;    // 1 - bep
    double f_mis_bep;      // 1 - (1 - bep)
    double one_minus_bep;  // FASTA mode multiplier              src/process_mapping.cpp:73
    double log_err_scale;  // -ln(10)/10: log(get_p_seq_error(Q)) for Q > 2 is Q * this
    int32_t is_fasta;
    int32_t pad;
};

// The read kernel's constant tables live in global memory as one image with exactly the layout they have in shared
// memory; every CTA fetches it with a single TMA bulk copy.
//   [ HcSmallTables | node_class u16[n_rows + 1] (entry n_rows = VGL_HC_CLASS_NONE) | mode 0: {1 - pcm, pcm * K} double2[100][row]
//                                                                          mode 1: class table double2[n_class + 1] ]
// row = an odd number of 16-byte entries >= n_class + 1, so that one class in rows of different mapping qualities falls
// into different banks; entry n_class of every row is the neutral {1, 0} (class table: {0, 0}).
struct HcImageLayout {
    uint32_t off_small, off_class, off_ac, ac_row_bytes, bytes;
};
inline HcImageLayout hc_image_layout(int n_rows, int n_class, int ac_mode) {
    HcImageLayout L;
    L.off_small = 0;
    L.off_class = (uint32_t)((sizeof(HcSmallTables) + 15) & ~(size_t)15);
    L.off_ac = L.off_class + (uint32_t)((((size_t)n_rows + 1) * sizeof(uint16_t) + 15) & ~(size_t)15);
    L.ac_row_bytes = 16u * (uint32_t)((n_class + 1) | 1);
    L.bytes = L.off_ac + (ac_mode == 0 ? 100u * L.ac_row_bytes : ac_mode == 1 ? 16u * (uint32_t)(n_class + 1) : 0u);
    return L;
}

struct HcTablesDev {
    int32_t n_rows;
    int32_t n_paths;
    int32_t n_words;           // number of 32-node words = ceil(n_rows / 32)
    int32_t n_class;
    const double2* class_tab;   // [n_class + 1]  {mappability, match_prob}; entry n_class = {0, 0}
    const unsigned char* image;  // see HcImageLayout
    HcImageLayout layout;
    const uint32_t* node_bits;  // [n_words][n_paths]  bit k of word (w, p) = pathsgo[p] of node 32w+k
    const double* logterm;      // log-term tables (see kLogTermEntries), or nullptr when there are too many classes
};

struct HcBatchDev {
    uint32_t n_groups;
    const uint32_t* grp_col_blk;
    const uint32_t* grp_map_blk;
    const uint32_t* col;
    const uint32_t* map;
    const uint16_t* n_cols;
    const uint16_t* n_close;
    const uint16_t* n_extra;
    const uint8_t* mapq;
};

// What the host knows about a batch's value ranges (vgl_hc_batch_meta): decides whether the log-term kernel can run.
struct HcBatchMeta {
    bool valid = false;
    uint32_t qual_max = 255;
    uint32_t n_segs = 0;                      // 0: groups are not ordered by mapping quality
    uint8_t seg_mapq[VGL_HC_MAX_SEGS] = {};   // mapping quality of segment k
    uint32_t seg_end[VGL_HC_MAX_SEGS] = {};   // group index one past segment k
};

// Log-term tables (the read kernel's fastest mode): for every (mapping quality, node class) pair the per-column term
//   log((1 - pcm) * bg[base] + pcm * K * (1 - eps(quality, mismatch)))            src/process_mapping.cpp:59-77
// for quality 0..63, indexed by bits 0-8 of the column record (quality-major).  Global layout:
//   logterm[(mapq * (n_class + 1) + class) * 512 + (record & 0x1ff)], class n_class = the neutral entry of unknown nodes.
// A launch stages the rows of the mapping qualities PRESENT in the batch, cut at the batch's largest quality, in shared memory.
constexpr int kLogTermEntries = 512;
// One launch: every CTA works on ONE segment (= one mapping quality) of the batch and stages that mapping quality's tables;
// the CTAs are dealt out to the segments in proportion to their groups.
struct HcLogTermLaunch {
    int n_segs;
    int q_rows;                               // qualities 0 .. q_rows - 1
    uint8_t seg_mapq[VGL_HC_MAX_SEGS];
    uint32_t seg_begin[VGL_HC_MAX_SEGS + 1];  // groups [seg_begin[k], seg_begin[k + 1])
    uint32_t cta_begin[VGL_HC_MAX_SEGS + 1];  // CTAs   [cta_begin[k], cta_begin[k + 1]) work on segment k
};

// Read kernel launch geometry (one persistent CTA per SM).
struct HcReadLaunch {
    int grid;
    int block;
    size_t smem;
    int ac_mode;  // 0: {1 - pcm, pcm * K} per (mapq, class) in shared memory; 1: class table in shared memory; 2: in global memory
};

// force_mode < 0: the fastest mode that fits the shared-memory limit.
HcReadLaunch hc_read_launch_config(int n_rows, int n_class, int sm_count, int max_smem_optin, int force_mode);
cudaError_t hc_read_configure(const HcReadLaunch& cfg);

// K1: per-read, per-mapping scalars added into VGL_HC_ACC_COPIES node vectors (CTA b adds into copy b % copies).
//   node_acc[copy][n_words * 32]  += sum over the copy's mappings at node n of (B_i - A_i)
//   a_total[copy]                 += sum over the copy's mappings of A_i
cudaError_t hc_launch_read_kernel(const HcReadLaunch& cfg, const HcBatchDev& batch, const HcTablesDev& tab,
                                  double* cta_node_acc, double* cta_a_total, int* err_flag, cudaStream_t s);

// K1, log-term mode: one table lookup and two fp64 adds per aligned column, mappings finalised in line (no log, no queue).
// hc_logterm_plan: can this batch run in that mode on this device?  Fills `plan` and returns the dynamic shared memory it
// needs, or 0 when it cannot (quality above 63, too many (mapping quality, class) pairs for shared memory, no tables).
size_t hc_logterm_plan(const HcBatchMeta& meta, uint32_t n_groups, int n_ctas, int n_rows, int n_class, bool have_tables, int max_smem_optin,
                       HcLogTermLaunch& plan);
cudaError_t hc_logterm_configure(int max_smem_optin);
cudaError_t hc_launch_read_logterm_kernel(int grid, size_t smem, const HcLogTermLaunch& plan, const HcBatchDev& batch, const HcTablesDev& tab,
                                          double* cta_node_acc, double* cta_a_total, int* err_flag, cudaStream_t s);

// K2b: sums the copies and runs the bit-GEMV over the node x path membership: chunk_ll[chunk][P].
int hc_path_chunks(int n_words);
int hc_path_tiles(int n_paths);
// K2b + K2c: chunk_ll[chunk][P], then (by the last CTA of every path tile) ll[p] = sum_copies a_total[copy] + sum_chunks
// chunk_ll[chunk][p] in a fixed order that is the same for every p.  tile_done: hc_path_tiles(P) counters, zero on entry
// (the kernel leaves them zero).
// (K2a first: the copies of the node vector -- and of the A total, which follow them in node_acc -- collapse into node_sum,
// hc_node_sum_doubles(n_words) doubles per sample)
size_t hc_node_sum_doubles(int n_words);
cudaError_t hc_launch_path_kernel(const HcTablesDev& tab, const double* node_acc, double* node_sum, double* chunk_ll, double* ll,
                                  unsigned int* tile_done, cudaStream_t s);

// K2 / K3 over S samples at once (vgl_hc_finalize_samples).  Sample s: node vectors at node_acc + s * acc_stride doubles (single-
// sample layout), chunk partials chunk_ll[s][chunks][P], vector ll[s][P], counters tile_done[s][tiles] (zero on entry, left zero),
// result block results + s * result_stride doubles ({total, ll[argmax]}, {argmax, n_post, err}, post[post_cap]), error word
// err_flag + s * err_stride ints.
cudaError_t hc_launch_path_samples_kernel(const HcTablesDev& tab, int n_samples, const double* node_acc, size_t acc_stride, double* node_sum,
                                          double* chunk_ll, double* ll, unsigned int* tile_done, cudaStream_t s);
cudaError_t hc_launch_posterior_samples(int n_samples, const double* ll, int n_paths, int row_words, const uint32_t* clade_bits,
                                        const int32_t* anc_off, const int32_t* anc_clade, double* results, int post_cap, size_t result_stride,
                                        const int* err_flag, size_t err_stride, cudaStream_t s);

// Multi-GPU: ll[p] = sum_k parts[k][p] in k order.
cudaError_t hc_launch_sum_parts(const double* parts, int n_parts, int n_paths, double* ll, cudaStream_t s);

// Multi-GPU exchange through peer memory: peer_bufs[k] = rank k's exchange buffer as mapped into this process
// (layout: parts[2][world][n_pad] doubles, flags[2][world] u64).  Writes this rank's ll into every buffer, waits for
// all ranks' vectors in the local buffer, leaves their rank-ordered sum in ll.
cudaError_t hc_launch_peer_exchange(double* const* peer_bufs, int rank, int world, int n_paths, int n_pad,
                                    unsigned long long epoch, int n_tiles, double* ll, int* err_flag, cudaStream_t s);

// The exchange fused into K2 / K3 (what vgl_hc_exchange_paths + vgl_hc_finalize launch): K2's last CTA of every path tile stores
// its 128 paths into slot `rank` of every rank's buffer and releases a per-rank tile counter; K3's prologue acquires the
// world counters of the local buffer and adds the slots in rank order.  No separate exchange kernels.
cudaError_t hc_launch_path_kernel_fused(const HcTablesDev& tab, const double* node_acc, double* node_sum, double* chunk_ll, double* ll,
                                        unsigned int* tile_done, double* const* peer_bufs, int rank, int world, int n_pad,
                                        unsigned long long epoch, cudaStream_t s);
cudaError_t hc_launch_posterior_fused(double* ll, int n_paths, int row_words, const uint32_t* clade_bits, const int32_t* anc_off,
                                      const int32_t* anc_clade, double* post, int post_cap, double* out_scalars, int32_t* out_idx,
                                      const int* err_flag, double* const* peer_bufs, int rank, int world, int n_pad,
                                      unsigned long long epoch, double timeout_s, cudaStream_t s);

// K3: argmax (first max wins), log-sum-exp, posterior of the call and of its listed clades.
//   out_scalars[0] = total (log-sum-exp), out_scalars[1] = ll[argmax]; out_idx[0] = argmax, out_idx[1] = n_post,
//   out_idx[2] = *err_flag
size_t hc_posterior_smem(int n_paths);
cudaError_t hc_posterior_configure(size_t smem);
cudaError_t hc_launch_posterior(const double* ll, int n_paths, int row_words, const uint32_t* clade_bits,
                                const int32_t* anc_off, const int32_t* anc_clade, double* post, int post_cap,
                                double* out_scalars, int32_t* out_idx, const int* err_flag, cudaStream_t s);

// ---------------------------------------------------------------------------------------------------------
// device-side reconstruction and packing (recon_kernels.cu)
// ---------------------------------------------------------------------------------------------------------
struct ReconGraphDev {
    uint32_t min_id, n_nodes;
    const uint32_t* seq_off;  // [n_nodes + 1]
    const uint8_t* seq;
};
struct ReconRawDev {  // device copy of a vgl_hc_raw
    uint32_t n_reads;
    const uint32_t* qual_off;
    const uint8_t* qual;
    const uint8_t* mapq;
    const uint32_t* str_off;
    const uint32_t* map_off;
    const uint32_t* map_node;
    const uint32_t* map_offset;
    const uint32_t* edit_off;
    const uint16_t* edit_from;
    const uint16_t* edit_to;
    const uint32_t* edit_seq_off;
    const uint8_t* edit_seq;
    const uint16_t* seq_len;      // [n_reads] |Alignment.sequence| (euka / soibean only)
    const uint32_t* read_index;   // [n_reads] position of the read in the caller's order (euka / soibean only)
};
struct ReconWorkDev {  // scratch
    uint8_t* gstr;      // aligned graph strings, read r at str_off[r]
    uint8_t* rstr;      // aligned read strings
    uint16_t* sizes;    // [n_edits] one size per edit (src/vgan_utils.h:72), by absolute edit index
    uint16_t* n_cols;   // [n_reads] in caller order
    uint16_t* n_close;
    uint16_t* n_extra;  // HaploCart: empty mappings; euka / soibean: mapping / edit records of the read
    uint32_t* hist;     // [n_slots][4096] reads per (mapping-quality slot, column count)
    uint32_t* start;    // [n_slots][4096]
    uint32_t* cursor;   // [n_slots][4096]
    uint32_t* order;    // [n_lanes] lane -> read id: by slot, then by column count; kReconNoRead for an unused lane
    uint32_t* grp_cb;   // [n_groups + 1] -> the batch's grp_col_blk
    uint32_t* grp_mb;   // [n_groups + 1] -> the batch's grp_map_blk
    // HaploCart: reads are dealt to slots by mapping quality (a group never mixes two: vgl_hc_batch_meta segments); slot s
    // owns the lanes from lane_base[s] on, a multiple of 32.  euka / soibean / unsegmented: one slot, lane_base[0] = 0.
    uint32_t n_slots;
    uint32_t n_lanes;   // 32 * n_groups
    uint32_t lane_base[VGL_HC_MAX_SEGS];
    uint8_t slot_of_mapq[128];
};
constexpr uint32_t kReconNoRead = 0xffffffffu;
// hist and cursor must be zero and order all kReconNoRead on entry.
// ---------------------------------------------------------------------------------------------------------
// GAM wire decode on the device (gam_kernels.cu): serialized vg.Alignment messages -> the raw-alignment arrays of ReconRawDev
// ---------------------------------------------------------------------------------------------------------
struct GamOutDev {   // where G2 writes; the layout ReconRawDev reads
    uint32_t* qual_off; uint8_t* qual; uint8_t* mapq; uint32_t* str_off; uint32_t* map_off; uint32_t* map_node; uint32_t* map_offset;
    uint32_t* edit_off; uint16_t* edit_from; uint16_t* edit_to; uint32_t* edit_seq_off; uint8_t* edit_seq;
    uint16_t* seq_len;      // may be nullptr
    uint32_t* read_index;   // may be nullptr: index of the message in the chunk
};
// G1 + scan: counts[6][n] (reads, mappings, edits, quality bytes, edit-sequence bytes, string capacity; turned into exclusive
// offsets in place), keep[n], mapq_hist[128] and qual_max (zero on entry), totals[6].
cudaError_t gam_launch_count(const uint8_t* bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint32_t n, uint64_t byte_base,
                             uint32_t* counts, uint8_t* keep, uint32_t* mapq_hist, uint32_t* qual_max, uint32_t* totals, int* err_flag,
                             cudaStream_t s);
cudaError_t gam_launch_fill(const uint8_t* bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint32_t n, uint64_t byte_base,
                            const uint32_t* offs, const uint8_t* keep, const uint32_t* totals, const GamOutDev& out, int* err_flag,
                            cudaStream_t s);

// Duplicate removal on the device: table_size a power of two >= 2 n, tab_key all ones and tab_first all ones on entry, *n_dup zero.
cudaError_t dedup_launch(uint32_t n, const unsigned long long* key, const uint8_t* paired, uint32_t table_size, unsigned long long* tab_key,
                         uint32_t* tab_first, uint8_t* is_dup, unsigned int* n_dup, cudaStream_t s);

// The narrow arrays of a vgl_hc_raw, widened on the device: out[i] = base + counts[0] + ... + counts[i-1] for i in [0, n]
// (tile_scratch: n / 4096 + 1 words); 16-bit offsets to 32 bits.
cudaError_t raw_launch_offsets_from_counts(const uint8_t* counts, uint32_t n, uint32_t base, uint32_t* tile_scratch, uint32_t* out,
                                           cudaStream_t s);
cudaError_t raw_launch_widen_u16(const uint16_t* in, uint32_t n, uint32_t* out, cudaStream_t s);
cudaError_t raw_launch_widen_u8(const uint8_t* in, uint32_t n, uint16_t* out, cudaStream_t s);
// map_offset_abs[map_off[r]] = first16[r] for the n reads of the chunk (map_off holds absolute mapping indices; the rest of map_offset is zero)
cudaError_t raw_launch_first_offsets(const uint32_t* map_off, const uint16_t* first16, uint32_t n, uint32_t* map_offset_abs, cudaStream_t s);
cudaError_t raw_launch_widen_node16(const uint16_t* in, uint32_t n, uint32_t* out, cudaStream_t s);  // node | rev << 15 -> node | rev << 31
// mode: 0 HaploCart, 1 euka, 2 soibean (which checks apply, what the record count means, 4- or 8-column blocks)
cudaError_t hc_launch_recon(const ReconRawDev& raw, const ReconGraphDev& G, const ReconWorkDev& W, int mode, int* err_flag, cudaStream_t s);
// euka / soibean batch from the reconstructed strings (after hc_launch_recon with mode 1 / 2)
cudaError_t sym_launch_pack(const ReconRawDev& raw, const ReconWorkDev& W, int soibean, uint16_t* col, uint32_t* map, uint16_t* n_cols,
                            uint16_t* n_maps, uint16_t* seq_len, uint8_t* mapq, uint8_t* flags, uint32_t* read_index, cudaStream_t s);
// xmap (the extra records) must be zero on entry.
cudaError_t hc_launch_pack(const ReconRawDev& raw, const ReconWorkDev& W, uint32_t col_cap_blks, uint32_t map_cap_blks, uint32_t* col,
                           uint32_t* xmap, uint16_t* n_cols, uint16_t* n_close, uint16_t* n_extra, uint8_t* mapq, int* err_flag,
                           cudaStream_t s);

// Wire form of a chunk (vgl_hc_wire) -> 32-bit column records + the chunk-relative grp_col_blk of the 32-bit layout.
// The 16-bit arrays are addressed with ABSOLUTE block indices (grp_c16 / grp_n16 as shipped); col32 is chunk relative.
cudaError_t hc_launch_expand(uint32_t n_groups, const uint32_t* grp_c16, const uint32_t* grp_n16, const uint16_t* col16,
                             const uint8_t* col8, const uint32_t* flag4, const uint16_t* n_owned, const uint16_t* node16,
                             const uint16_t* n_cols, uint32_t* col32, uint32_t* grp_c32, cudaStream_t s);  // col8 != NULL: narrow columns

// ---------------------------------------------------------------------------------------------------------
// euka
// ---------------------------------------------------------------------------------------------------------
struct EukaSmallTables {
    double one_minus_e[256];      // 1 - qscore[Q] by raw quality byte (negative as signed char -> 0.25, as get_p_seq_error)
    double e_third[256];          // qscore[Q] / 3
    double log_one_minus_e[256];  // soft-clip column counted as a match    (src/readGAM_Euka.h:272)
    double log_e_third[256];      // soft-clip column counted as a mismatch (src/readGAM_Euka.h:276)
    double base_freq_log[16];     // by read symbol code: A C G T N, everything else 0
    double tt[4][4];              // t_T_ratio
    double log_gap_in, log_gap_not, log_rare_not, log_quarter, log_match2, log_mis2;
};

struct EukaTablesDev {
    int32_t n_clades, n_bins, max_node, k5, k3;
    uint32_t minimum_mq;
    const uint16_t* node_clade;  // [max_node + 1]
    const double* clade_dist;
    const int32_t* bin_off;
    const int32_t* bin_lo;
    const int32_t* bin_hi;
    const double* deam;          // [k5 * k3 * 16]
    const EukaSmallTables* small;
};

struct EukaBatchDev {
    uint32_t n_groups;
    const uint32_t* grp_col_blk;
    const uint32_t* grp_map_blk;
    const uint16_t* col;
    const uint32_t* map;
    const uint16_t* n_cols;
    const uint16_t* n_maps;
    const uint16_t* seq_len;
    const uint8_t* mapq;
    const uint8_t* flags;
    const uint32_t* read_index;
};

struct EukaOutDev {
    uint32_t n_out;
    int32_t* clade;
    double* in_lik;
    double* not_lik;
    double* like;
    uint8_t* pass;
    unsigned long long* clade_count;  // [n_clades]
    double* bin_cov;                  // [n_bins]
    unsigned long long* like_n;       // [n_clades] entries appended to Clade::clade_like (every scored read, src/readGAM_Euka.h:491)
    double* like_logsum;              // [n_clades] sum of log(clade_like) over them: all the abundance MCMC needs (src/MCMC.cpp:1175-1213)
};

cudaError_t euka_launch_read_kernel(const EukaBatchDev& b, const EukaTablesDev& t, const EukaOutDev& o, cudaStream_t s);

// ---------------------------------------------------------------------------------------------------------
// soibean (same batch layout as euka; "mapping" records are edit records)
// ---------------------------------------------------------------------------------------------------------
struct SoibeanTablesDev {
    int32_t n_rows, min_id, n_paths, row_words, k5, k3, penalty;
    const uint32_t* member_bits;
    const double* deam;  // [k5 * k3][4] row sums of the damage matrices
    const EukaSmallTables* small;  // one_minus_e / e_third / their logs by quality byte
};

// K5a: per read and edit, the "supported" and "unsupported" sums over the edit's slice -> sup / uns (indexed like the
//      edit records), mcode = the edit's match-pair counts for the MCMC kernel; K5b: out[read][p] = sum over edits of (path p goes through the edit's node ? sup : uns).
cudaError_t soibean_launch(const EukaBatchDev& b, const SoibeanTablesDev& t, double* sup, double* uns, uint32_t* mcode, uint32_t n_out,
                           double* out, uint8_t* used, int* err_flag, cudaStream_t s);

// soibean MCMC: loglike[q] = sum over the batch's reads of the per-read log-likelihood under proposal q (sources[q][n_sources]),
// loglike[n_proposals] = the error flag; per_read (optional) [n_proposals][n_out]; scratch: H [n_proposals][n_sources][50],
// partials [n_proposals][soibean_mcmc_blocks(b)].  sup / uns as K5a left them.  Launches 2 + number of 32768-tile slices kernels.
int soibean_mcmc_blocks(const EukaBatchDev& b);
cudaError_t soibean_mcmc_launch(const EukaBatchDev& b, const SoibeanTablesDev& t, const double* sup, const double* uns, const uint32_t* mcode,
                                const vgl_soibean_source* sources, int n_proposals, int n_sources, const double* freqs7, double con,
                                uint32_t n_out, double* per_read, double* H, double* partials, double* loglike, int* err_flag,
                                cudaStream_t s);

}  // namespace vgl
