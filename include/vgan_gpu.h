/*
 * vgan_gpu.h -- C ABI of the B200-native read x path likelihood engine for vgan.
 *
 * The reference (grenaud/vgan) has no plugin / FFI interface; the drop-in boundary is the set of
 * C++ call sites listed in SURVEY.md 8(b).  Every entry point below names the reference interface
 * it replaces (file:line relative to the reference root).  Plain pointers and sizes only; no C++,
 * STL or torch types cross this boundary; nothing throws (the reference's call sites are
 * `noexcept`, src/update_likelihood.cpp:10) -- every call returns a status and
 * vgl_last_error() describes the last failure on the calling thread.
 *
 * Threading / process model: one caller thread per context; the CUDA context is
 * created lazily inside vgl_create(), never in a static initialiser, so a host that fork()s its
 * mapping children first (src/HaploCart.cpp:313-378) is safe as long as it calls vgl_create()
 * afterwards.
 */
#ifndef VGAN_GPU_H
#define VGAN_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGL_ABI_VERSION 2

/* status codes */
enum {
    VGL_OK = 0,
    VGL_ERR_INVALID = 1,    /* bad argument                                                            */
    VGL_ERR_CUDA = 2,       /* CUDA runtime failure (message in vgl_last_error)                        */
    VGL_ERR_STATE = 3,      /* call order violated (e.g. accumulate before set_tables)                 */
    VGL_ERR_BAD_NODE = 4,   /* a mapping names a node with no pangenome_map entry / outside the table:
                               the reference would throw from map::at inside a noexcept function
                               (src/process_mapping.cpp:33) and terminate                              */
    VGL_ERR_BAD_READ = 5,   /* malformed alignment (see vgh_* packer)                                  */
    VGL_ERR_NO_DEVICE = 6,  /* no CUDA device: there is NO CPU fallback                                */
    VGL_ERR_PEER = 7        /* multi-GPU exchange: a peer did not deliver its vector in time            */
};

typedef struct vgl_ctx vgl_ctx;

/* ------------------------------------------------------------------------------------------- */
/* context                                                                                     */
/* ------------------------------------------------------------------------------------------- */

/* Create a context on CUDA device `device_id` (first CUDA call of the library happens here).   */
int vgl_create(int device_id, vgl_ctx** out);
void vgl_destroy(vgl_ctx* ctx);
const char* vgl_last_error(void);
int vgl_abi_version(void);

/* Run all work of this context on an externally owned CUDA stream (cudaStream_t passed as a
 * pointer).  NULL restores the context's own stream.  Lets a host that already owns device
 * memory and streams (e.g. torch.distributed glue) order its collectives against our kernels. */
int vgl_set_stream(vgl_ctx* ctx, void* cuda_stream);
int vgl_synchronize(vgl_ctx* ctx);

/* Pinned host memory for batches (optional; pageable pointers are accepted everywhere).        */
int vgl_host_alloc(void** out, size_t bytes);
int vgl_host_free(void* p);
/* Raw device memory, for callers that keep batches resident in HBM.                            */
int vgl_device_alloc(vgl_ctx* ctx, void** out, size_t bytes);
int vgl_device_free(vgl_ctx* ctx, void* p);
int vgl_memcpy_h2d(vgl_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int vgl_memcpy_d2h(vgl_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ------------------------------------------------------------------------------------------- */
/* HaploCart tables                                                                            */
/* ------------------------------------------------------------------------------------------- */

/* Immutable per-process tables (reference: built once at src/HaploCart.cpp:279-286).
 *   support_bits : the node x path membership matrix NodeInfo::pathsgo (src/NodeInfo.h:28-42,
 *                  filled at src/readPathHandleGraph.cpp:24-34 from path_supports,
 *                  src/load.cpp:283-300), bit-packed: bit (p & 31) of word
 *                  support_bits[row * row_words + (p >> 5)] is pathsgo[p] of node id `row`.
 *                  Row index == node id, exactly as the reference indexes path_supports[i].
 *   node_mappability[row] : mappabilities[pangenome_map.at(id)]      (src/process_mapping.cpp:33-35)
 *   node_match_prob[row]  : pow(1 - 30*mu(pangenome_base), 8)        (src/get_p_obs_base.cpp:43-64)
 *   node_valid[row]       : 1 iff pangenome_map holds the node id    (src/load.cpp:27-41)
 *   qscore[Q]             : get_qscore_vec()                          (src/miscfunc.h:198-212)
 *   p_incorrect_map[Q]    : precompute_incorrect_mapping_probs()      (src/haplocart_functions.cpp:100-107)
 *   bg_freq[0..3]         : get_background_freq('A','C','G','T')      (src/haplocart_functions.cpp:81-97)
 *   background_error_prob : -e                                        (src/HaploCart.cpp:87-209)
 *   is_consensus_fasta    : consensus-FASTA mode (also forces the background error mode,
 *                           src/HaploCart.cpp:395-401)
 */
typedef struct vgl_hc_tables {
    int32_t n_rows;
    int32_t n_paths;
    int32_t row_words; /* >= (n_paths + 31) / 32 */
    int32_t reserved0;
    const uint32_t* support_bits;
    const double* node_mappability;
    const double* node_match_prob;
    const uint8_t* node_valid;
    double qscore[100];
    double p_incorrect_map[100];
    double bg_freq[4];
    double background_error_prob;
    int32_t is_consensus_fasta;
    int32_t reserved1;
} vgl_hc_tables;

int vgl_hc_set_tables(vgl_ctx* ctx, const vgl_hc_tables* tables);

/* Haplogroup tree for the posterior (reference: parents.txt / children.txt, src/load.cpp:303-345,
 * consumed by get_posterior, src/get_posterior.cpp:87-127).
 *   clade_bits[c * row_words + w] : bit-set over paths of the STRICT descendants of clade c that are
 *                                   path names (what get_posterior_of_clade collects, :51-76)
 *   anc_off[p] .. anc_off[p+1]    : for path p, indices into anc_clade listing the clades reported
 *                                   for a call of p, in parents[p] order with consecutive duplicates
 *                                   removed (:110-118)                                           */
int vgl_hc_set_tree(vgl_ctx* ctx, int32_t n_clades, const uint32_t* clade_bits, const int32_t* anc_off,
                    const int32_t* anc_clade);

/* ------------------------------------------------------------------------------------------- */
/* HaploCart read batch (warp-striped SoA)                                                     */
/* ------------------------------------------------------------------------------------------- */

/* One batch = the reads of one call to the hot loop (src/HaploCart.cpp:408-421) after
 * reconstruct_graph_sequence (src/vgan_utils.h:6-79) has produced, per read, the aligned graph
 * string G, the aligned read string `algnseq` and the per-edit sizes, marshalled the way
 * update_likelihood hands them to process_mapping (src/update_likelihood.cpp:33-49):
 *
 *   mapping i of the read (i < number of vg Mappings):
 *       size_i = mppg_sizes[i]            -- the size of the i-th EDIT, indexed by mapping number
 *       pos_i  = min(sum_{k<i} size_k, n) -- position_in_read, n = |algnseq| = number of columns
 *       it compares graph_seq = G[pos_i .. pos_i+size_i) with mapping_seq = algnseq[0 ..) (the WHOLE
 *       aligned read string, indexed from 0) and quality_scores[j] = quality[pos_i + j].
 *
 *   Mapping i therefore owns the columns [pos_i, pos_{i+1}) (pos_M := min(pos_{M-1} + size_{M-1}, n)); the
 *   columns from pos_M on belong to no mapping and only their quality bytes count.
 *
 *   column record c of a read (0 <= c < n), 32 bits:
 *       bit 0      MISMATCH: G[c] != the read base process_mapping pairs with it, algnseq[c - pos_i] of the
 *                  mapping i that owns column c (src/get_p_obs_base.cpp:15,20)
 *       bits 1-2   that read base (A=0 C=1 G=2 T=3)
 *       bits 3-10  quality byte quality[c] (raw PHRED as in vg.Alignment.quality; 0 when c is past the end of the
 *                  quality string: the reference indexes the quality string by aligned column), CANONICALISED: every
 *                  byte the reference treats as "error probability 0.25" -- Q <= 2 in get_p_seq_error and in
 *                  HaploCart's qscore_vec (src/miscfunc.h:182-186,198-212), and bytes >= 128, negative as signed char --
 *                  is stored as 0; bytes 3..127 as they are
 *       bit 11     LATCHED: the owning mapping runs in background-error mode: a real quality byte >= 90 lies
 *                  before the end of its quality window [pos_i, pos_i + n) (src/update_likelihood.cpp:40-44)
 *       bit 12     VALID: a mapping owns the column and both bases are A/C/G/T (the isValidDNA / 'N'
 *                  filter of src/process_mapping.cpp:62-63)
 *       bit 13     OPEN: last column of its mapping (the kernel walks a read from its last column to its first
 *                  and picks up the node's constants here)
 *       bit 14     CLOSE: some mapping has pos_i == c: the sum over the columns walked since the previous
 *                  CLOSE and the quality sums from c on are what that mapping needs.  Set also where only
 *                  EMPTY mappings (size 0) start, including c = pos_M when the trailing mappings are empty.
 *       bits 15-31 Mapping.position.node_id of the owning mapping (row of the membership matrix, <=
 *                  VGL_HC_MAX_NODE), VGL_HC_NO_NODE for columns no mapping owns.
 *   The graph base itself is not shipped: get_p_obs_base ignores its base arguments
 *   (src/get_p_obs_base.cpp:38-69), only the comparison matters.  Bits 0-11 index the kernel's table of per-column
 *   factors directly; when no quality exceeds 63 and nothing is latched, bits 0-8 (quality-major: mismatch, base,
 *   quality) index a table of per-column LOG terms, one table per (mapping quality, node class) pair.
 *
 *   "extra" record, 32 bits, one per EMPTY mapping (size_i == 0 or pos_i == n; it owns no column but still
 *   contributes log(err) over its quality window): bits 0-19 node id, bits 20-31 pos_i (< 4096), in mapping
 *   order.  Real alignments hardly ever have one.
 *
 *   per read lane: n_cols (n), n_close (number of CLOSE bits), n_extra, mapq (0..99).
 *
 * Reads are stored in groups of 32 "lanes" (one per thread of a warp); the sum over reads is order
 * independent, so the packer is free to group reads of similar shape.  Records are blocked so that every
 * lane fetches 16 bytes at a time and a warp fetches 512 contiguous bytes:
 *   column block = 32 lanes x 4 consecutive column records (col[(blk * 32 + lane) * 4 + (c & 3)], c >> 2 = block)
 *   extra block  = 32 lanes x 4 consecutive extra records  (map[(blk * 32 + lane) * 4 + (i & 3)], i >> 2 = block)
 * Group g owns column blocks grp_col_blk[g] .. grp_col_blk[g+1] and extra blocks grp_map_blk[g] ..
 * grp_map_blk[g+1] (as many as its longest read needs); records past a lane's own length, and every
 * record of the unused lanes of the last group, are zero (n_cols = n_close = n_extra = 0 for unused lanes).
 * All arrays are host pointers for vgl_hc_accumulate and device pointers for vgl_hc_accumulate_device.
 * Algorithmic bytes per read: 4 n + 7 (+ 4 per empty mapping) plus padding (a few percent when lanes are
 * grouped by shape). */
#define VGL_HC_COL_MISMATCH 0x0001u
#define VGL_HC_COL_BASE_SHIFT 1
#define VGL_HC_COL_BASE(rec) (((rec) >> VGL_HC_COL_BASE_SHIFT) & 3u)
#define VGL_HC_COL_QUAL_SHIFT 3
#define VGL_HC_COL_QUAL(rec) (((rec) >> VGL_HC_COL_QUAL_SHIFT) & 0xffu)
#define VGL_HC_COL_LATCHED 0x0800u
/* the byte as stored: 0 for everything whose error probability the reference takes as 0.25 */
#define VGL_HC_CANON_QUAL(byte) ((((byte) & 0x80u) || ((byte) & 0xffu) <= 2u) ? 0u : ((byte) & 0xffu))
#define VGL_HC_COL_VALID 0x1000u
#define VGL_HC_COL_OPEN 0x2000u
#define VGL_HC_COL_CLOSE 0x4000u
#define VGL_HC_COL_NODE_SHIFT 15
#define VGL_HC_COL_NODE(rec) ((rec) >> VGL_HC_COL_NODE_SHIFT)
#define VGL_HC_NO_NODE 0x1FFFFu
#define VGL_HC_MAX_NODE 0x1FFFEu
#define VGL_HC_MAP_NODE_BITS 20
#define VGL_HC_MAP_MAX_NODE ((1u << VGL_HC_MAP_NODE_BITS) - 1u)
#define VGL_HC_MAP_MAX_POS 4095u
#define VGL_HC_MAP_REC(node, pos) ((uint32_t)(node) | ((uint32_t)(pos) << VGL_HC_MAP_NODE_BITS))
#define VGL_HC_LANES 32
#define VGL_HC_COL_BLK 4
#define VGL_HC_MAP_BLK 4

/* What the producer of a batch knows about its value ranges; lets the library take its fastest read kernel (per-column
 * log terms from shared-memory tables) when no quality exceeds 63 and the tables of the (mapping quality, class) pairs
 * present fit.  valid == 0: unknown (the general kernel runs).  Filled by vgh_hc_pack / vgh_hc_wire_from_packed /
 * vgh_hc_raw_build / vgh_hc_feed; copy it along when a batch is moved to the device.                                */
#define VGL_HC_MAX_SEGS 16
typedef struct vgl_hc_batch_meta {
    uint32_t valid;        /* 1: the fields below hold for every read of the batch */
    uint32_t qual_max;     /* largest CANONICAL quality (bits 3-10 of the column record) */
    uint32_t mapq_mask[4]; /* bit q set: some read has mapping quality q (0..99) */
    /* Segments: the batch's groups are ordered by mapping quality and no group mixes two (a segment's last group is filled up
     * with unused lanes), so that a CTA of the log-term kernel needs the tables of ONE mapping quality.  n_segs == 0: not so
     * (more than VGL_HC_MAX_SEGS distinct mapping qualities, or a producer that does not sort); the general kernel runs.
     *   packed / wire batch : seg_end[k] = group index one past segment k (segment k starts where k - 1 ends)
     *   raw batch           : seg_end[k] = number of reads with mapping quality seg_mapq[k]; the device sorts and pads */
    uint32_t n_segs;
    uint8_t seg_mapq[VGL_HC_MAX_SEGS];
    uint32_t seg_end[VGL_HC_MAX_SEGS];
} vgl_hc_batch_meta;

typedef struct vgl_hc_batch {
    uint32_t n_reads;            /* real reads (unused lanes of the last group excluded) */
    uint32_t n_groups;           /* groups of 32 lanes */
    uint64_t n_col_blks;         /* = grp_col_blk[n_groups] */
    uint64_t n_map_blks;         /* = grp_map_blk[n_groups] */
    const uint32_t* grp_col_blk; /* [n_groups + 1] */
    const uint32_t* grp_map_blk; /* [n_groups + 1] */
    const uint32_t* col;         /* [n_col_blks][32][4] column records (16-byte aligned) */
    const uint32_t* map;         /* [n_map_blks][32][4] extra records (16-byte aligned) */
    const uint16_t* n_cols;      /* [n_groups * 32] */
    const uint16_t* n_close;     /* [n_groups * 32] */
    const uint16_t* n_extra;     /* [n_groups * 32] */
    const uint8_t* mapq;         /* [n_groups * 32] */
    vgl_hc_batch_meta meta;
} vgl_hc_batch;

/* Bytes the batch occupies (what crosses PCIe / is read from HBM once per pass).               */
uint64_t vgl_hc_batch_bytes(const vgl_hc_batch* b);

/* The same batch in its WIRE form, for host -> device transfers (PCIe is the bound of vgl_hc_accumulate): the node
 * id, which the 32-bit column record repeats on every column of a mapping, travels once per mapping.
 *   col  : 16-bit column records, blocks of 8 per lane (32 lanes x 16 bytes): bits 0-14 as in the 32-bit record,
 *          bit 15 OWNED (a mapping owns the column)
 *   node : 16-bit node ids (graphs with node ids <= 65535, e.g. HaploCart's 11825 rows) of the read's non-empty
 *          mappings, LAST mapping first (the order in which a walk from the read's last column meets their OPEN marks),
 *          blocks of 8 per lane
 *   extras (map), per-lane arrays: as in vgl_hc_batch.
 * vgl_hc_accumulate_wire copies it chunk by chunk, expands each chunk into 32-bit records on the device
 * (hc_expand_kernel) and runs the read kernel on it; results are identical to vgl_hc_accumulate of the 32-bit batch.
 * Bytes per read: 2 n + 2 M' + 7 (M' non-empty mappings) -- 295 instead of 407 at config 2.                         */
#define VGL_HC_WIRE_OWNED 0x8000u
#define VGL_HC_WIRE_BLK 8
typedef struct vgl_hc_wire {
    uint32_t n_reads;
    uint32_t n_groups;
    uint64_t n_col_blks;          /* = grp_col_blk[n_groups]  */
    uint64_t n_node_blks;         /* = grp_node_blk[n_groups] */
    uint64_t n_map_blks;          /* = grp_map_blk[n_groups]  */
    const uint32_t* grp_col_blk;  /* [n_groups + 1] */
    const uint32_t* grp_node_blk; /* [n_groups + 1] */
    const uint32_t* grp_map_blk;  /* [n_groups + 1] */
    const uint16_t* col;          /* [n_col_blks][32][8]  */
    const uint16_t* node;         /* [n_node_blks][32][8] */
    const uint32_t* map;          /* [n_map_blks][32][4] extra records */
    const uint16_t* n_cols;       /* [n_groups * 32] */
    const uint16_t* n_close;      /* [n_groups * 32] */
    const uint16_t* n_extra;      /* [n_groups * 32] */
    const uint8_t* mapq;          /* [n_groups * 32] */
    /* NARROW column records (optional: all three NULL, or all three present and then `col` is not read).  When every
     * quality byte of the batch is <= 63 -- which also means that no mapping is LATCHED -- a column is 12 bits instead of 16:
     *   col8  [n_col_blks][32][8]  one byte per column: bits 0-5 quality, bits 6-7 read base
     *   flag4 [n_col_blks][32]     one 32-bit word per lane and block, 4 bits per column (column j in bits 4j .. 4j+3):
     *                              bit 0 MISMATCH, bit 1 OPEN, bit 2 CLOSE, bit 3 VALID
     *   n_owned [n_groups * 32]    columns a mapping owns: the first n_owned of the read (the columns past the last
     *                              mapping's end are the only ones nobody owns)
     * Bytes per read: 1.5 n + 2 M' + 9 -- 269 instead of 319 at config 2.                                            */
    const uint8_t* col8;
    const uint32_t* flag4;
    const uint16_t* n_owned;
    vgl_hc_batch_meta meta;
} vgl_hc_wire;
uint64_t vgl_hc_wire_bytes(const vgl_hc_wire* w);

/* ------------------------------------------------------------------------------------------- */
/* HaploCart hot path                                                                          */
/* ------------------------------------------------------------------------------------------- */

/* Start a new sample: zero the accumulators (reference: fill(final_vec, 0), HaploCart.cpp:481). */
int vgl_hc_reset(vgl_ctx* ctx);

/* Replaces the `#pragma omp parallel for` body + critical-section sum of src/HaploCart.cpp:408-421
 * (Haplocart::update -> update_likelihood -> process_mapping -> get_p_obs_base, SURVEY.md 8a rows
 * a3-a9) for every read of the batch.  May be called repeatedly; contributions add up.
 * Asynchronous with respect to the host (stream ordered); errors surface at finalize/synchronize. */
int vgl_hc_accumulate(vgl_ctx* ctx, const vgl_hc_batch* host_batch);
int vgl_hc_accumulate_device(vgl_ctx* ctx, const vgl_hc_batch* device_batch);
int vgl_hc_accumulate_wire(vgl_ctx* ctx, const vgl_hc_wire* host_wire);

/* ------------------------------------------------------------------------------------------- */
/* Device-side reconstruction and packing                                */
/* ------------------------------------------------------------------------------------------- */

/* Instead of a packed vgl_hc_batch the host may hand over the RAW alignments: the library then runs
 * reconstruct_graph_sequence (src/vgan_utils.h:6-79) and the marshalling of update_likelihood
 * (src/update_likelihood.cpp:33-49) on the device and feeds the resulting batch to the read kernel.
 * vgl_hc_raw is a compact copy of the fields of vector<AlignmentInfo*> the path reads (built by
 * vgh_hc_raw_build, one linear pass):
 *   qual_off[r] .. qual_off[r+1]   quality bytes of read r (raw PHRED)
 *   str_off[r+1] - str_off[r]      capacity reserved for read r's aligned strings: sum over its edits of
 *                                  from_length + to_length + |sequence|, rounded up to a multiple of 4
 *   map_node[m]                    Mapping.position.node_id | is_reverse << 31;  map_offset[m] = position.offset
 *   edit_from / edit_to [e]        Edit.from_length / to_length (<= 65535); edit_seq_off[e] .. [e+1] its sequence
 * Only reads with identity >= 1e-10 are listed (src/HaploCart.cpp:411).  All pointers are HOST pointers.      */
#define VGL_HC_RAW_NARROW_ONLY 1u /* vgl_hc_raw.reserved: only the narrow arrays are present (first_offset16 form; node ids in
                                   * map_node16, or in map_node when one exceeds 32767); the batch is at most one device
                                   * chunk (131072 reads); edit_off and edit_seq_off hold just their two end points
                                   * {first, last}; map_offset / edit_from / edit_to are NULL.  What vgh_hc_feed and
                                   * vgh_euka_feed / vgh_soibean_feed submit.                                                */
typedef struct vgl_hc_raw {
    uint32_t n_reads, n_maps, n_edits, reserved;
    const uint32_t* qual_off;     /* [n_reads + 1] */
    const uint8_t* qual;
    const uint8_t* mapq;          /* [n_reads]     */
    const uint32_t* str_off;      /* [n_reads + 1] */
    const uint32_t* map_off;      /* [n_reads + 1] */
    const uint32_t* map_node;     /* [n_maps]      */
    const uint32_t* map_offset;   /* [n_maps]      */
    const uint32_t* edit_off;     /* [n_maps + 1]  */
    const uint16_t* edit_from;    /* [n_edits]     */
    const uint16_t* edit_to;      /* [n_edits]     */
    const uint32_t* edit_seq_off; /* [n_edits + 1] */
    const uint8_t* edit_seq;
    const uint16_t* seq_len;      /* [n_reads] |Alignment.sequence|            (euka / soibean; may be NULL for HaploCart) */
    const uint32_t* read_index;   /* [n_reads] position in the caller's order  (euka / soibean; may be NULL for HaploCart) */
    /* Narrow twins (NULL, or all three present): what vgl_hc_accumulate_raw ships INSTEAD of map_offset / edit_off /
     * edit_seq_off when every value fits -- the offsets are rebuilt on the device with prefix sums.  The wide arrays stay
     * valid on the host (the library reads a few entries to cut the batch into chunks).                               */
    const uint16_t* map_offset16; /* [n_maps]  Mapping.position.offset                        */
    const uint8_t* map_n_edits8;  /* [n_maps]  edits of the mapping (edit_off differences)    */
    const uint8_t* edit_slen8;    /* [n_edits] length of the edit's sequence                  */
    const uint8_t* edit_from8;    /* [n_edits] from_length, to_length: NULL, or both present (every length <= 255) */
    const uint8_t* edit_to8;
    const uint16_t* map_node16;   /* [n_maps] node_id | is_reverse << 15: NULL, or present (every node id <= 32767) */
    const uint16_t* first_offset16; /* [n_reads] offset of the read's FIRST mapping: NULL, or present when every other mapping starts
                                     * at offset 0 of its node (what a graph aligner emits); shipped instead of map_offset16 then */
    vgl_hc_batch_meta meta;       /* HaploCart only (valid == 0 otherwise): ranges of the quality bytes and mapping qualities */
} vgl_hc_raw;
uint64_t vgl_hc_raw_bytes(const vgl_hc_raw* raw);

/* Node sequences of the graph (what reconstruct_graph_sequence asks of bdsg::ODGI::get_sequence): node
 * min_id + k has the forward sequence seq[seq_off[k] .. seq_off[k+1]).  Once per process.        */
int vgl_hc_set_graph(vgl_ctx* ctx, int64_t min_id, int64_t n_nodes, const int64_t* seq_off, const char* seq);

/* Reconstruct and pack `raw` on the device; *device_batch receives device pointers that stay valid until
 * the next vgl_hc_pack_raw / vgl_hc_accumulate_raw on this context (for inspection and tests).    */
int vgl_hc_pack_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, vgl_hc_batch* device_batch);
/* vgl_hc_accumulate for raw alignments: H2D of `raw`, reconstruction, packing, read kernel.  Returns when the host
 * arrays have been copied (the caller may reuse them); the kernels run asynchronously.  Malformed reads surface as VGL_ERR_BAD_READ / VGL_ERR_BAD_NODE at vgl_hc_finalize.              */
int vgl_hc_accumulate_raw(vgl_ctx* ctx, const vgl_hc_raw* raw);

/* Collapse the accumulators into the per-path log-likelihood vector final_vec[P]
 * (src/HaploCart.cpp:297-299,419-420) on the device.  ll_dev (device pointer, P doubles) receives
 * this context's (this GPU's) vector; pass NULL to keep it internal.                            */
int vgl_hc_paths(vgl_ctx* ctx, double* ll_dev);

/* Multi-GPU: replace the internal vector by the fixed-order sum of `n_parts` device vectors of P
 * doubles each (parts[k * P + p], e.g. the result of an all-gather over ranks).  Summing in rank
 * order keeps paths with identical support bit-identical, so std::max_element ties resolve as in
 * the single-GPU run.                                                                           */
int vgl_hc_set_paths_from_parts(vgl_ctx* ctx, const double* parts_dev, int32_t n_parts);

/* Multi-GPU without a library collective: the ranks of one node exchange their vectors through NVLink peer memory.
 * Every rank owns an exchange buffer (two generations of `world` slots of P doubles plus arrival flags) that its
 * peers map through CUDA IPC.  vgl_hc_exchange_paths launches two kernels on the context's stream: the first writes
 * this rank's final_vec[P] into slot `rank` of EVERY rank's buffer (peer stores over NVLink) and raises the flags; the
 * second waits for the `world` flags of its own buffer and sums the slots in rank order -- the same fixed-order sum
 * as vgl_hc_set_paths_from_parts, so calls stay bit-identical to a single-GPU run.  All ranks must call it once
 * per sample, in the same order.  A peer that does not deliver within ~10 s raises VGL_ERR_PEER at finalize.
 *   vgl_hc_peer_export : allocate the buffer for `world` ranks, return its IPC handle (VGL_PEER_HANDLE_BYTES bytes)
 *   vgl_hc_peer_attach : handles[k * VGL_PEER_HANDLE_BYTES ...] = the handle exported by rank k (own slot ignored)  */
#define VGL_PEER_HANDLE_BYTES 64
int vgl_hc_peer_export(vgl_ctx* ctx, int32_t world, unsigned char* handle_out);
int vgl_hc_peer_attach(vgl_ctx* ctx, int32_t rank, int32_t world, const unsigned char* handles);
int vgl_hc_exchange_paths(vgl_ctx* ctx);

/* Replaces std::max_element (src/HaploCart.cpp:423-424) and Haplocart::get_posterior
 * (src/get_posterior.cpp:87-127): argmax with first-max-wins, log-sum-exp over all paths, posterior
 * of the call and of every clade on its ancestor list.  Host outputs:
 *   ll_out[P] (may be NULL), *argmax_out, *total_out (log-sum-exp),
 *   post_out[0] = posterior of the call, post_out[1 + k] = posterior of its k-th listed clade,
 *   *n_post_out = 1 + number of listed clades (post_cap bounds the array).
 * If no tree was set only post_out[0] is produced.                                             */
int vgl_hc_finalize(vgl_ctx* ctx, double* ll_out, int32_t* argmax_out, double* total_out, double* post_out,
                    int32_t post_cap, int32_t* n_post_out);
/* The same in two halves, for a caller that loops over samples (src/HaploCart.cpp:303) and wants the GPU busy across them:
 * _begin enqueues the collapse, the posterior kernel and the copy of the result block into one of two page-locked slots and
 * returns at once -- the next sample's vgl_hc_reset / accumulate calls may follow immediately, on the same stream --; _end waits
 * for that slot alone and hands the result out (outputs as vgl_hc_finalize without ll_out).  At most two results in flight.  */
int vgl_hc_finalize_begin(vgl_ctx* ctx, int32_t* ticket_out);
int vgl_hc_finalize_end(vgl_ctx* ctx, int32_t ticket, int32_t* argmax_out, double* total_out, double* post_out, int32_t post_cap,
                        int32_t* n_post_out);

/* ------------------------------------------------------------------------------------------- */
/* Several samples per pass                                                                    */
/* ------------------------------------------------------------------------------------------- */
/* Haplocart::run loops over samples (multifasta / several inputs: src/HaploCart.cpp:303, fill + loop + posterior per sample,
 * :408-424, :459-462, :480-482), and a real sample is at most ~2 x 10^4 reads after duplicate removal -- microseconds of GPU
 * work -- so the per-sample fixed costs dominate a one-sample-per-call use.  A pass over S samples:
 *     vgl_hc_samples_begin(ctx, S);                      zeroes S accumulator sets (instead of vgl_hc_reset)
 *     for s: vgl_hc_sample_select(ctx, s); any vgl_hc_accumulate* call(s), vgh_hc_feed, ...    (adds into sample s)
 *     vgl_hc_finalize_samples(ctx, ...);                 ONE collapse over all samples (the node x path bit matrix is read once
 *                                                        per slice of samples, not once per sample), one posterior kernel with a
 *                                                        CTA per sample, one copy back, one synchronisation
 * Results are those of S separate reset / accumulate / finalize sequences, bit for bit (same summation orders).
 * Outputs (each may be NULL): ll_out[S][P], argmax_out[S], total_out[S], post_out[S][post_cap], n_post_out[S], status_out[S]
 * (VGL_OK or the error vgl_hc_finalize would have returned for that sample; with status_out == NULL the first such error is the
 * return value).  Leaves the context in single-sample mode.                                                                    */
int vgl_hc_samples_begin(vgl_ctx* ctx, int32_t n_samples);
int vgl_hc_sample_select(vgl_ctx* ctx, int32_t sample);
int vgl_hc_finalize_samples(vgl_ctx* ctx, double* ll_out, int32_t* argmax_out, double* total_out, double* post_out, int32_t post_cap,
                            int32_t* n_post_out, int32_t* status_out);

/* ------------------------------------------------------------------------------------------- */
/* GAM wire decode on the device                                                               */
/* ------------------------------------------------------------------------------------------- */
/* readGAM (src/readGAM.h:20-68) hands every vg::Alignment of the stream to a lambda that keeps sequence, path, mapping_quality,
 * quality, read_paired and identity and drops identity == 0; it is single-threaded protobuf parsing.  Here the host only
 * inflates the stream and walks its framing (vgh_gam_frame, include/vgan_host.h: offset and length of every serialized
 * vg.Alignment in the inflated bytes); the messages are parsed ON THE DEVICE, one thread each, straight into the raw-alignment
 * arrays the reconstruction kernels read, and flow on through reconstruction, packing and the read kernel like
 * vgl_hc_accumulate_raw.  Reads with identity < 1e-10 are left out, as the hot loop leaves them out (src/HaploCart.cpp:411).
 * All pointers are HOST pointers; *n_reads_out (optional) receives the number of reads that entered the sample.
 * A message that does not parse surfaces as VGL_ERR_BAD_READ at vgl_hc_finalize.                                         */
int vgl_hc_accumulate_gam(vgl_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* msg_off, const uint32_t* msg_len,
                          uint64_t n_msgs, uint64_t* n_reads_out);
/* The decoded alignments of at most one chunk of messages (131072), copied back for inspection and tests (every output may be
 * NULL).  totals6 = {reads kept, mappings, edits, quality bytes, edit-sequence bytes, aligned-string capacity}; arrays as in
 * vgl_hc_raw (map_node carries is_reverse in bit 31); read_index[r] = index of the message read r came from.            */
int vgl_gam_decode(vgl_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* msg_off, const uint32_t* msg_len, uint64_t n_msgs,
                   uint32_t* totals6, uint32_t* qual_off, uint8_t* qual, uint8_t* mapq, uint32_t* map_off, uint32_t* map_node,
                   uint32_t* map_offset, uint32_t* edit_off, uint16_t* edit_from, uint16_t* edit_to, uint32_t* edit_seq_off,
                   uint8_t* edit_seq, uint16_t* seq_len, uint32_t* read_index, uint32_t* str_off);

/* ------------------------------------------------------------------------------------------- */
/* Duplicate removal                                                                           */
/* ------------------------------------------------------------------------------------------- */
/* Dup_Remover::remove_duplicates_internal (src/rmdup.cpp:68-110; update_for_single, :20-39) runs directly in front of the hot
 * loop (src/HaploCart.cpp:390) as an O(n^2) scan over the position-sorted reads: read j is a duplicate when an EARLIER read i
 * that is not paired has the same start coordinates.  One hash pass on the device instead (two kernels):
 *   start_key[r] = node_id << 32 | offset of the read's first mapping (host array; vgh_dedup_keys builds it)
 *   paired[r]    = Alignment.read_paired (NULL: none).  A paired read marks nothing: update_for_paired indexes
 *                  mapping()[n_mappings], one past the end (src/rmdup.cpp:48-59), so its result is undefined in the reference.
 *   is_dup[r]    (host, n_reads bytes) receives 1 for duplicates; *n_kept the number of survivors.               */
int vgl_dedup(vgl_ctx* ctx, uint32_t n_reads, const uint64_t* start_key, const uint8_t* paired, uint8_t* is_dup, uint64_t* n_kept);

/* Number of kernels this context has launched since creation (bench.py reports it).            */
uint64_t vgl_launch_count(const vgl_ctx* ctx);
/* How many read-kernel launches of this context took the log-term kernel (the batch's vgl_hc_batch_meta allowed it). */
uint64_t vgl_hc_logterm_launches(const vgl_ctx* ctx);
/* Device time in milliseconds spent in the read kernel since the last vgl_hc_reset (CUDA events
 * on the launching stream); valid after vgl_synchronize / vgl_hc_finalize.                      */
int vgl_hc_read_kernel_ms(vgl_ctx* ctx, float* ms_out, int32_t* launches_out);

/* ------------------------------------------------------------------------------------------- */
/* euka: per-read "in clade" vs "random alignment" likelihoods with ancient-DNA damage         */
/* ------------------------------------------------------------------------------------------- */

/* Replaces the body of the per-alignment lambda of readGAM3 (src/readGAM_Euka.h:103-549, caller
 * src/Euka.cpp:534-537): clade of the read from its first node, per aligned column the damage-aware
 * in-clade likelihood and the random-alignment likelihood, the per-read clade_like value, the pass
 * filter, the per-clade count and the per-bin coverage.  euka scores every read against exactly ONE
 * clade, so a "read x clade evaluation" is one read.  The baseshift counters
 * (src/readGAM_Euka.h:160-173) are an integer histogram outside the likelihood path and stay on the host.
 *
 *   clade_dist[c]        clade_vec->at(c*6+1)->dist                       (src/load.cpp:108-157)
 *   bin_off/bin_lo/hi    chunks.at(c): node-id ranges of the clade's bins (src/load.cpp:71-95)
 *   deam[i5][i3][a][b]   subDeamDiNuc[L][l].p[a][b] with i5 = min(l, k5-1), i3 = min(L-l-1, k3-1):
 *                        Damage::initDeamProbabilities repeats the last profile row to the end of the
 *                        fragment, so the matrix depends on (L, l) only through those two indices
 *                        (src/damage.cpp:41-258); k5/k3 = rows of the --deam5p/--deam3p profiles (1 if none)
 *   qscore[Q]            Euka::get_qscore_vec: qs[2] = 10^-0.2, unlike HaploCart's (src/Euka.cpp:38-51)
 *   base_freq_log[5]     base_freq['A','C','G','T','N'] (already logs)    (src/Euka.cpp:446-450)
 *   tt_ratio[a][b]       t_T_ratio over A,C,G,T                           (src/Euka.cpp:453-468)
 *   minimum_mq           --minMQ (default 29)                             (src/Euka.cpp:184,282-284)     */
typedef struct vgl_euka_tables {
    int32_t n_clades;
    int32_t n_bins;
    int32_t max_node;     /* largest node id of the graph (the node -> clade table covers 0..max_node) */
    int32_t k5, k3;
    uint32_t minimum_mq;
    const double* clade_dist; /* [n_clades] */
    const int32_t* bin_off;   /* [n_clades + 1] */
    const int32_t* bin_lo;    /* [n_bins] */
    const int32_t* bin_hi;    /* [n_bins] */
    const double* deam;       /* [k5 * k3 * 16] */
    double qscore[100];
    double base_freq_log[5];
    double tt_ratio[4][4];
} vgl_euka_tables;

int vgl_euka_set_tables(vgl_ctx* ctx, const vgl_euka_tables* tables);

/* euka read batch, warp-striped like vgl_hc_batch (groups of 32 lanes, column blocks of VGL_EUKA_COL_BLK = 8
 * 16-bit records, mapping blocks of 4).  Column record m of a read (m < n_cols = |aligned graph string|), 16 bits:
 *   bits 0-7   quality byte at ALIGNED COLUMN m (src/readGAM_Euka.h:264,286; 0 past the end)
 *   bits 8-11  symbol of the aligned graph string, bits 12-15 symbol of the aligned read string:
 *              0-3 = A C G T, 4 = N, 5 = '-', 6 = an IUPAC code flagged in rare_bases (W M K R Y B D H V),
 *              7 = 'S' (soft clip marker), 8 = anything else
 * Mapping record: the node id.  Per lane: n_cols, n_maps, seq_len = |Alignment.sequence|, mapq,
 * flags bit 0 = first mapping on the reverse strand, read_index = position of the read in the caller's
 * order (per-read results are written there).                                                  */
#define VGL_EUKA_COL_BLK 8
#define VGL_EUKA_SYM_N 4
#define VGL_EUKA_SYM_GAP 5
#define VGL_EUKA_SYM_RARE 6
#define VGL_EUKA_SYM_SOFTCLIP 7
#define VGL_EUKA_SYM_OTHER 8
typedef struct vgl_euka_batch {
    uint32_t n_reads;
    uint32_t n_groups;
    uint64_t n_col_blks;
    uint64_t n_map_blks;
    const uint32_t* grp_col_blk; /* [n_groups + 1] */
    const uint32_t* grp_map_blk; /* [n_groups + 1] */
    const uint16_t* col;         /* [n_col_blks][32][8] */
    const uint32_t* map;         /* [n_map_blks][32][4] */
    const uint16_t* n_cols;      /* [n_groups * 32] */
    const uint16_t* n_maps;      /* [n_groups * 32] */
    const uint16_t* seq_len;     /* [n_groups * 32] */
    const uint8_t* mapq;         /* [n_groups * 32] */
    const uint8_t* flags;        /* [n_groups * 32] */
    const uint32_t* read_index;  /* [n_groups * 32] */
} vgl_euka_batch;

/* Zero the per-clade counts and bin coverage (start of a sample). */
int vgl_euka_reset(vgl_ctx* ctx);
/* Score a host batch.  Per-read outputs (host arrays indexed by read_index, each may be NULL):
 *   clade[r]   clade id c_n;  in_lik[r] / not_lik[r] the two log-likelihoods;
 *   like[r]    map_q * exp(in - logsumexp(in, not))  (clade_like; clade_not_like is 1 - like);
 *   pass[r]    1 if in - not > 1 and mapq > minimum_mq (the read is counted and adds bin coverage).   */
int vgl_euka_run(vgl_ctx* ctx, const vgl_euka_batch* host_batch, uint32_t n_out, int32_t* clade, double* in_lik,
                 double* not_lik, double* like, uint8_t* pass);
/* The same from RAW alignments (vgl_hc_raw built with vgh_euka_raw_build: reads with identity != 0, seq_len and read_index
 * filled): reconstruct_graph_sequence and the packing run on the device (vgl_hc_set_graph first).               */
int vgl_euka_run_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik,
                     double* like, uint8_t* pass);
/* Score AGAIN the batch the last vgl_euka_run / vgl_euka_run_raw left in HBM (no host -> device traffic): what a caller does
 * that keeps a sample resident, and what bench.py times as the device-resident figure.  Outputs as vgl_euka_run (each may be NULL). */
int vgl_euka_rerun(vgl_ctx* ctx, uint32_t n_out, int32_t* clade, double* in_lik, double* not_lik, double* like, uint8_t* pass);
/* Device time (ms, CUDA events on the launching stream) of the last euka read kernel / soibean kernel pair this context ran. */
int vgl_sibling_kernel_ms(vgl_ctx* ctx, float* ms_out);
/* Per-clade count (Clade::count) and per-bin coverage (get<3>(chunks[c][j])) accumulated since the reset. */
int vgl_euka_results(vgl_ctx* ctx, int64_t* clade_count, double* bin_cov);
/* euka's abundance MCMC (MCMC::run / get_proposal_likelihood, src/MCMC.cpp:1175-1366) re-reads every clade's clade_like vector
 * for each proposal: sum over the listed clades j of sum_{k >= 1} log(frac_j * clade_like[k] + clade_not_like[k] * (1/334)).
 * `1/334` is an integer division, i.e. 0, and index 0 is the 0.0 the vector is created with (src/load.cpp:128), so the sum is
 * like_n[c] * log(frac_j) + like_logsum[c] with, per clade, like_n = the number of reads scored for it and like_logsum = the
 * sum of log(clade_like) over them -- accumulated by the read kernel since vgl_euka_reset.  vgl_euka_like_sums returns the two
 * [n_clades] vectors (sum them over batches / GPUs), vgl_euka_proposal_loglike evaluates one proposal on the host (O(n_listed)). */
int vgl_euka_like_sums(vgl_ctx* ctx, int64_t* like_n, double* like_logsum);
double vgl_euka_proposal_loglike(int32_t n_listed, const int32_t* clade_ids, const double* frac, const int64_t* like_n,
                                 const double* like_logsum);
uint64_t vgl_euka_batch_bytes(const vgl_euka_batch* b);

/* ------------------------------------------------------------------------------------------- */
/* soibean: per-read x per-path log-likelihoods                                                */
/* ------------------------------------------------------------------------------------------- */

/* Replaces the per-alignment body of analyse_GAM (src/getLCAfromGAM.h:102-604, caller src/soibean.cpp:558-561):
 * AlignmentInfo::pathMap of every read, i.e. the R x P matrix that convertMapsToVector (src/soibean.cpp:131-155)
 * hands to the MCMC.  For edit i of the read the reference computes, for every path separately, either the
 * "supported" sum or the "unsupported" sum over the slice of aligned columns -- neither depends on the path --
 * so the device computes the two scalars once per edit and selects per path from the membership bits.
 *   member_bits[(node - min_id) * row_words + (p >> 5)] bit (p & 31): path p goes through the node
 *       (the reference keeps the path NAMES per node, src/soibean.cpp:476-491)
 *   deam / k5 / k3 as in vgl_euka_tables;  qscore = Euka::get_qscore_vec (src/soibean.cpp:471);
 *   penalty = PENALTY (default 7).                                                                  */
typedef struct vgl_soibean_tables {
    int32_t n_rows;
    int32_t min_id;
    int32_t n_paths;
    int32_t row_words;
    int32_t k5, k3;
    int32_t penalty;
    int32_t reserved;
    const uint32_t* member_bits;
    const double* deam;
    double qscore[100];
} vgl_soibean_tables;

/* The batch has the layout of vgl_euka_batch with one difference: the "mapping" records are EDIT records, one per
 * entry of mppg_sizes (analyse_GAM loops over edits, src/getLCAfromGAM.h:144): bits 20-31 the size of the edit,
 * bits 0-19 the node id of the MAPPING with the same index, or 0xFFFFF when there is none ("No_support",
 * :156-160); n_maps counts edits.  Column records, seq_len, flags and read_index are as for euka.        */
typedef vgl_euka_batch vgl_soibean_batch;
#define VGL_SOIBEAN_NO_NODE 0xFFFFFu

int vgl_soibean_set_tables(vgl_ctx* ctx, const vgl_soibean_tables* tables);
/* path_ll[r * n_paths + p] = pathMap[pathNames[p]] of read r (r = read_index); used[r] = 1 for reads of the batch. */
int vgl_soibean_run(vgl_ctx* ctx, const vgl_soibean_batch* host_batch, uint32_t n_out, double* path_ll, uint8_t* used);
int vgl_soibean_run_raw(vgl_ctx* ctx, const vgl_hc_raw* raw, uint32_t n_out, double* path_ll, uint8_t* used);
/* The same kernels again on the batch the last vgl_soibean_run / _run_raw left in HBM; with path_ll == NULL the R x P matrix stays
 * on the device (where vgl_soibean_mcmc_loglike reads what it needs), i.e. no device -> host traffic at all.           */
int vgl_soibean_rerun(vgl_ctx* ctx, uint32_t n_out, double* path_ll, uint8_t* used);

/* ------------------------------------------------------------------------------------------- */
/* soibean MCMC: log-likelihood of proposals                                                   */
/* ------------------------------------------------------------------------------------------- */

/* Replaces the read loop of MCMC::run_tree_proportion (src/MCMC.cpp:738-993) with computeBaseLogLike
 * (src/MCMC.h:111-296) and calculateLogWeightedAverage (src/MCMC.h:299-316): the log-likelihood of a proposal, summed
 * over the reads of the batch the LAST vgl_soibean_run / vgl_soibean_run_raw call scored -- that batch and its per-edit
 * sums stay resident in HBM until the next euka / soibean call on the context, so an MCMC iteration moves nothing but
 * the proposal in and one double out (sum over batches / GPUs on the caller's side when the sample spans several).
 * A proposal is n_sources sources (state_t_1.positions_tree[y], src/MCMC.cpp:697-709):
 *   child      index of pos->longname in pathNames        parent   index of pos->parent->longname (child's own at the root)
 *   pos_branch position on the branch                      dist     pos->dist (0 becomes 1e-5 as in :757-759)
 *   proportion state_t_1.proportions[y] (unused when n_sources == 1, as in the reference)
 * freqs7 = params.freqs['A','C','G','T','R','Y','M'] (src/soibean.cpp:618-640); con = the reference's `con`
 * (src/soibean.cpp:596-598).  Reproduced as written: kappa = 1/22 is an integer division, i.e. 0 (src/MCMC.h:66).
 * Stated limit: a read or graph base outside ACGT in a supported regular column has base frequency 0 (the reference
 * indexes freqs by the character, so the IUPAC letters R, Y, M would pick up the purine / pyrimidine / mu entries).
 * sources[q * n_sources + y]; loglike[q] for q < n_proposals; per_read: NULL or [n_proposals][n_out] by read_index.
 * VGL_ERR_BAD_READ where the reference would throw (a read's log-likelihood positive, nan or infinite). */
#define VGL_SOIBEAN_MAX_SOURCES 16
typedef struct vgl_soibean_source {
    int32_t child, parent;
    double pos_branch, dist, proportion;
} vgl_soibean_source;
int vgl_soibean_mcmc_loglike(vgl_ctx* ctx, int32_t n_proposals, int32_t n_sources, const vgl_soibean_source* sources,
                             const double* freqs7, double con, uint32_t n_out, double* per_read, double* loglike);

#ifdef __cplusplus
}
#endif
#endif /* VGAN_GPU_H */
