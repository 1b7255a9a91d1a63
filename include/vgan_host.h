/*
 * vgan_host.h -- C ABI of libvgan_host.so: the HOST side of the drop-in boundary.
 *
 * These are the pieces of the reference that sit directly on either side of the likelihood hot
 * path and that a C++ host (vgan itself) or a test harness needs in order to feed
 * include/vgan_gpu.h:
 *   - a GAM reader                     (reference: readGAM, src/readGAM.h:20-68)
 *   - alignment reconstruction         (reference: reconstruct_graph_sequence, src/vgan_utils.h:6-79)
 *   - the SoA batch packer             (new: AlignmentInfo -> vgl_hc_batch)
 *   - HaploCart table construction     (reference: src/load.cpp:6-58,283-345,
 *                                       src/readPathHandleGraph.cpp:14-37, src/miscfunc.h:180-216,
 *                                       src/haplocart_functions.cpp:81-107, src/get_p_obs_base.cpp:43-64)
 *   - the haplogroup tree for the posterior (reference: src/load.cpp:303-345, src/get_posterior.cpp:36-76)
 * It is plain C++17 with no CUDA dependency except optional pinned allocation through
 * vgl_host_alloc.  No likelihood arithmetic happens here; that is all on the device.
 */
#ifndef VGAN_HOST_H
#define VGAN_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "vgan_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    VGH_OK = 0,
    VGH_ERR_INVALID = 101,
    VGH_ERR_IO = 102,
    VGH_ERR_FORMAT = 103,   /* malformed GAM / table text                                          */
    VGH_ERR_BAD_READ = 104, /* an alignment the reference could not process either (it would throw
                               from std::string::substr / vector::at inside a noexcept function)    */
    VGH_ERR_BAD_NODE = 105,
    VGH_ERR_DEVICE = 106    /* the likelihood library (include/vgan_gpu.h) reported an error; see vgl_last_error() */
};

const char* vgh_last_error(void);

/* ---- graph node sequences (stands in for bdsg::ODGI::get_sequence) ------------------------- */
typedef struct vgh_graph vgh_graph;
/* node ids min_id .. min_id + n_nodes - 1; seq_off[n_nodes + 1] into forward-strand `seq`.      */
int vgh_graph_create(int64_t min_id, int64_t n_nodes, const int64_t* seq_off, const char* seq, vgh_graph** out);
void vgh_graph_destroy(vgh_graph* g);
/* The node table as stored (for vgl_hc_set_graph): ids min_id .. min_id + n_nodes - 1, node k = seq[seq_off[k] .. seq_off[k+1]). */
int vgh_graph_view(const vgh_graph* g, int64_t* min_id, int64_t* n_nodes, const int64_t** seq_off, const char** seq);

/* ---- alignments (flat view of vector<AlignmentInfo*>, reference src/AlignmentInfo.h:16-46) --- */
typedef struct vgh_alignments {
    int64_t n_reads;
    const int64_t* seq_off;  /* [n_reads + 1] into seq                                    */
    const char* seq;         /* Alignment.sequence                                        */
    const int64_t* qual_off; /* [n_reads + 1] into qual                                   */
    const char* qual;        /* Alignment.quality, raw PHRED bytes                        */
    const int32_t* mapq;     /* Alignment.mapping_quality                                 */
    const double* identity;  /* Alignment.identity                                        */
    const uint8_t* paired;   /* Alignment.read_paired (may be NULL)                       */
    const int64_t* map_off;  /* [n_reads + 1] into the mapping arrays                     */
    const int64_t* map_node; /* Mapping.position.node_id                                  */
    const int64_t* map_offset;
    const uint8_t* map_rev;
    const int64_t* edit_off; /* [n_maps + 1] into the edit arrays                         */
    const int32_t* edit_from;
    const int32_t* edit_to;
    const int64_t* edit_seq_off; /* [n_edits + 1] into edit_seq                           */
    const char* edit_seq;
} vgh_alignments;

/* GAM reader: gzip/BGZF or plain stream of framed vg.Alignment messages.  Like the reference's
 * readGAM it drops alignments whose identity is exactly 0 when drop_unmapped != 0
 * (src/readGAM.h:47).                                                                         */
typedef struct vgh_gam vgh_gam;
int vgh_gam_read_file(const char* path, int drop_unmapped, vgh_gam** out);
int vgh_gam_parse(const void* bytes, size_t n, int drop_unmapped, vgh_gam** out);
const vgh_alignments* vgh_gam_view(const vgh_gam* g);
void vgh_gam_destroy(vgh_gam* g);

/* GAM framing for the DEVICE-side decoder (vgl_hc_accumulate_gam, include/vgan_gpu.h): inflates the stream (BGZF members in
 * parallel) and walks the group / item framing; the vg.Alignment messages themselves are parsed on the GPU.
 * vgh_gam_frames_view returns the number of messages and the inflated bytes plus every message's offset and length.
 * An UNCOMPRESSED stream is not copied: the view points into `bytes`, which the caller keeps alive as long as the frames.   */
typedef struct vgh_gam_frames vgh_gam_frames;
int vgh_gam_frame(const void* bytes, size_t n, vgh_gam_frames** out);
int64_t vgh_gam_frames_view(const vgh_gam_frames* f, const uint8_t** bytes, uint64_t* n_bytes, const uint64_t** msg_off, const uint32_t** msg_len);
void vgh_gam_frames_destroy(vgh_gam_frames* f);

/* reconstruct_graph_sequence for read r: aligned graph string, aligned read string, one size per
 * EDIT.  Buffers hold `cap` entries.                                                           */
int vgh_reconstruct(const vgh_graph* g, const vgh_alignments* a, int64_t r, char* graph_seq, char* read_seq,
                    int32_t* sizes, int32_t cap, int32_t* n_graph, int32_t* n_read, int32_t* n_sizes);

/* ---- duplicate removal (reference: Dup_Remover::remove_duplicates_internal, src/rmdup.cpp:20-110) ---------- */
/* The reference marks read j a duplicate when an EARLIER read i that is not paired starts at the same
 * (node id, offset) of its first mapping, strand ignored (update_for_single, src/rmdup.cpp:20-39) -- an O(n^2)
 * scan directly in front of the hot loop (src/HaploCart.cpp:390).  This is the same predicate in one hash pass.
 * Paired reads: update_for_paired indexes mapping()[n_mappings], one past the end (src/rmdup.cpp:48-49,55-59), so
 * its result is undefined in the reference; here a paired read i marks nothing.
 * is_dup[r] (n_reads bytes) receives 1 for duplicates; returns the number of reads kept via *n_kept.        */
int vgh_dedup(const vgh_alignments* a, uint8_t* is_dup, int64_t* n_kept);
/* The start coordinates vgl_dedup (include/vgan_gpu.h) keys on: keys[r] = node_id << 32 | offset of read r's first mapping. */
int vgh_dedup_keys(const vgh_alignments* a, uint64_t* keys);

/* ---- batch packer --------------------------------------------------------------------------- */
typedef struct vgh_hc_packed vgh_hc_packed;
/* Pack reads [r0, r1) into one vgl_hc_batch.  Reads with identity < 1e-10 are skipped exactly as
 * the hot loop skips them (src/HaploCart.cpp:411).  `pinned` != 0 allocates the arrays with
 * vgl_host_alloc (needs a CUDA device).  n_threads <= 0: all cores.                            */
int vgh_hc_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                vgh_hc_packed** out);
const vgl_hc_batch* vgh_hc_packed_batch(const vgh_hc_packed* p);
int64_t vgh_hc_packed_skipped(const vgh_hc_packed* p);
void vgh_hc_packed_destroy(vgh_hc_packed* p);

/* ---- wire form of a packed batch (vgl_hc_wire, include/vgan_gpu.h) ---------------------------- */
typedef struct vgh_hc_wire vgh_hc_wire;
/* Re-encodes a packed batch for the host -> device transfer: 16-bit column records plus one 16-bit node id per
 * non-empty mapping.  Fails with VGH_ERR_INVALID when a node id exceeds 65535 (use the 32-bit batch then).  */
int vgh_hc_wire_from_packed(const vgh_hc_packed* p, int n_threads, int pinned, vgh_hc_wire** out);
const vgl_hc_wire* vgh_hc_wire_view(const vgh_hc_wire* w);
void vgh_hc_wire_destroy(vgh_hc_wire* w);

/* ---- raw alignments for the device-side reconstruction (vgl_hc_raw, include/vgan_gpu.h) ------ */
typedef struct vgh_hc_raw vgh_hc_raw;
/* One linear pass over reads [r0, r1): keeps those with identity >= 1e-10 (src/HaploCart.cpp:411) and copies the
 * fields the path reads into one (optionally pinned) block.  No reconstruction happens here.              */
int vgh_hc_raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out);
/* The same for euka / soibean: keeps the reads readGAM3 keeps (identity != 0, src/readGAM_Euka.h:73). */
int vgh_euka_raw_build(const vgh_alignments* a, int64_t r0, int64_t r1, int pinned, vgh_hc_raw** out);
const vgl_hc_raw* vgh_hc_raw_view(const vgh_hc_raw* p);
int64_t vgh_hc_raw_skipped(const vgh_hc_raw* p);
void vgh_hc_raw_destroy(vgh_hc_raw* p);

/* ---- feeder: flat alignment fields -> GPU, chunked and overlapped ------------------------------ */
/* Replaces the `#pragma omp parallel for` of src/HaploCart.cpp:408-421 for a caller that holds the flat alignment fields:
 * vgh_hc_feed adds reads [r0, r1) to the sample open on `ctx` (between vgl_hc_reset and vgl_hc_finalize).  The reads travel
 * in chunks of `chunk_reads` (<= 131072): while the device copies, reconstructs (src/vgan_utils.h:6-79) and scores one
 * chunk, `n_threads` host threads marshal the next into the other of two page-locked blocks the feeder keeps.  Returns
 * when every chunk has been copied (the kernels run on; errors in the reads surface at vgl_hc_finalize).
 * n_threads = 0: three quarters of the OpenMP threads (the submitting thread and the driver's threads need cores too). */
typedef struct vgh_hc_feeder vgh_hc_feeder;
struct vgl_ctx;
int vgh_hc_feeder_create(int n_threads, int64_t chunk_reads, vgh_hc_feeder** out);
int vgh_hc_feed(vgh_hc_feeder* f, struct vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1);
/* One chunk (r1 - r0 <= chunk_reads) marshalled without submitting it, for inspection and tests; *out is NULL when the
 * chunk does not fit the narrow arrays (vgh_hc_feed then takes vgh_hc_raw_build for it).                            */
int vgh_hc_feeder_marshal(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, const vgl_hc_raw** out);
/* The same pipeline for the sibling tools (the loops of src/Euka.cpp:534-537 and src/soibean.cpp:558-561): reads [r0, r1) with
 * identity != 0 (readGAM3's rule) go to vgl_euka_run_raw / vgl_soibean_run_raw chunk by chunk; the per-read outputs are indexed
 * by the read's position in [r0, r1) (each may be NULL; path_ll is (r1 - r0) x n_paths doubles) and are complete on return.     */
int vgh_euka_feed(vgh_hc_feeder* f, struct vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t* clade, double* in_lik,
                  double* not_lik, double* like, uint8_t* pass);
int vgh_soibean_feed(vgh_hc_feeder* f, struct vgl_ctx* ctx, const vgh_alignments* a, int64_t r0, int64_t r1, int32_t n_paths, double* path_ll,
                     uint8_t* used);
int vgh_euka_feeder_marshal(vgh_hc_feeder* f, const vgh_alignments* a, int64_t r0, int64_t r1, const vgl_hc_raw** out);
void vgh_hc_feeder_stats(const vgh_hc_feeder* f, double* build_seconds, double* wait_seconds, uint64_t* h2d_bytes, uint64_t* chunks,
                         uint64_t* slow_chunks);
void vgh_hc_feeder_destroy(vgh_hc_feeder* f);

/* ---- HaploCart tables ----------------------------------------------------------------------- */
typedef struct vgh_hc_tables vgh_hc_tables;
/* Inputs as the reference holds them in memory:
 *   supports[n_rows][n_paths] bytes (NodeInfo::pathsgo; row index == node id),
 *   pangenome_map as (node id, stored value = file coordinate + 1) pairs (src/load.cpp:27-41),
 *   mappabilities expanded per base (src/load.cpp:6-24).                                        */
int vgh_hc_tables_build(int32_t n_rows, int32_t n_paths, const uint8_t* supports, int64_t n_pmap,
                        const int64_t* pmap_node, const int32_t* pmap_val, int64_t n_mappab, const double* mappab,
                        double background_error_prob, int is_consensus_fasta, vgh_hc_tables** out);
const vgl_hc_tables* vgh_hc_tables_view(const vgh_hc_tables* t);
void vgh_hc_tables_destroy(vgh_hc_tables* t);
/* pow(1 - 30*mu(pangenome_base), 8): the only surviving term of get_p_obs_base
 * (src/get_p_obs_base.cpp:43-67).                                                              */
double vgh_hc_region_match_prob(int32_t pangenome_base);

/* ---- haplogroup tree ------------------------------------------------------------------------ */
typedef struct vgh_hc_tree vgh_hc_tree;
/* path_names_txt: one name per line in path_supports column order (src/load.cpp:44-58);
 * parents_txt / children_txt: the reference's file formats (src/load.cpp:303-345).             */
int vgh_hc_tree_build(int32_t n_paths, const char* path_names_txt, const char* parents_txt, const char* children_txt,
                      vgh_hc_tree** out);
int32_t vgh_hc_tree_n_clades(const vgh_hc_tree* t);
int32_t vgh_hc_tree_row_words(const vgh_hc_tree* t);
const uint32_t* vgh_hc_tree_clade_bits(const vgh_hc_tree* t);
const int32_t* vgh_hc_tree_anc_off(const vgh_hc_tree* t);
const int32_t* vgh_hc_tree_anc_clade(const vgh_hc_tree* t);
const char* vgh_hc_tree_clade_name(const vgh_hc_tree* t, int32_t clade);
void vgh_hc_tree_destroy(vgh_hc_tree* t);

/* ---- euka: tables and batch packer -------------------------------------------------------------- */
typedef struct vgh_euka_tables vgh_euka_tables;
/* deam5p_prof / deam3p_prof: CONTENTS of the --deam5p / --deam3p profile files (src/miscfunc.h:84-136),
 * NULL or "" for none.  Builds the combined substitution matrices of Damage::initDeamProbabilities
 * (src/damage.cpp:41-258) for every distinct (5' row, 3' row) pair.                                 */
int vgh_euka_tables_build(int32_t n_clades, const double* clade_dist, const int32_t* bin_off, const int32_t* bin_lo,
                          const int32_t* bin_hi, int32_t max_node, const char* deam5p_prof, const char* deam3p_prof,
                          uint32_t minimum_mq, vgh_euka_tables** out);
const vgl_euka_tables* vgh_euka_tables_view(const vgh_euka_tables* t);
void vgh_euka_tables_destroy(vgh_euka_tables* t);

typedef struct vgh_euka_packed vgh_euka_packed;
/* Pack reads [r0, r1) (identity != 0, as readGAM3 keeps them, src/readGAM_Euka.h:73); read_index is relative
 * to r0.  Fragments outside [15, 1000] bp are rejected: the reference indexes subDeamDiNuc out of range. */
int vgh_euka_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                  vgh_euka_packed** out);
const vgl_euka_batch* vgh_euka_packed_batch(const vgh_euka_packed* p);
void vgh_euka_packed_destroy(vgh_euka_packed* p);

/* ---- soibean: tables and batch packer ----------------------------------------------------------- */
typedef struct vgh_soibean_tables vgh_soibean_tables;
/* member[(node - min_id) * n_paths + p] != 0: path p goes through the node (soibean's nodepaths, src/soibean.cpp:476-491) */
int vgh_soibean_tables_build(int32_t n_rows, int32_t min_id, int32_t n_paths, const uint8_t* member, const char* deam5p_prof,
                             const char* deam3p_prof, int32_t penalty, vgh_soibean_tables** out);
const vgl_soibean_tables* vgh_soibean_tables_view(const vgh_soibean_tables* t);
void vgh_soibean_tables_destroy(vgh_soibean_tables* t);
/* Same packer as vgh_euka_pack, with edit records instead of mapping records (see vgl_soibean_batch). */
int vgh_soibean_pack(const vgh_graph* g, const vgh_alignments* a, int64_t r0, int64_t r1, int n_threads, int pinned,
                     vgh_euka_packed** out);

/* ---- synthetic HaploCart-shaped workloads (benchmark / test inputs) --- */
/* The real tables (share/vgan/hcfiles) are FTP downloads; tests and bench.py generate inputs of the
 * same shapes instead.  Pure functions of their seeds.                                            */
typedef struct vgh_synth_hc vgh_synth_hc;
int vgh_synth_hc_create(uint64_t seed, int32_t n_paths, int32_t n_rows, int32_t n_nodes, vgh_synth_hc** out);
void vgh_synth_hc_destroy(vgh_synth_hc* s);
const vgh_graph* vgh_synth_hc_graph(const vgh_synth_hc* s);
int64_t vgh_synth_hc_graph_seq(const vgh_synth_hc* s, const int64_t** seq_off, const char** seq);
const uint8_t* vgh_synth_hc_supports(const vgh_synth_hc* s);
int64_t vgh_synth_hc_n_pmap(const vgh_synth_hc* s);
const int64_t* vgh_synth_hc_pmap_node(const vgh_synth_hc* s);
const int32_t* vgh_synth_hc_pmap_val(const vgh_synth_hc* s);
int64_t vgh_synth_hc_n_mappab(const vgh_synth_hc* s);
const double* vgh_synth_hc_mappab(const vgh_synth_hc* s);
const char* vgh_synth_hc_names(const vgh_synth_hc* s);
const char* vgh_synth_hc_parents(const vgh_synth_hc* s);
const char* vgh_synth_hc_children(const vgh_synth_hc* s);
int32_t vgh_synth_hc_total_len(const vgh_synth_hc* s);
int vgh_synth_hc_reads(const vgh_synth_hc* s, uint64_t seed, int64_t n_reads, int32_t fixed_len, double sub_rate,
                       double indel_rate, int damage, int32_t origin_path, vgh_gam** out);

/* euka-shaped workload: clades own consecutive node ranges, 10-15 overlapping bins each. */
typedef struct vgh_synth_euka vgh_synth_euka;
int vgh_synth_euka_create(uint64_t seed, int32_t n_clades, int32_t genome_len, vgh_synth_euka** out);
/* clade table and bins verbatim (share/vgan/euka_dir/euka_db.clade / .bins: dist, snode..enode per
 * clade, bins as node-id ranges); node sequences synthetic; a read's clade drawn in proportion to its node-range size. */
int vgh_synth_euka_create_from_tables(uint64_t seed, int32_t n_clades, const double* dist, const int32_t* snode, const int32_t* enode,
                                      const int32_t* bin_off, const int32_t* bin_lo, const int32_t* bin_hi, vgh_synth_euka** out);
void vgh_synth_euka_destroy(vgh_synth_euka* s);
const vgh_graph* vgh_synth_euka_graph(const vgh_synth_euka* s);
int64_t vgh_synth_euka_graph_seq(const vgh_synth_euka* s, const int64_t** seq_off, const char** seq);
int32_t vgh_synth_euka_n_clades(const vgh_synth_euka* s);
const double* vgh_synth_euka_dist(const vgh_synth_euka* s);
const int32_t* vgh_synth_euka_bin_off(const vgh_synth_euka* s);
const int32_t* vgh_synth_euka_bin_lo(const vgh_synth_euka* s);
const int32_t* vgh_synth_euka_bin_hi(const vgh_synth_euka* s);
int vgh_synth_euka_reads(const vgh_synth_euka* s, uint64_t seed, int64_t n_reads, int32_t fixed_len, double indel_rate,
                         int damage, double n_rate, double softclip_rate, vgh_gam** out);

#ifdef __cplusplus
}
#endif
#endif /* VGAN_HOST_H */
