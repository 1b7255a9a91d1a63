"""ctypes bindings of the two in-tree C-ABI libraries (include/vgan_gpu.h, include/vgan_host.h).

There is no CPU fallback: if the libraries are missing, importing this module raises, and on a box
without a CUDA device `vgl_create` fails with VGL_ERR_NO_DEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("VGAN_B200_LIB_DIR") or os.path.join(_HERE, "lib")  # (override: kernel A/B builds)
GPU_LIB_PATH = os.path.join(LIB_DIR, "libvgan_gpu.so")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libvgan_host.so")


class VglError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vgan_gpu error %d: %s" % (code, msg))
        self.code = code


class vgl_hc_tables(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_paths", C.c_int32), ("row_words", C.c_int32), ("reserved0", C.c_int32),
                ("support_bits", C.c_void_p), ("node_mappability", C.c_void_p), ("node_match_prob", C.c_void_p),
                ("node_valid", C.c_void_p), ("qscore", C.c_double * 100), ("p_incorrect_map", C.c_double * 100),
                ("bg_freq", C.c_double * 4), ("background_error_prob", C.c_double), ("is_consensus_fasta", C.c_int32),
                ("reserved1", C.c_int32)]


class vgl_hc_batch_meta(C.Structure):
    _fields_ = [("valid", C.c_uint32), ("qual_max", C.c_uint32), ("mapq_mask", C.c_uint32 * 4), ("n_segs", C.c_uint32),
                ("seg_mapq", C.c_uint8 * 16), ("seg_end", C.c_uint32 * 16)]


class vgl_hc_wire(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_groups", C.c_uint32), ("n_col_blks", C.c_uint64), ("n_node_blks", C.c_uint64),
                ("n_map_blks", C.c_uint64), ("grp_col_blk", C.c_void_p), ("grp_node_blk", C.c_void_p), ("grp_map_blk", C.c_void_p),
                ("col", C.c_void_p), ("node", C.c_void_p), ("map", C.c_void_p), ("n_cols", C.c_void_p), ("n_close", C.c_void_p),
                ("n_extra", C.c_void_p), ("mapq", C.c_void_p), ("col8", C.c_void_p), ("flag4", C.c_void_p), ("n_owned", C.c_void_p),
                ("meta", vgl_hc_batch_meta)]


class vgl_hc_raw(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_maps", C.c_uint32), ("n_edits", C.c_uint32), ("reserved", C.c_uint32),
                ("qual_off", C.c_void_p), ("qual", C.c_void_p), ("mapq", C.c_void_p), ("str_off", C.c_void_p),
                ("map_off", C.c_void_p), ("map_node", C.c_void_p), ("map_offset", C.c_void_p), ("edit_off", C.c_void_p),
                ("edit_from", C.c_void_p), ("edit_to", C.c_void_p), ("edit_seq_off", C.c_void_p), ("edit_seq", C.c_void_p),
                ("seq_len", C.c_void_p), ("read_index", C.c_void_p),
                ("map_offset16", C.c_void_p), ("map_n_edits8", C.c_void_p), ("edit_slen8", C.c_void_p),
                ("edit_from8", C.c_void_p), ("edit_to8", C.c_void_p), ("map_node16", C.c_void_p), ("first_offset16", C.c_void_p),
                ("meta", vgl_hc_batch_meta)]


class vgl_hc_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_groups", C.c_uint32), ("n_col_blks", C.c_uint64), ("n_map_blks", C.c_uint64),
                ("grp_col_blk", C.c_void_p), ("grp_map_blk", C.c_void_p), ("col", C.c_void_p), ("map", C.c_void_p),
                ("n_cols", C.c_void_p), ("n_close", C.c_void_p), ("n_extra", C.c_void_p), ("mapq", C.c_void_p),
                ("meta", vgl_hc_batch_meta)]


# column / mapping record fields (include/vgan_gpu.h)
COL_MISMATCH, COL_LATCHED, COL_VALID, COL_OPEN, COL_CLOSE = 0x0001, 0x0800, 0x1000, 0x2000, 0x4000
COL_BASE_SHIFT, COL_QUAL_SHIFT, COL_NODE_SHIFT, NO_NODE = 1, 3, 15, 0x1FFFF
MAP_NODE_BITS = 20
LANES, COL_BLK, MAP_BLK = 32, 4, 4
PEER_HANDLE_BYTES = 64
EUKA_COL_BLK = 8


class vgl_euka_tables(C.Structure):
    _fields_ = [("n_clades", C.c_int32), ("n_bins", C.c_int32), ("max_node", C.c_int32), ("k5", C.c_int32), ("k3", C.c_int32),
                ("minimum_mq", C.c_uint32), ("clade_dist", C.c_void_p), ("bin_off", C.c_void_p), ("bin_lo", C.c_void_p),
                ("bin_hi", C.c_void_p), ("deam", C.c_void_p), ("qscore", C.c_double * 100), ("base_freq_log", C.c_double * 5),
                ("tt_ratio", (C.c_double * 4) * 4)]


class vgl_euka_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_groups", C.c_uint32), ("n_col_blks", C.c_uint64), ("n_map_blks", C.c_uint64),
                ("grp_col_blk", C.c_void_p), ("grp_map_blk", C.c_void_p), ("col", C.c_void_p), ("map", C.c_void_p),
                ("n_cols", C.c_void_p), ("n_maps", C.c_void_p), ("seq_len", C.c_void_p), ("mapq", C.c_void_p),
                ("flags", C.c_void_p), ("read_index", C.c_void_p)]


class vgl_soibean_source(C.Structure):
    _fields_ = [("child", C.c_int32), ("parent", C.c_int32), ("pos_branch", C.c_double), ("dist", C.c_double), ("proportion", C.c_double)]


class vgl_soibean_tables(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("min_id", C.c_int32), ("n_paths", C.c_int32), ("row_words", C.c_int32), ("k5", C.c_int32),
                ("k3", C.c_int32), ("penalty", C.c_int32), ("reserved", C.c_int32), ("member_bits", C.c_void_p), ("deam", C.c_void_p),
                ("qscore", C.c_double * 100)]


class vgh_alignments(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("qual_off", C.c_void_p),
                ("qual", C.c_void_p), ("mapq", C.c_void_p), ("identity", C.c_void_p), ("paired", C.c_void_p),
                ("map_off", C.c_void_p), ("map_node", C.c_void_p), ("map_offset", C.c_void_p), ("map_rev", C.c_void_p),
                ("edit_off", C.c_void_p), ("edit_from", C.c_void_p), ("edit_to", C.c_void_p),
                ("edit_seq_off", C.c_void_p), ("edit_seq", C.c_void_p)]


# every symbol include/vgan_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
GPU_SYMBOLS = {
    "vgl_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vgl_destroy": (None, [_P]),
    "vgl_last_error": (C.c_char_p, []),
    "vgl_abi_version": (C.c_int, []),
    "vgl_set_stream": (C.c_int, [_P, _P]),
    "vgl_synchronize": (C.c_int, [_P]),
    "vgl_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "vgl_host_free": (C.c_int, [_P]),
    "vgl_device_alloc": (C.c_int, [_P, C.POINTER(_P), C.c_size_t]),
    "vgl_device_free": (C.c_int, [_P, _P]),
    "vgl_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "vgl_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "vgl_hc_set_tables": (C.c_int, [_P, C.POINTER(vgl_hc_tables)]),
    "vgl_hc_set_tree": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "vgl_hc_batch_bytes": (C.c_uint64, [C.POINTER(vgl_hc_batch)]),
    "vgl_hc_reset": (C.c_int, [_P]),
    "vgl_hc_accumulate": (C.c_int, [_P, C.POINTER(vgl_hc_batch)]),
    "vgl_hc_accumulate_device": (C.c_int, [_P, C.POINTER(vgl_hc_batch)]),
    "vgl_hc_paths": (C.c_int, [_P, _P]),
    "vgl_hc_set_paths_from_parts": (C.c_int, [_P, _P, C.c_int32]),
    "vgl_hc_wire_bytes": (C.c_uint64, [_P]),
    "vgl_hc_accumulate_wire": (C.c_int, [_P, _P]),
    "vgl_hc_raw_bytes": (C.c_uint64, [_P]),
    "vgl_hc_set_graph": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "vgl_hc_pack_raw": (C.c_int, [_P, _P, _P]),
    "vgl_hc_accumulate_raw": (C.c_int, [_P, _P]),
    "vgl_hc_peer_export": (C.c_int, [_P, C.c_int32, _P]),
    "vgl_hc_peer_attach": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "vgl_hc_exchange_paths": (C.c_int, [_P]),
    "vgl_hc_finalize": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_double), _P, C.c_int32,
                                  C.POINTER(C.c_int32)]),
    "vgl_hc_finalize_begin": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "vgl_hc_finalize_end": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double), _P, C.c_int32, C.POINTER(C.c_int32)]),
    "vgl_euka_set_tables": (C.c_int, [_P, C.POINTER(vgl_euka_tables)]),
    "vgl_euka_reset": (C.c_int, [_P]),
    "vgl_euka_run": (C.c_int, [_P, C.POINTER(vgl_euka_batch), C.c_uint32, _P, _P, _P, _P, _P]),
    "vgl_euka_results": (C.c_int, [_P, _P, _P]),
    "vgl_euka_like_sums": (C.c_int, [_P, _P, _P]),
    "vgl_euka_proposal_loglike": (C.c_double, [C.c_int32, _P, _P, _P, _P]),
    "vgl_euka_batch_bytes": (C.c_uint64, [C.POINTER(vgl_euka_batch)]),
    "vgl_soibean_set_tables": (C.c_int, [_P, C.POINTER(vgl_soibean_tables)]),
    "vgl_soibean_run": (C.c_int, [_P, C.POINTER(vgl_euka_batch), C.c_uint32, _P, _P]),
    "vgl_soibean_run_raw": (C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    "vgl_soibean_rerun": (C.c_int, [_P, C.c_uint32, _P, _P]),
    "vgl_euka_rerun": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P]),
    "vgl_sibling_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "vgl_soibean_mcmc_loglike": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_double, C.c_uint32, _P, _P]),
    "vgl_euka_run_raw": (C.c_int, [_P, _P, C.c_uint32, _P, _P, _P, _P, _P]),
    "vgl_launch_count": (C.c_uint64, [_P]),
    "vgl_hc_logterm_launches": (C.c_uint64, [_P]),
    "vgl_hc_accumulate_gam": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "vgl_gam_decode": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64] + [_P] * 15),
    "vgl_dedup": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.POINTER(C.c_uint64)]),
    "vgl_hc_samples_begin": (C.c_int, [_P, C.c_int32]),
    "vgl_hc_sample_select": (C.c_int, [_P, C.c_int32]),
    "vgl_hc_finalize_samples": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "vgl_hc_read_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

HOST_SYMBOLS = {
    "vgh_last_error": (C.c_char_p, []),
    "vgh_graph_create": (C.c_int, [C.c_int64, C.c_int64, _P, _P, C.POINTER(_P)]),
    "vgh_graph_destroy": (None, [_P]),
    "vgh_graph_view": (C.c_int, [_P, _P, _P, _P, _P]),
    "vgh_hc_wire_from_packed": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "vgh_hc_wire_view": (C.POINTER(vgl_hc_wire), [_P]),
    "vgh_hc_wire_destroy": (None, [_P]),
    "vgh_euka_raw_build": (C.c_int, [C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int, C.POINTER(_P)]),
    "vgh_hc_raw_build": (C.c_int, [C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int, C.POINTER(_P)]),
    "vgh_gam_frame": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "vgh_gam_frames_view": (C.c_int64, [_P, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(_P), C.POINTER(_P)]),
    "vgh_gam_frames_destroy": (None, [_P]),
    "vgh_dedup_keys": (C.c_int, [C.POINTER(vgh_alignments), _P]),
    "vgh_hc_feeder_create": (C.c_int, [C.c_int, C.c_int64, C.POINTER(_P)]),
    "vgh_hc_feed": (C.c_int, [_P, _P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64]),
    "vgh_hc_feeder_marshal": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.POINTER(C.POINTER(vgl_hc_raw))]),
    "vgh_euka_feeder_marshal": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.POINTER(C.POINTER(vgl_hc_raw))]),
    "vgh_euka_feed": (C.c_int, [_P, _P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, _P, _P, _P, _P, _P]),
    "vgh_soibean_feed": (C.c_int, [_P, _P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "vgh_hc_feeder_stats": (None, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64)]),
    "vgh_hc_feeder_destroy": (None, [_P]),
    "vgh_hc_raw_view": (C.POINTER(vgl_hc_raw), [_P]),
    "vgh_hc_raw_skipped": (C.c_int64, [_P]),
    "vgh_hc_raw_destroy": (None, [_P]),
    "vgh_gam_read_file": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "vgh_gam_parse": (C.c_int, [_P, C.c_size_t, C.c_int, C.POINTER(_P)]),
    "vgh_gam_view": (C.POINTER(vgh_alignments), [_P]),
    "vgh_gam_destroy": (None, [_P]),
    "vgh_reconstruct": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, _P, _P, _P, C.c_int32,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vgh_dedup": (C.c_int, [C.POINTER(vgh_alignments), _P, C.POINTER(C.c_int64)]),
    "vgh_hc_pack": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "vgh_hc_packed_batch": (C.POINTER(vgl_hc_batch), [_P]),
    "vgh_hc_packed_skipped": (C.c_int64, [_P]),
    "vgh_hc_packed_destroy": (None, [_P]),
    "vgh_hc_tables_build": (C.c_int, [C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, C.c_int64, _P, C.c_double, C.c_int,
                                      C.POINTER(_P)]),
    "vgh_hc_tables_view": (C.POINTER(vgl_hc_tables), [_P]),
    "vgh_hc_tables_destroy": (None, [_P]),
    "vgh_hc_region_match_prob": (C.c_double, [C.c_int32]),
    "vgh_hc_tree_build": (C.c_int, [C.c_int32, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "vgh_hc_tree_n_clades": (C.c_int32, [_P]),
    "vgh_hc_tree_row_words": (C.c_int32, [_P]),
    "vgh_hc_tree_clade_bits": (_P, [_P]),
    "vgh_hc_tree_anc_off": (_P, [_P]),
    "vgh_hc_tree_anc_clade": (_P, [_P]),
    "vgh_hc_tree_clade_name": (C.c_char_p, [_P, C.c_int32]),
    "vgh_hc_tree_destroy": (None, [_P]),
    "vgh_euka_tables_build": (C.c_int, [C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(_P)]),
    "vgh_euka_tables_view": (C.POINTER(vgl_euka_tables), [_P]),
    "vgh_euka_tables_destroy": (None, [_P]),
    "vgh_euka_pack": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "vgh_euka_packed_batch": (C.POINTER(vgl_euka_batch), [_P]),
    "vgh_euka_packed_destroy": (None, [_P]),
    "vgh_soibean_tables_build": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(_P)]),
    "vgh_soibean_tables_view": (C.POINTER(vgl_soibean_tables), [_P]),
    "vgh_soibean_tables_destroy": (None, [_P]),
    "vgh_soibean_pack": (C.c_int, [_P, C.POINTER(vgh_alignments), C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "vgh_synth_euka_create": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "vgh_synth_euka_create_from_tables": (C.c_int, [C.c_uint64, C.c_int32, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "vgh_synth_euka_destroy": (None, [_P]),
    "vgh_synth_euka_graph": (_P, [_P]),
    "vgh_synth_euka_graph_seq": (C.c_int64, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "vgh_synth_euka_n_clades": (C.c_int32, [_P]),
    "vgh_synth_euka_dist": (_P, [_P]),
    "vgh_synth_euka_bin_off": (_P, [_P]),
    "vgh_synth_euka_bin_lo": (_P, [_P]),
    "vgh_synth_euka_bin_hi": (_P, [_P]),
    "vgh_synth_euka_reads": (C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int32, C.c_double, C.c_int, C.c_double, C.c_double,
                                       C.POINTER(_P)]),
    "vgh_synth_hc_create": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "vgh_synth_hc_destroy": (None, [_P]),
    "vgh_synth_hc_graph": (_P, [_P]),
    "vgh_synth_hc_graph_seq": (C.c_int64, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "vgh_synth_hc_supports": (_P, [_P]),
    "vgh_synth_hc_n_pmap": (C.c_int64, [_P]),
    "vgh_synth_hc_pmap_node": (_P, [_P]),
    "vgh_synth_hc_pmap_val": (_P, [_P]),
    "vgh_synth_hc_n_mappab": (C.c_int64, [_P]),
    "vgh_synth_hc_mappab": (_P, [_P]),
    "vgh_synth_hc_names": (C.c_char_p, [_P]),
    "vgh_synth_hc_parents": (C.c_char_p, [_P]),
    "vgh_synth_hc_children": (C.c_char_p, [_P]),
    "vgh_synth_hc_total_len": (C.c_int32, [_P]),
    "vgh_synth_hc_reads": (C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int, C.c_int32,
                                     C.POINTER(_P)]),
}


def _load(path, symbols):
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `make -C vgan_b200/csrc` (or __graft_entry__.build()); "
                          "there is no CPU fallback" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


gpu = _load(GPU_LIB_PATH, GPU_SYMBOLS)
host = _load(HOST_LIB_PATH, HOST_SYMBOLS)


def check(rc):
    """Raise on a non-zero status from either library."""
    if rc == 0:
        return
    if rc >= 100:
        raise VglError(rc, host.vgh_last_error().decode(errors="replace"))
    raise VglError(rc, gpu.vgl_last_error().decode(errors="replace"))
