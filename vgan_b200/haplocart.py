"""Host-side mirror of the reference's HaploCart hot-path interface over the C ABI.

The reference's call sites are
    Haplocart::update(...) per read + `final_vec[j] +=`          src/HaploCart.cpp:408-421
    std::max_element -> predicted haplotype                       src/HaploCart.cpp:423-424
    Haplocart::get_posterior(final_vec, path_names, parents, ...) src/get_posterior.cpp:87-127
`Haplocart` below keeps those names; every method forwards to libvgan_gpu.so (include/vgan_gpu.h).
There is no CPU path here: construction fails without a CUDA device.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, gpu


class Haplocart:
    def __init__(self, device=0):
        self.ctx = C.c_void_p()
        check(gpu.vgl_create(int(device), C.byref(self.ctx)))
        self.device = int(device)
        self.n_paths = 0
        self.tree = None
        self.path_names = None
        self._max_post = 1

    def close(self):
        if self.ctx:
            gpu.vgl_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables (reference: loaded once, src/HaploCart.cpp:279-286) -------------------------------
    def load(self, tables, tree=None, path_names=None):
        check(gpu.vgl_hc_set_tables(self.ctx, C.byref(tables.view)))
        self.n_paths = int(tables.n_paths)
        self.path_names = list(path_names) if path_names is not None else (tree.path_names if tree else None)
        self.tree = tree
        if tree is not None:
            bits = np.ascontiguousarray(tree.clade_bits())
            check(gpu.vgl_hc_set_tree(self.ctx, tree.n_clades, bits.ctypes.data_as(C.c_void_p),
                                      tree.anc_off.ctypes.data_as(C.c_void_p),
                                      tree.anc_clade.ctypes.data_as(C.c_void_p) if tree.anc_clade.size else
                                      np.zeros(1, np.int32).ctypes.data_as(C.c_void_p)))
            self._max_post = 1 + int(np.max(np.diff(tree.anc_off))) if tree.anc_off.size > 1 else 1

    def load_graph(self, graph):
        """Node sequences for the device-side reconstruction (update_raw)."""
        mid, n = C.c_int64(), C.c_int64()
        off, seq = C.c_void_p(), C.c_void_p()
        check(_lib.host.vgh_graph_view(graph.handle, C.byref(mid), C.byref(n), C.byref(off), C.byref(seq)))
        check(gpu.vgl_hc_set_graph(self.ctx, mid.value, n.value, off, seq))

    def set_stream(self, cuda_stream_ptr):
        check(gpu.vgl_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    # ---- per-sample hot loop ------------------------------------------------------------------------
    def reset(self):
        """fill(final_vec, 0) -- start of a sample (src/HaploCart.cpp:481)."""
        check(gpu.vgl_hc_reset(self.ctx))

    def update(self, packed):
        """Haplocart::update for every read of a packed host batch, summed into final_vec."""
        check(gpu.vgl_hc_accumulate(self.ctx, C.byref(packed.batch)))

    def update_wire(self, wire):
        """Haplocart::update for a batch in its wire form (WireBatch): fewer bytes over PCIe, expanded on the device."""
        check(gpu.vgl_hc_accumulate_wire(self.ctx, C.byref(wire.wire)))

    def update_raw(self, raw):
        """Haplocart::update for the raw alignments of a RawBatch: reconstruction and packing happen on the device."""
        check(gpu.vgl_hc_accumulate_raw(self.ctx, C.byref(raw.raw)))

    def pack_raw(self, raw):
        """Reconstruct and pack on the device only; returns the device batch (a `_lib.vgl_hc_batch` of device pointers,
        valid until the next raw call on this object) for inspection."""
        b = _lib.vgl_hc_batch()
        check(gpu.vgl_hc_pack_raw(self.ctx, C.byref(raw.raw), C.byref(b)))
        return b

    def download(self, dev_ptr, n, dtype):
        """Copy n items of dtype from device memory (tests)."""
        out = np.zeros(n, dtype=dtype)
        if n:
            check(gpu.vgl_memcpy_d2h(self.ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(dev_ptr), out.nbytes))
        return out

    def update_device(self, batch_struct):
        """Same, for a batch whose arrays already live in HBM (a `_lib.vgl_hc_batch` of device pointers)."""
        check(gpu.vgl_hc_accumulate_device(self.ctx, C.byref(batch_struct)))

    def paths(self, ll_dev_ptr=None):
        """Collapse into this GPU's final_vec[P] on the device (optionally copied to `ll_dev_ptr`)."""
        check(gpu.vgl_hc_paths(self.ctx, C.c_void_p(ll_dev_ptr) if ll_dev_ptr else None))

    def set_paths_from_parts(self, parts_dev_ptr, n_parts):
        check(gpu.vgl_hc_set_paths_from_parts(self.ctx, C.c_void_p(parts_dev_ptr), int(n_parts)))

    # ---- multi-GPU exchange through NVLink peer memory (include/vgan_gpu.h) ---------------------------
    def peer_export(self, world):
        """Allocate this rank's exchange buffer; returns its IPC handle (bytes) for the other ranks."""
        buf = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        check(gpu.vgl_hc_peer_export(self.ctx, int(world), buf))
        return bytes(buf)

    def peer_attach(self, rank, world, handles):
        """handles: the `world` exported handles in rank order."""
        blob = b"".join(handles)
        assert len(blob) == world * _lib.PEER_HANDLE_BYTES
        check(gpu.vgl_hc_peer_attach(self.ctx, int(rank), int(world), C.c_char_p(blob)))

    def exchange_paths(self):
        """Every rank's final_vec summed in rank order, on every rank (all ranks call it once per sample)."""
        check(gpu.vgl_hc_exchange_paths(self.ctx))

    def get_posterior(self, want_ll=True):
        """argmax + Haplocart::get_posterior.  Returns a dict with final_vec, the call and its clade posteriors."""
        ll = np.zeros(self.n_paths, dtype=np.float64) if want_ll else None
        arg, npost, total = C.c_int32(), C.c_int32(), C.c_double()
        post = np.zeros(max(self._max_post, 1), dtype=np.float64)
        check(gpu.vgl_hc_finalize(self.ctx, ll.ctypes.data_as(C.c_void_p) if want_ll else None, C.byref(arg),
                                  C.byref(total), post.ctypes.data_as(C.c_void_p), post.size, C.byref(npost)))
        res = {"final_vec": ll, "maxh_index": arg.value, "total_ll": total.value,
               "confidence_vec": post[:npost.value].copy()}
        if self.path_names is not None:
            res["predicted_haplotype"] = self.path_names[arg.value]
            if self.tree is not None:
                res["clade_vec"] = [res["predicted_haplotype"]] + self.tree.clades_of(arg.value)
        return res

    def posterior_begin(self):
        """First half of get_posterior(): enqueue the collapse, the posterior kernel and the copy of the result; returns a ticket.
        The next sample's reset() / update*() may follow at once (at most two results in flight)."""
        t = C.c_int32()
        check(gpu.vgl_hc_finalize_begin(self.ctx, C.byref(t)))
        return int(t.value)

    def posterior_end(self, ticket):
        """Second half: wait for that result alone and return it (as get_posterior(want_ll=False))."""
        arg, npost, total = C.c_int32(), C.c_int32(), C.c_double()
        post = np.zeros(max(self._max_post, 1), dtype=np.float64)
        check(gpu.vgl_hc_finalize_end(self.ctx, int(ticket), C.byref(arg), C.byref(total), post.ctypes.data_as(C.c_void_p), post.size,
                                      C.byref(npost)))
        res = {"final_vec": None, "maxh_index": arg.value, "total_ll": total.value, "confidence_vec": post[:npost.value].copy()}
        if self.path_names is not None:
            res["predicted_haplotype"] = self.path_names[arg.value]
            if self.tree is not None:
                res["clade_vec"] = [res["predicted_haplotype"]] + self.tree.clades_of(arg.value)
        return res

    # ---- several samples per pass (the reference loops over samples, src/HaploCart.cpp:303) --------------
    def samples_begin(self, n_samples):
        """Start a pass over n_samples samples (instead of reset()); sample 0 is selected."""
        check(gpu.vgl_hc_samples_begin(self.ctx, int(n_samples)))
        self._n_samples = int(n_samples)

    def sample_select(self, sample):
        """The update*() calls / Feeder.feed that follow add into this sample."""
        check(gpu.vgl_hc_sample_select(self.ctx, int(sample)))

    def finalize_samples(self, want_ll=False):
        """One collapse + posterior over all samples of the pass, one copy back.  Returns a list of per-sample dicts as
        get_posterior() gives them (plus "status": 0, or the error code that sample alone would have raised)."""
        S, P, cap = self._n_samples, self.n_paths, max(self._max_post, 1)
        ll = np.zeros((S, P), dtype=np.float64) if want_ll else None
        arg, npost, status = np.zeros(S, dtype=np.int32), np.zeros(S, dtype=np.int32), np.zeros(S, dtype=np.int32)
        total, post = np.zeros(S), np.zeros((S, cap))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(gpu.vgl_hc_finalize_samples(self.ctx, p(ll) if want_ll else None, p(arg), p(total), p(post), cap, p(npost), p(status)))
        out = []
        for k in range(S):
            res = {"final_vec": ll[k] if want_ll else None, "maxh_index": int(arg[k]), "total_ll": float(total[k]),
                   "confidence_vec": post[k, :npost[k]].copy(), "status": int(status[k])}
            if self.path_names is not None and status[k] == 0:
                res["predicted_haplotype"] = self.path_names[int(arg[k])]
                if self.tree is not None:
                    res["clade_vec"] = [res["predicted_haplotype"]] + self.tree.clades_of(int(arg[k]))
            out.append(res)
        return out

    def run_sample(self, batches, want_ll=True):
        """One sample end to end: reset, update over the packed host batches, posterior."""
        self.reset()
        for b in batches:
            self.update(b)
        return self.get_posterior(want_ll=want_ll)

    # ---- instrumentation ----------------------------------------------------------------------------
    def synchronize(self):
        check(gpu.vgl_synchronize(self.ctx))

    def update_gam(self, frames):
        """Haplocart::update for every alignment of a GAM stream whose framing the host has walked (host.GamFrames): the messages
        are parsed on the device.  Returns the number of reads that entered the sample."""
        n = C.c_uint64()
        check(gpu.vgl_hc_accumulate_gam(self.ctx, frames.bytes_ptr, frames.n_bytes, frames.off_ptr, frames.len_ptr, frames.n_msgs, C.byref(n)))
        return int(n.value)

    def decode_gam(self, frames):
        """The device decoder's output for (at most one chunk of) a GAM stream, as a dict of flat numpy arrays (tests)."""
        tot = np.zeros(6, dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        cap_m = max(frames.n_bytes, 16)   # generous: every mapping / edit / byte of the output costs at least one input byte
        n = frames.n_msgs
        out = {"qual_off": np.zeros(n + 1, np.uint32), "qual": np.zeros(cap_m, np.uint8), "mapq": np.zeros(max(n, 1), np.uint8),
               "map_off": np.zeros(n + 1, np.uint32), "map_node": np.zeros(cap_m, np.uint32), "map_offset": np.zeros(cap_m, np.uint32),
               "edit_off": np.zeros(cap_m + 1, np.uint32), "edit_from": np.zeros(cap_m, np.uint16), "edit_to": np.zeros(cap_m, np.uint16),
               "edit_seq_off": np.zeros(cap_m + 1, np.uint32), "edit_seq": np.zeros(cap_m, np.uint8), "seq_len": np.zeros(max(n, 1), np.uint16),
               "read_index": np.zeros(max(n, 1), np.uint32), "str_off": np.zeros(n + 1, np.uint32)}
        check(gpu.vgl_gam_decode(self.ctx, frames.bytes_ptr, frames.n_bytes, frames.off_ptr, frames.len_ptr, n, p(tot), p(out["qual_off"]),
                                 p(out["qual"]), p(out["mapq"]), p(out["map_off"]), p(out["map_node"]), p(out["map_offset"]),
                                 p(out["edit_off"]), p(out["edit_from"]), p(out["edit_to"]), p(out["edit_seq_off"]), p(out["edit_seq"]),
                                 p(out["seq_len"]), p(out["read_index"]), p(out["str_off"])))
        R, M, E, Q, Sq = (int(x) for x in tot[:5])
        cut = {"qual_off": R + 1, "qual": Q, "mapq": R, "map_off": R + 1, "map_node": M, "map_offset": M, "edit_off": M + 1, "edit_from": E,
               "edit_to": E, "edit_seq_off": E + 1, "edit_seq": Sq, "seq_len": R, "read_index": R, "str_off": R + 1}
        res = {k: v[:cut[k]].copy() for k, v in out.items()}
        res["n_reads"], res["capacity"] = R, int(tot[5])
        return res

    def dedup(self, alns):
        """Dup_Remover::remove_duplicates_internal (src/rmdup.cpp:20-110) on the device: (is_dup mask, reads kept)."""
        n = len(alns)
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        check(_lib.host.vgh_dedup_keys(C.byref(alns.view), keys.ctypes.data_as(C.c_void_p)))
        pz = alns.view.paired if n else None          # the library's own array (may be NULL: no read is paired)
        mask = np.zeros(max(n, 1), dtype=np.uint8)
        kept = C.c_uint64()
        check(gpu.vgl_dedup(self.ctx, n, keys.ctypes.data_as(C.c_void_p), pz, mask.ctypes.data_as(C.c_void_p), C.byref(kept)))
        return mask[:n].astype(bool), int(kept.value)

    def logterm_launches(self):
        """Read-kernel launches so far that took the log-term kernel (the batch's meta allowed it)."""
        return int(gpu.vgl_hc_logterm_launches(self.ctx))

    def launch_count(self):
        return int(gpu.vgl_launch_count(self.ctx))

    def read_kernel_ms(self):
        ms, n = C.c_float(), C.c_int32()
        check(gpu.vgl_hc_read_kernel_ms(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- device-resident batches --------------------------------------------------------------------
    def upload(self, packed):
        """Copy a packed host batch into HBM once; returns a DeviceBatch usable with update_device."""
        return DeviceBatch(self, packed)


class DeviceBatch:
    def __init__(self, hc, packed):
        self.hc = hc
        b = packed.batch
        G, L = int(b.n_groups), _lib.LANES
        sizes = {"grp_col_blk": (G + 1) * 4, "grp_map_blk": (G + 1) * 4, "col": int(b.n_col_blks) * 512,
                 "map": int(b.n_map_blks) * 512, "n_cols": G * L * 2, "n_close": G * L * 2, "n_extra": G * L * 2,
                 "mapq": G * L}
        self.batch = _lib.vgl_hc_batch()
        self.batch.n_reads, self.batch.n_groups = b.n_reads, b.n_groups
        self.batch.n_col_blks, self.batch.n_map_blks = b.n_col_blks, b.n_map_blks
        self.batch.meta = b.meta     # value ranges of the batch: lets the library take the log-term read kernel
        self.ptrs = []
        self.nbytes = 0
        for name, n in sizes.items():
            p = C.c_void_p()
            check(gpu.vgl_device_alloc(hc.ctx, C.byref(p), n + 16))
            self.ptrs.append(p)
            if n:
                check(gpu.vgl_memcpy_h2d(hc.ctx, p, C.c_void_p(getattr(b, name)), n))
            setattr(self.batch, name, p.value)
            self.nbytes += n

    def free(self):
        for p in self.ptrs:
            gpu.vgl_device_free(self.hc.ctx, p)
        self.ptrs = []

    def __del__(self):
        try:
            if self.hc.ctx:
                self.free()
        except Exception:
            pass
