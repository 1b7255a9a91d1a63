"""soibean per-read x per-path likelihood path over the C ABI (reference: analyse_GAM, src/getLCAfromGAM.h:102-604).

`Soibean.analyse` returns the R x P matrix of AlignmentInfo::pathMap values (what convertMapsToVector,
src/soibean.cpp:131-155, hands to the MCMC), columns in pathNames order.  `Soibean.mcmc_loglike` scores MCMC proposals on the batch the last
analyse() / analyse_raw() left in HBM (reference: read loop of MCMC::run_tree_proportion, src/MCMC.cpp:738-993).
No CPU path."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, gpu, host


class SoibeanTables:
    def __init__(self, member, min_id=0, deam5p_prof="", deam3p_prof="", penalty=7):
        member = np.ascontiguousarray(member, dtype=np.uint8)
        self.n_rows, self.n_paths = member.shape
        self.handle = C.c_void_p()
        check(host.vgh_soibean_tables_build(self.n_rows, int(min_id), self.n_paths, member.ctypes.data_as(C.c_void_p),
                                            (deam5p_prof or "").encode(), (deam3p_prof or "").encode(), int(penalty),
                                            C.byref(self.handle)))
        self.view = host.vgh_soibean_tables_view(self.handle).contents

    def __del__(self):
        if self.handle:
            host.vgh_soibean_tables_destroy(self.handle)
            self.handle = None


class SoibeanPackedBatch:
    def __init__(self, graph, alns, r0=0, r1=None, n_threads=0, pinned=False):
        r1 = len(alns) if r1 is None else r1
        self.n_span = int(r1 - r0)
        self.handle = C.c_void_p()
        check(host.vgh_soibean_pack(graph.handle, C.byref(alns.view), int(r0), int(r1), int(n_threads), int(pinned), C.byref(self.handle)))
        self.batch = host.vgh_euka_packed_batch(self.handle).contents

    @property
    def n_reads(self):
        return int(self.batch.n_reads)

    @property
    def nbytes(self):
        return int(gpu.vgl_euka_batch_bytes(C.byref(self.batch)))   # (the soibean batch has the euka batch's layout)

    def __del__(self):
        if self.handle:
            host.vgh_euka_packed_destroy(self.handle)
            self.handle = None


class Soibean:
    def __init__(self, device=0):
        self.ctx = C.c_void_p()
        check(gpu.vgl_create(int(device), C.byref(self.ctx)))
        self.tables = None

    def close(self):
        if self.ctx:
            gpu.vgl_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, tables):
        check(gpu.vgl_soibean_set_tables(self.ctx, C.byref(tables.view)))
        self.tables = tables

    def load_graph(self, graph):
        """Node sequences for analyse_raw (reconstruction on the device)."""
        mid, n = C.c_int64(), C.c_int64()
        off, seq = C.c_void_p(), C.c_void_p()
        check(_lib.host.vgh_graph_view(graph.handle, C.byref(mid), C.byref(n), C.byref(off), C.byref(seq)))
        check(gpu.vgl_hc_set_graph(self.ctx, mid.value, n.value, off, seq))

    def analyse_raw(self, raw):
        """As analyse(), from a RawBatch(euka_rule=True): reconstruction and packing happen on the device."""
        n, P = raw.n_span, self.tables.n_paths
        out = np.zeros((n, P), dtype=np.float64)
        used = np.zeros(n, dtype=np.uint8)
        check(gpu.vgl_soibean_run_raw(self.ctx, C.byref(raw.raw), n, out.ctypes.data_as(C.c_void_p), used.ctypes.data_as(C.c_void_p)))
        return out, used

    def analyse(self, packed, pinned=False):
        """pathMap of every read of the packed range: (matrix [n, P], used [n]).  pinned=True returns arrays in page-locked
        memory (the R x P matrix then comes back at PCIe speed)."""
        n, P = packed.n_span, self.tables.n_paths
        if pinned:
            from .host import pinned_empty
            out, used = pinned_empty((n, P), np.float64), pinned_empty((n,), np.uint8)
            out[...] = 0.0
            used[...] = 0
        else:
            out = np.zeros((n, P), dtype=np.float64)
            used = np.zeros(n, dtype=np.uint8)
        check(gpu.vgl_soibean_run(self.ctx, C.byref(packed.batch), n, out.ctypes.data_as(C.c_void_p), used.ctypes.data_as(C.c_void_p)))
        return out, used

    def set_stream(self, cuda_stream_ptr):
        check(gpu.vgl_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def analyse_into(self, packed, out, used):
        """analyse() into caller-provided (page-locked) arrays [n, P] float64 / [n] uint8."""
        check(gpu.vgl_soibean_run(self.ctx, C.byref(packed.batch), packed.n_span, out.ctypes.data_as(C.c_void_p), used.ctypes.data_as(C.c_void_p)))
        return out, used

    def rerun(self, n_out, out=None, used=None):
        """The kernels again on the batch the last analyse*() left in HBM.  With out=None the R x P matrix stays on the device
        (where mcmc_loglike reads what it needs): no device -> host traffic."""
        check(gpu.vgl_soibean_rerun(self.ctx, int(n_out), out.ctypes.data_as(C.c_void_p) if out is not None else None,
                                    used.ctypes.data_as(C.c_void_p) if used is not None else None))

    def kernel_ms(self):
        """Device time of the last soibean kernel pair (edit + path kernels; CUDA events on their stream)."""
        ms = C.c_float()
        check(gpu.vgl_sibling_kernel_ms(self.ctx, C.byref(ms)))
        return float(ms.value)

    def mcmc_loglike(self, proposals, freqs7, con, n_out=0, per_read=False):
        """Log-likelihood of every proposal over the reads of the last analyse() / analyse_raw() batch.
        proposals: [n_proposals][n_sources] of (child path, parent path, pos_branch, dist, proportion);
        freqs7 = params.freqs['A','C','G','T','R','Y','M'] (`soibean_freqs7`); con as src/soibean.cpp:596-598.
        Returns loglike[n_proposals] (and the per-read contributions [n_proposals, n_out] with per_read=True)."""
        nq = len(proposals)
        K = len(proposals[0]) if nq else 1
        arr = (_lib.vgl_soibean_source * max(nq * K, 1))()
        for q, prop in enumerate(proposals):
            if len(prop) != K:
                raise ValueError("every proposal needs the same number of sources")
            for y, (c, p, b, d, w) in enumerate(prop):
                arr[q * K + y] = _lib.vgl_soibean_source(int(c), int(p), float(b), float(d), float(w))
        f7 = np.ascontiguousarray(freqs7, dtype=np.float64)
        if f7.shape != (7,):
            raise ValueError("freqs7 is A, C, G, T, R, Y, M")
        out = np.zeros(nq, dtype=np.float64)
        pr = np.zeros((nq, n_out), dtype=np.float64) if per_read else None
        check(gpu.vgl_soibean_mcmc_loglike(self.ctx, nq, K, C.cast(arr, C.c_void_p), f7.ctypes.data_as(C.c_void_p), float(con), int(n_out),
                                           pr.ctypes.data_as(C.c_void_p) if per_read else None, out.ctypes.data_as(C.c_void_p)))
        return (out, pr) if per_read else out


def soibean_freqs7(a, c, g, t):
    """params.freqs['A','C','G','T','R','Y','M'] from the four base frequencies, as src/soibean.cpp:629-640 derives them
    (the 22 in mu is the transition / transversion ratio the reference hard-codes)."""
    R, Y = a + g, c + t
    M = 1 / (2 * (((22) * (a * g)) + ((22) * (c * t)) + (a * c + (a * t) + (g * c + (g * t)))))
    return np.array([a, c, g, t, R, Y, M], dtype=np.float64)
