"""vgan_b200 -- B200-native read x path likelihood engine for vgan (HaploCart hot path).

Layout: csrc/ (CUDA kernels + C ABI + C++ host side), lib/ (built .so files, in-tree),
_lib.py (ctypes), host.py (handles over the host library), haplocart.py (the reference-shaped
interface over the device library).  No CPU fallback exists."""
from . import _lib  # noqa: F401  (raises if the libraries are not built)
from .haplocart import DeviceBatch, Haplocart  # noqa: F401
from .host import Alignments, Feeder, GamFrames, Graph, HcTables, HcTree, PackedBatch, RawBatch, WireBatch, SynthHaplocart, reconstruct  # noqa: F401
from .euka import Euka, EukaPackedBatch, EukaTables, SynthEuka, DHIGH5P_PROF, DHIGH3P_PROF  # noqa: F401
from .soibean import Soibean, SoibeanPackedBatch, SoibeanTables, soibean_freqs7  # noqa: F401
