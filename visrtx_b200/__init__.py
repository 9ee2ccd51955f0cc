"""visrtx_b200 — B200-native direct-volume-rendering path behind VisRTX's ANARI surface.

The compute path lives in ``libdvr_b200.so`` (hand-written sm_100a CUDA behind the C-ABI declared in
``include/dvr_b200.h``) and the ANARI device in ``libanari_library_visrtx_b200.so``.  This Python
package is a thin ctypes mirror used by the tests and ``bench.py``; it contains no compute and
no CPU fallback: importing :mod:`visrtx_b200.capi` fails loudly when the native library is missing,
and every compute entry fails with ``DVR_ERR_NO_DEVICE`` without a GPU.
"""

__version__ = "0.1.0"
