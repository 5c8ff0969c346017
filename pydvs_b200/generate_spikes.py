"""Drop-in for ``pydvs.generate_spikes`` (``pydvs/generate_spikes.pyx``) on B200.

Same module-level names and positional signatures as the Cython module; the
arithmetic runs in the sm_100a kernels of ``libpydvs_b200.so``.  Hand it CUDA tensors (``[H, W]`` or
``[S, H, W]``) and it returns CUDA tensors; hand it NumPy arrays and it uploads, computes on the GPU
and downloads, so ``stand_alone*.py``, ``VirtualCam`` and ``external_dvs_emulator_device`` run unchanged:

    import pydvs_b200.generate_spikes as gs
"""
from ._api import (DTYPE, DTYPE_FLOAT, DTYPE_IDX, DTYPE_U8, DTYPE_U16, KEY_SPINNAKER, KEY_XYP,  # noqa: F401
                   build_api)
from ._helpers import build_helpers

UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY = 0, 1, 2, 3
DTYPE_str = 'h'

globals().update(build_api(uncoded=False))
globals().update(build_helpers(uncoded=False))
