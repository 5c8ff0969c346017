"""Drop-in for ``pydvs.generate_spikes_uncoded`` (``pydvs/generate_spikes_uncoded.pyx``) on B200.

The "uncoded twin": no RECTIFIED polarity, ``split_spikes`` returns uint16, the rate/time encoders emit
``[row, col, +1|-1]`` triples, un-inverted rate slot rule, its own key layout, guarded un-clipped
adaptive threshold and reversed ``time_bin_thr`` slot order.
"""
from ._api import DTYPE, DTYPE_FLOAT, DTYPE_IDX, DTYPE_U8, DTYPE_U16, build_api  # noqa: F401
from ._helpers import build_helpers

UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY = 0, 1, 2
DTYPE_str = 'h'

globals().update(build_api(uncoded=True))
globals().update(build_helpers(uncoded=True))
