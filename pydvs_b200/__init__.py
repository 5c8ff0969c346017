"""pydvs_b200 -- B200-native (sm_100a) implementation of pyDVS's spike-generation hot path.

    pydvs_b200.generate_spikes           drop-in for pydvs.generate_spikes          (Cython, coded)
    pydvs_b200.generate_spikes_uncoded   drop-in for pydvs.generate_spikes_uncoded  (Cython, uncoded)
    pydvs_b200.DVSStreams                batched fused pipeline, state resident in HBM
    pydvs_b200.dvs_op                    float32 DVSOperator mode (pydvs/cpp/dvs_op.cpp)

Importing the package loads ``libpydvs_b200.so``; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .spike_lists import SpikeLists  # noqa: F401
from .streams import (DVSConfig, DVSStreams, StepResult, OUTPUT_RATE, OUTPUT_TIME, OUTPUT_TIME_BIN,  # noqa: F401
                      OUTPUT_TIME_BIN_THR, UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY)
from .float_mode import DVSOperatorF32, dvs_op  # noqa: F401
from . import sharding, synth  # noqa: F401

__version__ = "0.1.0"
