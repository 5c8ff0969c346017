"""pydvs_b200 -- B200-native (sm_100a) implementation of pyDVS's spike-generation hot path.

    pydvs_b200.generate_spikes           drop-in for pydvs.generate_spikes          (Cython, coded)
    pydvs_b200.generate_spikes_uncoded   drop-in for pydvs.generate_spikes_uncoded  (Cython, uncoded)
    pydvs_b200.DVSStreams                batched fused pipeline, state resident in HBM
    pydvs_b200.dvs_op                    float32 DVSOperator mode (pydvs/cpp/dvs_op.cpp)
    pydvs_b200.HostFeed                  pinned host frames -> device, one set ahead, optional exact narrowing to bytes

Importing the package loads ``libpydvs_b200.so``; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .spike_lists import SpikeLists  # noqa: F401
from .streams import (DVSConfig, DVSStreams, StepResult, OUTPUT_RATE, OUTPUT_TIME, OUTPUT_TIME_BIN,  # noqa: F401
                      OUTPUT_TIME_BIN_THR, UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY)
from .float_mode import DVSOperatorF32, dvs_op  # noqa: F401
from .host_feed import HostFeed  # noqa: F401
from . import sharding, synth  # noqa: F401

__version__ = "0.2.0"


def install_as_pydvs(force=False):
    """Make ``import pydvs.generate_spikes as gs`` (``stand_alone.py:9``, ``mnist_to_spikes.py:10``,
    ``pydvs/virtual_cam.py:14``, ``external_dvs_emulator_device.py:41``) resolve to this package, so that the
    reference's callers run unchanged on the CUDA path: registers ``pydvs`` and its submodules ``generate_spikes``,
    ``generate_spikes_uncoded``, ``decode_spikes``, ``numba_nvs`` and ``virtual_cam`` in ``sys.modules``.
    Refuses to shadow an already imported ``pydvs`` unless ``force`` is set.  Returns the alias package."""
    import sys
    import types
    from . import generate_spikes, generate_spikes_uncoded, decode_spikes, numba_nvs, virtual_cam
    have = sys.modules.get("pydvs")
    if have is not None and not getattr(have, "__pydvs_b200_alias__", False) and not force:
        raise ImportError("a different `pydvs` package is already imported; pass force=True to replace it")
    pkg = types.ModuleType("pydvs")
    pkg.__doc__ = "alias of pydvs_b200 (pydvs_b200.install_as_pydvs)"
    pkg.__path__ = []            # a package: `import pydvs.generate_spikes` consults sys.modules first
    pkg.__pydvs_b200_alias__ = True
    subs = {"generate_spikes": generate_spikes, "generate_spikes_uncoded": generate_spikes_uncoded,
            "decode_spikes": decode_spikes, "numba_nvs": numba_nvs, "virtual_cam": virtual_cam}
    for name, mod in subs.items():
        setattr(pkg, name, mod)
        sys.modules["pydvs." + name] = mod
    sys.modules["pydvs"] = pkg
    return pkg


def uninstall_pydvs_alias():
    """Undo :func:`install_as_pydvs`."""
    import sys
    have = sys.modules.get("pydvs")
    if have is not None and getattr(have, "__pydvs_b200_alias__", False):
        for k in [k for k in sys.modules if k == "pydvs" or k.startswith("pydvs.")]:
            del sys.modules[k]
