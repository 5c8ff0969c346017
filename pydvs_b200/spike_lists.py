"""``SpikeLists``: the list-of-time-slot-lists the reference's ``make_spike_lists_*`` return
(``generate_spikes.pyx:783-1099``), as a read-only ``Sequence`` over the CSR buffers the CUDA path
produces.  It behaves like ``list[list[int]]`` for ``len``, indexing, iteration, ``==`` and pickling
(what ``external_dvs_emulator_device.py:540-542``, ``mnist_to_spikes.py:290-298`` and the
``pickle.dump`` at ``external_dvs_emulator_device.py:414`` need) and only copies to the host when a
slot is first read.
"""
from collections.abc import Sequence

import numpy as np
import torch


class SpikeLists(Sequence):
    """slot_offsets: int64[n_slots + 1] (host or device); payload: per-event values on the device.

    ``payload`` is either an int16 key tensor [N] (coded modules, and the uncoded ``_bin`` encoders)
    or an int32 tensor [N, 3] of (row, col, +1|-1) triples (uncoded rate / time encoders,
    ``generate_spikes_uncoded.pyx:748-757``).
    """

    def __init__(self, slot_offsets, payload):
        self._offsets = slot_offsets
        self._payload = payload
        self._lists = None

    # -- materialisation ----------------------------------------------------------------------
    def _materialise(self):
        if self._lists is None:
            off = self._offsets
            if isinstance(off, torch.Tensor):
                off = off.cpu().numpy()
            off = np.asarray(off, dtype=np.int64)
            flat = self._payload
            if isinstance(flat, torch.Tensor):
                flat = flat.cpu().numpy()
            flat = flat.tolist()  # ints, or [row, col, p] lists
            self._lists = [flat[int(off[j]) - int(off[0]):int(off[j + 1]) - int(off[0])]
                           for j in range(len(off) - 1)]
        return self._lists

    def tolist(self):
        return self._materialise()

    # -- Sequence -----------------------------------------------------------------------------
    def __len__(self):
        return int(self._offsets.shape[0]) - 1

    def __getitem__(self, idx):
        return self._materialise()[idx]

    def __iter__(self):
        return iter(self._materialise())

    def __eq__(self, other):
        if isinstance(other, SpikeLists):
            other = other.tolist()
        return self._materialise() == other

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return "SpikeLists(%r)" % (self._materialise(),)

    def __reduce__(self):
        # pickles as a plain list of lists, exactly what the reference's callers store
        return (list, (self._materialise(),))

    # -- device views -------------------------------------------------------------------------
    @property
    def slot_offsets(self):
        return self._offsets

    @property
    def payload(self):
        return self._payload

    def slot_counts(self):
        off = self._offsets
        if isinstance(off, torch.Tensor):
            off = off.cpu().numpy()
        return np.diff(np.asarray(off, dtype=np.int64))
