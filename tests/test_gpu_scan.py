"""GPU: K2 (dvs_scan_tiles / dvs_scan_tiles_ticket) against NumPy prefix sums for every kernel shape it dispatches
to: one warp per row, one CTA per row (more than 256 tiles per stream, several 2048-tile rounds), row pointers by the
last CTA (ticket), as a launch of their own, and the three-launch form for more than 8192 segments."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _expected(counts, S, Cn, tps, n_slots):
    c = counts.reshape(S, Cn, tps).astype(np.int64)
    excl = np.cumsum(c, axis=2) - c
    totals = c.sum(axis=2)
    seg = np.concatenate([[0], np.cumsum(totals[:, 2:].reshape(-1))])
    return excl.reshape(-1), totals.reshape(-1), seg


@pytest.mark.parametrize("S,tps,n_slots", [(1, 1, 1), (3, 31, 4), (5, 255, 8), (4, 256, 8), (7, 257, 8), (32, 1080, 8),
                                            (2, 2048, 3), (3, 2049, 2), (2, 5000, 1), (300, 270, 8), (700, 3, 8),
                                            (4096, 1, 16), (513, 300, 8), (4096, 1, 6), (5, 8, 2), (9000, 2, 1), (33, 7, 40)])
@pytest.mark.parametrize("with_ticket", [False, True])
def test_scan_matches_numpy(S, tps, n_slots, with_ticket):
    from pydvs_b200 import _lib
    lib = _lib.lib
    Cn = 2 + 2 * n_slots
    rng = np.random.RandomState(S * 1000 + tps + n_slots)
    counts = rng.randint(0, 2000, size=S * Cn * tps).astype(np.int32)
    counts[rng.rand(counts.size) < 0.5] = 0
    dev = "cuda"
    d_counts = torch.from_numpy(counts).to(dev)
    d_excl = torch.full_like(d_counts, -1)
    d_tot = torch.full((S * Cn,), -1, dtype=torch.int32, device=dev)
    d_seg = torch.full((S * 2 * n_slots + 1,), -1, dtype=torch.int64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):   # twice: the ticket must come back to zero
        if with_ticket:
            rc = lib.dvs_scan_tiles_ticket(S, tps, n_slots, _ptr(d_counts), _ptr(d_excl), _ptr(d_tot), _ptr(d_seg),
                                           _ptr(ticket), st)
        else:
            rc = lib.dvs_scan_tiles(S, tps, n_slots, _ptr(d_counts), _ptr(d_excl), _ptr(d_tot), _ptr(d_seg), st)
        assert rc == 0, _lib.last_error()
        excl, totals, seg = _expected(counts, S, Cn, tps, n_slots)
        assert np.array_equal(d_excl.cpu().numpy().astype(np.int64), excl)
        assert np.array_equal(d_tot.cpu().numpy().astype(np.int64), totals)
        assert np.array_equal(d_seg.cpu().numpy(), seg)
        assert int(ticket.item()) == 0
        d_seg.fill_(-1)
