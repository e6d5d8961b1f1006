"""Invariants of the fused-MLP tile list (csrc/gemm_tcgen05.cu build_mlp_tiles; no GPU needed): every tile of both
GEMMs exactly once, every wait points to an earlier round (deadlock freedom of the persistent kernel), c_proj tiles
spread evenly over the CTA pairs, and a replay with per-tile durations finishes close to the work-conserving bound."""
import ctypes
import os
import sys

import numpy as np
import pytest

from aihab_clip_b200 import _lib

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools", "probes"))
import mlp_tiles_sim  # noqa: E402


def tile_list(P, nfc, nproj, U, ring):
    lib = _lib.load()
    fn = lib.aihab_debug_mlp_tiles
    fn.restype = ctypes.c_long
    fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_long]
    n = fn(P, nfc, nproj, U, ring, None, 0)
    assert n > 0 and n % U == 0
    buf = np.empty(n, dtype=np.uint32)
    assert fn(P, nfc, nproj, U, ring, buf.ctypes.data, n) == n
    return buf.reshape(-1, U)


CASES = [(197, 12, 3, 74, 40), (197, 12, 3, 74, 32), (34, 4, 1, 74, 32), (34, 4, 1, 74, 6), (50, 16, 4, 74, 12),
         (5, 12, 3, 74, 4), (197, 12, 3, 10, 40), (64, 12, 3, 66, 9)]


@pytest.mark.parametrize("P,nfc,nproj,U,ring", CASES)
def test_tile_list_invariants(P, nfc, nproj, U, ring):
    t = tile_list(P, nfc, nproj, U, ring)
    R = t.shape[0]
    fc_round = -np.ones((P, nfc), dtype=np.int64)
    pj_round = -np.ones((P, nproj), dtype=np.int64)
    for r in range(R):
        for u in range(U):
            d = int(t[r, u])
            if d == 0xFFFFFFFF:
                continue
            pr, n = (d >> 8) & 0x7FFFFF, d & 0xFF
            tab = pj_round if d >> 31 else fc_round
            assert pr < P and n < tab.shape[1] and tab[pr, n] < 0      # in range, dealt once
            tab[pr, n] = r
    assert (fc_round >= 0).all() and (pj_round >= 0).all()               # every tile is there
    # a c_proj tile reads what the pair-row's c_fc tiles stored: all of them in EARLIER rounds
    assert (pj_round.min(axis=1) > fc_round.max(axis=1)).all()
    # a c_fc tile overwrites the ring slot of pair-row pr - ring: its c_proj tiles were dealt in earlier rounds
    for pr in range(ring, P):
        assert fc_round[pr].min() > pj_round[pr - ring].max()
    per_unit = [(t[:, u][t[:, u] != 0xFFFFFFFF] >> 31).sum() for u in range(U)]
    if P * nproj >= 2 * U and ring >= 32:                                # (a small ring forces c_proj tiles in bursts)
        assert max(per_unit) - min(per_unit) <= 2                        # c_proj tiles spread evenly


@pytest.mark.parametrize("P,nfc,nproj,U,ring", CASES[:5])
def test_tile_list_replay_has_no_deadlock_and_is_balanced(P, nfc, nproj, U, ring):
    t = tile_list(P, nfc, nproj, U, ring)
    tiles = [[None if int(d) == 0xFFFFFFFF else (int(d) >> 31, (int(d) >> 8) & 0x7FFFFF, int(d) & 0xFF) for d in row]
             for row in t]
    makespan, ideal, _ = mlp_tiles_sim.simulate(tiles, P, nfc, nproj, U, ring)  # raises on deadlock
    if (P, ring) == (197, 40):
        assert ideal / makespan > 0.97                                   # the headline shape: balanced tail
