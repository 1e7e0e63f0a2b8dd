"""Host restatement of the TMA decode-attention kernel's ring / table index arithmetic (commu/engine/decode.py):
every visible slot is visited exactly once, masked slots are exactly the invisible ones, and a slot tile pairs with
64 consecutive rows of the reversed, doubled relative-position table - also across the age wrap."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))

from commu.engine.decode import rtab2_row, slot_age, visible_slot_tiles  # noqa: E402


def _cases():
    rng = random.Random(7)
    yield 2112, 2047, 2048          # the bench shape (mem_len 2048 -> ring of 2112 slots), ring just filled
    yield 2112, 5, 2048             # wrapped
    yield 64, 0, 1
    yield 64, 63, 64
    yield 128, 64, 65
    for _ in range(300):
        C = 64 * rng.randint(1, 40)
        yield C, rng.randrange(C), rng.randint(1, C)


def test_tiles_cover_visible_slots_exactly_once():
    for C, cur, n_vis in _cases():
        tiles = visible_slot_tiles(C, cur, n_vis)
        assert len(set(tiles)) == len(tiles), (C, cur, n_vis)
        assert all(0 <= t < C // 64 for t in tiles)
        visible = {(cur - a) % C for a in range(n_vis)}
        covered = {t * 64 + i for t in tiles for i in range(64)}
        assert visible <= covered, (C, cur, n_vis)
        # the kernel keeps a slot of a visited tile iff its age is below n_vis: exactly the visible set
        kept = {s for s in covered if slot_age(C, cur, s) < n_vis}
        assert kept == visible, (C, cur, n_vis)
        # no tile is visited without a visible slot in it (no wasted HBM traffic beyond tile granularity)
        assert all(any(slot_age(C, cur, t * 64 + i) < n_vis for i in range(64)) for t in tiles), (C, cur, n_vis)


def test_doubled_reversed_table_rows_are_consecutive_per_tile():
    for C, cur, n_vis in _cases():
        for t in visible_slot_tiles(C, cur, n_vis):
            j0 = rtab2_row(C, cur, t * 64)
            assert 0 <= j0 < C and j0 + 63 < 2 * C
            for i in range(64):
                slot = t * 64 + i
                j = j0 + i                                # what one TMA box of 64 rows delivers for this slot
                age_from_table = C - 1 - (j % C)          # rt2[j] = R[C-1 - (j mod C)]
                assert age_from_table == slot_age(C, cur, slot), (C, cur, t, i)
