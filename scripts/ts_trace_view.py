#!/usr/bin/env python
"""Render a cycle trace written by scripts/ts_trace.py as a three-column timeline (MMA warp | epilogue part 0 | part 1)."""
import sys

import numpy as np

rec = np.loadtxt(sys.argv[1], dtype=np.int64)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 300)
# the buffer holds the LAST launch that wrote each slot; clocks are 32-bit
t0 = rec[rec[:, 0] == 1][:, 3].min()
names = {1: 'acc_wait', 2: 'acc_ready', 4: 'silu_begin', 5: 'silu_done', 6: 'published', 7: 'tile_bar', 8: 'tile_bar_done',
         9: 'split_done', 10: 'st_issued', 11: 'st_done', 12: 'token_wait', 13: 'token_got', 14: 'next_acc_ready', 15: 'next_ld_issued', 16: 'seg_begin', 17: 'seg_done', 40: 'acc_half0_loaded', 41: 'acc_half1_loaded', 42: 'free_wait', 43: 'free_got', 25: 'cp_async_waited', 26: 'next_requested', 20: 'first_begin', 24: 'first_pq_here', 21: 'first_silu_done', 22: 'first_split_done', 23: 'dot_done',
         100: 'A_wait0', 102: 'A_wait2', 110: 'A_got0', 112: 'A_got2', 120: 'h0_committed', 121: 'h1_committed'}
evs = sorted((int((c - t0) & 0xffffffff), int(w), int(e), int(a)) for w, e, a, c in rec)
for t, w, e, a in evs[lo:hi]:
    col = {1: 0, 4: 34, 8: 68}.get(w, 100)
    print(f"{t:8d} " + " " * col + f"{names.get(e, e)} {a}")
