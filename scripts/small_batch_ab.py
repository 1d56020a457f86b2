#!/usr/bin/env python
"""Config-2 train step at small per-GPU batches (the reference trains at 32 / 64): A/B of the small-batch launch
shapes (fov_debug_seq_spread, fov_debug_lstm_small_tiles, fov_debug_wgrad_rows_full_grid).  usage: scripts/small_batch_ab.py [B ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data

lib = fov._lib.load()
for fn in ("fov_debug_seq_spread", "fov_debug_lstm_small_tiles", "fov_debug_wgrad_rows_full_grid"):
    getattr(lib, fn).argtypes = [ctypes.c_int]
    getattr(lib, fn).restype = None
Bs = [int(a) for a in sys.argv[1:]] or [32, 64, 256, 1110]


def step_ms(m, xs, ys, reps=30):
    for _ in range(5):
        m.train_step_device(xs, ys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.train_step_device(xs, ys)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


from longterm360fov_b200 import ops

for B in Bs:
    x, y = data.make_m3_batch(B, 34, seed=0)
    row = []
    for spread, side, wave in ((0, False, False), (1, False, False), (1, True, False), (1, True, True)):
        lib.fov_debug_seq_spread(spread)
        lib.fov_debug_lstm_small_tiles(spread)
        lib.fov_debug_wgrad_rows_full_grid(1 - spread)
        ops.set_layer_wavefront(wave)
        for graphs in (False, True):
            m = fov.others_lstm_span_whole(num_user=34, seed=1).compile("Adam", ["mean_squared_error"] * 3, [1, 1, 1])
            m.wgrad_side_stream = side
            m.layer_wavefront = wave
            if graphs:
                m.enable_cuda_graphs()
            xs, ys = m._to_dev(x), m._to_dev(y)
            row.append(step_ms(m, xs, ys))
    print("B=%5d  eager / graph ms:  large-batch shapes %.3f / %.3f | small-batch shapes %.3f / %.3f | + weight gradients "
          "on a side stream %.3f / %.3f | + layer wavefront %.3f / %.3f   (best: %.0f -> %.0f -> %.0f -> %.0f seq/s)"
          % ((B,) + tuple(row) + tuple(B / min(row[2 * i], row[2 * i + 1]) * 1e3 for i in range(4))))
