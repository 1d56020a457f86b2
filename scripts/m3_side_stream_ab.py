#!/usr/bin/env python
"""Config 2 at large batches: weight gradients on a side stream, on / off, repeated (run-to-run spread)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data

def step_ms(m, xs, ys, reps=20, warm=5):
    for _ in range(warm):
        m.train_step_device(xs, ys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.train_step_device(xs, ys)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for B in [int(a) for a in sys.argv[1:]] or [4440, 8880]:
    pool = 296
    x, y = data.make_m3_batch(pool, 34, seed=0)
    tile = lambda a: np.tile(a, (B // pool,) + (1,) * (a.ndim - 1))
    x, y = [tile(a) for a in x], [tile(a) for a in y]
    out = []
    for rep in range(3):
        for side in (False, True):
            m = fov.others_lstm_span_whole(num_user=34, seed=1).compile("Adam", ["mean_squared_error"] * 3, [1, 1, 1])
            m.wgrad_side_stream = side
            xs, ys = m._to_dev(x), m._to_dev(y)
            out.append((side, step_ms(m, xs, ys)))
            del m, xs, ys
            torch.cuda.empty_cache()
    print("B=%d  off: %s ms | on: %s ms   (reserved %.1f GB)" % (
        B, ["%.2f" % t for s, t in out if not s], ["%.2f" % t for s, t in out if s], torch.cuda.memory_reserved() / 1e9))
