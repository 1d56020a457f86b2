#!/usr/bin/env python
"""A few steps of one model for ncu launch lists / full captures.
usage: scripts/profile_step.py m1|m3|m4 BATCH [compute] [train|infer] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data

which = sys.argv[1]
B = int(sys.argv[2])
compute = sys.argv[3] if len(sys.argv) > 3 else "bf16x2"
what = sys.argv[4] if len(sys.argv) > 4 else "train"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
pool = min(B, 256)
tile = lambda a: np.tile(a, (B // pool,) + (1,) * (a.ndim - 1))
if which == "m1":
    m = fov.fov_seq2seq(seed=4).compile("Adam", "mean_squared_error")
    e, d, t, _ = data.make_m1_batch(pool, seed=9)
    x, y = [tile(e), tile(d)], [tile(t)]
elif which == "m3":
    m = fov.others_lstm_span_whole(num_user=34, seed=1).compile("Adam", ["mean_squared_error"] * 3, [1, 1, 1])
    x, y = data.make_m3_batch(pool, 34, seed=0)
    x, y = [tile(a) for a in x], [tile(a) for a in y]
else:
    m = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
    x, y = data.make_m4_batch(B, seed=7)
m.set_compute(compute)
xs, ys = m._to_dev(x), m._to_dev(y)
for _ in range(steps):
    if what == "train":
        m.train_step_device(xs, ys)
    else:
        with torch.no_grad():
            fov.ops.set_math(compute) if hasattr(fov, "ops") else None
            from longterm360fov_b200 import ops
            ops.set_math(compute)
            m._forward(xs, False)
torch.cuda.synchronize()
print("done")
