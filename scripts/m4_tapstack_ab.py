#!/usr/bin/env python
"""Config 5 (convlstm_seq2seq heatmap form): A/B of the tap-stacked narrow head convolution (ops.set_tapstack).
usage: scripts/m4_tapstack_ab.py [B] [compute]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
compute = sys.argv[2] if len(sys.argv) > 2 else "bf16"


def t(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x, y = data.make_m4_batch(B, seed=7)
for on in (False, True):
    ops.set_tapstack(on)
    m = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
    m.set_compute(compute)
    xs, ys = m._to_dev(x), m._to_dev(y)

    def infer():
        with torch.no_grad():
            ops.set_math(compute)
            m._forward(xs, False)
    ms_i = t(infer, 5)
    ms_t = t(lambda: m.train_step_device(xs, ys), 3)
    print("tapstack=%d B=%d %s: infer %.3f ms (%.0f heatmaps/s), train %.3f ms (%.0f heatmaps/s)"
          % (on, B, compute, ms_i, B * 10 / ms_i * 1e3, ms_t, B * 10 / ms_t * 1e3))
