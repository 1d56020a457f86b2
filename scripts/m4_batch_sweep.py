#!/usr/bin/env python
"""Config 5 (convlstm_seq2seq heatmap form): heatmaps/s of the train step over batch sizes and arithmetic modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data

m = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
for B in (32, 37, 64, 74, 128):
    x, y = data.make_m4_batch(B, seed=7)
    xs, ys = m._to_dev(x), m._to_dev(y)
    for mode in ("bf16x2", "bf16"):
        m.set_compute(mode)
        for _ in range(2):
            m.train_step_device(xs, ys)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m.train_step_device(xs, ys)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("B=%d %s: %.1f ms/step, %.0f heatmaps/s" % (B, mode, ms, B * 10 / ms * 1e3), flush=True)
    del xs, ys
