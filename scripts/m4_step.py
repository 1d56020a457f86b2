#!/usr/bin/env python
"""A few training steps of the config-5 model (convlstm_seq2seq heatmap form, B=32) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
x, y = data.make_m4_batch(B, seed=7)
xs, ys = m._to_dev(x), m._to_dev(y)
for _ in range(steps):
    m.train_step_device(xs, ys)
torch.cuda.synchronize()
print("done")
