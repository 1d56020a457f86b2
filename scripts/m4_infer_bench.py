#!/usr/bin/env python
"""config 5 (convlstm_seq2seq heatmaps) forward / train step times per arithmetic mode; A/B of the 256-row conv tiles."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import longterm360fov_b200 as fov
from longterm360fov_b200 import data, _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m4 = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
x, y = data.make_m4_batch(B, seed=7)
xs, ys = m4._to_dev(x), m4._to_dev(y)
res = {}
for mode in ("bf16", "bf16x2"):
    m4.set_compute(mode)
    for mt2 in (1, 0):
        lib.fov_debug_conv_mt2(mt2)
        from longterm360fov_b200 import ops
        ops.set_math(mode)
        with torch.no_grad():
            ms_i = bench._time_cuda(lambda: m4._forward(xs, False), reps=5, warm=2)
        ms_t = bench._time_cuda(lambda: m4.train_step_device(xs, ys), reps=3, warm=2)
        res["%s_mt2=%d" % (mode, mt2)] = {"infer_ms": ms_i, "infer_heatmaps_s": B * 10 / ms_i * 1e3, "train_ms": ms_t,
                                          "train_heatmaps_s": B * 10 / ms_t * 1e3}
lib.fov_debug_conv_mt2(1)
print(json.dumps(res, indent=1))
