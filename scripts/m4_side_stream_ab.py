import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data, ops
B = 32
x, y = data.make_m4_batch(B, seed=7)
def t(fn, reps, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for side in (False, True, None):
    m = fov.convlstm_seq2seq(seed=2).compile("RMSprop", "mean_squared_error")
    m.set_compute("bf16")
    m.wgrad_side_stream = side
    xs, ys = m._to_dev(x), m._to_dev(y)
    ms = t(lambda: m.train_step_device(xs, ys), 4)
    print("M4 B=32 bf16 train, wgrad side stream %s: %.3f ms (%.0f heatmaps/s)" % (side, ms, B * 10 / ms * 1e3))
