import sys, json, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import longterm360fov_b200 as fov
from longterm360fov_b200 import data, _lib
sys.path.insert(0, '/root/repo')
import bench
lib = _lib.load()
dev = torch.device('cuda', 0)
res = {}
for Bt in (8192, 32768):
    m1 = fov.fov_seq2seq(seed=4, device=dev).compile("Adam", "mean_squared_error")
    e, d, t, _ = data.make_m1_batch(512, seed=9)
    rep = Bt // 512
    xs = m1._to_dev([np.tile(e, (rep, 1, 1)), np.tile(d, (rep, 1, 1))])
    ys = m1._to_dev([np.tile(t, (rep, 1, 1))])
    for mode in (0, -1):
        lib.fov_debug_lstm_tc(mode)
        ms = bench._time_cuda(lambda: m1.train_step_device(xs, ys), reps=10, warm=3)
        with torch.no_grad():
            msf = bench._time_cuda(lambda: m1._forward(xs, False), reps=10, warm=3)
        res["B%d_tc%d" % (Bt, mode)] = {"train_ms": ms, "train_seq_s": Bt / ms * 1e3, "fwd_ms": msf, "infer_seq_s": Bt / msf * 1e3}
lib.fov_debug_lstm_tc(0)
print(json.dumps(res, indent=1))
