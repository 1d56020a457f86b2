"""A/B of the fc-LSTM BPTT kernels (FOV_LSTM_BPTT_TC=0/1 in the environment): train step of the mu/var model."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import longterm360fov_b200 as fov  # noqa: E402

dev = torch.device("cuda")
for tf in (True, False):
    for B in (8880, 37888):
        m = fov.fov_seq2seq_mu_var(seed=3, device=dev, teacher_forcing=tf).compile("Adam", "mean_squared_error")
        enc = torch.randn(B, 10, 6, device=dev) * 0.3
        dec = torch.randn(B, 10 if tf else 1, 6, device=dev) * 0.3
        tgt = torch.randn(B, 10, 6, device=dev) * 0.3
        for _ in range(3):
            m.train_step_device([enc, dec], [tgt])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m.train_step_device([enc, dec], [tgt])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("BPTT_TC=%s tf=%d B=%d: %.3f ms/step, %.2f M seq/s" % (os.environ.get("FOV_LSTM_BPTT_TC", "default"), tf, B, ms, B / ms / 1e3))
