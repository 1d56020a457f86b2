#!/usr/bin/env python
"""Training step of the mu/var model (config 3's model) and of FoV_seq2seq (config 1) per batch size:
tensor-core recurrence forced on (forward + BPTT) vs the fp32 kernels - where does the crossover sit?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import longterm360fov_b200 as fov
from longterm360fov_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
res = {}
for name, in_enc in (("mu_var(in=6)", 6), ("fov_seq2seq(in=90)", 90)):
    for B in (2048, 4096, 8192, 8880, 16384, 37888):
        m = fov.fov_seq2seq(num_encoder_tokens=in_enc, seed=3, device=dev).compile("Adam", "mean_squared_error")
        e = torch.randn(B, 10, in_enc, device=dev) * 0.3
        d = torch.randn(B, 10, 6, device=dev) * 0.3
        t = torch.randn(B, 10, 6, device=dev) * 0.3
        out = {}
        for mode, tag in ((1, "tc"), (-1, "fp32"), (0, "auto")):
            lib.fov_debug_lstm_tc(mode)
            ms = bench._time_cuda(lambda: m.train_step_device([e, d], [t]), reps=10, warm=3)
            out[tag] = round(ms, 4)
        lib.fov_debug_lstm_tc(0)
        res["%s B=%d" % (name, B)] = out
        print(name, B, out, flush=True)
        del m
