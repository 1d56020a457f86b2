"""Config 3 (mu/var fc-LSTM, autoregressive decode, B = 75 776) a few times, for ncu:
   ncu --set full --clock-control none --import-source on -k regex:lstm_tc --launch-skip 3 -c 1 python scripts/lstm_tc_profile.py"""
import sys

import torch

sys.path.insert(0, ".")
import longterm360fov_b200 as fov                      # noqa: E402

dev = torch.device("cuda")
Bi = 148 * 64 * 8
m2 = fov.fov_seq2seq_mu_var(seed=3, device=dev)
enc = torch.randn(Bi, 10, 6, device=dev) * 0.3
last = enc[:, -1:, :].contiguous()
with torch.no_grad():
    for _ in range(5):
        y = m2._forward([enc, last], False, teacher_forcing=False, steps=10)
torch.cuda.synchronize()
print("ok", float(y[0].abs().mean()))
