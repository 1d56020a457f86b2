#!/usr/bin/env python
"""One shape of the TMA-fed weight gradient, a few calls (for ncu): usage wgrad_planes_one.py Cin Cout [mode] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from longterm360fov_b200 import _lib, ops
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
Cin, Cout = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
math = _lib.MATH[mode]
x = torch.randn(B, 36, 18, Cin, device="cuda")
dy = torch.randn(B, 36, 18, Cout, device="cuda")
gw = torch.zeros(5, 5, Cin, Cout, device="cuda")
gb = torch.zeros(Cout, device="cuda")
cfg = ops._conv_cfg(B, 36, 18, Cin, Cout, 5, 5, (1, 1), None, 0.0, 648 * Cin, Cin, 648 * Cout, Cout)
nws = lib.fov_conv_wgrad_ws_bytes(C.byref(cfg), math)
ws = torch.empty(int(nws) + 256, dtype=torch.uint8, device="cuda")
for _ in range(3):
    _lib.check(lib.fov_conv2d_bwd_weight_tc_ws(C.byref(cfg), x.data_ptr(), dy.data_ptr(), gw.data_ptr(), gb.data_ptr(),
                                               ws.data_ptr(), math, st))
torch.cuda.synchronize()
print("done")
