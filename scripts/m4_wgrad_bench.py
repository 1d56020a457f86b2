#!/usr/bin/env python
"""Weight gradient of the three head convolutions of config 5 (B=32): TMA-fed plane kernel vs the general gather kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from longterm360fov_b200 import _lib, ops
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
res = {}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for mode in ("bf16", "bf16x2"):
    math = _lib.MATH[mode]
    for name, Cin, Cout in (("conv0_56_512", 56, 512), ("conv1_512_1024", 512, 1024), ("conv2_1024_30", 1024, 30)):
        x = torch.randn(B, 36, 18, Cin, device="cuda")
        dy = torch.randn(B, 36, 18, Cout, device="cuda")
        gw = torch.zeros(5, 5, Cin, Cout, device="cuda")
        gb = torch.zeros(Cout, device="cuda")
        cfg = ops._conv_cfg(B, 36, 18, Cin, Cout, 5, 5, (1, 1), None, 0.0, 648 * Cin, Cin, 648 * Cout, Cout)
        out = {}
        for planes in (1, 0):
            lib.fov_debug_wgrad_planes(planes)
            nws = lib.fov_conv_wgrad_ws_bytes(C.byref(cfg), math)
            ws = torch.empty(int(nws) + 256, dtype=torch.uint8, device="cuda") if nws else None
            fn = lambda: _lib.check(lib.fov_conv2d_bwd_weight_tc_ws(C.byref(cfg), x.data_ptr(), dy.data_ptr(), gw.data_ptr(),
                                                                    gb.data_ptr(), ws.data_ptr() if ws is not None else None,
                                                                    math, st))
            ms = bench._time_cuda(fn, reps=5, warm=2)
            out["planes" if planes else "gather"] = {"ms": ms, "tflops": 2.0 * B * 648 * 25 * Cin * Cout / (ms * 1e-3) / 1e12}
        lib.fov_debug_wgrad_planes(1)
        res["%s_%s" % (mode, name)] = out
print(json.dumps(res, indent=1))
