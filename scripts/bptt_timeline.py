#!/usr/bin/env python
"""In-kernel clock64 timeline of the persistent ConvLSTM BPTT of M3 layer 0 (CTA 0), fused weight gradient on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from longterm360fov_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8880
T, W_, Cin, F = 20, 33, 6, 32
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
math = _lib.MATH["bf16x2"]
x0 = torch.randn(B, T, 1, W_, Cin, device=dev)
hseq = torch.empty(B, T, 1, W_, 56, device=dev)
dcat = torch.randn(B, T, 1, W_, 56, device=dev) * 0.01
K = torch.randn(1, 5, Cin, 4 * F, device=dev) * 0.1
R = torch.randn(1, 5, F, 4 * F, device=dev) * 0.1
b = torch.zeros(4 * F, device=dev)
gates = torch.empty(B, T, 1, W_, 4 * F, device=dev)
cseq = torch.empty(B, T, 1, W_, F, device=dev)
hT, cT = torch.empty(B, 1, W_, F, device=dev), torch.empty(B, 1, W_, F, device=dev)
lcfg = _lib.ConvLstmCfg(B, T, 1, W_, Cin, F, 1, 5, 1, 1, 0, T * W_ * Cin, W_ * Cin, Cin, T * W_ * 56, W_ * 56, 56, 1, math)
ws = torch.empty(int(lib.fov_convlstm_fwd_ws_bytes(C.byref(lcfg))) + 256, dtype=torch.uint8, device=dev)
io = _lib.ConvLstmIO(x0.data_ptr(), K.data_ptr(), R.data_ptr(), b.data_ptr(), None, None, hseq.data_ptr(), gates.data_ptr(),
                     cseq.data_ptr(), hT.data_ptr(), cT.data_ptr(), ws.data_ptr())
_lib.check(lib.fov_convlstm_fwd(C.byref(lcfg), C.byref(io), st))
gK, gR, gb = torch.zeros_like(K), torch.zeros_like(R), torch.zeros_like(b)
bws = torch.empty(int(lib.fov_convlstm_bwd_ws_floats(C.byref(lcfg))), device=dev)
gr = _lib.ConvLstmGrads(dcat.data_ptr(), None, None, None, None, None, gK.data_ptr(), gR.data_ptr(), gb.data_ptr(), bws.data_ptr(), 0)
names = ["worker wait tmem_full", "worker compute", "worker total", "worker wait wg_done", "mma wait a_full", "mma issue", "epilogue", "-"]
for fused in (1, 0):
    lib.fov_debug_seq_bwd_wgrad(fused)
    fn = lambda: _lib.check(lib.fov_convlstm_bwd(C.byref(lcfg), C.byref(io), C.byref(gr), st))
    ms = bench._time_cuda(fn, reps=5, warm=2)
    lib.fov_debug_seq_bwd_enable(1)
    fn()
    torch.cuda.synchronize()
    lib.fov_debug_seq_bwd_enable(0)
    buf = (C.c_ulonglong * 8)()
    lib.fov_debug_seq_bwd_read(buf)
    print("fused=%d  %.3f ms per call; CTA 0 cycles:" % (fused, ms), {n: int(v) for n, v in zip(names, buf)})
lib.fov_debug_seq_bwd_wgrad(0)
