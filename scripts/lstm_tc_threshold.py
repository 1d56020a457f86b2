import ctypes as C, sys, torch
sys.path.insert(0, ".")
import longterm360fov_b200 as fov
from longterm360fov_b200 import _lib
lib = _lib.load(); lib.fov_debug_lstm_tc.argtypes = [C.c_int]
dev = torch.device("cuda")
for Bi in (512, 1024, 2048, 3072):
    m2 = fov.fov_seq2seq_mu_var(seed=3, device=dev)
    enc = torch.randn(Bi, 10, 6, device=dev) * 0.3
    last = enc[:, -1:, :].contiguous()
    for mode, name in ((1, "tc"), (-1, "fp32")):
        lib.fov_debug_lstm_tc(mode)
        with torch.no_grad():
            for _ in range(5): m2._forward([enc, last], False, teacher_forcing=False, steps=10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): m2._forward([enc, last], False, teacher_forcing=False, steps=10)
            e1.record(); torch.cuda.synchronize()
        print("B=%d %s: %.4f ms" % (Bi, name, e0.elapsed_time(e1) / 20), flush=True)
