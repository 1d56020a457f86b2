"""Bring-up check of the tensor-core fc-LSTM forward (csrc/lstm_seq2seq_tc.cu) on a B200:
parity against the CPU oracle for small and multi-CTA batches, then A/B timing against the fp32 kernel.
Usage: python scripts/lstm_tc_check.py [--quick]"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import longterm360fov_b200 as fov                      # noqa: E402
from longterm360fov_b200 import _lib, ops              # noqa: E402
from oracle import keras_numpy as kn                   # noqa: E402  (checker only)

lib = _lib.load()
lib.fov_debug_lstm_tc.argtypes = [C.c_int]


def perturb(w, seed, scale=0.05):
    rng = np.random.default_rng(seed)
    return {k: (v + rng.normal(size=v.shape) * scale).astype(np.float32) for k, v in w.items()}


def parity():
    worst = 0.0
    for math in ("bf16x2", "bf16x3", "bf16"):
        ops.set_math(math)
        for B in (1, 130, 300, 19100):
            for tf in (True, False):
                for kw in ({}, {"recurrent_activation": "sigmoid"}, {"decoder_no_init_state": True}):
                    if B > 300 and (kw or math != "bf16x2"):
                        continue
                    rng = np.random.default_rng(B + 7)
                    w = perturb(kn.init_fov_seq2seq(seed=2, num_encoder_tokens=6), 3)
                    enc = rng.uniform(-1, 1, (B, 10, 6)).astype(np.float32)
                    dec = rng.uniform(-1, 1, (B, 10 if tf else 1, 6)).astype(np.float32)
                    w64 = {k: v.astype(np.float64) for k, v in w.items()}
                    ref = kn.fov_seq2seq_forward(w64, enc.astype(np.float64), dec.astype(np.float64),
                                                 teacher_forcing=tf, **kw)
                    m = fov.fov_seq2seq_mu_var(teacher_forcing=tf, weights=w, **kw)
                    m.set_compute(math)
                    out = {}
                    for mode, name in ((1, "tc"), (-1, "fp32")):
                        lib.fov_debug_lstm_tc(mode)
                        out[name] = m.predict([enc, dec], batch_size=B)
                    e_tc = np.abs(out["tc"] - ref).max()
                    e_32 = np.abs(out["fp32"] - ref).max()
                    print("math=%s B=%d tf=%d %s: tc err %.2e  fp32 err %.2e" % (math, B, tf, kw, e_tc, e_32), flush=True)
                    if math != "bf16":
                        worst = max(worst, e_tc)
    lib.fov_debug_lstm_tc(0)
    ops.set_math("bf16x2")
    print("worst tc error (bf16x2/x3): %.3e" % worst)
    return worst


def timing():
    dev = torch.device("cuda")
    res = {}
    for Bi in (148 * 64 * 8, 148 * 256, 4096):
        m2 = fov.fov_seq2seq_mu_var(seed=3, device=dev)
        enc = torch.randn(Bi, 10, 6, device=dev) * 0.3
        last = enc[:, -1:, :].contiguous()
        for mode, name in ((1, "tc"), (2, "tc_wpg4"), (-1, "fp32")):
            lib.fov_debug_lstm_tc(1 if mode == 2 else mode)
            lib.fov_debug_lstm_tc_wpg(4 if mode == 2 else 8)
            with torch.no_grad():
                for _ in range(3):
                    m2._forward([enc, last], False, teacher_forcing=False, steps=10)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    m2._forward([enc, last], False, teacher_forcing=False, steps=10)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["ar_B%d_%s" % (Bi, name)] = {"ms": ms, "seq_per_s": Bi / ms * 1e3}
            print("AR decode B=%d %s: %.3f ms  %.1f M seq/s" % (Bi, name, ms, Bi / ms / 1e3), flush=True)
    # in-kernel timeline of CTA 0 at the large batch
    lib.fov_debug_lstm_tc(1)
    lib.fov_debug_lstm_tc_enable(1)
    Bi = 148 * 64 * 8
    m2 = fov.fov_seq2seq_mu_var(seed=3, device=dev)
    enc = torch.randn(Bi, 10, 6, device=dev) * 0.3
    last = enc[:, -1:, :].contiguous()
    with torch.no_grad():
        m2._forward([enc, last], False, teacher_forcing=False, steps=10)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)()
    lib.fov_debug_lstm_tc_read(buf)
    res["timeline_cycles"] = dict(zip(["wait_acc", "gate_algebra", "head_stores", "worker_total", "mma_wait", "mma_issue"],
                                      [int(v) for v in buf[:6]]))
    print("timeline", res["timeline_cycles"])
    lib.fov_debug_lstm_tc_enable(0)
    lib.fov_debug_lstm_tc(0)
    return res


if __name__ == "__main__":
    t0 = time.time()
    worst = parity()
    res = timing() if "--quick" not in sys.argv else {}
    res["worst_err"] = worst
    print(json.dumps(res))
    print("elapsed %.1f s" % (time.time() - t0))
    sys.exit(0 if worst < 1e-4 else 1)
