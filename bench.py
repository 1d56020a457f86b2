#!/usr/bin/env python
"""bench.py - the driver's measurement contract.

Workload (BASELINE.json configs[1]): the concat-state others_LSTM_span_whole seq2seq
(target past + other viewers' whole-span FoV), ONE training step =
forward + BPTT + 3xMSE + Keras-form Adam over one batch of synthetic windows.  All tensors, gates,
cell state, losses and optimiser state are fp32; the GEMM/conv arithmetic is `--compute` (default
bf16x2: tensor cores, two bf16 terms per fp32 operand ~ 16 mantissa bits, fp32 accumulate; the line
also carries the step time in fp32 and bf16x3 and the forward error measured at the benchmarked shape).
Metric: sequences/s (whole job, all GPUs).  `value` is measured with the inputs
resident in HBM; `e2e` goes through the public API (`model.train_on_batch`) with
pinned HOST buffers, H2D copies and the D2H loss read inside the timed region.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU oracle (Keras-equivalent stand-in) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "seq2seq FoV sequences/sec, training step (concat-state others_LSTM_span_whole)"
UNIT = "sequences/s"
# SURVEY.md 8(d) / BASELINE.md section 3: algorithmic work per sequence of config 2
FWD_FLOP_PER_SEQ = 82_307_200
TRAIN_FLOP_PER_SEQ = 3 * FWD_FLOP_PER_SEQ
NUM_USER = 34
# arithmetic type of the GEMM/conv path per --compute mode (fp32 accumulation everywhere; gates, cell state,
# losses and optimiser in fp32)
DTYPE_OF = {"fp32": "fp32 (CUDA cores)",
            "bf16": "bf16 (tensor cores, 1 term per operand, fp32 accumulate)",
            "bf16x2": "bf16x2 (tensor cores, 2 bf16 terms per fp32 operand ~ 16 mantissa bits, fp32 accumulate)",
            "bf16x3": "bf16x3 (tensor cores, 3 bf16 terms per fp32 operand ~ 24 mantissa bits = fp32-grade, fp32 accumulate)"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "which": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms while running."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.active = False        # samples are kept only while a timed leg is running

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for n, bit in names.items():
                        if r & bit:
                            self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover - NVML missing
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _time_cuda(fn, reps=20, warm=3):
    """Average device time of `fn` (ms) over `reps` back-to-back launches, CUDA events on torch's current
    stream (the stream the C-ABI calls are given)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, written by scripts/ncu_raw_summary.py): {kernel key: {batch: bytes}}."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


M3_LAYERS = [(6, 32, 0), (32, 16, 32), (16, 8, 48)]          # (Cin, F, channel offset in the 56-wide concat buffer)


def kernel_rooflines(lib, dev, B, compute, peaks, seq_per_s_per_gpu, step_ms):
    """`roofline` of the bench line.
    Top level = SURVEY.md 8(d)'s definition for config 2: algorithmic train FLOPs (246.9 MFLOP/sequence) x
    sequences/s against the bf16 tensor peak (sustained figure: the step is a long run).  `dominant_kernel` follows the
    contract's per-kernel definition (algorithmic bytes per launch / its CUDA-event duration) for the C-ABI call with
    the largest share of the step, and `kernels` lists EVERY C-ABI call family of the step, each timed alone at the
    step's own shapes: algorithmic bytes and FLOPs, achieved GB/s and TFLOP/s, the ncu DRAM traffic of the committed
    capture where one exists and its ratio to the algorithmic bytes (> 1 = wasted re-reads or saved-tensor traffic)."""
    import ctypes as C
    import torch
    from longterm360fov_b200 import _lib
    math = _lib.MATH[compute]
    st = torch.cuda.current_stream().cuda_stream
    W_, T = NUM_USER - 1, 20
    npix = B * T * W_
    hbm, tens = peaks["hbm_gbs"], peaks["bf16"]
    traffic = _ncu_traffic()
    issue = {0: 1, 1: 1, 2: 3, 3: 6}[math]
    out = []

    def entry(name, key, ms, byts, flop, note=None):
        gbs, tf = byts / (ms * 1e-3) / 1e9, flop / (ms * 1e-3) / 1e12
        tr = traffic.get(key, {}).get(str(B))
        e = {"call": name, "ms": ms, "share_of_step": ms / step_ms,
             "algorithmic_bytes": byts, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm,
             "algorithmic_flop": flop, "achieved_tflops": tf, "frac_of_bf16_peak": tf / tens,
             "issued_tflops": tf * issue, "ncu_dram_bytes": tr, "traffic_over_algorithmic": (tr / byts) if tr else None}
        if note:
            e["note"] = note
        out.append(e)
        return e

    # ---- the three stacked ConvLSTM layers of the others branch, through the C-ABI calls the model makes ----
    x0 = torch.randn(B, T, 1, W_, 6, device=dev)
    hseq = torch.empty(B, T, 1, W_, 56, device=dev)
    dcat = torch.randn(B, T, 1, W_, 56, device=dev) * 0.01
    for l, (Cin, F, off) in enumerate(M3_LAYERS):
        K = torch.randn(1, 5, Cin, 4 * F, device=dev) * 0.1
        R = torch.randn(1, 5, F, 4 * F, device=dev) * 0.1
        b = torch.zeros(4 * F, device=dev)
        gates = torch.empty(B, T, 1, W_, 4 * F, device=dev)
        cseq = torch.empty(B, T, 1, W_, F, device=dev)
        hT, cT = torch.empty(B, 1, W_, F, device=dev), torch.empty(B, 1, W_, F, device=dev)
        if l == 0:
            xp, xb, xt, xpix = x0.data_ptr(), T * W_ * 6, W_ * 6, 6
        else:
            xp, xb, xt, xpix = hseq.data_ptr() + 4 * M3_LAYERS[l - 1][2], T * W_ * 56, W_ * 56, 56
        lcfg = _lib.ConvLstmCfg(B, T, 1, W_, Cin, F, 1, 5, 1, 1, 0, xb, xt, xpix, T * W_ * 56, W_ * 56, 56, 1, math)
        wsb = lib.fov_convlstm_fwd_ws_bytes(C.byref(lcfg))
        ws = torch.empty(int(wsb) + 256, dtype=torch.uint8, device=dev) if wsb else None
        io = _lib.ConvLstmIO(xp, K.data_ptr(), R.data_ptr(), b.data_ptr(), None, None, hseq.data_ptr() + 4 * off,
                             gates.data_ptr(), cseq.data_ptr(), hT.data_ptr(), cT.data_ptr(),
                             ws.data_ptr() if ws is not None else None)
        n0 = lib.fov_launch_count()
        _lib.check(lib.fov_convlstm_fwd(C.byref(lcfg), C.byref(io), st))
        nl = int(lib.fov_launch_count() - n0)
        ms_f = _time_cuda(lambda: _lib.check(lib.fov_convlstm_fwd(C.byref(lcfg), C.byref(io), st)), reps=10, warm=2)
        flop_f = 2.0 * npix * 5 * (Cin + F) * 4 * F
        # x_t in; h_t, c_t and the 4 activated gates out (h_{t-1}, c_{t-1} never leave the SM on the persistent path)
        entry("fov_convlstm_fwd L%d (Cin=%d,F=%d): %d launch(es), persistent over %d timesteps" % (l, Cin, F, nl, T),
              "convlstm_fwd_L%d" % l, ms_f, npix * (Cin + 6 * F) * 4, flop_f)
        gK, gR, gb = torch.zeros_like(K), torch.zeros_like(R), torch.zeros_like(b)
        bws = torch.empty(int(lib.fov_convlstm_bwd_ws_floats(C.byref(lcfg))), device=dev)
        dxp = None if l == 0 else dcat.data_ptr() + 4 * M3_LAYERS[l - 1][2]
        gr = _lib.ConvLstmGrads(dcat.data_ptr() + 4 * off, None, None, dxp, None, None, gK.data_ptr(), gR.data_ptr(),
                                gb.data_ptr(), bws.data_ptr(), 1)
        n0 = lib.fov_launch_count()
        _lib.check(lib.fov_convlstm_bwd(C.byref(lcfg), C.byref(io), C.byref(gr), st))
        nl = int(lib.fov_launch_count() - n0)
        # (the BPTT overwrites the saved gates with dZ; re-running it on dZ is the same memory traffic)
        ms_b = _time_cuda(lambda: _lib.check(lib.fov_convlstm_bwd(C.byref(lcfg), C.byref(io), C.byref(gr), st)),
                          reps=10, warm=2)
        # gates 4F + c_t F + c_{t-1} F + dh F in, dx Cin read-modify-write (layers above 0); weight gradient needs
        # dZ 4F, h_{t-1} F, x_t Cin
        byts_b = npix * (7 * F + (2 * Cin if l else 0)) * 4
        if math:
            byts_b += npix * 4 * F * 4          # dZ written for the separate weight-gradient launch ...
            byts_b += npix * (5 * F + Cin) * 4  # ... which reads dZ, h and x
        entry("fov_convlstm_bwd L%d: %d launches (persistent BPTT%s + weight gradient)" % (l, nl, " + dx" if l else ""),
              "convlstm_bwd_L%d" % l, ms_b, byts_b, 2 * flop_f)
        del K, R, b, gates, cseq, bws
    del x0, dcat
    # ---- the dense layers on the flattened concat buffer: Dense(198) on 20 slices, Dense(256) on 10 ----
    flat = hseq.view(B, T, 1848)
    dx = torch.empty(B, T, 1848, device=dev)

    def gemm_calls(cfg, xp, dyp, w, bias, yp, dxp, gw, gb):
        """forward / backward-data / weight-gradient closures of one dense layer through the C ABI"""
        if math == 0:
            wsd = torch.empty(w.numel(), device=dev)
            return (lambda: _lib.check(lib.fov_conv2d_fwd(C.byref(cfg), xp, w.data_ptr(), bias.data_ptr(), yp, st)),
                    lambda: _lib.check(lib.fov_conv2d_bwd_data(C.byref(cfg), dyp, w.data_ptr(), dxp, wsd.data_ptr(), st)),
                    lambda: _lib.check(lib.fov_conv2d_bwd_weight(C.byref(cfg), xp, dyp, gw.data_ptr(), gb.data_ptr(), st)))
        ws1 = torch.empty(int(lib.fov_conv_tc_ws_bytes(C.byref(cfg), math, 0)) + 256, dtype=torch.uint8, device=dev)
        ws2 = torch.empty(int(lib.fov_conv_tc_ws_bytes(C.byref(cfg), math, 1)) + 256, dtype=torch.uint8, device=dev)
        return (lambda: _lib.check(lib.fov_conv2d_fwd_tc(C.byref(cfg), xp, w.data_ptr(), bias.data_ptr(), yp, ws1.data_ptr(), math, st)),
                lambda: _lib.check(lib.fov_conv2d_bwd_data_tc(C.byref(cfg), dyp, w.data_ptr(), dxp, ws2.data_ptr(), math, st)),
                lambda: _lib.check(lib.fov_conv2d_bwd_weight_tc(C.byref(cfg), xp, dyp, gw.data_ptr(), gb.data_ptr(), math, st)))

    Tf = 10
    w198 = torch.randn(1, 1, 1848, 198, device=dev) * 0.02
    w256 = torch.randn(1, 1, 1848, 256, device=dev) * 0.02
    wcat = torch.cat([w198, w256], dim=-1)
    b198, b256, bcat = torch.zeros(198, device=dev), torch.zeros(256, device=dev), torch.zeros(454, device=dev)
    y198 = torch.empty(B * T, 198, device=dev)
    y256 = torch.empty(B * Tf, 256, device=dev)
    ycat = torch.empty(B * Tf, 454, device=dev)
    g198, g256 = torch.zeros_like(w198), torch.zeros_like(w256)
    gb198, gb256 = torch.zeros_like(b198), torch.zeros_like(b256)
    fut_off = 4 * (T - Tf) * 1848
    c_full = _lib.ConvCfg(B * T, 1, 1, 1848, 198, 1, 1, 1, 1, 0, 0, 1848, 1848, 198, 198, 0, 0.0)
    c_fut = _lib.ConvCfg(B, 1, Tf, 1848, 256, 1, 1, 1, 1, 0, 0, T * 1848, 1848, Tf * 256, 256, 0, 0.0)
    c_past = _lib.ConvCfg(B, 1, T - Tf, 1848, 198, 1, 1, 1, 1, 0, 0, T * 1848, 1848, T * 198, 198, 0, 0.0)
    c_cat = _lib.ConvCfg(B, 1, Tf, 1848, 454, 1, 1, 1, 1, 0, 0, T * 1848, 1848, Tf * 454, 454, 0, 0.0)
    f198 = gemm_calls(c_full, flat.data_ptr(), y198.data_ptr(), w198, b198, y198.data_ptr(), dx.data_ptr(), g198, gb198)
    f256 = gemm_calls(c_fut, flat.data_ptr() + fut_off, y256.data_ptr(), w256, b256, y256.data_ptr(), dx.data_ptr() + fut_off,
                      g256, gb256)
    fpast = gemm_calls(c_past, flat.data_ptr(), y198.data_ptr(), w198, b198, y198.data_ptr(), dx.data_ptr(), g198, gb198)
    fcat = gemm_calls(c_cat, flat.data_ptr() + fut_off, ycat.data_ptr(), wcat, bcat, ycat.data_ptr(), dx.data_ptr() + fut_off,
                      g198, gb198)
    r20, r10 = B * T, B * Tf
    io = lambda rows, cout: rows * (1848 + cout) * 4
    fl = lambda rows, cout: 2.0 * rows * 1848 * cout
    entry("Dense 1848->198 (reconstruction head, 20 slices): forward", "dense198_fwd", _time_cuda(f198[0], reps=10, warm=2),
          io(r20, 198), fl(r20, 198))
    entry("Dense 1848->256 (concat-state fusion, 10 future slices): forward", "dense256_fwd",
          _time_cuda(f256[0], reps=10, warm=2), io(r10, 256), fl(r10, 256))
    entry("Dense backward-data, 10 past slices: dx = dy198 . W198^T", "dense_bwd_data_past",
          _time_cuda(fpast[1], reps=10, warm=2), io(r10, 198), fl(r10, 198))
    entry("Dense backward-data, 10 future slices, both layers in ONE GEMM: dx = [dy198 | dy256] . [W198 | W256]^T",
          "dense_bwd_data_future_fused", _time_cuda(fcat[1], reps=10, warm=2), io(r10, 454), fl(r10, 454))
    entry("Dense 1848->198: weight gradient", "dense198_wgrad", _time_cuda(f198[2], reps=10, warm=2), io(r20, 198), fl(r20, 198))
    entry("Dense 1848->256: weight gradient", "dense256_wgrad", _time_cuda(f256[2], reps=10, warm=2), io(r10, 256), fl(r10, 256))
    del w198, w256, wcat, y198, y256, ycat, dx
    del hseq
    out.sort(key=lambda e: -e["ms"])
    dom = out[0]
    step_tflops = seq_per_s_per_gpu * TRAIN_FLOP_PER_SEQ / 1e12
    step_traffic = sum(e["ncu_dram_bytes"] for e in out if e["ncu_dram_bytes"]) or None
    return {
        "bound": "tensor", "achieved": step_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": step_tflops / peaks["bf16_sustained"], "traffic": step_traffic,
        "definition": "SURVEY.md 8(d), config 2: seq/s x %d algorithmic train FLOP per sequence (matmul/conv MACs x 2, "
                      "fwd x 3) / bf16 tensor peak; the sustained figure of MEASURED_PEAKS.json because the kernels are "
                      "timed inside a long step (%.1f %% of the burst peak %.1f)" % (
                          TRAIN_FLOP_PER_SEQ, 100 * step_tflops / peaks["bf16"], peaks["bf16"]),
        "peak_is": "%s (MEASURED_PEAKS.json)" % peaks["which"],
        "issued_mma_per_algorithmic_mac": issue,
        "frac_issued": step_tflops * issue / peaks["bf16_sustained"],
        "io_bytes_per_seq": 32424,
        "traffic_note": "sum of the ncu DRAM bytes of the listed calls (profiles/ncu_traffic.json) per step: the design "
                        "saves gates + cell state for BPTT in fp32, which is why traffic is far above the %.2f GB of "
                        "algorithmic I/O per step" % (B * 32424 / 1e9),
        "dominant_kernel": {"call": dom["call"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": hbm,
                            "unit": "GB/s", "frac": dom["frac_of_hbm_peak"], "traffic": dom["ncu_dram_bytes"],
                            "ms_per_call": dom["ms"], "share_of_step": dom["share_of_step"],
                            "algorithmic_bytes_per_call": dom["algorithmic_bytes"]},
        "kernels": out,
    }


def other_workloads(dev, peaks, compute):
    """The other BASELINE.json configs, a few steps each (reported, not the headline): ConvLSTM heatmaps/s
    (config 5), large-batch autoregressive mu/var inference (config 3), teacher-forced fc-LSTM training
    (config 1's model on the GPU).  Inputs resident in HBM; CUDA events."""
    import torch
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import data
    res = {}

    # ---- config 5: convlstm_seq2seq heatmap form, one heatmap = one predicted (36,18,30) second ----
    Bh = 32
    m4 = fov.convlstm_seq2seq(seed=2, device=dev).compile("RMSprop", "mean_squared_error")
    x, y = data.make_m4_batch(Bh, seed=7)
    xs, ys = m4._to_dev(x), m4._to_dev(y)
    flop_fwd = 10 * 381.5e6 + 10 * (381.5e6 + 18.9e9)         # encoder + decoder steps + heads
    # the default arithmetic (two-term bf16 split, ~16 mantissa bits) and the plain bf16 tensor-core mode BASELINE.json's config 5
    # names ("bf16 tensor-core gate convolutions"; forward tolerance 3e-2 max-abs, tests/test_gpu_parity.py)
    for mode in dict.fromkeys([compute, "bf16"]):
        m4.set_compute(mode)
        ms_t = _time_cuda(lambda: m4.train_step_device(xs, ys), reps=3, warm=2)
        with torch.no_grad():
            ms_i = _time_cuda(lambda: m4._forward(xs, False), reps=3, warm=1)
        entry = {
            "batch": Bh, "unit": "heatmaps/s", "train": Bh * 10 / (ms_t * 1e-3), "infer": Bh * 10 / (ms_i * 1e-3),
            "train_ms_per_step": ms_t, "infer_ms_per_step": ms_i,
            "train_algorithmic_tflops": 3 * flop_fwd * Bh / (ms_t * 1e-3) / 1e12,
            "infer_algorithmic_tflops": flop_fwd * Bh / (ms_i * 1e-3) / 1e12,
            "tensor_peak_tflops": peaks["bf16"], "compute": mode}
        res["convlstm_seq2seq_heatmap" if mode == compute else "convlstm_seq2seq_heatmap_" + mode] = entry
    del m4, xs, ys
    torch.cuda.empty_cache()

    # ---- config 3: FoV_seq2seq_mu_var, autoregressive decode without teacher forcing, large batch:
    #      ONE persistent launch = 10 encoder + 10 decoder steps (Dense+tanh inside the recurrence) ----
    Bi = 148 * 64 * 8
    m2 = fov.fov_seq2seq_mu_var(seed=3, device=dev)
    m2.set_compute(compute)
    from longterm360fov_b200 import ops as _ops
    _ops.set_math(compute)
    enc = torch.randn(Bi, 10, 6, device=dev) * 0.3
    last = enc[:, -1:, :].contiguous()
    with torch.no_grad():
        ms = _time_cuda(lambda: m2._forward([enc, last], False, teacher_forcing=False, steps=10), reps=10, warm=3)
    flop = Bi * (10 * 2 * (6 + 64) * 256 + 10 * (2 * (6 + 64) * 256 + 2 * 64 * 6))
    byts = Bi * (10 * 6 + 6 + 10 * 6) * 4                     # inputs + outputs
    res["fov_seq2seq_mu_var_autoregressive_infer"] = {
        "batch": Bi, "unit": "sequences/s", "value": Bi / (ms * 1e-3), "ms_per_launch": ms,
        "us_per_timestep_per_cta_wave": 1e3 * ms / 20 / max(1, -(-Bi // (64 if compute == "fp32" else 256)) / 148.0),
        "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "hbm_gbs": byts / (ms * 1e-3) / 1e9,
        "issued_bf16_tflops": (3 if compute == "bf16x2" else {"bf16": 1, "bf16x3": 6}.get(compute, 0)) *
                              Bi * 20 * 2 * (16 + 64) * 256 / (ms * 1e-3) / 1e12,
        "compute": compute,
        "note": "persistent recurrence, ONE launch: tcgen05 gate GEMM per step (weights + h resident on the SM, "
                "Dense+tanh head and re-feed in the epilogue) unless compute=fp32; bound by the gate algebra "
                "(MUFU/issue), not by HBM or the tensor pipe"}

    # ---- config 3's model, training at a batch that fills the machine with 128-sequence tiles: tcgen05 forward (saved
    #      tensors written) + tcgen05 BPTT + one tensor-core weight-gradient launch per LSTM ----
    Bl = 148 * 256
    for tf in (True, False):
        m2t = fov.fov_seq2seq_mu_var(seed=3, device=dev, teacher_forcing=tf).compile("Adam", "mean_squared_error")
        m2t.set_compute(compute)
        e_ = torch.randn(Bl, 10, 6, device=dev) * 0.3
        d_ = torch.randn(Bl, 10 if tf else 1, 6, device=dev) * 0.3
        t_ = torch.randn(Bl, 10, 6, device=dev) * 0.3
        ms = _time_cuda(lambda: m2t.train_step_device([e_, d_], [t_]), reps=10, warm=3)
        res["fov_seq2seq_mu_var_train_" + ("teacher_forced" if tf else "autoregressive")] = {
            "batch": Bl, "unit": "sequences/s", "value": Bl / (ms * 1e-3), "ms_per_step": ms, "compute": compute}
        del m2t

    # ---- config 2 at the reference's own batch size (32): launch-bound, eager vs CUDA-graph replay ----
    Bs = 32
    px, py = data.make_m3_batch(Bs, NUM_USER, seed=11)
    small = {}
    for graphed in (False, True):
        ms3 = fov.others_lstm_span_whole(num_user=NUM_USER, seed=1, device=dev)
        ms3.compile(optimizer="Adam", loss=["mean_squared_error"] * 3, loss_weights=[1, 1, 1])
        ms3.set_compute(compute)
        ms3.enable_cuda_graphs(graphed)
        xs, ys = ms3._to_dev(px), ms3._to_dev(py)
        ms = _time_cuda(lambda: ms3.train_step_device(xs, ys), reps=30, warm=5)
        small["cuda_graph" if graphed else "eager"] = {"value": Bs / (ms * 1e-3), "ms_per_step": ms}
        del ms3
    res["others_lstm_span_whole_train_batch32"] = dict(small, batch=Bs, unit="sequences/s",
                                                       note="the reference's batch size; latency bound (3 x 20 dependent recurrence steps forward and "
                                                            "backward): small-batch launch shapes (one image per CTA, 4-sequence fc-LSTM tiles); "
                                                            "model.enable_cuda_graphs() replays forward + BPTT from one graph, in which the three ConvLSTM "
                                                            "layers run as a wavefront and the weight gradients on a side stream")

    # ---- sample builders: windows of a full-size video, one-hot heatmaps; HBM-bound ----
    from longterm360fov_b200 import ops as _o
    SEC, NHM = 3000, 16384                                    # outputs of 1.5 GB / 1.3 GB: well beyond the 126 MB L2
    vid = torch.rand(48, SEC, 90, device=dev) * 2 - 1
    ms = _time_cuda(lambda: _o.reshape2second_stacks(vid, collapse_user=False, stride=1), reps=10, warm=3)
    nwin = SEC - 10 + 1 - 10
    wbytes = 3 * nwin * 48 * 10 * 90 * 4                      # three window tensors written; the source stays in L2
    fr = torch.nn.functional.normalize(torch.randn(NHM, 30, 3, device=dev), dim=-1)
    ms_h = _time_cuda(lambda: _o.one_hot_heatmaps(fr), reps=10, warm=3)
    hbytes = NHM * 36 * 18 * 30 * 4
    res["sample_builders"] = {
        "window_stacks": {"shape": "48 viewers x %d s x 90, stride 1" % SEC, "windows_per_s": nwin * 48 / (ms * 1e-3),
                          "ms": ms, "hbm_gbs": wbytes / (ms * 1e-3) / 1e9, "frac_of_copy_peak": wbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "one_hot_heatmaps": {"shape": "%d seconds x 30 frames -> (36,18,30)" % NHM, "heatmaps_per_s": NHM / (ms_h * 1e-3),
                             "ms": ms_h, "hbm_gbs": hbytes / (ms_h * 1e-3) / 1e9,
                             "frac_of_copy_peak": hbytes / (ms_h * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    del vid, fr

    # ---- config 1 model (FoV_seq2seq, teacher forcing) training on the GPU ----
    m1 = fov.fov_seq2seq(seed=4, device=dev).compile("Adam", "mean_squared_error")
    e, d, t, _ = data.make_m1_batch(512, seed=9)
    cfg1 = {"unit": "sequences/s", "io_bytes_per_seq": 4320,
            "note": "SURVEY.md 8(d) config 1 is HBM-classified on 4 080 + 240 B of input / target per sequence; the "
                    "train step also writes and re-reads ~41 KB of saved tensors per sequence"}
    for Bt in (8192, 65536):               # SURVEY.md 8(d) lists B = 32 (reference), 4 096 and 65 536 for this config
        rep = Bt // 512
        xs = m1._to_dev([np.tile(e, (rep, 1, 1)), np.tile(d, (rep, 1, 1))])
        ys = m1._to_dev([np.tile(t, (rep, 1, 1))])
        ms = _time_cuda(lambda: m1.train_step_device(xs, ys), reps=10, warm=3)
        val = Bt / (ms * 1e-3)
        cfg1["batch_%d" % Bt] = {"value": val, "ms_per_step": ms,
                                 "io_frac_of_hbm_peak": val * 4320 / 1e9 / peaks["hbm_gbs"],
                                 "algorithmic_tflops": val * 3.46e6 / 1e12}
        if Bt == 8192:
            cfg1.update(batch=Bt, value=val, ms_per_step=ms)
        del xs, ys
    res["fov_seq2seq_teacher_forced_train"] = cfg1
    return res


def _cpu_oracle_step_fn(batch, dtype_name="float32"):
    """One training step of config 2 on the CPU oracle (torch restatement of Keras semantics)."""
    import torch
    from longterm360fov_b200 import data
    from oracle import keras_numpy as kn
    from oracle import keras_torch as kt
    dt = getattr(torch, dtype_name)
    w = kt.to_torch(kn.init_others_lstm_span_whole(seed=1, num_user=NUM_USER), dtype=dt)
    x, y = data.make_m3_batch(batch, NUM_USER, seed=0)
    xs = [torch.tensor(a, dtype=dt) for a in x]
    ys = [torch.tensor(a, dtype=dt) for a in y]
    opt = kt.KerasAdam(w)

    def step():
        loss, _, grads = kt.loss_and_grads(kt.others_lstm_span_whole_forward, w, xs, ys, [kt.mse] * 3)
        opt.step(grads)
        return float(loss)
    return step


def cpu_baseline(seconds=12.0, batch=32):
    import torch
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    step = _cpu_oracle_step_fn(batch)
    step()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds or n < 3:
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": batch * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d training steps of batch %d (the reference batch size) of the same config-2 model, "
                      "torch-CPU oracle (Keras-equivalent stand-in; Keras/TF1 not installable), fp32, %.1f s" % (n, batch, dt)}


def _time_cpu(fn, seconds, min_reps=2):
    fn()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds or n < min_reps:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n, n


def cpu_baseline_more(budget=1.0):
    """CPU-oracle points BASELINE.md section 4 asks for beside config 2 at batch 32: a large-batch config-2 point,
    config 1 (teacher-forced training) and config 3 (autoregressive inference) at the reference batch and a large
    batch, config 5 (heatmaps) train + inference.  torch-CPU restatement of the Keras graphs, all host threads;
    `budget` scales the seconds spent per point."""
    import torch
    from longterm360fov_b200 import data
    from oracle import keras_numpy as kn
    from oracle import keras_torch as kt
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    f32 = torch.float32
    res = {"cores": torch.get_num_threads(), "kind": "port"}
    for b in (64, 1024):
        step = _cpu_oracle_step_fn(b)
        dt, n = _time_cpu(step, 3.0 * budget)
        res["config2_train_batch%d" % b] = {"value": b / dt, "unit": UNIT, "steps": n}
    # config 1: FoV_seq2seq teacher-forced training; config 3: mu/var model, autoregressive inference
    for b in (32, 4096):
        w = kt.to_torch(kn.init_fov_seq2seq(seed=1), dtype=f32)
        e, d, t, _ = data.make_m1_batch(min(b, 512), seed=0)
        rep = b // len(e)
        xs = [torch.tensor(np.tile(a, (rep, 1, 1)), dtype=f32) for a in (e, d)]
        ys = [torch.tensor(np.tile(t, (rep, 1, 1)), dtype=f32)]
        opt = kt.KerasAdam(w)

        def step1():
            _, _, g = kt.loss_and_grads(kt.fov_seq2seq_forward, w, xs, ys, [kt.mse])
            opt.step(g)
        dt, n = _time_cpu(step1, 2.0 * budget)
        res["config1_train_batch%d" % b] = {"value": b / dt, "unit": UNIT, "steps": n}
    for b in (64, 16384):
        w = kt.to_torch(kn.init_fov_seq2seq(seed=2, num_encoder_tokens=6), dtype=f32)
        enc = torch.randn(b, 10, 6) * 0.3
        last = enc[:, -1:].clone()

        def infer3():
            with torch.no_grad():
                kt.fov_seq2seq_forward(w, enc, last, teacher_forcing=False)
        dt, n = _time_cpu(infer3, 2.0 * budget)
        res["config3_autoregressive_infer_batch%d" % b] = {"value": b / dt, "unit": UNIT, "steps": n}
    # config 5: heatmap ConvLSTM seq2seq with the 512/1024 heads (196.7 GFLOP forward per sequence)
    b = 2
    w = kt.to_torch(kn.init_convlstm_seq2seq(seed=2), dtype=f32)
    x, y = data.make_m4_batch(b, seed=7)
    xs, ys = [torch.tensor(a, dtype=f32) for a in x], [torch.tensor(a, dtype=f32) for a in y]
    opt5 = kt.KerasRMSprop(w)

    def infer5():
        with torch.no_grad():
            kt.convlstm_seq2seq_forward(w, *xs)

    def train5():
        _, _, g = kt.loss_and_grads(kt.convlstm_seq2seq_forward, w, xs, ys, [kt.mse])
        opt5.step(g)
    dt, n = _time_cpu(infer5, 1.0 * budget, min_reps=1)
    res["config5_heatmaps_infer_batch%d" % b] = {"value": 10 * b / dt, "unit": "heatmaps/s", "steps": n}
    dt, n = _time_cpu(train5, 1.0 * budget, min_reps=1)
    res["config5_heatmaps_train_batch%d" % b] = {"value": 10 * b / dt, "unit": "heatmaps/s", "steps": n}
    return res


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port (Keras cannot run here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core it can
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    batch = args.ref_batch
    step = _cpu_oracle_step_fn(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = batch * args.steps / dt
    cores = torch.get_num_threads()
    sample = "each step = one training step on a bounded sample of %d sequences of the same workload" % batch
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "configs[1]: others_LSTM_span_whole concat-state seq2seq (enc (B,10,6), others "
                               "(B,20,1,33,6), dec0 (B,1,6)), train step = fwd + BPTT + 3xMSE + Adam; CPU oracle, fp32",
                   "per_step_batch": batch,
                   "sample": "each step trains on %d sequences = 1/%d of the GPU arm's per-GPU batch (%d); CPU seq/s is "
                             "flat in the batch beyond 64 (cpu_baseline_more of the GPU arm's line)" % (
                                 batch, max(1, args.batch // batch), args.batch)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import data, ops, _lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B = args.batch
    peaks = _peaks()

    model = fov.others_lstm_span_whole(num_user=NUM_USER, seed=1, device=dev)
    model.compile(optimizer="Adam", loss=["mean_squared_error"] * 3, loss_weights=[1, 1, 1])
    model.set_compute(args.compute)
    if world > 1:
        model.distribute()

    # synthetic windows; a pool of distinct windows, resampled to the batch size
    pool = min(B, 512)
    px, py = data.make_m3_batch(pool, NUM_USER, seed=rank)
    rng = np.random.default_rng(100 + rank)
    host_batches = []
    for _ in range(2):
        idx = rng.integers(0, pool, B)
        host_batches.append(([torch.from_numpy(np.ascontiguousarray(a[idx])).pin_memory() for a in px],
                             [torch.from_numpy(np.ascontiguousarray(a[idx])).pin_memory() for a in py]))
    dev_batches = [([t.to(dev) for t in xs], [t.to(dev) for t in ys]) for xs, ys in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        xs, ys = dev_batches[i % 2]
        model.train_step_device(xs, ys)
    barrier()
    sampler.active = True
    n0 = lib.fov_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        xs, ys = dev_batches[i % 2]
        loss = model.train_step_device(xs, ys)
    e1.record()
    torch.cuda.synchronize()
    launches = int(lib.fov_launch_count() - n0)
    ms = e0.elapsed_time(e1)
    sampler.active = False
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    value = world * B * args.steps / (ms / 1e3)
    final_loss = float(loss.item())

    if args.profile_only:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / args.steps,
                              "gpu_launches": launches, "profile_only": True}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- end-to-end through the public API ----------------
    # model.fit_generator(...) is the call the reference's scripts make (mycode/convlstm_heatmap.py:415-418; model.fit
    # at mycode/others_LSTM_span_whole.py:778-785 runs the same loop): every step copies its inputs and targets from
    # pinned HOST memory and reads its loss back; the copy of batch i+1 overlaps the kernels of step i.
    # at least 40 steps: the first step of a fit_generator call cannot overlap its own H2D copy and batch build, and at
    # N = 8 that un-overlapped start is shared by 8 ranks on one PCIe / host memory system (reported in e2e.steps)
    e2e_steps = max(40, min(args.steps, 100))

    def host_gen():
        i = 0
        while True:
            yield host_batches[i % 2]
            i += 1

    def timed_fit(gen, builder=None):
        model.fit_generator(gen, steps_per_epoch=2, epochs=1, batch_builder=builder)
        barrier()
        e0.record()
        model.fit_generator(gen, steps_per_epoch=e2e_steps, epochs=1, batch_builder=builder)   # per step: H2D, step, D2H of the loss
        e1.record()
        torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms

    sampler.active = True
    # (a) the host prepares the featurised tensors, as the reference's scripts do (get_data -> get_gt_target_xyz ->
    #     get_whole_span on NumPy): 32 KB per sequence cross PCIe
    ms_host = timed_fit(host_gen())
    h2d_host = sum(t.numel() * 4 for t in host_batches[0][0] + host_batches[0][1])
    # (b) HEADLINE e2e: the host hands over RAW video chunks (viewers x seconds x 90 xyz floats, pinned) and the mean/var
    #     featuriser, the windowing and the target/others split run on the GPU (pipeline.M3VideoBatches): every viewer is
    #     the target once on the same raw data, so only ~3.6 KB per training sequence cross PCIe
    secs = 10 * ((B + NUM_USER - 1) // NUM_USER - 1) + 20
    raw_chunks = [torch.from_numpy(np.ascontiguousarray(
        data.synth_trajectories(1, NUM_USER, secs, seed=1000 + 10 * rank + i)[0].reshape(NUM_USER, secs, 90))).pin_memory()
        for i in range(2)]
    builder = fov.M3VideoBatches(num_user=NUM_USER, limit=B)

    def raw_gen():
        i = 0
        while True:
            yield raw_chunks[i % 2]
            i += 1
    ms_e2e = timed_fit(raw_gen(), builder)
    e2e_val = world * B * e2e_steps / (ms_e2e / 1e3)
    e2e_host_val = world * B * e2e_steps / (ms_host / 1e3)
    h2d = raw_chunks[0].numel() * 4

    # ---------------- inference throughput (forward only, resident inputs) ----------------
    with torch.no_grad():
        for _ in range(2):
            model._forward(dev_batches[0][0], False)
        barrier()
        e0.record()
        for i in range(args.steps):
            model._forward(dev_batches[i % 2][0], False)
        e1.record()
        torch.cuda.synchronize()
    ms_inf = e0.elapsed_time(e1)
    sampler.active = False
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms_inf], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_inf = float(t.item())
    infer_val = world * B * args.steps / (ms_inf / 1e3)

    # ---------------- parity AT THE BENCHMARKED SHAPE (rank 0): rows of the B-sequence forward vs the float64 oracle ----
    parity = None
    if rank == 0 and not args.no_parity:
        parity = bench_parity(model, dev_batches[0], B)

    # ---------------- the same step in the other arithmetic modes (rank 0 reports; every rank runs: DP collectives) ----
    modes = {args.compute: {"ms_per_step": ms / args.steps, "value": value}}
    if not args.no_modes:
        for mode in ("fp32", "bf16x3", "bf16x2"):
            if mode in modes:
                continue
            model.set_compute(mode)
            for i in range(2):
                model.train_step_device(*dev_batches[i % 2])
            barrier()
            nrep = 3
            e0.record()
            for i in range(nrep):
                model.train_step_device(*dev_batches[i % 2])
            e1.record()
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([t_ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_ms = float(t.item())
            modes[mode] = {"ms_per_step": t_ms / nrep, "value": world * B * nrep / (t_ms / 1e3)}
        model.set_compute(args.compute)
        barrier()

    # ---------------- strong scaling: the GLOBAL batch fixed at 8880 and at 1024, split over the ranks ----
    strong = None
    if not args.no_modes:
        strong = {"unit": UNIT, "note": "global batch fixed, per-rank batch = global / n_gpus; the driver's runs at "
                                        "N = 1, 2, 4, 8 give the strong-scaling curves"}
        for gb in (8880, 1024):
            per = gb // world
            if per < 1:
                continue
            sub = [([t[:per].contiguous() for t in xs], [t[:per].contiguous() for t in ys]) for xs, ys in dev_batches]
            for i in range(3):
                model.train_step_device(*sub[i % 2])
            barrier()
            nrep = 5
            e0.record()
            for i in range(nrep):
                model.train_step_device(*sub[i % 2])
            e1.record()
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([t_ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_ms = float(t.item())
            strong["global_batch_%d" % gb] = {"per_gpu_batch": per, "ms_per_step": t_ms / nrep,
                                              "value": per * world * nrep / (t_ms / 1e3)}
        barrier()

    # ---------------- config 5 under data parallelism: 60.7 MB gradient bucket per step (every rank runs; world > 1) ------
    m4_dp = None
    if world > 1 and not args.no_extras:
        m4 = fov.convlstm_seq2seq(seed=2, device=dev).compile("RMSprop", "mean_squared_error")
        m4.set_compute("bf16")
        m4.distribute(model.comm)
        hx, hy = data.make_m4_batch(32, seed=7 + rank)
        mxs, mys = m4._to_dev(hx), m4._to_dev(hy)
        for _ in range(2):
            m4.train_step_device(mxs, mys)
        barrier()
        e0.record()
        for _ in range(3):
            m4.train_step_device(mxs, mys)
        e1.record()
        torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1)
        t = torch.tensor([t_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
        m4_dp = {"unit": "heatmaps/s", "per_gpu_batch": 32, "compute": "bf16", "ms_per_step": t_ms / 3,
                 "value": world * 32 * 10 * 3 / (t_ms / 1e3), "allreduce_bytes": int(m4.n_flat * 4),
                 "note": "weak scaling; the gradient allreduce (fov_dp_allreduce, NCCL over NVLink) runs after the "
                         "backward pass, not overlapped with it"}
        del m4, mxs, mys
        torch.cuda.empty_cache()
        barrier()

    # ---------------- every C-ABI call family of the step, each timed alone with CUDA events (rank 0) ----------------
    roofline = None
    if rank == 0:
        roofline = kernel_rooflines(lib, dev, B, args.compute, peaks, value / world, ms / args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    extras = other_workloads(dev, peaks, args.compute) if world == 1 and not args.no_extras else None
    cpu = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
    cpu_more = cpu_baseline_more() if world == 1 and not args.no_cpu_baseline and not args.no_extras else None
    equal = equal_batch_e2e(dev, args.compute, cpu, cpu_more) if world == 1 and not args.no_extras else None
    saved_gb = B * (20 * 33 * (56 * 5 + 56) * 4 + 20 * 1848 * 4) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_OF[args.compute], "data": "synthetic",
        "config": {"workload": "configs[1]: others_LSTM_span_whole concat-state seq2seq (enc (B,10,6), others "
                               "(B,20,1,33,6), dec0 (B,1,6)), train step = fwd + BPTT + 3xMSE + Adam; compute=%s" % args.compute,
                   "per_gpu_batch": B, "global_batch": world * B, "parallelism": "dp%d" % world,
                   "batch_choice": "multiples of 148 SMs x 6 sequences per CTA fill whole waves of the persistent "
                                   "ConvLSTM kernels (4096: 391k, 4440: 407k, 8880: 424k, 17760: 436k seq/s on one B200)",
                   "params": model.count_params(),
                   "l2": "no flush needed: per-step working set (saved activations ~%.1f GB) >> 126 MB L2; "
                         "two alternating input batches" % saved_gb},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "api": "model.fit_generator(gen of pinned RAW video chunks (%d viewers x %d s x 90 xyz floats), "
                       "batch_builder=M3VideoBatches): H2D of chunk i+1 on a side stream under the kernels of step i; the "
                       "mean/var featuriser, windowing and target/others split (the reference's NumPy get_data / "
                       "get_gt_target_xyz / get_whole_span) run on the GPU; loss read back every step" % (NUM_USER, secs),
                "host_featurised": {"value": e2e_host_val, "h2d_bytes_per_step": h2d_host, "ms_per_step": ms_host / e2e_steps,
                                    "api": "model.fit_generator(gen of pinned, already featurised host batches) - round 1's "
                                           "e2e definition: 32 KB per sequence cross PCIe"}},
        "gpu_launches": launches,
        "infer": {"value": infer_val, "unit": UNIT, "ms_per_step": ms_inf / args.steps},
        "final_loss": final_loss,
        "parity": parity,
        "compute_modes": {"unit": UNIT, "note": "the same train step at the same batch in every arithmetic mode: "
                          "fp32 = CUDA-core kernels; bf16x3 = 3-term split (6 MMAs per MAC, ~24 mantissa bits); bf16x2 = "
                          "2-term split (3 MMAs per MAC, ~16 mantissa bits; measured forward error in `parity`)",
                          **modes},
        "strong_scaling": strong,
        "roofline": roofline,
        "clocks": sampler.result(),
    }
    if m4_dp is not None:
        out["convlstm_seq2seq_heatmap_data_parallel"] = m4_dp
    if extras is not None:
        out["other_workloads"] = extras
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if cpu_more is not None:
        out["cpu_baseline_more"] = cpu_more
    if equal is not None:
        out["equal_batch"] = equal
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_parity(model, dev_batch, B, n_rows=32):
    """Forward of the FULL benchmarked batch on the GPU (current weights, after the timed steps), `n_rows` of its rows
    against the float64 oracle run on exactly those rows: max-abs error per output (north-star bar 1e-4) and the
    per-row loss.  The oracle is the checker only."""
    import torch
    from longterm360fov_b200 import ops
    from oracle import keras_numpy as kn
    xs, ys = dev_batch
    with torch.no_grad():
        ops.set_math(model.compute)
        outs = model._forward(xs, False)
    rows = np.unique(np.r_[0:8, B - 8:B, np.random.default_rng(0).integers(0, B, n_rows)])[:n_rows]
    idx = torch.as_tensor(rows, device=xs[0].device)
    w64 = {k: v.astype(np.float64) for k, v in model.get_weights_dict().items()}
    ref = kn.others_lstm_span_whole_forward(w64, *[t[idx].cpu().numpy().astype(np.float64) for t in xs])
    errs, loss_err = [], 0.0
    for o, r, y in zip(outs, ref, ys):
        got = o[idx].cpu().numpy().astype(np.float64)
        errs.append(float(np.abs(got - r).max()))
        yt = y[idx].cpu().numpy().astype(np.float64)
        l_got = ((got - yt) ** 2).reshape(len(rows), -1).mean(1)
        l_ref = ((r - yt) ** 2).reshape(len(rows), -1).mean(1)
        loss_err = max(loss_err, float(np.abs(l_got - l_ref).max()))
    ok = max(errs) < 1e-4
    res = {"rows_checked": int(len(rows)), "batch": B, "compute": model.compute, "fwd_max_abs_err": errs,
           "bar": 1e-4, "per_row_loss_max_abs_err": loss_err, "ok": bool(ok),
           "against": "float64 NumPy oracle (oracle/keras_numpy.py) on the same rows, weights after the timed steps"}
    if not ok:
        raise SystemExit("bench parity FAILED at the benchmarked shape: %s" % json.dumps(res))
    return res


def equal_batch_e2e(dev, compute, cpu, cpu_more):
    """The reference's own batch sizes (32, 64), END TO END through model.fit_generator from pinned host batches, next
    to the CPU oracle at the same batch: the equal-batch ratio (the headline batch fills the GPU; these do not)."""
    import torch
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import data
    res = {}
    for b in (32, 64):
        m = fov.others_lstm_span_whole(num_user=NUM_USER, seed=1, device=dev)
        m.compile(optimizer="Adam", loss=["mean_squared_error"] * 3, loss_weights=[1, 1, 1])
        m.set_compute(compute)
        px, py = data.make_m3_batch(b, NUM_USER, seed=5)
        hb = ([torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in px],
              [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in py])

        def gen():
            while True:
                yield hb
        out = {}
        for graphed in (False, True):
            m.enable_cuda_graphs(graphed)
            m.fit_generator(gen(), steps_per_epoch=5, epochs=1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 100
            m.fit_generator(gen(), steps_per_epoch=n, epochs=1)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out["e2e_cuda_graph" if graphed else "e2e_eager"] = b * n / dt
        cpu_v = None
        if b == 32 and cpu is not None:
            cpu_v = cpu["value"]
        elif cpu_more is not None and ("config2_train_batch%d" % b) in cpu_more:
            cpu_v = cpu_more["config2_train_batch%d" % b]["value"]
        out["cpu_oracle"] = cpu_v
        out["ratio_e2e_over_cpu"] = (max(out["e2e_eager"], out["e2e_cuda_graph"]) / cpu_v) if cpu_v else None
        res["batch%d" % b] = out
        del m
    res["unit"] = UNIT
    res["note"] = "wall-clock around fit_generator (H2D of every batch + loss read-back every step inside)"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8880,
                    help="sequences per GPU per step (default 8880 = 148 SMs x 6 sequences per persistent-forward CTA x "
                         "10 waves: no partial last wave in the persistent ConvLSTM kernels; 4096 leaves 8 %% of one idle)")
    ap.add_argument("--ref-batch", type=int, default=1110,
                    help="sequences per step of the CPU reference arm (default: 1/8 of the GPU arm's per-GPU batch, "
                         "~2 s per CPU step)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check against the oracle")
    ap.add_argument("--no-modes", action="store_true", help="skip timing the step in the other arithmetic modes")
    ap.add_argument("--compute", default="bf16x2", choices=["fp32", "bf16", "bf16x2", "bf16x3"],
                    help="arithmetic of the conv/dense/ConvLSTM kernels: fp32 = CUDA cores; bf16x2 (default) = "
                         "tcgen05 with two bf16 terms per operand, fp32 accumulate (~16 mantissa bits; measured forward error in `parity`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other_workloads legs (configs 1, 3, 5)")
    ap.add_argument("--profile-only", action="store_true",
                    help="only the device-resident timed region (for ncu runs): no e2e / infer / cpu legs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
