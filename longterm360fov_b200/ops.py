"""Thin torch.autograd.Functions over the C ABI of libfov360.so.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every
FLOP on the hot path is one of our kernels.  Weight gradients are accumulated by
the kernels straight into the flat gradient bucket (the ``sink`` tensors), so the
Functions return ``None`` for weight inputs.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT, REC, ptr


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.FovError("fov ops need CUDA tensors; there is no CPU fallback")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _expect(cond, msg, *args):
    """Host-side operand checks: the C ABI gets raw pointers, so a shape that disagrees with the weights would be
    a silent out-of-bounds read of the flat parameter bucket, not an error."""
    if not cond:
        raise _lib.FovError(msg % args if args else msg)


# arithmetic used by the conv / dense / ConvLSTM Functions built from now on (see _lib.MATH);
# each Function records the mode of its forward for its backward.
_MATH = [_lib.MATH["bf16x2"]]


def set_math(mode):
    """mode: 'fp32' (CUDA cores), 'bf16', 'bf16x2' (default: 2 bf16 terms per operand, ~16 mantissa bits), 'bf16x3' (3 terms, fp32-grade)."""
    _MATH[0] = _lib.MATH[mode] if isinstance(mode, str) else int(mode)


def get_math():
    return _MATH[0]


def _ws(nbytes, device):
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- #
# weight gradients on a side stream
# --------------------------------------------------------------------------- #
# Nothing on the backward chain reads a weight gradient: only the optimiser does.  At small batches (the reference
# trains at 32 / 64) the step is a chain of dependent launches that leaves most SMs idle, so the weight-gradient
# launches can leave that chain: they fork onto a side stream after the kernel that produced their operands and the
# caller joins the stream before the optimiser step (join_wgrad_stream).  Tensors they read are held in a keep-list
# until that join (so nothing is recycled under them; record_stream would defer their release to an allocator event
# poll and make the pool churn at multi-GB batches).  Graph-capture safe (event fork / join).
_WG = {"on": False, "streams": {}, "used": False, "keep": []}


def set_wgrad_side_stream(on):
    _WG["on"] = bool(on) and os.environ.get("FOV_NO_WGRAD_STREAM", "0") in ("", "0")     # FOV_NO_WGRAD_STREAM=1: never
    if not on and (_WG["used"] or _WG["keep"]):      # a pass that ended without its join (an exception on the way)
        join_wgrad_stream()


def _wg_stream():
    """The side stream of the current device, or None when the feature is off."""
    if not _WG["on"]:
        return None
    dev = torch.cuda.current_device()
    st = _WG["streams"].get(dev)
    if st is None:
        st = _WG["streams"][dev] = torch.cuda.Stream(device=dev)
    _WG["used"] = True
    return st


def _wg_keep(side, *tensors):
    _WG["keep"].extend(t for t in tensors if t is not None)


class _WgFork:
    """``with _WgFork(x, dy) as side:`` - inside, the current stream is the side stream (forked after everything
    enqueued so far) when the feature is on, else nothing changes."""

    def __init__(self, *reads):
        self.reads = reads
        self.side = _wg_stream()
        self.ctx = None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            _wg_keep(self.side, *self.reads)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self.side

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join_wgrad_stream():
    """The current stream waits for the side-stream weight gradients of this pass (call before the optimiser)."""
    if _WG["used"]:
        cur = torch.cuda.current_stream()
        for st in _WG["streams"].values():
            if st.device == cur.device:
                cur.wait_stream(st)
        _WG["used"] = False
    _WG["keep"].clear()


# --------------------------------------------------------------------------- #
# layer wavefront of stacked ConvLSTMs at small batches
# --------------------------------------------------------------------------- #
# At the reference's batch sizes the three stacked recurrences (20 dependent steps each, ~8 us per step) run back to
# back on a few SMs.  When every CTA of the whole stack fits on the GPU at once, the layers are launched on separate
# streams instead and run CONCURRENTLY: layer l publishes a flag per (image group, timestep) once h_t is in global memory,
# layer l+1 waits for it before it reads x_t (fov_convlstm_io.wave_set / wave_wait); the BPTT kernels do the same in the
# other direction through the fused input gradient (fov_convlstm_grads.wave_set / wave_wait).
_WAVE = {"on": os.environ.get("FOV_NO_WAVEFRONT", "0") in ("", "0"), "streams": {}}      # FOV_NO_WAVEFRONT=1: never


def set_layer_wavefront(on):
    _WAVE["on"] = bool(on) and os.environ.get("FOV_NO_WAVEFRONT", "0") in ("", "0")


def _wave_streams(n):
    dev = torch.cuda.current_device()
    sts = _WAVE["streams"].setdefault(dev, [])
    while len(sts) < n:
        sts.append(torch.cuda.Stream(device=dev))
    return sts[:n]


def _wave_groups(lib, cfgs, backward, with_dx):
    """Image groups per layer when EVERY layer of the stack takes a persistent one-group-per-CTA kernel with the same
    grouping and all their CTAs can be resident at once; else 0."""
    if not _WAVE["on"] or len(cfgs) < 2 or cfgs[0].T < 2:
        return 0
    g0 = 0
    for l, cfg in enumerate(cfgs):
        cfg.wave_layers = len(cfgs)               # group size such that the whole stack fits on the SMs
        g = lib.fov_convlstm_wave_groups(C.byref(cfg), int(backward), int(with_dx[l]))
        if g <= 0 or (l and g != g0):
            g0 = 0
            break
        g0 = g
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    if g0 and g0 * len(cfgs) <= sms:
        return g0
    for cfg in cfgs:
        cfg.wave_layers = 0
    return 0


# --------------------------------------------------------------------------- #
# persistent fc-LSTM encoder-decoder
# --------------------------------------------------------------------------- #


class LSTMSeq2SeqFn(torch.autograd.Function):
    """y, enc_hseq, dec_hseq = f(x_enc, x_dec, extra | enc/dec/head weights).

    opts: dict(T_dec, teacher_forcing, head_act, rec_act, dec_zero_init, training, need_enc_hseq, need_dec_hseq)
    sinks: dict name -> gradient view for the 8 weights (or None at inference).
    Wo / bo may be None (no head: a lower layer of a stacked model, mycode/Fov_seq2seq_2layers.py:232-272, whose
    hidden sequences feed the layer above); gradients flow back into x_enc / x_dec (teacher forcing) when they need them.
    """

    @staticmethod
    def forward(ctx, opts, sinks, x_enc, x_dec, extra, We, Ue, be, Wd, Ud, bd, Wo, bo):
        lib = _lib.load()
        _require_cuda(x_enc, x_dec, We)
        x_enc, x_dec, extra = _f32c(x_enc), _f32c(x_dec), _f32c(extra)
        _expect(x_enc.dim() == 3 and x_dec.dim() == 3, "LSTMSeq2SeqFn: x_enc / x_dec must be (B,T,in)")
        B, T_enc, in_enc = x_enc.shape
        in_dec = x_dec.shape[2]
        H = Ue.shape[0]
        out_dim = Wo.shape[1] if Wo is not None else 0
        T_dec = opts["T_dec"]
        _expect(x_dec.shape[0] == B, "LSTMSeq2SeqFn: x_dec has %d sequences, x_enc %d", x_dec.shape[0], B)
        _expect(tuple(We.shape) == (in_enc, 4 * H) and tuple(Ue.shape) == (H, 4 * H) and be.numel() == 4 * H,
                "LSTMSeq2SeqFn: encoder weights %s/%s/%s do not fit input width %d, H=%d", tuple(We.shape),
                tuple(Ue.shape), tuple(be.shape), in_enc, H)
        _expect(tuple(Wd.shape) == (in_dec, 4 * H) and tuple(Ud.shape) == (H, 4 * H) and bd.numel() == 4 * H,
                "LSTMSeq2SeqFn: decoder weights %s/%s/%s do not fit input width %d, H=%d", tuple(Wd.shape),
                tuple(Ud.shape), tuple(bd.shape), in_dec, H)
        _expect(Wo is None or (Wo.shape[0] == H and bo.numel() == out_dim),
                "LSTMSeq2SeqFn: head weights do not fit H=%d", H)
        if opts["teacher_forcing"]:
            _expect(x_dec.shape[1] == T_dec, "LSTMSeq2SeqFn: teacher forcing needs x_dec (B,%d,in), got %s", T_dec,
                    tuple(x_dec.shape))
        else:
            _expect(x_dec.shape[1] == 1, "LSTMSeq2SeqFn: autoregressive decoding takes x_dec (B,1,in), got %s",
                    tuple(x_dec.shape))
            _expect(out_dim == in_dec, "LSTMSeq2SeqFn: re-feeding needs head width %d == decoder input width %d",
                    out_dim, in_dec)
        if extra is not None:
            _expect(tuple(extra.shape) == (B, T_dec, out_dim), "LSTMSeq2SeqFn: extra must be (B,T_dec,out), got %s",
                    tuple(extra.shape))
        training = bool(opts.get("training", False))
        dev = x_enc.device
        cfg = _lib.LstmCfg(B, T_enc, T_dec, in_enc, in_dec, H, out_dim,
                           int(opts["teacher_forcing"]), ACT[opts.get("head_act")],
                           REC[opts.get("rec_act", "hard_sigmoid")],
                           int(opts.get("dec_zero_init", False)), int(training), _MATH[0])
        w = _lib.LstmWeights(ptr(We), ptr(Ue), ptr(be), ptr(Wd), ptr(Ud), ptr(bd), ptr(Wo), ptr(bo))
        y = torch.empty(B, T_dec, out_dim, device=dev)
        # the encoder's h sequence is an output only for callers that read it (M3's target-past head)
        enc_hseq = torch.empty(B, T_enc, H, device=dev) if opts.get("need_enc_hseq", True) else None
        dec_hseq = torch.empty(B, T_dec, H, device=dev) if (training or opts.get("need_dec_hseq", False)) else None
        saved = {}
        if training:
            saved["enc_xh"] = torch.empty(B, T_enc, (H + in_enc + 3) // 4 * 4, device=dev)   # [h | x | 0] rows
            saved["enc_gates"] = torch.empty(B, T_enc, 4 * H, device=dev)
            saved["enc_c"] = torch.empty(B, T_enc, H, device=dev)
            saved["dec_xh"] = torch.empty(B, T_dec, (H + in_dec + 3) // 4 * 4, device=dev)
            saved["dec_gates"] = torch.empty(B, T_dec, 4 * H, device=dev)
            saved["dec_c"] = torch.empty(B, T_dec, H, device=dev)
            saved["dec_hseq"] = dec_hseq
        # workspace of the time-batched input projection (inputs wider than 16 features on the tensor-core path)
        nws = lib.fov_lstm_fwd_ws_bytes(C.byref(cfg))
        ws = _ws(nws, dev) if nws else None
        io = _lib.LstmIO(ptr(x_enc), ptr(x_dec), ptr(extra), None, None, ptr(y), None, None,
                         _lib.LstmSaved(ptr(saved.get("enc_xh")), ptr(saved.get("enc_gates")),
                                        ptr(saved.get("enc_c")), ptr(enc_hseq)),
                         _lib.LstmSaved(ptr(saved.get("dec_xh")), ptr(saved.get("dec_gates")),
                                        ptr(saved.get("dec_c")), ptr(dec_hseq)), ptr(ws))
        _lib.check(lib.fov_lstm_seq2seq_fwd(C.byref(cfg), C.byref(w), C.byref(io), _stream()),
                   "fov_lstm_seq2seq_fwd")
        if training:
            ctx.cfg, ctx.sinks, ctx.saved = cfg, sinks, saved
            ctx.has_extra = extra is not None
            ctx.need_dx = (ctx.needs_input_grad[2], ctx.needs_input_grad[3] and bool(opts["teacher_forcing"]))
            ctx.save_for_backward(x_enc, x_dec, extra, We, Ue, be, Wd, Ud, bd, Wo, bo, y, enc_hseq)
        nondiff = []
        if enc_hseq is None:
            enc_hseq = y.new_empty(0)
            nondiff.append(enc_hseq)
        if dec_hseq is None or not opts.get("need_dec_hseq", False):
            dec_hseq = y.new_empty(0)                 # saved tensor only: not an output of this call
            nondiff.append(dec_hseq)
        if nondiff:
            ctx.mark_non_differentiable(*nondiff)
        return y, enc_hseq, dec_hseq

    @staticmethod
    def backward(ctx, dy, dhseq_enc, dhseq_dec):
        lib = _lib.load()
        x_enc, x_dec, extra, We, Ue, be, Wd, Ud, bd, Wo, bo, y, enc_hseq = ctx.saved_tensors
        cfg, s, sv = ctx.cfg, ctx.sinks, ctx.saved
        dy = _f32c(dy) if dy is not None else torch.zeros_like(y)
        dhseq_enc = _f32c(dhseq_enc) if enc_hseq is not None else None
        dhseq_dec = _f32c(dhseq_dec) if (dhseq_dec is not None and dhseq_dec.numel()) else None
        dev = y.device
        dz_enc = torch.empty(cfg.B, cfg.T_enc, 4 * cfg.H, device=dev)
        dz_dec = torch.empty(cfg.B, cfg.T_dec, 4 * cfg.H, device=dev)
        dpre = torch.empty_like(y)
        ws = torch.empty(int(lib.fov_lstm_bwd_ws_floats(C.byref(cfg))), device=dev)
        w = _lib.LstmWeights(ptr(We), ptr(Ue), ptr(be), ptr(Wd), ptr(Ud), ptr(bd), ptr(Wo), ptr(bo))
        io = _lib.LstmIO(ptr(x_enc), ptr(x_dec), ptr(extra), None, None, ptr(y), None, None,
                         _lib.LstmSaved(ptr(sv["enc_xh"]), ptr(sv["enc_gates"]), ptr(sv["enc_c"]),
                                        ptr(enc_hseq)),
                         _lib.LstmSaved(ptr(sv["dec_xh"]), ptr(sv["dec_gates"]), ptr(sv["dec_c"]),
                                        ptr(sv["dec_hseq"])), None)
        side = _wg_stream()
        if side is not None:     # the weight-gradient GEMMs read these after this node has returned
            _wg_keep(side, dz_enc, dz_dec, dpre, ws, sv["enc_xh"], sv["dec_xh"], sv["dec_hseq"])
        g = _lib.LstmGrads(ptr(dy), ptr(dhseq_enc), ptr(y), ptr(dz_enc), ptr(dz_dec), ptr(dpre),
                           ptr(s["enc_kernel"]), ptr(s["enc_recurrent"]), ptr(s["enc_bias"]),
                           ptr(s["dec_kernel"]), ptr(s["dec_recurrent"]), ptr(s["dec_bias"]),
                           ptr(s.get("head_kernel")), ptr(s.get("head_bias")), ptr(ws), ptr(dhseq_dec),
                           side.cuda_stream if side is not None else None)
        _lib.check(lib.fov_lstm_seq2seq_bwd(C.byref(cfg), C.byref(w), C.byref(io), C.byref(g), _stream()),
                   "fov_lstm_seq2seq_bwd")
        d_extra = dpre if ctx.has_extra else None
        # gradients w.r.t. the inputs (a layer below in a stacked model): dx = dZ . W^T, one Dense backward-data GEMM
        dxs = [None, None]
        for i, (need, xin, dz, Wk) in enumerate(((ctx.need_dx[0], x_enc, dz_enc, We), (ctx.need_dx[1], x_dec, dz_dec, Wd))):
            if not need:
                continue
            rows, cin = xin.shape[0] * xin.shape[1], xin.shape[2]
            ccfg = _conv_cfg(rows, 1, 1, cin, 4 * cfg.H, 1, 1, (1, 1), None, 0.0, cin, cin, 4 * cfg.H, 4 * cfg.H)
            dx = torch.empty_like(xin)
            math = cfg.math
            if math == 0:
                wsd = torch.empty(Wk.numel(), device=dev)
                _lib.check(lib.fov_conv2d_bwd_data(C.byref(ccfg), ptr(dz), ptr(Wk), ptr(dx), ptr(wsd), _stream()),
                           "fov_conv2d_bwd_data")
            else:
                wsd = _ws(lib.fov_conv_tc_ws_bytes(C.byref(ccfg), math, 1), dev)
                _lib.check(lib.fov_conv2d_bwd_data_tc(C.byref(ccfg), ptr(dz), ptr(Wk), ptr(dx), ptr(wsd), math, _stream()),
                           "fov_conv2d_bwd_data_tc")
            dxs[i] = dx
        return (None, None, dxs[0], dxs[1], d_extra) + (None,) * 8


def lstm_states(x_enc, We, Ue, be, rec_act="hard_sigmoid", h0=None, c0=None):
    """encoder_model.predict: final [h, c] of the encoder LSTM (mycode/FoV_seq2seq.py:137,156)."""
    lib = _lib.load()
    _require_cuda(x_enc)
    x_enc = _f32c(x_enc)
    B, T, in_enc = x_enc.shape
    H = Ue.shape[0]
    cfg = _lib.LstmCfg(B, T, 0, in_enc, 0, H, 0, 1, 0, REC[rec_act], 0, 0, _MATH[0])
    w = _lib.LstmWeights(ptr(We), ptr(Ue), ptr(be), None, None, None, None, None)
    hT = torch.empty(B, H, device=x_enc.device)
    cT = torch.empty(B, H, device=x_enc.device)
    nws = lib.fov_lstm_fwd_ws_bytes(C.byref(cfg))
    ws = _ws(nws, x_enc.device) if nws else None
    io = _lib.LstmIO(ptr(x_enc), None, None, ptr(_f32c(h0)), ptr(_f32c(c0)), None, ptr(hT), ptr(cT),
                     _lib.LstmSaved(), _lib.LstmSaved(), ptr(ws))
    _lib.check(lib.fov_lstm_seq2seq_fwd(C.byref(cfg), C.byref(w), C.byref(io), _stream()), "lstm_states")
    return hT, cT


def lstm_decode_steps(x_dec, h0, c0, Wd, Ud, bd, Wo, bo, T_dec, teacher_forcing, head_act="tanh",
                      rec_act="hard_sigmoid", extra=None):
    """decoder_model.predict: run T_dec decoder steps from given states
    (mycode/FoV_seq2seq.py:139-148,167).  Returns (y, h, c)."""
    lib = _lib.load()
    _require_cuda(x_dec)
    x_dec = _f32c(x_dec)
    B, _, in_dec = x_dec.shape
    H = Ud.shape[0]
    out_dim = Wo.shape[1]
    cfg = _lib.LstmCfg(B, 0, T_dec, 0, in_dec, H, out_dim, int(teacher_forcing), ACT[head_act],
                       REC[rec_act], 0, 0, _MATH[0])
    w = _lib.LstmWeights(None, None, None, ptr(Wd), ptr(Ud), ptr(bd), ptr(Wo), ptr(bo))
    dev = x_dec.device
    y = torch.empty(B, T_dec, out_dim, device=dev)
    hT = torch.empty(B, H, device=dev)
    cT = torch.empty(B, H, device=dev)
    io = _lib.LstmIO(None, ptr(x_dec), ptr(_f32c(extra)), ptr(_f32c(h0)), ptr(_f32c(c0)), ptr(y), ptr(hT),
                     ptr(cT), _lib.LstmSaved(), _lib.LstmSaved(), None)
    _lib.check(lib.fov_lstm_seq2seq_fwd(C.byref(cfg), C.byref(w), C.byref(io), _stream()), "lstm_decode_steps")
    return y, hT, cT


# --------------------------------------------------------------------------- #
# conv / dense
# --------------------------------------------------------------------------- #


def _conv_cfg(N, H, W, Cin, Cout, kh, kw, dil, act, beta, x_img, x_pix, y_img, y_pix):
    return _lib.ConvCfg(N, H, W, Cin, Cout, kh, kw, dil[0], dil[1],
                        ((kh - 1) * dil[0]) // 2, ((kw - 1) * dil[1]) // 2,
                        x_img, x_pix, y_img, y_pix, ACT[act], float(beta))


class Conv2DFn(torch.autograd.Function):
    """keras Conv2D(padding='same')/Dense: x (N,H,W,Cin) NHWC, kernel (kh,kw,Cin,Cout).
    sinks = (gw_view, gb_view) or None."""

    @staticmethod
    def forward(ctx, opts, sinks, x, kernel, bias):
        lib = _lib.load()
        _require_cuda(x, kernel)
        x = _f32c(x)
        _expect(x.dim() == 4 and kernel.dim() == 4, "Conv2DFn: x must be (N,H,W,Cin) and kernel (kh,kw,Cin,Cout)")
        N, H, W, Cin = x.shape
        kh, kw, kc, Cout = kernel.shape
        _expect(kc == Cin, "Conv2DFn: kernel expects %d input channels, x has %d", kc, Cin)
        _expect(bias is None or bias.numel() == Cout, "Conv2DFn: bias does not have Cout=%d elements", Cout)
        dil = opts.get("dilation", (1, 1))
        act = opts.get("activation")
        y = torch.empty(N, H, W, Cout, device=x.device)
        cfg = _conv_cfg(N, H, W, Cin, Cout, kh, kw, dil, act, 0.0, H * W * Cin, Cin, H * W * Cout, Cout)
        math = opts.get("math", _MATH[0])
        if math == 0:
            _lib.check(lib.fov_conv2d_fwd(C.byref(cfg), ptr(x), ptr(kernel), ptr(bias), ptr(y), _stream()),
                       "fov_conv2d_fwd")
        else:
            ws = _packed_weights(lib, opts.get("pack_cache"), cfg, kernel, math, 0)
            _lib.check(lib.fov_conv2d_fwd_tc_packed(C.byref(cfg), ptr(x), ptr(bias), ptr(y), ptr(ws), math, _stream()),
                       "fov_conv2d_fwd_tc_packed")
        if opts.get("training", False):
            ctx.math = math
            ctx.pack_cache = opts.get("pack_cache")
            ctx.cfg, ctx.sinks, ctx.act = cfg, sinks, act
            ctx.need_dx = ctx.needs_input_grad[2]
            ctx.save_for_backward(x, kernel, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, kernel, y = ctx.saved_tensors
        cfg = ctx.cfg
        dy = _f32c(dy)
        st = _stream()
        if ctx.act not in (None, "linear"):
            dpre = torch.empty_like(y)
            rows = y.numel() // y.shape[-1]
            cols = y.shape[-1]
            _lib.check(lib.fov_act_bwd(ACT[ctx.act], rows, cols, ptr(y), cols, ptr(dy), cols, ptr(dpre),
                                       cols, st), "fov_act_bwd")
        else:
            dpre = dy
        cfg.act = 0
        gw, gb = ctx.sinks
        math = ctx.math
        with _WgFork(x, dpre):
            wst = _stream()
            if math == 0:
                _lib.check(lib.fov_conv2d_bwd_weight(C.byref(cfg), ptr(x), ptr(dpre), ptr(gw), ptr(gb), wst),
                           "fov_conv2d_bwd_weight")
            else:
                nws = lib.fov_conv_wgrad_ws_bytes(C.byref(cfg), math)   # > 0: wide k x k conv, TMA-fed plane kernel
                wws = _ws(nws, x.device) if nws else None
                _lib.check(lib.fov_conv2d_bwd_weight_tc_ws(C.byref(cfg), ptr(x), ptr(dpre), ptr(gw), ptr(gb), ptr(wws),
                                                           math, wst), "fov_conv2d_bwd_weight_tc_ws")
        dx = None
        if ctx.need_dx:
            dx = torch.empty_like(x)
            cfg.beta = 0.0
            if math == 0:
                ws = torch.empty(kernel.numel(), device=x.device)
                _lib.check(lib.fov_conv2d_bwd_data(C.byref(cfg), ptr(dpre), ptr(kernel), ptr(dx), ptr(ws), st),
                           "fov_conv2d_bwd_data")
            else:
                ws = _packed_weights(lib, ctx.pack_cache, cfg, kernel, math, 1)
                _lib.check(lib.fov_conv2d_bwd_data_tc_packed(C.byref(cfg), ptr(dpre), ptr(dx), ptr(ws), math, st),
                           "fov_conv2d_bwd_data_tc_packed")
        return None, None, dx, None, None


def _packed_weights(lib, cache, cfg, kernel, math, bwd_data):
    """bf16-term image of a layer's weights for the tensor-core conv kernels (fov_conv_tc_pack).  ``cache``: optional
    dict owned by the caller and valid while the weights do not change (one forward + backward pass): layers that run
    several times per pass - the head convolutions of the 10 decoder steps - are packed once."""
    key = None
    if cache is not None:
        key = (kernel.data_ptr(), tuple(kernel.shape), int(math), int(bwd_data), cfg.H, cfg.W, cfg.N)
        ws = cache.get(key)
        if ws is not None:
            return ws
    ws = _ws(lib.fov_conv_tc_ws_bytes(C.byref(cfg), math, bwd_data), kernel.device)
    _lib.check(lib.fov_conv_tc_pack(C.byref(cfg), ptr(kernel), ptr(ws), math, bwd_data, _stream()), "fov_conv_tc_pack")
    if key is not None:
        cache[key] = ws
    return ws


class TapStackConvFn(torch.autograd.Function):
    """Narrow convolution (Cout <= 32 over >= 128 input channels: the 1024 -> 30 Conv2D head of
    mycode/convlstm_seq2seq.py:179-181, the 1024 -> 3 Conv1D head of :187-189) on the tensor-core conv kernels with
    the kw column taps STACKED along the output channels: a kh x kw convolution with N = 32 issues kh*kw small
    tcgen05.mma per k-step, each at the ~45-cycle floor of a narrow MMA; the kh x 1 convolution onto kw*Cp stacked
    columns (tx, c) issues kw x fewer, kw x wider ones, and fov_tapstack_reduce folds the taps
    (y[n,h,w,c] = act(b[c] + sum_tx Y'[n,h,w+tx-pw,(tx,c)])).  Backward: the transpose unfold, then the ordinary
    backward-data / weight-gradient kernels on the stacked problem; the stacked weight gradient is folded back into the
    Keras-layout sink.  Same operands, semantics and sinks as Conv2DFn."""

    @staticmethod
    def stacked_kernel(kernel, Cp, cache):
        """(kh,kw,Cin,Cout) -> (kh,1,Cin,kw*Cp), columns (tx, c), zero columns for c >= Cout; cached per pass."""
        key = ("tapstack", kernel.data_ptr(), tuple(kernel.shape), Cp)
        if cache is not None and key in cache:
            return cache[key]
        kh, kw, Cin, Cout = kernel.shape
        ks = torch.zeros(kh, 1, Cin, kw * Cp, device=kernel.device)
        ks.view(kh, Cin, kw, Cp)[..., :Cout] = kernel.detach().permute(0, 2, 1, 3)
        if cache is not None:
            cache[key] = ks
        return ks

    @staticmethod
    def forward(ctx, opts, sinks, x, kernel, bias):
        lib = _lib.load()
        _require_cuda(x, kernel)
        x = _f32c(x)
        N, H, W, Cin = x.shape
        kh, kw, kc, Cout = kernel.shape
        _expect(kc == Cin, "TapStackConvFn: kernel expects %d input channels, x has %d", kc, Cin)
        _expect(bias is None or bias.numel() == Cout, "TapStackConvFn: bias does not have Cout=%d elements", Cout)
        act, math, cache = opts.get("activation"), opts.get("math", _MATH[0]), opts.get("pack_cache")
        Cp = (Cout + 15) // 16 * 16
        CW = kw * Cp
        ks = TapStackConvFn.stacked_kernel(kernel, Cp, cache)
        cfg = _conv_cfg(N, H, W, Cin, CW, kh, 1, (1, 1), None, 0.0, H * W * Cin, Cin, H * W * CW, CW)
        yp = torch.empty(N, H, W, CW, device=x.device)
        ws = _packed_weights(lib, cache, cfg, ks, math, 0)
        _lib.check(lib.fov_conv2d_fwd_tc_packed(C.byref(cfg), ptr(x), None, ptr(yp), ptr(ws), math, _stream()),
                   "fov_conv2d_fwd_tc_packed")
        y = torch.empty(N, H, W, Cout, device=x.device)
        pad_w = (kw - 1) // 2
        _lib.check(lib.fov_tapstack_reduce(N * H, W, kw, pad_w, Cp, Cout, ptr(yp), ptr(bias), ACT[act], ptr(y), _stream()),
                   "fov_tapstack_reduce")
        if opts.get("training", False):
            ctx.cfg, ctx.sinks, ctx.act, ctx.math, ctx.cache = cfg, sinks, act, math, cache
            ctx.geom = (N, H, W, Cin, kh, kw, Cout, Cp, pad_w)
            ctx.need_dx = ctx.needs_input_grad[2]
            ctx.save_for_backward(x, kernel, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, kernel, y = ctx.saved_tensors
        N, H, W, Cin, kh, kw, Cout, Cp, pad_w = ctx.geom
        cfg, math, st = ctx.cfg, ctx.math, _stream()
        dy = _f32c(dy)
        if ctx.act not in (None, "linear"):
            dpre = torch.empty_like(y)
            _lib.check(lib.fov_act_bwd(ACT[ctx.act], N * H * W, Cout, ptr(y), Cout, ptr(dy), Cout, ptr(dpre), Cout, st),
                       "fov_act_bwd")
        else:
            dpre = dy
        CW = kw * Cp
        dyp = torch.empty(N, H, W, CW, device=x.device)
        _lib.check(lib.fov_tapstack_expand(N * H, W, kw, pad_w, Cp, Cout, ptr(dpre), ptr(dyp), st), "fov_tapstack_expand")
        gw, gb = ctx.sinks
        with _WgFork(x, dyp):               # off the backward chain when the side stream is on (ops._WgFork)
            gws = torch.zeros(kh, 1, Cin, CW, device=x.device)
            gbs = torch.zeros(CW, device=x.device) if gb is not None else None
            nws = lib.fov_conv_wgrad_ws_bytes(C.byref(cfg), math)
            wws = _ws(nws, x.device) if nws else None
            _lib.check(lib.fov_conv2d_bwd_weight_tc_ws(C.byref(cfg), ptr(x), ptr(dyp), ptr(gws), ptr(gbs), ptr(wws), math,
                                                       _stream()), "fov_conv2d_bwd_weight_tc_ws")
            gw.view(kh, kw, Cin, Cout).add_(gws.view(kh, Cin, kw, Cp)[..., :Cout].permute(0, 2, 1, 3))
            if gb is not None:              # the centre tap's columns are dpre itself (shift 0): their sums are the bias gradient
                gb += gbs[pad_w * Cp:pad_w * Cp + Cout]
        dx = None
        if ctx.need_dx:
            dx = torch.empty_like(x)
            ks = TapStackConvFn.stacked_kernel(kernel, Cp, ctx.cache)
            ws = _packed_weights(lib, ctx.cache, cfg, ks, math, 1)
            _lib.check(lib.fov_conv2d_bwd_data_tc_packed(C.byref(cfg), ptr(dyp), ptr(dx), ptr(ws), math, st),
                       "fov_conv2d_bwd_data_tc_packed")
        return None, None, dx, None, None


_TAPSTACK = [True]          # A/B switch (set_tapstack)


def set_tapstack(on):
    """Diagnostics: route narrow wide-input convolutions through TapStackConvFn (default) or the plain kernel."""
    _TAPSTACK[0] = bool(on)


def conv2d(x, kernel, bias, activation=None, dilation=(1, 1), sinks=None, training=False, pack_cache=None):
    opts = {"activation": activation, "dilation": dilation, "training": training, "pack_cache": pack_cache}
    kh, kw, Cin, Cout = kernel.shape
    if (_TAPSTACK[0] and _MATH[0] != 0 and tuple(dilation) == (1, 1) and kw >= 3 and Cout <= 32 and Cin >= 128 and
            kw * ((Cout + 15) // 16 * 16) <= 256):
        return TapStackConvFn.apply(opts, sinks, x, kernel, bias)
    return Conv2DFn.apply(opts, sinks, x, kernel, bias)


def dense(x, kernel, bias, activation=None, sinks=None, training=False):
    """keras Dense on the last axis: the 1x1 case of the conv family."""
    shp = x.shape
    x2 = x.reshape(-1, 1, 1, shp[-1])
    y = Conv2DFn.apply({"activation": activation, "training": training}, sinks, x2,
                       kernel.view(1, 1, *kernel.shape), bias)
    return y.reshape(*shp[:-1], kernel.shape[1])


class DualDenseFn(torch.autograd.Function):
    """Two Dense layers reading the SAME (B,T,C) sequence: y1 = x.W1 + b1 on every time slice and
    y2 = x[:, t0:].W2 + b2 on the slices from t0 on (the reconstruction head and the concat-state
    fusion of mycode/others_LSTM_span_whole.py:105-109,261-271).  One Function so that the two input
    gradients are accumulated by the kernels into ONE buffer (beta = 1 on the strided view) instead of
    slice-backward zero fills, copies and adds; the strided view is read in place (no contiguous copy).

    forward(opts, sinks1, sinks2, x, W1, b1, W2, b2) -> (y1 (B,T,C1), y2 (B,T-t0,C2))"""

    @staticmethod
    def _cfgs(B, T, t0, Cin, C1, C2):
        full = _conv_cfg(B * T, 1, 1, Cin, C1, 1, 1, (1, 1), None, 0.0, Cin, Cin, C1, C1)
        Ts = T - t0
        part = _conv_cfg(B, 1, Ts, Cin, C2, 1, 1, (1, 1), None, 0.0, T * Cin, Cin, Ts * C2, C2)
        return full, part

    @staticmethod
    def forward(ctx, opts, sinks1, sinks2, x, W1, b1, W2, b2):
        lib = _lib.load()
        _require_cuda(x, W1, W2)
        x = _f32c(x)
        _expect(x.dim() == 3, "DualDenseFn: x must be (B,T,C)")
        B, T, Cin = x.shape
        t0 = int(opts["t0"])
        C1, C2 = W1.shape[-1], W2.shape[-1]
        _expect(W1.shape[-2] == Cin and W2.shape[-2] == Cin, "DualDenseFn: kernels expect %d / %d inputs, x has %d",
                W1.shape[-2], W2.shape[-2], Cin)
        _expect(b1.numel() == C1 and b2.numel() == C2 and 0 <= t0 < T, "DualDenseFn: bias / t0 mismatch")
        math = opts.get("math", _MATH[0])
        full, part = DualDenseFn._cfgs(B, T, t0, Cin, C1, C2)
        y1 = torch.empty(B, T, C1, device=x.device)
        y2 = torch.empty(B, T - t0, C2, device=x.device)
        xs = x.data_ptr() + 4 * t0 * Cin
        st = _stream()
        for cfg, xp, W, b, y in ((full, x.data_ptr(), W1, b1, y1), (part, xs, W2, b2, y2)):
            if math == 0:
                _lib.check(lib.fov_conv2d_fwd(C.byref(cfg), xp, ptr(W), ptr(b), ptr(y), st), "fov_conv2d_fwd")
            else:
                ws = _ws(lib.fov_conv_tc_ws_bytes(C.byref(cfg), math, 0), x.device)
                _lib.check(lib.fov_conv2d_fwd_tc(C.byref(cfg), xp, ptr(W), ptr(b), ptr(y), ptr(ws), math, st),
                           "fov_conv2d_fwd_tc")
        if opts.get("training", False):
            ctx.math, ctx.t0, ctx.sinks = math, t0, (sinks1, sinks2)
            ctx.need_dx = ctx.needs_input_grad[3]
            ctx.save_for_backward(x, W1, W2)
        return y1, y2

    @staticmethod
    def backward(ctx, dy1, dy2):
        lib = _lib.load()
        x, W1, W2 = ctx.saved_tensors
        B, T, Cin = x.shape
        t0, math = ctx.t0, ctx.math
        C1, C2 = W1.shape[-1], W2.shape[-1]
        full, part = DualDenseFn._cfgs(B, T, t0, Cin, C1, C2)
        dy1 = _f32c(dy1) if dy1 is not None else torch.zeros(B, T, C1, device=x.device)
        dy2 = _f32c(dy2) if dy2 is not None else torch.zeros(B, T - t0, C2, device=x.device)
        st = _stream()
        xs = x.data_ptr() + 4 * t0 * Cin
        dx = torch.empty_like(x) if ctx.need_dx else None
        with _WgFork(x, dy1, dy2):
            wst = _stream()
            for cfg, xp, W, dy, sinks in ((full, x.data_ptr(), W1, dy1, ctx.sinks[0]), (part, xs, W2, dy2, ctx.sinks[1])):
                gw, gb = sinks
                if math == 0:
                    _lib.check(lib.fov_conv2d_bwd_weight(C.byref(cfg), xp, ptr(dy), ptr(gw), ptr(gb), wst),
                               "fov_conv2d_bwd_weight")
                else:
                    _lib.check(lib.fov_conv2d_bwd_weight_tc(C.byref(cfg), xp, ptr(dy), ptr(gw), ptr(gb), math, wst),
                               "fov_conv2d_bwd_weight_tc")
        if dx is not None:
            # dx[:, :t0] = dy1[:, :t0] . W1^T ;  dx[:, t0:] = [dy1[:, t0:] | dy2] . [W1 | W2]^T  - ONE GEMM (K = C1 + C2) for the
            # slices both layers read, each slice of dx written exactly once.  (Two accumulating GEMMs re-read and
            # re-wrote the 10 future slices: 1.3 GB of extra traffic per step at B = 8880, 1.16 ms -> 0.6 ms.)
            Ts = T - t0
            jobs = []
            if t0 > 0:
                past = _conv_cfg(B, 1, t0, Cin, C1, 1, 1, (1, 1), None, 0.0, T * Cin, Cin, T * C1, C1)
                jobs.append((past, dy1.data_ptr(), W1, dx.data_ptr()))
            dcat = torch.cat([dy1[:, t0:], dy2], dim=-1)                       # (B, Ts, C1 + C2), small next to dx
            Wcat = torch.cat([W1, W2], dim=-1)
            fut = _conv_cfg(B, 1, Ts, Cin, C1 + C2, 1, 1, (1, 1), None, 0.0, T * Cin, Cin, Ts * (C1 + C2), C1 + C2)
            jobs.append((fut, dcat.data_ptr(), Wcat, dx.data_ptr() + 4 * t0 * Cin))
            for cfg, dyp, W, dxp in jobs:
                if math == 0:
                    ws = torch.empty(W.numel(), device=x.device)
                    _lib.check(lib.fov_conv2d_bwd_data(C.byref(cfg), dyp, ptr(W), dxp, ptr(ws), st), "fov_conv2d_bwd_data")
                else:
                    ws = _ws(lib.fov_conv_tc_ws_bytes(C.byref(cfg), math, 1), x.device)
                    _lib.check(lib.fov_conv2d_bwd_data_tc(C.byref(cfg), dyp, ptr(W), dxp, ptr(ws), math, st),
                               "fov_conv2d_bwd_data_tc")
        return None, None, None, dx, None, None, None, None


def dual_dense(x, W1, b1, W2, b2, t0, sinks1=None, sinks2=None, training=False):
    """(x.W1 + b1 on all T slices, x[:, t0:].W2 + b2): see DualDenseFn."""
    return DualDenseFn.apply({"t0": t0, "training": training}, sinks1, sinks2, x, W1.view(1, 1, *W1.shape), b1,
                             W2.view(1, 1, *W2.shape), b2)


def conv1d(x, kernel, bias, activation=None, sinks=None, training=False):
    """keras Conv1D(padding='same'): x (B,L,Cin), kernel (k,Cin,Cout)."""
    y = conv2d(x.unsqueeze(1), kernel.unsqueeze(0), bias, activation, (1, 1), sinks, training)
    return y.squeeze(1)


# --------------------------------------------------------------------------- #
# ConvLSTM2D stack
# --------------------------------------------------------------------------- #


class ConvLSTMStackFn(torch.autograd.Function):
    """L stacked ConvLSTM2D layers (return_sequences) writing straight into the
    channel-concatenated sequence (mycode/others_LSTM_span_whole.py:88-102,
    mycode/convlstm_seq2seq.py:100-126,213-220).

    forward(opts, sinks, x, *flat) with flat = [K_l, R_l, b_l]*L + [h0_l, c0_l]*L (None allowed)
    returns (concat_seq (B,T,H,W,sumF), hT_0, cT_0, ..., hT_{L-1}, cT_{L-1})
    """

    @staticmethod
    def forward(ctx, opts, sinks, x, *flat):
        lib = _lib.load()
        _require_cuda(x)
        x = _f32c(x)
        L = opts["layers"]
        weights = [flat[3 * l:3 * l + 3] for l in range(L)]
        states = [tuple(_f32c(s) for s in flat[3 * L + 2 * l:3 * L + 2 * l + 2]) for l in range(L)]
        _expect(x.dim() == 5, "ConvLSTMStackFn: x must be (B,T,H,W,C)")
        B, T, H, W, Cin0 = x.shape
        Fs = [w[1].shape[2] for w in weights]
        cprev = Cin0
        for l, (K, R, b) in enumerate(weights):
            F = Fs[l]
            _expect(K.dim() == 4 and R.dim() == 4 and K.shape[:2] == R.shape[:2],
                    "ConvLSTMStackFn: layer %d kernel / recurrent kernel shapes %s / %s", l, tuple(K.shape), tuple(R.shape))
            _expect(K.shape[2] == cprev, "ConvLSTMStackFn: layer %d kernel expects %d input channels, gets %d", l,
                    K.shape[2], cprev)
            _expect(K.shape[3] == 4 * F and R.shape[3] == 4 * F and b.numel() == 4 * F,
                    "ConvLSTMStackFn: layer %d gate width is not 4F = %d", l, 4 * F)
            for s_ in states[l]:
                _expect(s_ is None or tuple(s_.shape) == (B, H, W, F),
                        "ConvLSTMStackFn: layer %d initial state must be (B,H,W,F) = %s, got %s", l, (B, H, W, F),
                        None if s_ is None else tuple(s_.shape))
            _expect((states[l][0] is None) == (states[l][1] is None), "ConvLSTMStackFn: give both h0 and c0 or neither")
            cprev = F
        Fsum = sum(Fs)
        dev = x.device
        dil = opts.get("dilation", (1, 1))
        rec = REC[opts.get("rec_act", "hard_sigmoid")]
        training = bool(opts.get("training", False))
        math = opts.get("math", _MATH[0])
        cat = torch.empty(B, T, H, W, Fsum, device=dev)
        HW = H * W
        st = _stream()
        outs, saved, cfgs = [], [], []
        cin, off = Cin0, 0
        cur, cur_b, cur_t, cur_pix = x, T * HW * Cin0, HW * Cin0, Cin0
        cur_ptr = ptr(x)
        masks = opts.get("dropout_masks") or [None] * L
        drop = []
        # layer wavefront (small batches, no dropout): every layer on its own stream, per-step flags between them
        wave_g, wflags, wstreams, main_s = 0, None, None, torch.cuda.current_stream()
        wave_keep = []
        if not (training and any(m is not None for m in masks)):
            probe, c_ = [], Cin0
            for l in range(L):
                K_ = weights[l][0]
                probe.append(_lib.ConvLstmCfg(B, T, H, W, c_, Fs[l], K_.shape[0], K_.shape[1], dil[0], dil[1], rec,
                                              (T * HW * Cin0) if l == 0 else T * HW * Fsum, (HW * Cin0) if l == 0 else HW * Fsum,
                                              Cin0 if l == 0 else Fsum, T * HW * Fsum, HW * Fsum, Fsum, int(training), math, 0))
                c_ = Fs[l]
            wave_g = _wave_groups(lib, probe, False, [False] * L)
            if wave_g:
                wflags = torch.zeros(max(L - 1, 1), wave_g, T, dtype=torch.int32, device=dev)
                wstreams = [main_s] + _wave_streams(L - 1)
                for s_ in wstreams[1:]:
                    s_.wait_stream(main_s)
        if (not training and not wave_g and _WAVE["on"] and L >= 2 and T >= 2 and math != 0
                and all(lib.fov_convlstm_fwd_ws_bytes(C.byref(c_)) > 0 and not lib.fov_convlstm_fwd_persistent(C.byref(c_))
                        for c_ in probe)):
            # inference on images larger than one MMA tile (config 5's 36 x 18 heat maps): every (layer, timestep) is its
            # own launch, so the layers can run as a wavefront at LAUNCH level - layer l step t next to layer l+1 step
            # t-1 - on one stream per layer with an event per (layer, step): (T + L - 1) launch slots instead of T * L
            return ConvLSTMStackFn._forward_step_wavefront(lib, opts, x, weights, states, probe, Fs, math)
        for l in range(L):
            K, R, b = weights[l]
            kh, kw = K.shape[0], K.shape[1]
            F = Fs[l]
            mask = masks[l] if training else None
            x4 = K4 = None
            if mask is not None:
                # keras ConvLSTM2D(dropout=p): one input mask per gate, constant over time == a 4x wider input
                # [x*m_i | x*m_f | x*m_c | x*m_o] with a block kernel (include/fov360.h fov_dropout_expand)
                mask = _f32c(mask)
                assert tuple(mask.shape) == (4, B, H, W, cin), (mask.shape, (4, B, H, W, cin))
                x4 = torch.empty(B, T, H, W, 4 * cin, device=dev)
                K4 = torch.empty(kh, kw, 4 * cin, 4 * F, device=dev)
                _lib.check(lib.fov_dropout_expand(B, T, HW, cin, cur_ptr, cur_b, cur_t, cur_pix, ptr(mask), ptr(x4), st),
                           "fov_dropout_expand")
                _lib.check(lib.fov_gate_kernel_expand(kh * kw, cin, F, ptr(K), ptr(K4), st), "fov_gate_kernel_expand")
                lx_ptr, lx_b, lx_t, lx_pix, lcin, lK = ptr(x4), T * HW * 4 * cin, HW * 4 * cin, 4 * cin, 4 * cin, K4
            else:
                lx_ptr, lx_b, lx_t, lx_pix, lcin, lK = cur_ptr, cur_b, cur_t, cur_pix, cin, K
            drop.append((mask, x4, K4))
            cfg = _lib.ConvLstmCfg(B, T, H, W, lcin, F, kh, kw, dil[0], dil[1], rec,
                                   lx_b, lx_t, lx_pix, T * HW * Fsum, HW * Fsum, Fsum, int(training), math, 0,
                                   L if wave_g else 0)
            # saved gates exist only for BPTT; the fused tensor-core step never materialises them otherwise
            fws_bytes = lib.fov_convlstm_fwd_ws_bytes(C.byref(cfg))     # > 0: the fused tensor-core step runs
            gates = torch.empty((B, T, H, W, 4 * F) if (training or fws_bytes == 0) else (1,), device=dev)
            fws = None
            if fws_bytes:
                # packed gate weights: shared by the calls of one pass when the caller hands in a cache (the one-step
                # decoder calls of convlstm_seq2seq run the same three layers 10 times)
                pcache = opts.get("pack_cache")
                pkey = (lK.data_ptr(), R.data_ptr(), tuple(lK.shape), math, B, H, W) if (pcache is not None and mask is None) else None
                fws = pcache.get(pkey) if pkey is not None else None
                if fws is not None:
                    cfg.ws_prepacked = 1
                else:
                    fws = _ws(fws_bytes, dev)
                    if pkey is not None:
                        pcache[pkey] = fws
            cseq = torch.empty(B, T, H, W, F, device=dev)
            hT = torch.empty(B, H, W, F, device=dev)
            cT = torch.empty(B, H, W, F, device=dev)
            h0, c0 = states[l]
            hptr = cat.data_ptr() + 4 * off
            w_wait = wflags[l - 1].data_ptr() if (wave_g and l > 0) else None
            w_set = wflags[l].data_ptr() if (wave_g and l < L - 1) else None
            # wavefront: this layer runs on its own stream while the loop goes on - its workspace must outlive the loop
            # iteration (a block freed here could be handed to the next layer's allocation on the main stream)
            wave_keep.append(fws)
            io = _lib.ConvLstmIO(lx_ptr, ptr(lK), ptr(R), ptr(b), ptr(h0), ptr(c0), hptr,
                                 ptr(gates), ptr(cseq), ptr(hT), ptr(cT), ptr(fws), w_wait, w_set)
            _lib.check(lib.fov_convlstm_fwd(C.byref(cfg), C.byref(io), wstreams[l].cuda_stream if wave_g else st),
                       "fov_convlstm_fwd")
            outs += [hT, cT]
            saved.append((gates, cseq))
            cfgs.append((cfg, cur_ptr, hptr, off, F, (cur_b, cur_t, cur_pix, cin)))
            cur_ptr, cur_b, cur_t, cur_pix, cin = hptr, T * HW * Fsum, HW * Fsum, Fsum, F
            off += F
        if wave_g:
            for s_ in wstreams[1:]:
                main_s.wait_stream(s_)
        del wave_keep
        if training:
            ctx.sinks, ctx.cfgs, ctx.L, ctx.drop = sinks, cfgs, L, drop
            ctx.x_needs_grad = ctx.needs_input_grad[2]
            ctx.state_needs_grad = [ctx.needs_input_grad[3 + 3 * L + 2 * l] for l in range(L)]
            tensors = [x, cat]
            for l in range(L):
                tensors += list(weights[l]) + list(states[l]) + list(saved[l])
            ctx.save_for_backward(*tensors)
        return (cat,) + tuple(outs)

    @staticmethod
    def _forward_step_wavefront(lib, opts, x, weights, states, probe, Fs, math):
        B, T, H, W, Cin0 = x.shape
        L, HW, Fsum, dev = len(weights), H * W, sum(Fs), x.device
        cat = torch.empty(B, T, H, W, Fsum, device=dev)
        main_s = torch.cuda.current_stream()
        streams = [main_s] + _wave_streams(L - 1)
        for s_ in streams[1:]:
            s_.wait_stream(main_s)
        pcache = opts.get("pack_cache")
        lay, off, cin = [], 0, Cin0
        for l in range(L):
            K, R, b = weights[l]
            F = Fs[l]
            c1 = probe[l]
            cfg = _lib.ConvLstmCfg(B, 1, H, W, cin, F, K.shape[0], K.shape[1], c1.dil_h, c1.dil_w, c1.rec_act,
                                   c1.x_b_stride, c1.x_t_stride, c1.x_pix_stride, T * HW * Fsum, HW * Fsum, Fsum, 0, math, 0, 0)
            fws_bytes = lib.fov_convlstm_fwd_ws_bytes(C.byref(cfg))
            pkey = (K.data_ptr(), R.data_ptr(), tuple(K.shape), math, B, H, W) if pcache is not None else None
            fws = pcache.get(pkey) if pkey is not None else None
            packed = fws is not None
            if fws is None:
                fws = _ws(fws_bytes, dev)
                if pkey is not None:
                    pcache[pkey] = fws
            x_base = x.data_ptr() if l == 0 else cat.data_ptr() + 4 * (off - Fs[l - 1])
            lay.append(dict(cfg=cfg, K=K, R=R, b=b, fws=fws, packed=packed, x_base=x_base, x_t=4 * c1.x_t_stride,
                            h_base=cat.data_ptr() + 4 * off, h_t=4 * HW * Fsum,
                            st=[(torch.empty(B, H, W, F, device=dev), torch.empty(B, H, W, F, device=dev)) for _ in range(2)],
                            cseq=torch.empty(B, 1, H, W, F, device=dev), gates=torch.empty(1, device=dev),
                            h=states[l][0], c=states[l][1]))
            off += F
            cin = F
        ev = [[torch.cuda.Event() for _ in range(T)] for _ in range(L)]
        for t in range(T):
            for l, d in enumerate(lay):
                s_ = streams[l]
                if l > 0:
                    s_.wait_event(ev[l - 1][t])
                d["cfg"].ws_prepacked = 1 if (d["packed"] or t > 0) else 0
                hT, cT = d["st"][t & 1]
                io = _lib.ConvLstmIO(d["x_base"] + t * d["x_t"], ptr(d["K"]), ptr(d["R"]), ptr(d["b"]), ptr(d["h"]), ptr(d["c"]),
                                     d["h_base"] + t * d["h_t"], ptr(d["gates"]), ptr(d["cseq"]), ptr(hT), ptr(cT), ptr(d["fws"]))
                _lib.check(lib.fov_convlstm_fwd(C.byref(d["cfg"]), C.byref(io), s_.cuda_stream), "fov_convlstm_fwd")
                ev[l][t].record(s_)
                d["h"], d["c"] = hT, cT
        for s_ in streams[1:]:
            main_s.wait_stream(s_)
        outs = []
        for d in lay:
            outs += [d["h"], d["c"]]
        return (cat,) + tuple(outs)

    @staticmethod
    def backward(ctx, dcat, *dstates):
        lib = _lib.load()
        L, cfgs = ctx.L, ctx.cfgs
        tensors = ctx.saved_tensors
        x, cat = tensors[0], tensors[1]
        per = [tensors[2 + 7 * l:2 + 7 * l + 7] for l in range(L)]   # K,R,b,h0,c0,gates,cseq
        dev = x.device
        st = _stream()
        # layers accumulate dx of the layer above into the incoming gradient buffer in place (it is owned by the
        # autograd engine and handed to this node only; no retain_graph use on this path)
        dcat = torch.zeros_like(cat) if dcat is None else dcat.contiguous()
        dx0 = torch.empty_like(x) if ctx.x_needs_grad else None
        dstate_out = [None] * (2 * L)
        # layer wavefront of the BPTT kernels: layer l waits per step for the fused dx of layer l+1
        wave_g, wflags, wstreams, main_s = 0, None, None, torch.cuda.current_stream()
        wave_keep = []          # per-layer temporaries stay alive until the layers' streams have been joined
        if all(d[0] is None for d in ctx.drop) and not any(ctx.state_needs_grad[l] and per[l][3] is not None for l in range(L)):
            wave_g = _wave_groups(lib, [c[0] for c in cfgs], True, [l > 0 or ctx.x_needs_grad for l in range(L)])
            if wave_g:
                wflags = torch.zeros(max(L - 1, 1), wave_g, cfgs[0][0].T, dtype=torch.int32, device=dev)
                wstreams = [main_s] + _wave_streams(L - 1)
                for s_ in wstreams[1:]:
                    s_.wait_stream(main_s)
        for l in reversed(range(L)):
            cfg, _xptr, _hptr, off, F, xgeo = cfgs[l]
            K, R, b, h0, c0, gates, cseq = per[l]
            mask, x4, K4 = ctx.drop[l]
            xptr = x.data_ptr() if l == 0 else cat.data_ptr() + 4 * cfgs[l - 1][3]
            hptr = cat.data_ptr() + 4 * off
            dhT, dcT = _f32c(dstates[2 * l]), _f32c(dstates[2 * l + 1])
            nws = lib.fov_convlstm_bwd_ws_floats(C.byref(cfg))
            ws = torch.empty(nws, device=dev)
            wave_keep += [ws, dhT, dcT]
            dh0 = dc0 = None
            if ctx.state_needs_grad[l] and h0 is not None:
                dh0 = torch.empty_like(h0)
                dc0 = torch.empty_like(c0)
            if l == 0:
                dxp, acc = ptr(dx0), 0
            else:
                dxp, acc = dcat.data_ptr() + 4 * cfgs[l - 1][3], 1
            gk, gr_, gb = ctx.sinks[l]
            if mask is None:
                io = _lib.ConvLstmIO(xptr, ptr(K), ptr(R), ptr(b), ptr(h0), ptr(c0), hptr,
                                     ptr(gates), ptr(cseq), None, None, None)
                side = _wg_stream()
                if side is not None:     # the layer's weight gradient reads these after this node has returned
                    _wg_keep(side, gates, cat, x, h0)
                # producer of layer l's dhseq = layer l+1 (flags[l]); this layer publishes into flags[l-1]
                w_wait = wflags[l].data_ptr() if (wave_g and l < L - 1) else None
                w_set = wflags[l - 1].data_ptr() if (wave_g and l > 0) else None
                g = _lib.ConvLstmGrads(dcat.data_ptr() + 4 * off, ptr(dhT), ptr(dcT), dxp, ptr(dh0), ptr(dc0),
                                       ptr(gk), ptr(gr_), ptr(gb), ptr(ws), acc,
                                       side.cuda_stream if side is not None else None, w_wait, w_set)
                _lib.check(lib.fov_convlstm_bwd(C.byref(cfg), C.byref(io), C.byref(g),
                                                wstreams[l].cuda_stream if wave_g else st), "fov_convlstm_bwd")
            else:
                # widened-input form: gradients w.r.t. x4 / K4, folded back through the masks / the gate blocks
                xb_, xt_, xpix_, cin_ = xgeo
                gk4 = torch.zeros_like(K4)
                dx4 = torch.empty_like(x4) if dxp else None
                io = _lib.ConvLstmIO(ptr(x4), ptr(K4), ptr(R), ptr(b), ptr(h0), ptr(c0), hptr,
                                     ptr(gates), ptr(cseq), None, None, None)
                g = _lib.ConvLstmGrads(dcat.data_ptr() + 4 * off, ptr(dhT), ptr(dcT), ptr(dx4), ptr(dh0), ptr(dc0),
                                       ptr(gk4), ptr(gr_), ptr(gb), ptr(ws), 0, None, None, None)
                _lib.check(lib.fov_convlstm_bwd(C.byref(cfg), C.byref(io), C.byref(g), st), "fov_convlstm_bwd")
                _lib.check(lib.fov_gate_kernel_reduce(K.shape[0] * K.shape[1], cin_, F, ptr(gk4), ptr(gk), st),
                           "fov_gate_kernel_reduce")
                if dxp:
                    _lib.check(lib.fov_dropout_reduce(cfg.B, cfg.T, cfg.H * cfg.W, cin_, ptr(dx4), ptr(mask), dxp,
                                                      xb_, xt_, xpix_, acc, st), "fov_dropout_reduce")
            dstate_out[2 * l], dstate_out[2 * l + 1] = dh0, dc0
        if wave_g:
            for s_ in wstreams[1:]:
                main_s.wait_stream(s_)
        del wave_keep
        return (None, None, dx0) + (None,) * (3 * L) + tuple(dstate_out)


def convlstm_stack(x, weights, states=None, sinks=None, dilation=(1, 1), rec_act="hard_sigmoid",
                   training=False, dropout_masks=None, pack_cache=None):
    """weights: [(K,R,b)]*L ; states: [(h0,c0)]*L or None; dropout_masks: per layer None or a
    (4,B,H,W,Cin_l) tensor of keep-masks already scaled by 1/(1-rate) (training only).
    Returns (concat_seq, [(hT,cT)]*L)."""
    L = len(weights)
    flat = []
    for w in weights:
        flat += list(w)
    for l in range(L):
        flat += list(states[l]) if states is not None else [None, None]
    out = ConvLSTMStackFn.apply({"layers": L, "dilation": dilation, "rec_act": rec_act,
                                 "training": training, "dropout_masks": dropout_masks, "pack_cache": pack_cache},
                                sinks, x, *flat)
    return out[0], [(out[1 + 2 * l], out[2 + 2 * l]) for l in range(L)]


# --------------------------------------------------------------------------- #
# softmax / losses
# --------------------------------------------------------------------------- #


class SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        _require_cuda(x)
        x = _f32c(x)
        y = torch.empty_like(x)
        C_ = x.shape[-1]
        _lib.check(lib.fov_softmax_fwd(x.numel() // C_, C_, ptr(x), ptr(y), _stream()), "fov_softmax_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(y)
        C_ = y.shape[-1]
        _lib.check(lib.fov_softmax_bwd(y.numel() // C_, C_, ptr(y), ptr(dy), ptr(dx), _stream()),
                   "fov_softmax_bwd")
        return dx


_UNIT_LOSS_GRAD = [False]


def set_unit_loss_grad(on):
    """The caller promises that the loss tensors are back-propagated with gradient exactly 1 (Model's own train step:
    ``total = sum(losses); total.backward()``): LossFn.backward then returns the fused gradient as it is instead of
    multiplying the whole (B,T,out) tensor by the incoming scalar (140 MB for config 2's reconstruction head at
    B = 8880, ~0.1 ms per step for a multiplication by 1.0).  Off by default: any other use of the loss tensor
    (scaling, averaging over micro-batches, a weighted data-parallel seed) keeps the multiply."""
    _UNIT_LOSS_GRAD[0] = bool(on)


class LossFn(torch.autograd.Function):
    """kind in {'mse','nll','cce'}: returns a 1-element loss tensor; gradient fused."""

    @staticmethod
    def forward(ctx, kind, weight, y_pred, y_true, running_length):
        lib = _lib.load()
        _require_cuda(y_pred, y_true)
        y_pred, y_true = _f32c(y_pred), _f32c(y_true)
        loss = torch.zeros(1, device=y_pred.device)
        need = ctx.needs_input_grad[2]
        dy = torch.empty_like(y_pred) if need else None
        st = _stream()
        if kind == "mse":
            _lib.check(lib.fov_mse_fwd_bwd(y_pred.numel(), ptr(y_pred), ptr(y_true), weight, ptr(loss),
                                           ptr(dy), st), "fov_mse_fwd_bwd")
        elif kind == "nll":
            B, T = y_pred.shape[0], y_pred.shape[1]
            _lib.check(lib.fov_gauss_nll_fwd_bwd(B, T, running_length, ptr(y_pred), ptr(y_true), weight,
                                                 ptr(loss), ptr(dy), st), "fov_gauss_nll_fwd_bwd")
        elif kind == "cce":
            C_ = y_pred.shape[-1]
            _lib.check(lib.fov_cce_fwd_bwd(y_pred.numel() // C_, C_, ptr(y_pred), ptr(y_true), weight,
                                           ptr(loss), ptr(dy), st), "fov_cce_fwd_bwd")
        else:
            raise ValueError(kind)
        ctx.dy = dy
        return loss

    @staticmethod
    def backward(ctx, dl):
        # the loss weight is already folded into dy by the kernel; dl is whatever the caller did with the loss tensor
        # afterwards (1 for the plain sum of Keras compile(loss=[...]), the shard weight of a data-parallel step,
        # a micro-batch average, ...): one scalar multiply of the small (B,T,out) gradient.
        dy = ctx.dy
        if dy is not None and not _UNIT_LOSS_GRAD[0]:
            dy = dy * dl.reshape(())
        return None, None, dy, None, None


def loss(kind, y_true, y_pred, weight=1.0, running_length=10):
    return LossFn.apply(kind, float(weight), y_pred, y_true, int(running_length))


# --------------------------------------------------------------------------- #
# optimiser steps, featuriser, re-sampler (no autograd)
# --------------------------------------------------------------------------- #


def adam_step(p, g, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, grad_scale=1.0, grad_div=None):
    """grad_div: optional 1-element device tensor; g is divided by it inside the kernel (no host sync)."""
    lib = _lib.load()
    _require_cuda(p, g, m, v)
    _lib.check(lib.fov_adam_step(p.numel(), ptr(p), ptr(g), ptr(m), ptr(v), int(t), lr, beta1, beta2, eps,
                                 grad_scale, ptr(grad_div), _stream()), "fov_adam_step")


def rmsprop_step(p, g, a, lr=1e-3, rho=0.9, eps=1e-7, grad_scale=1.0, grad_div=None):
    lib = _lib.load()
    _require_cuda(p, g, a)
    _lib.check(lib.fov_rmsprop_step(p.numel(), ptr(p), ptr(g), ptr(a), lr, rho, eps, grad_scale, ptr(grad_div),
                                    _stream()), "fov_rmsprop_step")


def mean_var_xyz(frames):
    """get_gt_target_xyz on the device: (..., 90) interleaved xyz or (..., 30, 3) -> (..., 6)."""
    lib = _lib.load()
    _require_cuda(frames)
    frames = _f32c(frames)
    lead = frames.shape[:-1] if frames.shape[-1] == 90 else frames.shape[:-2]
    rows = 1
    for d in lead:
        rows *= d
    out = torch.empty(*lead, 6, device=frames.device)
    _lib.check(lib.fov_mean_var_xyz(rows, ptr(frames), ptr(out), _stream()), "fov_mean_var_xyz")
    return out


RESAMPLE_MODES = {"sqrt_floor": 0, "sqrt": 1, "var_as_std": 2}


class GaussResampleFn(torch.autograd.Function):
    """frames = mu + sd(var) * noise, differentiable in (mu, var) like K.random_normal(mean=mu, stddev=...)
    (mycode/convlstm_seq2seq.py:51-60,259-272; mycode/others_LSTM_span_whole.py:64-71,302-315)."""

    @staticmethod
    def forward(ctx, mode, muvar, noise):
        lib = _lib.load()
        _require_cuda(muvar, noise)
        muvar, noise = _f32c(muvar), _f32c(noise)
        rows = muvar.numel() // 6
        _expect(muvar.shape[-1] == 6 and noise.numel() == rows * 90,
                "gauss_resample: muvar (rows,6) and noise (rows,30,3) disagree: %s vs %s", tuple(muvar.shape),
                tuple(noise.shape))
        out = torch.empty_like(noise)
        _lib.check(lib.fov_gauss_resample(rows, mode, ptr(muvar), ptr(noise), ptr(out), _stream()),
                   "fov_gauss_resample")
        ctx.mode, ctx.rows = mode, rows
        ctx.save_for_backward(muvar, noise)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        muvar, noise = ctx.saved_tensors
        dmuvar = torch.empty_like(muvar)
        _lib.check(lib.fov_gauss_resample_bwd(ctx.rows, ctx.mode, ptr(muvar), ptr(noise), ptr(_f32c(dout)), ptr(dmuvar),
                                              _stream()), "fov_gauss_resample_bwd")
        return None, dmuvar, None


def gauss_resample(muvar, noise, mode="sqrt_floor"):
    """(rows,6) mean/var + (rows,30,3) N(0,1) noise -> (rows,30,3) frames; gradients flow to muvar."""
    return GaussResampleFn.apply(RESAMPLE_MODES[mode], muvar, noise)


def philox_normal(shape, seed, offset=0, device=None, return_words=False):
    """N(0,1) draws from the in-kernel Philox4x32-10 stream (fov_philox_normal): element i is word i%4 of
    philox(counter = offset + i//4, key = seed), Box-Muller on word pairs.  Depends on (seed, offset) only."""
    lib = _lib.load()
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = 1
    for d in shape:
        n *= int(d)
    out = torch.empty(shape, device=device)
    words = torch.empty(n, dtype=torch.int32, device=device) if return_words else None
    _lib.check(lib.fov_philox_normal(n, int(seed) & (2 ** 64 - 1), int(offset), ptr(words), ptr(out), _stream()),
               "fov_philox_normal")
    return (out, words) if return_words else out


# --------------------------------------------------------------------------- #
# sample builders: windows, whole spans, one-hot heatmaps on the device
# --------------------------------------------------------------------------- #


def reshape2second_stacks(per_video_db, collapse_user=False, stride=10, running_length=10, purelly_testing=False):
    """mycode/utility.py:264-305 on the device: (U,S,C) seconds of one video -> (past, future, future_input)
    windows, shapes (n*U, L, C) when ``collapse_user`` else (U, n, L, C)."""
    lib = _lib.load()
    _require_cuda(per_video_db)
    src = _f32c(per_video_db)
    U, S, Cc = src.shape
    n = lib.fov_window_count(S, running_length, stride, int(purelly_testing))
    if n <= 0:
        raise _lib.FovError("reshape2second_stacks: %d seconds give no complete window of 2 x %d" % (S, running_length))
    shape = (n * U, running_length, Cc) if collapse_user else (U, n, running_length, Cc)
    past, fut, fut_in = (torch.empty(shape, device=src.device) for _ in range(3))
    _lib.check(lib.fov_window_stacks(U, S, Cc, running_length, stride, int(purelly_testing), int(collapse_user),
                                     ptr(src), ptr(past), ptr(fut), ptr(fut_in), _stream()), "fov_window_stacks")
    return past, fut, fut_in


def get_whole_span(x):
    """mycode/others_LSTM_span_whole.py:403-419 on the device: (N, L, ...) -> (N, 2L, ...), row i = [x[i]; x[i+1]],
    last row zero (unshuffled data only, as the reference asserts)."""
    lib = _lib.load()
    _require_cuda(x)
    x = _f32c(x)
    N = x.shape[0]
    half = x[0].numel()
    out = torch.empty((N, 2 * x.shape[1]) + tuple(x.shape[2:]), device=x.device)
    _lib.check(lib.fov_whole_span(N, half, ptr(x), ptr(out), _stream()), "fov_whole_span")
    return out


def get_data(datadb, pick_user=False, num_user=48, draw=None, running_length=10, stride=10, fps=30, device=None):
    """mycode/utility.py:359-446 on the device (cfg.cut_data_head / cfg.time_shift at their defaults).  ``datadb``:
    {video: {'x','y','z': (viewers, frames) array or tensor}}.  Each video is windowed ONCE for all its viewers
    (fov_window_stacks); the target / others split with duplicate padding or truncation to num_user - 1 is a gather
    over the viewer axis of those windows (fov_pick_user_gather), written straight into the final tensors.
    ``draw(n)`` supplies the duplicate indices (default np.random.randint, as the reference draws them).
    Returns (past, future, future_input) (N,10,90), plus the others' three (num_user-1, N, 10, 90) when pick_user."""
    import numpy as np
    lib = _lib.load()
    device = device or torch.device("cuda", torch.cuda.current_device())
    draw = draw or (lambda n: int(np.random.randint(n)))
    K = num_user - 1
    vids = []
    for vid in datadb.keys():
        d = datadb[vid]
        xyz = torch.stack([torch.as_tensor(np.asarray(d[c]) if not isinstance(d[c], torch.Tensor) else d[c])
                           .to(device=device, dtype=torch.float32) for c in "xyz"], dim=-1)      # (U, frames, 3)
        U, S = xyz.shape[0], xyz.shape[1] // fps
        if S < 2 * running_length:
            continue                                                # 'video only has %d seconds. skip...'
        per = xyz[:, :S * fps].reshape(U, S, 3 * fps)
        vids.append((per, U))
    if not vids:
        raise _lib.FovError("get_data: no video has %d seconds" % (2 * running_length))
    if not pick_user:
        parts = [reshape2second_stacks(per, True, stride, running_length) for per, _ in vids]
        return tuple(torch.cat([p[k] for p in parts], dim=0) for k in range(3))
    wins = [reshape2second_stacks(per, False, stride, running_length) for per, _ in vids]     # (U, n, L, C) x 3
    n_total = sum(w[0].shape[0] * w[0].shape[1] for w in wins)
    L, Cc = wins[0][0].shape[2], wins[0][0].shape[3]
    row = L * Cc
    tar = [torch.cat([w[k].reshape(-1, L, Cc) for w in wins], dim=0) for k in range(3)]       # target t = viewer t, in order
    oth = [torch.empty(K, n_total, L, Cc, device=device) for _ in range(3)]
    base = 0
    for (per, U), w in zip(vids, wins):
        n = w[0].shape[1]
        idx = []
        for t in range(U):
            lst = [u for u in range(U) if u != t]
            while len(lst) < K:
                lst.append(lst[draw(len(lst))])
            idx += lst[:K]
        idx_t = torch.tensor(idx, dtype=torch.int32, device=device)
        for k in range(3):
            _lib.check(lib.fov_pick_user_gather(U, K, n, row, ptr(idx_t), ptr(w[k]), ptr(oth[k]), n_total * row, base,
                                                _stream()), "fov_pick_user_gather")
        base += U * n
    return tuple(tar) + tuple(oth)


def others_index(n_viewers, num_user, draw=None):
    """(U, num_user-1) int32 table: the others of every target viewer (mycode/utility.py:404-416): the video's other
    viewers in order, padded with duplicates drawn from the current list or truncated."""
    import numpy as np
    draw = draw or (lambda n: int(np.random.randint(n)))
    K = num_user - 1
    rows = []
    for t in range(n_viewers):
        lst = [u for u in range(n_viewers) if u != t]
        while len(lst) < K:
            lst.append(lst[draw(len(lst))])
        rows.append(lst[:K])
    return np.asarray(rows, dtype=np.int32)


def m3_batches_from_video(frames, num_user, stride=10, limit=None, idx=None, draw=None):
    """Training batches of others_lstm_span_whole from ONE raw video on the device: ``frames`` (U, S, 90) unit-sphere
    xyz per second (30 frames interleaved) -> per-second mean/var (fov_mean_var_xyz) -> every (target viewer, window)
    sequence (fov_m3_batches).  Returns (inputs [enc (B,10,6), oth (B,20,1,K,6), dec0 (B,1,6)], targets [fut (B,10,6),
    oth viewed (B,20,K*6), enc]) - the two reconstruction targets alias their inputs, nothing is copied.
    ``limit``: keep the first ``limit`` sequences."""
    lib = _lib.load()
    _require_cuda(frames)
    frames = _f32c(frames)
    U, S = frames.shape[0], frames.shape[1]
    K = num_user - 1
    mv = mean_var_xyz(frames)                                     # (U,S,6)
    n = (S - 20) // stride + 1
    if n < 1:
        raise _lib.FovError("m3_batches_from_video: %d seconds give no 20-second window" % S)
    B = U * n if limit is None else min(int(limit), U * n)
    if idx is None:
        idx = torch.as_tensor(others_index(U, num_user, draw), device=frames.device)
    dev = frames.device
    enc = torch.empty(B, 10, 6, device=dev)
    oth = torch.empty(B, 20, 1, K, 6, device=dev)
    dec0 = torch.empty(B, 1, 6, device=dev)
    fut = torch.empty(B, 10, 6, device=dev)
    _lib.check(lib.fov_m3_batches(U, S, stride, K, ptr(idx), ptr(mv), B, ptr(enc), ptr(oth), ptr(dec0), ptr(fut),
                                  _stream()), "fov_m3_batches")
    return [enc, oth, dec0], [fut, oth.view(B, 20, K * 6), enc]


def one_hot_heatmaps(frames, bin_size=10):
    """(..., F, 3) xyz frames -> (..., 360/bin, 180/bin, F) one-hot FoV-centre maps, frames as channels
    (mycode/dataIO.py:77-82, mycode/utility.py:533-556, mycode/data_generator_for_heatmap.py:32,65-67)."""
    lib = _lib.load()
    _require_cuda(frames)
    frames = _f32c(frames)
    lead, Fr = frames.shape[:-2], frames.shape[-2]
    rows = 1
    for d in lead:
        rows *= d
    out = torch.empty(*lead, 360 // bin_size, 180 // bin_size, Fr, device=frames.device)
    _lib.check(lib.fov_onehot_heatmaps(rows, Fr, bin_size, ptr(frames), ptr(out), _stream()), "fov_onehot_heatmaps")
    return out


def theta_phi_frames(frames):
    """(..., F, 3) xyz frames -> (..., F, 2) float64 [phi / pi, (theta + pi) / 2 / pi], the frame-level label array of
    mycode/data_generator_gaussian_FoV.py:21-55."""
    lib = _lib.load()
    _require_cuda(frames)
    frames = _f32c(frames)
    out = torch.empty(*frames.shape[:-1], 2, device=frames.device, dtype=torch.float64)
    _lib.check(lib.fov_theta_phi_frames(out.numel() // 2, ptr(frames), ptr(out), _stream()), "fov_theta_phi_frames")
    return out


def gaussian_fov_tiles(phi_theta, kind="fov", fps=30, return_peak=False):
    """Gaussian-FoV (``kind='fov'``: get_gaussianFoV_per_vid_per_target_giventhetaphi,
    mycode/data_generator_gaussian_FoV.py:57-138) or head-direction (``kind='head'``: :163-243) tiles:
    phi_theta (num_user, num_sec * fps, 2) float64 centres in [0, 1] -> (num_user, num_sec, 18, 36, fps) float32,
    normalised by the maximum over every full-resolution frame of the call like the reference."""
    lib = _lib.load()
    _require_cuda(phi_theta)
    _expect(phi_theta.dim() == 3 and phi_theta.shape[-1] == 2, "phi_theta must be (num_user, frames, 2), got %s",
            tuple(phi_theta.shape))
    _expect(kind in ("fov", "head"), "kind must be 'fov' or 'head', got %r", kind)
    U, n = phi_theta.shape[0], phi_theta.shape[1] // fps
    _expect(n > 0, "need at least one whole second of frames")
    pt = phi_theta[:, :n * fps].to(torch.float64).contiguous()
    out = torch.empty(U, n, 18, 36, fps, device=pt.device)
    peak = torch.empty(1, device=pt.device)
    _lib.check(lib.fov_gaussian_fov_tiles(U * n, fps, 0 if kind == "fov" else 1, ptr(pt), ptr(out), ptr(peak),
                                          _stream()), "fov_gaussian_fov_tiles")
    return (out, peak) if return_peak else out


def heatmap_sum(tiles):
    """heatmap_sum + normalize_to_distribution (mycode/data_generator_gaussian_FoV.py:246-261):
    (..., 18, 36, F) -> (..., 18, 36, 1), frame channels summed, each map scaled to sum 1."""
    lib = _lib.load()
    _require_cuda(tiles)
    tiles = _f32c(tiles)
    _expect(tiles.dim() >= 3, "tiles must be (..., H, W, F)")
    H, W, Fr = tiles.shape[-3:]
    maps = tiles.numel() // (H * W * Fr)
    out = torch.empty(*tiles.shape[:-1], 1, device=tiles.device)
    _lib.check(lib.fov_heatmap_sum(maps, H * W, Fr, ptr(tiles), ptr(out), _stream()), "fov_heatmap_sum")
    return out


def hit_rate(pred_theta_phi, gt_theta_phi, a=1.0, span_deg=120.0):
    """FoV hit rate per (theta, phi) centre pair (mycode/baseline_knn_mean.py:48-93,123-168): (..., 2) radians ->
    (...,); the predicted box is ``a`` x 120 degrees wide, the ground-truth box 120 degrees."""
    import math
    lib = _lib.load()
    _require_cuda(pred_theta_phi, gt_theta_phi)
    p, g = _f32c(pred_theta_phi), _f32c(gt_theta_phi)
    out = torch.empty(p.shape[:-1], device=p.device)
    sp = a * span_deg / 180.0 * math.pi
    gs = span_deg / 180.0 * math.pi
    _lib.check(lib.fov_hit_rate(out.numel(), ptr(p), ptr(g), sp, sp, gs, gs, ptr(out), _stream()), "fov_hit_rate")
    return out
