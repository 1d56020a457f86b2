"""torch-CPU restatement of the same Keras 2.2.x graphs as ``keras_numpy`` —
the second, independent restatement (F.conv2d / F.linear / autograd instead of
NumPy loops).  It supplies gradients (autograd) for the BPTT parity tests and is
the timed CPU baseline of ``bench.py`` (a Keras-equivalent stand-in: Keras/TF1
cannot be installed here).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **parity unpinned** for
the Keras layer numerics; anchored by agreement with ``keras_numpy`` (1e-6),
``torch.nn.LSTM`` (sigmoid mode), known-answer cases and finite differences.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FPS = 30


def hard_sigmoid(x):
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def _rec_act(name):
    return {"hard_sigmoid": hard_sigmoid, "sigmoid": torch.sigmoid}[name]


def _act(name):
    if name in (None, "linear"):
        return lambda v: v
    if name == "tanh":
        return torch.tanh
    if name == "relu":
        return torch.relu
    if name == "softmax":
        return lambda v: torch.softmax(v, dim=-1)
    raise ValueError(name)


def to_torch(w, dtype=torch.float64, requires_grad=False):
    out = {}
    for k, v in w.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def dense(x, kernel, bias, activation=None):
    return _act(activation)(x @ kernel + bias)


def lstm_step(x, h, c, kernel, rk, bias, ra="hard_sigmoid"):
    H = rk.shape[0]
    a = _rec_act(ra)
    z = x @ kernel + bias + h @ rk
    i, f, g, o = a(z[:, :H]), a(z[:, H:2 * H]), torch.tanh(z[:, 2 * H:3 * H]), a(z[:, 3 * H:])
    c = f * c + i * g
    return o * torch.tanh(c), c


def lstm(x, kernel, rk, bias, h0=None, c0=None, ra="hard_sigmoid"):
    B, T, _ = x.shape
    H = rk.shape[0]
    h = x.new_zeros(B, H) if h0 is None else h0
    c = x.new_zeros(B, H) if c0 is None else c0
    seq = []
    for t in range(T):
        h, c = lstm_step(x[:, t], h, c, kernel, rk, bias, ra)
        seq.append(h)
    return torch.stack(seq, 1), h, c


def conv2d_same(x, kernel, bias=None, dilation=(1, 1)):
    """NHWC in/out, kernel (kh,kw,Cin,Cout); TF 'same' (asymmetric for even k)."""
    kh, kw = kernel.shape[:2]
    dh, dw = dilation
    th, tw = (kh - 1) * dh, (kw - 1) * dw
    xp = F.pad(x.permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2))
    y = F.conv2d(xp, kernel.permute(3, 2, 0, 1), bias, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def conv2d(x, kernel, bias, activation=None, dilation=(1, 1)):
    return _act(activation)(conv2d_same(x, kernel, bias, dilation))


def conv1d(x, kernel, bias, activation=None):
    return _act(activation)(conv2d_same(x[:, None], kernel[None], bias)[:, 0])


def convlstm2d_step(x, h, c, kernel, rk, bias, dilation=(1, 1), ra="hard_sigmoid",
                    dropout_masks=None):
    Fh = rk.shape[-1] // 4
    a = _rec_act(ra)
    if dropout_masks is None:
        zx = conv2d_same(x, kernel, bias, dilation)
    else:
        zx = torch.cat([conv2d_same(x * dropout_masks[g], kernel[..., g * Fh:(g + 1) * Fh],
                                    bias[g * Fh:(g + 1) * Fh], dilation) for g in range(4)], -1)
    z = zx + conv2d_same(h, rk)
    i, f = a(z[..., :Fh]), a(z[..., Fh:2 * Fh])
    g, o = torch.tanh(z[..., 2 * Fh:3 * Fh]), a(z[..., 3 * Fh:])
    c = f * c + i * g
    return o * torch.tanh(c), c


def convlstm2d(x, kernel, rk, bias, h0=None, c0=None, dilation=(1, 1), ra="hard_sigmoid",
               dropout_masks=None):
    B, T, Hh, Ww, _ = x.shape
    Fh = rk.shape[-1] // 4
    h = x.new_zeros(B, Hh, Ww, Fh) if h0 is None else h0
    c = x.new_zeros(B, Hh, Ww, Fh) if c0 is None else c0
    seq = []
    for t in range(T):
        h, c = convlstm2d_step(x[:, t], h, c, kernel, rk, bias, dilation, ra, dropout_masks)
        seq.append(h)
    return torch.stack(seq, 1), h, c


def convlstm_stack(w, x, prefix, h0c0=None, dilation=(1, 1), ra="hard_sigmoid", n_layers=3, dropout_masks=None):
    seqs, states = [], []
    cur = x
    for l in range(n_layers):
        p = "%s%d" % (prefix, l)
        h0, c0 = (None, None) if h0c0 is None else h0c0[l]
        cur, h, c = convlstm2d(cur, w[p + "/kernel"], w[p + "/recurrent_kernel"], w[p + "/bias"],
                               h0, c0, dilation, ra, None if dropout_masks is None else dropout_masks[l])
        seqs.append(cur)
        states.append((h, c))
    return torch.cat(seqs, -1), states


# ------------------------------- losses ----------------------------------- #


def mse(y_true, y_pred):
    return torch.mean((y_pred - y_true) ** 2)


def gauss_nll(y_true, y_pred, running_length=10, fps=FPS, eps=1e-7):
    total = 0.0
    for a in range(3):
        u = y_pred[:, :, a:a + 1]
        v = torch.clamp(torch.abs(y_pred[:, :, 3 + a:4 + a]), 1e-4, 2.0)
        x = y_true[:, :, a::3]
        l = torch.log(v + eps) + (x - u) ** 2 / (v + eps)
        total = total + torch.clamp(l, -2000.0, 2000.0)
    return torch.mean(total.sum(2).sum(1)) / running_length / fps


def categorical_crossentropy(y_true, y_pred, eps=1e-7):
    p = y_pred / y_pred.sum(-1, keepdim=True)
    p = torch.clamp(p, eps, 1.0 - eps)
    return torch.mean(-(y_true * torch.log(p)).sum(-1))


# ------------------------------- models ----------------------------------- #


def fov_seq2seq_forward(w, enc_in, dec_in, teacher_forcing=True, decoder_no_init_state=False,
                        ra="hard_sigmoid", steps=10):
    _, h, c = lstm(enc_in, w["encoder/kernel"], w["encoder/recurrent_kernel"], w["encoder/bias"],
                   ra=ra)
    if decoder_no_init_state:
        h, c = torch.zeros_like(h), torch.zeros_like(c)
    if teacher_forcing:
        seq, _, _ = lstm(dec_in, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], h, c, ra)
        return dense(seq, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
    x = dec_in[:, 0]
    outs = []
    for _ in range(steps):
        h, c = lstm_step(x, h, c, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], ra)
        y = dense(h, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        outs.append(y)
        x = y
    return torch.stack(outs, 1)


def stacked_fov_seq2seq_forward(w, enc_in, dec_in, n_layers=2, share_last_decoder=None, ra="hard_sigmoid"):
    """mycode/Fov_seq2seq_2layers.py:232-272, mycode/3layers.py:223-275 (see oracle/keras_numpy.py)."""
    if share_last_decoder is None:
        share_last_decoder = n_layers == 3
    xe, xd = enc_in, dec_in
    for l in range(n_layers):
        pe = "encoder%d" % l
        pd = "decoder%d" % (l - 1 if (share_last_decoder and l == n_layers - 1) else l)
        xe, h, c = lstm(xe, w[pe + "/kernel"], w[pe + "/recurrent_kernel"], w[pe + "/bias"], ra=ra)
        xd, _, _ = lstm(xd, w[pd + "/kernel"], w[pd + "/recurrent_kernel"], w[pd + "/bias"], h, c, ra=ra)
    return dense(xd, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")


def given_others_seq2seq_forward(w, enc_in, oth_in, dec_in, variant="mlp_mixing", teacher_forcing=False,
                                 ra="hard_sigmoid"):
    """mycode/given_others_gt_mean_var_seq2seq.py:97-308 (see oracle/keras_numpy.py)."""
    B, T = oth_in.shape[0], oth_in.shape[1]
    L = lambda n: (w[n + "/kernel"], w[n + "/recurrent_kernel"], w[n + "/bias"])
    e1, h1, c1 = lstm(enc_in, *L("encoder0"), ra=ra)
    _, h2, c2 = lstm(e1, *L("encoder1"), ra=ra)
    if variant == "others_lstm":
        y, st = oth_in.reshape(B, T, -1), {"fwd": (None, None), "bwd": (None, None)}
        for l in range(2):
            f, fh, fc = lstm(y, *L("others_bilstm%d_fwd" % l), *st["fwd"], ra=ra)
            b, bh, bc = lstm(torch.flip(y, dims=[1]), *L("others_bilstm%d_bwd" % l), *st["bwd"], ra=ra)
            y, st = torch.cat([f, torch.flip(b, dims=[1])], dim=-1), {"fwd": (fh, fc), "bwd": (bh, bc)}
        oth_seq = y
    if teacher_forcing:
        d1, _, _ = lstm(dec_in, *L("decoder0"), h1, c1, ra=ra)
        d2, _, _ = lstm(d1, *L("decoder1"), h2, c2, ra=ra)
    x = dec_in[:, 0]
    outs = []
    for t in range(T):
        if teacher_forcing:
            s2 = d2[:, t]
        else:
            h1, c1 = lstm_step(x, h1, c1, *L("decoder0"), ra=ra)
            h2, c2 = lstm_step(h1, h2, c2, *L("decoder1"), ra=ra)
            s2 = h2
        flat = oth_in[:, t].reshape(B, -1)
        if variant == "target_only":
            y = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        elif variant == "others_lstm":
            s = d2[:, 0] if teacher_forcing else s2
            y = dense(torch.cat([oth_seq[:, t], s], dim=-1), w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        elif variant == "conv_mixing":
            pred = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
            img = torch.cat([oth_in[:, t], pred[:, None]], dim=1).permute(0, 2, 1)[:, None]
            for l in range(3):
                img = conv2d(img, w["mixing_conv%d/kernel" % l], w["mixing_conv%d/bias" % l], "relu")
            y = img[:, 0, :, 0]
        elif variant == "others_mlp":
            o = dense(flat, w["others_dense1/kernel"], w["others_dense1/bias"], "relu")
            o = dense(o, w["others_dense2/kernel"], w["others_dense2/bias"], "relu")
            y = dense(torch.cat([o, s2], dim=-1), w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        else:
            pred = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
            y = dense(torch.cat([flat, pred], dim=-1), w["mixing/kernel"], w["mixing/bias"], "tanh")
        outs.append(y)
        x = y
    return torch.stack(outs, dim=1)


def others_lstm_span_whole_forward(w, enc_in, oth_in, dec_in, ra="hard_sigmoid"):
    B, Tenc = enc_in.shape[:2]
    Tall = oth_in.shape[1]
    oth_seq, _ = convlstm_stack(w, oth_in, "oth_convlstm", ra=ra)
    flat = oth_seq.reshape(B, Tall, -1)
    r_oth = dense(flat, w["oth_recon_dense/kernel"], w["oth_recon_dense/bias"])
    enc_seq, h, c = lstm(enc_in, w["encoder/kernel"], w["encoder/recurrent_kernel"],
                         w["encoder/bias"], ra=ra)
    r_tar = dense(enc_seq, w["encoder_dense/kernel"], w["encoder_dense/bias"], "tanh")
    s_all = dense(flat[:, Tenc:], w["oth_flat_dense/kernel"], w["oth_flat_dense/bias"])
    x = dec_in[:, 0]
    outs = []
    for t in range(Tall - Tenc):
        h, c = lstm_step(x, h, c, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], ra)
        y = dense(torch.cat([h, s_all[:, t]], 1), w["decoder_dense/kernel"],
                  w["decoder_dense/bias"])
        outs.append(y)
        x = y
    return [torch.stack(outs, 1), r_oth, r_tar]


def others_convlstm_target_forward(w, enc_in, oth_in, dec_in, ra="hard_sigmoid"):
    """All-ConvLSTM form of mycode/others_LSTM_span_whole.py (see oracle/keras_numpy.py)."""
    Tenc = enc_in.shape[1]
    Tdec = oth_in.shape[1] - Tenc
    oth_seq, _ = convlstm_stack(w, oth_in, "oth_convlstm", ra=ra)
    r_oth = dense(oth_seq, w["oth_recon_dense/kernel"], w["oth_recon_dense/bias"])
    pst, states = convlstm_stack(w, enc_in, "tar_enc_convlstm", ra=ra)
    r_tar = dense(pst, w["encoder_dense/kernel"], w["encoder_dense/bias"], "tanh")
    x = dec_in
    outs = []
    for t in range(Tdec):
        cat_in = torch.cat([x, oth_seq[:, Tenc + t:Tenc + t + 1]], dim=-1)
        dstate, states = convlstm_stack(w, cat_in, "tar_dec_convlstm", h0c0=states, ra=ra)
        y = dense(dstate, w["decoder_dense/kernel"], w["decoder_dense/bias"])
        outs.append(y)
        x = y
    return [torch.cat(outs, dim=1), r_oth, r_tar]


def gaussian_resample(muvar, noise, mode="var_as_std"):
    """(B,6) [mu | var] + (B,30,3) N(0,1) noise -> (B,30,3); differentiable in muvar like K.random_normal(mean, stddev)
    (mycode/convlstm_seq2seq.py:51-60; others_LSTM_span_whole.py:64-71; utility.py:73-80)."""
    mu, var = muvar[:, :3], muvar[:, 3:]
    if mode == "sqrt_floor":
        std = torch.sqrt(torch.where(var < 0, torch.full_like(var, 1e-3), var))
    elif mode == "sqrt":
        std = torch.sqrt(var)
    elif mode == "var_as_std":
        std = var
    else:
        raise ValueError(mode)
    return mu[:, None, :] + std[:, None, :] * noise


def convlstm_seq2seq_forward(w, enc_in, dec_in, head_kind="conv2d", steps=10, dilation=(1, 1),
                             ra="hard_sigmoid", noise=None, resample_mode="var_as_std"):
    """noise (steps,B,30,3): cfg.sample_and_refeed - the dense head's (mu,var) is re-sampled into 30 frames that
    become the next decoder input (mycode/convlstm_seq2seq.py:259-272)."""
    B = enc_in.shape[0]
    _, states = convlstm_stack(w, enc_in, "enc_convlstm", dilation=dilation, ra=ra)
    x = dec_in[:, 0]
    outs = []
    for _ in range(steps):
        hs, new_states, cur = [], [], x
        for l in range(3):
            p = "dec_convlstm%d" % l
            h, c = convlstm2d_step(cur, states[l][0], states[l][1], w[p + "/kernel"],
                                   w[p + "/recurrent_kernel"], w[p + "/bias"], dilation, ra)
            new_states.append((h, c))
            hs.append(h)
            cur = h
        states = new_states
        d = torch.cat(hs, -1)
        if head_kind == "conv2d":
            y = conv2d(d, w["head_conv0/kernel"], w["head_conv0/bias"], "relu")
            y = conv2d(y, w["head_conv1/kernel"], w["head_conv1/bias"], "relu")
            y = conv2d(y, w["head_conv2/kernel"], w["head_conv2/bias"], "relu")
            y = torch.softmax(y, -1)
            x = y
        elif head_kind == "conv1d":
            y = conv1d(d[:, 0], w["head_conv0/kernel"], w["head_conv0/bias"], "relu")
            y = conv1d(y, w["head_conv1/kernel"], w["head_conv1/bias"], "relu")
            y = conv1d(y, w["head_conv2/kernel"], w["head_conv2/bias"], "softmax")
            y = y[:, None]
            x = y
        else:
            y = dense(d[:, 0].reshape(B, -1), w["head_dense/kernel"], w["head_dense/bias"])
            if noise is not None:
                x = gaussian_resample(y, noise[len(outs)], resample_mode)[:, None]     # (B,1,30,3)
            else:
                x = y[:, None, None, :]
        outs.append(y)
    return torch.stack(outs, 1)


# ------------------------- loss + gradient helpers ------------------------- #


def loss_and_grads(forward, w, inputs, targets, loss_fns, loss_weights=None):
    """Run ``forward(w, *inputs)``, total loss = sum_i weight_i * loss_i(target_i, out_i)
    (Keras multi-output compile, mycode/others_LSTM_span_whole.py:352-353); return
    (loss, outputs, {name: grad})."""
    for t in w.values():
        t.requires_grad_(True)
        t.grad = None
    outs = forward(w, *inputs)
    if not isinstance(outs, (list, tuple)):
        outs = [outs]
    if loss_weights is None:
        loss_weights = [1.0] * len(outs)
    total = 0.0
    for o, y, fn, lw in zip(outs, targets, loss_fns, loss_weights):
        total = total + lw * fn(y, o)
    total.backward()
    # a weight the loss does not depend on (e.g. the encoder under decoder_no_init_state) has a zero gradient
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in w.items()}
    return total.detach(), [o.detach() for o in outs], grads


class KerasAdam:
    """Keras-form Adam over a dict of tensors (eps outside the bias correction)."""

    def __init__(self, w, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.w, self.lr, self.b1, self.b2, self.eps = w, lr, beta1, beta2, eps
        self.m = {k: torch.zeros_like(v) for k, v in w.items()}
        self.v = {k: torch.zeros_like(v) for k, v in w.items()}
        self.t = 0

    def step(self, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for k, p in self.w.items():
                g = grads[k]
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


class KerasRMSprop:
    def __init__(self, w, lr=1e-3, rho=0.9, eps=1e-7):
        self.w, self.lr, self.rho, self.eps = w, lr, rho, eps
        self.a = {k: torch.zeros_like(v) for k, v in w.items()}

    def step(self, grads):
        with torch.no_grad():
            for k, p in self.w.items():
                g = grads[k]
                self.a[k].mul_(self.rho).addcmul_(g, g, value=1 - self.rho)
                p.sub_(self.lr * g / (self.a[k].sqrt() + self.eps))
