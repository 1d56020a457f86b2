"""Keras-default weight initialisers for the four model builders (host NumPy):
glorot_uniform kernels, orthogonal recurrent kernels, zero biases with the forget
block set to one (unit_forget_bias=True).  Weight names/layouts are Keras'."""
from __future__ import annotations

import numpy as np


def glorot_uniform(rng, shape):
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def orthogonal(rng, shape):
    rows, cols = int(np.prod(shape[:-1])), shape[-1]
    a = rng.normal(0.0, 1.0, (rows, cols))
    u, _, vt = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == (rows, cols) else vt
    return q.reshape(shape).astype(np.float32)


def _lstm_bias(units):
    b = np.zeros(4 * units, np.float32)
    b[units:2 * units] = 1.0
    return b


def _lstm(rng, in_dim, units, prefix, w):
    w[prefix + "/kernel"] = glorot_uniform(rng, (in_dim, 4 * units))
    w[prefix + "/recurrent_kernel"] = orthogonal(rng, (units, 4 * units))
    w[prefix + "/bias"] = _lstm_bias(units)


def _convlstm(rng, kh, kw, cin, f, prefix, w):
    w[prefix + "/kernel"] = glorot_uniform(rng, (kh, kw, cin, 4 * f))
    w[prefix + "/recurrent_kernel"] = orthogonal(rng, (kh, kw, f, 4 * f))
    w[prefix + "/bias"] = _lstm_bias(f)


def _dense(rng, i, o, prefix, w):
    w[prefix + "/kernel"] = glorot_uniform(rng, (i, o))
    w[prefix + "/bias"] = np.zeros(o, np.float32)


def _conv(rng, kshape, prefix, w):
    w[prefix + "/kernel"] = glorot_uniform(rng, kshape)
    w[prefix + "/bias"] = np.zeros(kshape[-1], np.float32)


def init_fov_seq2seq(seed=1, num_encoder_tokens=90, num_decoder_tokens=6, latent_dim=64):
    rng = np.random.default_rng(seed)
    w = {}
    _lstm(rng, num_encoder_tokens, latent_dim, "encoder", w)
    _lstm(rng, num_decoder_tokens, latent_dim, "decoder", w)
    _dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
    return w


def init_others_lstm_span_whole(seed=1, num_user=34, kernel_size=5, latent_dim=64, oth_filters=(32, 16, 8),
                                flat_dense=256):
    rng = np.random.default_rng(seed)
    w = {}
    cin = 6
    for l, f in enumerate(oth_filters):
        _convlstm(rng, 1, kernel_size, cin, f, "oth_convlstm%d" % l, w)
        cin = f
    flat = (num_user - 1) * sum(oth_filters)
    _dense(rng, flat, (num_user - 1) * 6, "oth_recon_dense", w)
    _dense(rng, flat, flat_dense, "oth_flat_dense", w)
    _lstm(rng, 6, latent_dim, "encoder", w)
    _lstm(rng, 6, latent_dim, "decoder", w)
    _dense(rng, latent_dim, 6, "encoder_dense", w)
    _dense(rng, latent_dim + flat_dense, 6, "decoder_dense", w)
    return w


def init_convlstm_seq2seq(seed=1, in_ch=30, filters=(32, 16, 8), kernel_size=5, head=(512, 1024, 30),
                          head_kind="conv2d", head_kernel=7, flat_dim=None):
    rng = np.random.default_rng(seed)
    w = {}
    for side in ("enc", "dec"):
        cin = in_ch
        for l, f in enumerate(filters):
            _convlstm(rng, kernel_size, kernel_size, cin, f, "%s_convlstm%d" % (side, l), w)
            cin = f
    cin = sum(filters)
    if head_kind == "conv2d":
        for l, f in enumerate(head):
            _conv(rng, (kernel_size, kernel_size, cin, f), "head_conv%d" % l, w)
            cin = f
    elif head_kind == "conv1d":
        for l, f in enumerate(head):
            _conv(rng, (head_kernel, cin, f), "head_conv%d" % l, w)
            cin = f
    else:
        _dense(rng, flat_dim, 6, "head_dense", w)
    return w


def init_stacked_fov_seq2seq(seed=1, n_layers=2, num_encoder_tokens=6, num_decoder_tokens=6, latent_dim=64):
    """2- / 3-layer target-only models (mycode/Fov_seq2seq_2layers.py:232-272, mycode/3layers.py:223-275): LSTMs of
    latent_dim // 2 units."""
    rng = np.random.default_rng(seed)
    units = latent_dim // 2
    w = {}
    for l in range(n_layers):
        _lstm(rng, num_encoder_tokens if l == 0 else units, units, "encoder%d" % l, w)
    for l in range(n_layers):
        _lstm(rng, num_decoder_tokens if l == 0 else units, units, "decoder%d" % l, w)
    _dense(rng, units, num_decoder_tokens, "decoder_dense", w)
    return w


def init_given_others_seq2seq(seed=1, num_user=34, latent_dim=32, num_encoder_tokens=6, num_decoder_tokens=6,
                              variant="mlp_mixing"):
    """mycode/given_others_gt_mean_var_seq2seq.py:97-166: two encoder + two decoder LSTMs of latent_dim units,
    Dense(6, tanh) and the variant's mixing layers."""
    rng = np.random.default_rng(seed)
    w = {}
    for l in range(2):
        _lstm(rng, num_encoder_tokens if l == 0 else latent_dim, latent_dim, "encoder%d" % l, w)
    for l in range(2):
        _lstm(rng, num_decoder_tokens if l == 0 else latent_dim, latent_dim, "decoder%d" % l, w)
    oth = (num_user - 1) * 6
    if variant == "others_lstm":
        for l in range(2):
            for d in ("fwd", "bwd"):
                _lstm(rng, oth if l == 0 else 2 * latent_dim, latent_dim, "others_bilstm%d_%s" % (l, d), w)
        _dense(rng, 3 * latent_dim, num_decoder_tokens, "decoder_dense", w)
    elif variant == "conv_mixing":
        _dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
        for l, (ci, co) in enumerate([(num_user, 8), (8, 8), (8, 1)]):
            _conv(rng, (1, 3, ci, co), "mixing_conv%d" % l, w)
    elif variant == "others_mlp":
        _dense(rng, oth, 256, "others_dense1", w)
        _dense(rng, 256, latent_dim, "others_dense2", w)
        _dense(rng, 2 * latent_dim, num_decoder_tokens, "decoder_dense", w)
    else:
        _dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
        if variant == "mlp_mixing":
            _dense(rng, oth + num_decoder_tokens, num_decoder_tokens, "mixing", w)
    return w


def init_others_convlstm_target(seed=1, num_user=34, kernel_size=5, oth_filters=(32, 16, 8), tar_filters=(8, 4, 2)):
    """All-ConvLSTM form of mycode/others_LSTM_span_whole.py (use_fclstm_tar=False, :133-199), raw xyz layout."""
    rng = np.random.default_rng(seed)
    w = {}
    cin = (num_user - 1) * 3
    for l, f in enumerate(oth_filters):
        _convlstm(rng, 1, kernel_size, cin, f, "oth_convlstm%d" % l, w)
        cin = f
    cin = 3
    for l, f in enumerate(tar_filters):
        _convlstm(rng, 1, kernel_size, cin, f, "tar_enc_convlstm%d" % l, w)
        cin = f
    cin = 3 + sum(oth_filters)
    for l, f in enumerate(tar_filters):
        _convlstm(rng, 1, kernel_size, cin, f, "tar_dec_convlstm%d" % l, w)
        cin = f
    _dense(rng, sum(oth_filters), (num_user - 1) * 3, "oth_recon_dense", w)
    _dense(rng, sum(tar_filters), 3, "encoder_dense", w)
    _dense(rng, sum(tar_filters), 3, "decoder_dense", w)
    return w
