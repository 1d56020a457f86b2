"""NumPy restatement of the Keras 2.2.x layer numerics used by the LongTerm360FoV
hot path, and of the four canonical model graphs (SURVEY.md section 8a').

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity status of this
file: **parity unpinned** for the Keras layer numerics (Keras/TF1 is not
installable here and the reference holds no golden vectors); the featuriser,
windowing, whole-span, one-hot heatmap and hit-rate functions are pinned against
the reference's own code via ``tests/golden/reference_numpy_golden.npz``.

Every function cites the reference call site it follows (paths relative to
``/root/reference``).  Weight layouts are Keras': LSTM ``kernel (in,4H)``,
``recurrent_kernel (H,4H)``, ``bias (4H)`` with gate column blocks i,f,c,o;
ConvLSTM2D ``kernel (kh,kw,Cin,4F)``, ``recurrent_kernel (kh,kw,F,4F)``; Dense
``kernel (in,out)``; Conv2D ``(kh,kw,Cin,Cout)``; Conv1D ``(k,Cin,Cout)``.
"""
from __future__ import annotations

import math

import numpy as np

FPS = 30

# --------------------------------------------------------------------------- #
# activations
# --------------------------------------------------------------------------- #


def hard_sigmoid(x):
    """Keras backend ``hard_sigmoid``: clip(0.2*x + 0.5, 0, 1).  Default
    ``recurrent_activation`` of LSTM / ConvLSTM2D in Keras 2.2.x (signature pasted
    at mycode/FOV_trj_pred_LSTM_concatState.py:110)."""
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _rec_act(name):
    return {"hard_sigmoid": hard_sigmoid, "sigmoid": sigmoid}[name]


def _act(name):
    if name in (None, "linear"):
        return lambda v: v
    if name == "tanh":
        return np.tanh
    if name == "relu":
        return lambda v: np.maximum(v, 0.0)
    if name == "softmax":
        return lambda v: softmax(v, axis=-1)
    raise ValueError(name)


def softmax(x, axis=-1):
    """keras.layers.Softmax(axis=-1) (mycode/convlstm_seq2seq.py:237)."""
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


# --------------------------------------------------------------------------- #
# Dense / LSTM
# --------------------------------------------------------------------------- #


def dense(x, kernel, bias, activation=None):
    """keras.layers.Dense on the last axis (mycode/FoV_seq2seq.py:96-97)."""
    return _act(activation)(x @ kernel + bias)


def lstm_step(x, h, c, kernel, recurrent_kernel, bias, recurrent_activation="hard_sigmoid"):
    """One Keras LSTMCell step (implementation=1 algebra): gate blocks i,f,c,o."""
    H = recurrent_kernel.shape[0]
    ra = _rec_act(recurrent_activation)
    z = x @ kernel + bias + h @ recurrent_kernel
    i = ra(z[:, 0 * H:1 * H])
    f = ra(z[:, 1 * H:2 * H])
    g = np.tanh(z[:, 2 * H:3 * H])
    o = ra(z[:, 3 * H:4 * H])
    c_new = f * c + i * g
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def lstm(x, kernel, recurrent_kernel, bias, h0=None, c0=None,
         recurrent_activation="hard_sigmoid"):
    """keras.layers.LSTM(H, return_sequences=True, return_state=True)
    (mycode/FoV_seq2seq.py:83-84,93-95).  x: (B,T,in).  Returns (seq, h, c)."""
    B, T, _ = x.shape
    H = recurrent_kernel.shape[0]
    h = np.zeros((B, H), x.dtype) if h0 is None else h0
    c = np.zeros((B, H), x.dtype) if c0 is None else c0
    seq = np.zeros((B, T, H), x.dtype)
    for t in range(T):
        h, c = lstm_step(x[:, t], h, c, kernel, recurrent_kernel, bias, recurrent_activation)
        seq[:, t] = h
    return seq, h, c


# --------------------------------------------------------------------------- #
# convolutions (channels-last, TF 'same' padding)
# --------------------------------------------------------------------------- #


def _same_pads(k, d):
    total = (k - 1) * d
    lo = total // 2
    return lo, total - lo


def conv2d_same(x, kernel, bias=None, dilation=(1, 1)):
    """TF/Keras conv2d, stride 1, padding='same', NHWC, kernel (kh,kw,Cin,Cout).
    It is a cross-correlation (no kernel flip), as in TF."""
    kh, kw, cin, cout = kernel.shape
    B, H, W, C = x.shape
    assert C == cin
    dh, dw = dilation
    pt, pb = _same_pads(kh, dh)
    pl, pr = _same_pads(kw, dw)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B, H, W, cout), np.result_type(x, kernel))
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i * dh:i * dh + H, j * dw:j * dw + W, :]
            out += patch @ kernel[i, j]
    if bias is not None:
        out = out + bias
    return out


def conv2d(x, kernel, bias, activation=None, dilation=(1, 1)):
    """keras.layers.Conv2D(padding='same') (mycode/convlstm_seq2seq.py:175-181)."""
    return _act(activation)(conv2d_same(x, kernel, bias, dilation))


def conv1d(x, kernel, bias, activation=None):
    """keras.layers.Conv1D(padding='same') (mycode/convlstm_seq2seq.py:184-189).
    x: (B,L,Cin), kernel (k,Cin,Cout)."""
    y = conv2d_same(x[:, None], kernel[None], bias)[:, 0]
    return _act(activation)(y)


def convlstm2d_step(x, h, c, kernel, recurrent_kernel, bias, dilation=(1, 1),
                    recurrent_activation="hard_sigmoid", dropout_masks=None):
    """One Keras ConvLSTM2DCell step.  Input conv: user padding 'same', given
    dilation, with bias; recurrent conv: always 'same', stride 1, no dilation, no
    bias.  ``dropout_masks`` (4 arrays, one per gate, already scaled by 1/(1-p)) is
    the training-time input dropout of mycode/others_LSTM_span_whole.py:89."""
    F = recurrent_kernel.shape[-1] // 4
    ra = _rec_act(recurrent_activation)
    if dropout_masks is None:
        zx = conv2d_same(x, kernel, bias, dilation)
    else:
        zx = np.concatenate([
            conv2d_same(x * dropout_masks[g], kernel[..., g * F:(g + 1) * F],
                        bias[g * F:(g + 1) * F], dilation) for g in range(4)], axis=-1)
    z = zx + conv2d_same(h, recurrent_kernel, None)
    i = ra(z[..., 0 * F:1 * F])
    f = ra(z[..., 1 * F:2 * F])
    g = np.tanh(z[..., 2 * F:3 * F])
    o = ra(z[..., 3 * F:4 * F])
    c_new = f * c + i * g
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def convlstm2d(x, kernel, recurrent_kernel, bias, h0=None, c0=None, dilation=(1, 1),
               recurrent_activation="hard_sigmoid", dropout_masks=None):
    """keras.layers.ConvLSTM2D(F, padding='same', return_sequences=True,
    return_state=True) (mycode/others_LSTM_span_whole.py:88-100,
    mycode/convlstm_seq2seq.py:100-126).  x: (B,T,H,W,Cin)."""
    B, T, Hh, Ww, _ = x.shape
    F = recurrent_kernel.shape[-1] // 4
    h = np.zeros((B, Hh, Ww, F), x.dtype) if h0 is None else h0
    c = np.zeros((B, Hh, Ww, F), x.dtype) if c0 is None else c0
    seq = np.zeros((B, T, Hh, Ww, F), x.dtype)
    for t in range(T):
        h, c = convlstm2d_step(x[:, t], h, c, kernel, recurrent_kernel, bias, dilation,
                               recurrent_activation, dropout_masks)
        seq[:, t] = h
    return seq, h, c


# --------------------------------------------------------------------------- #
# featuriser / re-sampler (pure NumPy in the reference)
# --------------------------------------------------------------------------- #


def get_gt_target_xyz(y):
    """mycode/utility.py:483-500: per-second mean and population variance of
    x,y,z.  (N,T,90) interleaved xyz or (N,T,30,3) -> (N,T,6)."""
    if y.shape[-1] == 3:
        assert y.ndim == 4
        tx, ty, tz = y[..., 0], y[..., 1], y[..., 2]
    else:
        assert y.shape[-1] == 3 * FPS and y.ndim == 3
        tx, ty, tz = y[:, :, 0::3], y[:, :, 1::3], y[:, :, 2::3]
    parts = [np.mean(tx, -1), np.mean(ty, -1), np.mean(tz, -1),
             np.var(tx, -1), np.var(ty, -1), np.var(tz, -1)]
    return np.stack(parts, axis=-1)


def get_gt_target_xyz_oth(y):
    """mycode/utility.py:505-517: (N,T,U,30,3) -> (N,T,U,6)."""
    tx, ty, tz = y[..., 0], y[..., 1], y[..., 2]
    parts = [np.mean(tx, -1), np.mean(ty, -1), np.mean(tz, -1),
             np.var(tx, -1), np.var(ty, -1), np.var(tz, -1)]
    return np.stack(parts, axis=-1)


def gaussian_resample(mu, var, noise, mode="sqrt_floor"):
    """Gaussian sample-and-refeed with the standard-normal draw made explicit.
    mu,var: (B,3); noise: (B,30,3) ~ N(0,1).  Returns (B,30,3) frames.
    mode 'sqrt_floor': mycode/utility.py:73-80 (var<0 -> 1e-3, std=sqrt(var));
    mode 'sqrt': mycode/others_LSTM_span_whole.py:64-69 (std = sqrt(var));
    mode 'var_as_std': mycode/convlstm_seq2seq.py:51-58 (std = var, as written)."""
    if mode == "sqrt_floor":
        v = np.where(var < 0, 1e-3, var)
        std = np.sqrt(v)
    elif mode == "sqrt":
        std = np.sqrt(var)
    elif mode == "var_as_std":
        std = var
    else:
        raise ValueError(mode)
    return mu[:, None, :] + std[:, None, :] * noise


def reshape2second_stacks(per_video_db, collapse_user=False, stride=10, running_length=10,
                          purelly_testing=False):
    """mycode/utility.py:264-305: sliding windows of ``running_length`` seconds with
    stride ``stride``; future = windows shifted by running_length//stride; decoder
    input = [last past second, future[:-1]]; ``purelly_testing`` appends
    running_length//stride zero seconds first (:274-277)."""
    L = running_length
    n_tok = per_video_db.shape[-1]
    assert per_video_db.shape[1] >= 2 * L
    shift = L // stride
    if purelly_testing:
        per_video_db = np.concatenate(
            (per_video_db, np.zeros((per_video_db.shape[0], shift, per_video_db.shape[2]))), axis=1)
    nrows = (per_video_db.shape[1] - L) // stride + 1
    idx = stride * np.arange(nrows)[:, None] + np.arange(L)
    win = per_video_db[:, idx, :].transpose(1, 0, 2, 3)        # (nrows, users, L, tok)
    fut = win[shift:]
    past = win[:-shift]
    last = past[:, :, -1:, :]
    fut_in = np.concatenate((last, fut[:, :, :-1, :]), axis=2)
    if collapse_user:
        return (past.reshape(-1, L, n_tok), fut.reshape(-1, L, n_tok),
                fut_in.reshape(-1, L, n_tok))
    return (past.transpose(1, 0, 2, 3), fut.transpose(1, 0, 2, 3),
            fut_in.transpose(1, 0, 2, 3))


def others_index_list(n_viewers, target, num_user, draw):
    """Viewer indices of the 'others' of one target (mycode/utility.py:404-416): the video's other viewers in order;
    fewer than num_user - 1 -> padded with duplicates drawn one at a time from the CURRENT (already padded) list,
    ``draw(len)`` standing for np.random.randint(len); more -> truncated."""
    lst = [u for u in range(n_viewers) if u != target]
    K = num_user - 1
    while len(lst) < K:
        lst.append(lst[draw(len(lst))])
    return lst[:K]


def get_data(datadb, pick_user=False, num_user=48, draw=None, running_length=10, stride=10, fps=FPS):
    """mycode/utility.py:359-446 (cfg.cut_data_head False, cfg.time_shift False).  datadb: {video: {'x','y','z':
    (viewers, frames)}}.  pick_user False: every (viewer, window) of every video, windows collapsed over viewers ->
    (past, future, future_input), each (N, 10, 90).  pick_user True: every viewer is the target once; returns the
    target's three tensors (N,10,90) and the others' three (num_user-1, N, 10, 90).  Videos shorter than 2 x
    running_length seconds are skipped.  ``draw(n)`` supplies the duplicate-padding indices (np.random.randint)."""
    draw = draw or (lambda n: int(np.random.randint(n)))
    tar, oth = [[], [], []], [[], [], []]
    for vid in datadb.keys():
        d = datadb[vid]
        per = np.stack((d["x"], d["y"], d["z"]), axis=-1)
        per = per[:, :per.shape[1] // fps * fps, :]
        per = per.reshape(per.shape[0], per.shape[1] // fps, 3 * fps)
        if per.shape[1] < 2 * running_length:
            continue
        if not pick_user:
            for k, a in enumerate(reshape2second_stacks(per, True, stride, running_length)):
                tar[k].append(a)
            continue
        for t in range(per.shape[0]):
            idx = others_index_list(per.shape[0], t, num_user, draw)
            for k, a in enumerate(reshape2second_stacks(per[t:t + 1], True, stride, running_length)):
                tar[k].append(a)
            for k, a in enumerate(reshape2second_stacks(per[idx], False, stride, running_length)):
                oth[k].append(a)
    out = [np.concatenate(a, axis=0) for a in tar]
    if pick_user:
        out += [np.concatenate(a, axis=1) for a in oth]
    return tuple(out)


def get_whole_span(x):
    """mycode/others_LSTM_span_whole.py:403-419: (N,L,...) -> (N,2L,...), row i =
    [x[i]; x[i+1]]; the last row stays zero."""
    out = np.zeros((x.shape[0], 2 * x.shape[1]) + x.shape[2:], x.dtype)
    out[:-1] = np.concatenate((x[:-1], x[1:]), axis=1)
    return out


def xyz2thetaphi(x, y, z):
    """mycode/dataIO.py:77-82: theta in [-pi,pi), phi in [0,pi)."""
    theta = np.mod(np.arctan2(y, x), 2 * np.pi) - np.pi
    phi = np.mod(np.arctan2(z, np.sqrt(x ** 2 + y ** 2)) + np.pi / 2, np.pi)
    return theta, phi


def theta_phi_index(frames, bin_size=10):
    """mycode/utility.py:520-539: (...,F,3) xyz -> integer (theta_index, phi_index)."""
    theta, phi = xyz2thetaphi(frames[..., 0], frames[..., 1], frames[..., 2])
    ti = np.floor((theta + np.pi) / np.pi * 180 / bin_size)
    ti[ti == 360 / bin_size] -= 1
    pj = np.floor(phi / np.pi * 180 / bin_size)
    pj[pj == 180 / bin_size] -= 1
    return ti.astype(np.int64), pj.astype(np.int64)


def one_hot_heatmaps(frames, bin_size=10):
    """_create_one_hot (mycode/utility.py:546-556) of the indices above, then frames
    as channels (mycode/data_generator_for_heatmap.py:32,65-67):
    (N,T,F,3) -> (N,T,360/bin,180/bin,F)."""
    ti, pj = theta_phi_index(frames, bin_size)
    N, T, Fr = ti.shape
    one_hot = np.zeros((N, T, Fr, 360 // bin_size, 180 // bin_size))
    n, t, f = np.meshgrid(np.arange(N), np.arange(T), np.arange(Fr), indexing="ij")
    one_hot[n, t, f, ti, pj] = 1
    return one_hot.transpose(0, 1, 3, 4, 2)


def theta_phi_frames(xyz):
    """get_theta_phi_array / get_theta_phi_array_per_user (mycode/data_generator_gaussian_FoV.py:21-55) for one
    viewer: (..., F, 3) xyz frames -> (..., F, 2) = [phi / pi, (theta + pi) / 2 / pi], both in [0, 1]."""
    theta, phi = xyz2thetaphi(xyz[..., 0], xyz[..., 1], xyz[..., 2])
    return np.stack([phi / np.pi, (theta + np.pi) / 2 / np.pi], axis=-1)


def _fov_rows(kind, xi, img_h):
    """Image rows visited by the row loop: crop_FoV_from_equirect (:71-74, 90 rows from xi - 44, wrapping over the
    poles) or blur_head_direction_equirect (:176-183, blur_h rows from max(int(xi - blur_h / 2), 0) + 1, no wrap).
    Returns (rows that exist in the image, rowx of the LAST visited row - the sigma below is taken from it)."""
    if kind == "fov":
        row = int(xi - img_h / 4 + img_h)
        n_it = 0
        while n_it < img_h / 2:
            n_it += 1
        rows = [(row + 1 + i) % img_h for i in range(n_it)]
    else:
        blur_h = 5 / 18 * img_h
        row = int(xi - blur_h / 2)
        n_it = 0
        while n_it < blur_h:
            n_it += 1
        if row <= 0:
            row = 0
        rows = [row + 1 + i for i in range(n_it)]
    return rows, rows[-1] / float(img_h)


def _fov_longitude(kind, row, img_h, img_w):
    """Half width (columns) of the FoV at an image row (:77-80 / :186-189); float index of the Python-2 code -> int."""
    rowx = row / float(img_h)
    half_span = int(img_w / 6) if kind == "fov" else 0.5 * (5 / 36) * img_w
    lon = int(half_span / (math.cos(math.pi * (abs(rowx - 0.5))) + 0.00001))
    if kind == "fov":
        if lon > img_w / 2 - 1:
            lon = int(img_w / 2 - 1)
    elif lon >= img_h - 1:
        lon = img_h - 1
    return lon


def gaussian_fov_frames(phi_theta, kind="fov", img_h=180, img_w=360, full=False):
    """get_gaussian_FoV (mycode/data_generator_gaussian_FoV.py:121-127 -> crop_FoV_from_equirect :57-118,
    kind="fov") and get_head_direction (:226-232 -> blur_head_direction_equirect :163-224, kind="head") restated:
    (n, 2) frame centres [phi, theta] in [0, 1] -> (n, 18, 36, 1) float32, every 10th row / column of the 180 x 360
    map, normalised by the maximum of the FULL-resolution maps of all n frames.  A frame's map is the indicator of a
    latitude-dependent column span around the centre (wrapping in theta) times exp(-d^2 / 2 * sigma^2) + its
    wrapped copy; sigma comes from the last visited row (as written: the value multiplies, it does not divide)."""
    phi_theta = np.asarray(phi_theta, np.float64).reshape(-1, 2)
    n = phi_theta.shape[0]
    shrink = img_h / 256
    Y, X = np.meshgrid(np.arange(img_h, dtype=np.float64), np.arange(img_w, dtype=np.float64), indexing="ij")
    maps = np.zeros((n, img_h, img_w), np.float32)
    for k in range(n):
        xi, zi = int(phi_theta[k, 0] * img_h), int(phi_theta[k, 1] * img_w)
        rows, rowx = _fov_rows(kind, xi, img_h)
        mask = np.zeros((img_h, img_w))
        for row in rows:
            if row >= img_h:
                continue
            lon = _fov_longitude(kind, row, img_h, img_w)
            zlow, zhigh = zi - lon, zi + lon
            if zlow < 0:
                mask[row, zlow + img_w:img_w] = 1
                zlow = 0
            if zhigh > img_w:
                mask[row, 0:zhigh % img_w] = 1
                zhigh = img_w - 1
            mask[row, zlow:zhigh] = 1
        sigma = ((0.01 if kind == "fov" else 0.05) + 0.008 * (math.pi * (abs(rowx - 0.5)))) / shrink
        G = np.exp(-((X - zi) ** 2 + (Y - xi) ** 2) / 2.0 * sigma ** 2)
        zoo = zi
        margin = int(img_w * 0.3)
        if zi + margin > img_w:
            zoo = zi - img_w
        if zi - margin < 0:
            zoo = zi + img_w
        G1 = np.exp(-((X - zoo) ** 2 + (Y - xi) ** 2) / 2.0 * sigma ** 2)
        if zoo == zi:
            G1 = 0.00001 * G1
        maps[k] = mask * (G + G1)
    maps = maps / np.max(maps)
    if full:
        return maps
    return maps[:, 0::10, 0::10, None]


def gaussian_fov_per_video(phi_theta, kind="fov", fps=FPS):
    """get_gaussianFoV_per_vid_per_target_giventhetaphi / get_headdirection_... (:130-138, :235-243):
    (num_user, num_sec * fps, 2) -> (num_user, num_sec, 18, 36, fps), frames of a second as channels."""
    U, n = phi_theta.shape[0], phi_theta.shape[1] // fps
    t = gaussian_fov_frames(phi_theta[:, :n * fps].reshape(-1, 2), kind)
    return t.reshape(U, n, fps, 18, 36).transpose(0, 1, 3, 4, 2)


def heatmap_sum(heat_frame):
    """heatmap_sum + normalize_to_distribution (mycode/data_generator_gaussian_FoV.py:246-261): sum the frame
    channels of every second and scale each (18, 36) map to sum 1."""
    s = np.sum(heat_frame, axis=-1)[..., np.newaxis]
    flat = s.reshape((-1,) + s.shape[-3:])
    for i in range(flat.shape[0]):
        flat[i] = flat[i] / np.sum(flat[i])
    return flat.reshape(s.shape)


def hit_rate(pred, gt, a=1.0, span_deg=120.0):
    """mycode/baseline_knn_mean.py:48-93 (boundary_cases, get_iou_or_hitrate,
    bbox_overlaps_hit_rate) and :131-132 (spans), vectorised: (...,2) (theta, phi)
    centres -> box overlap over the ground-truth box area."""
    ct, cp = np.array(pred[..., 0], np.float64), np.array(pred[..., 1], np.float64)
    gt_t, gt_p = np.array(gt[..., 0], np.float64), np.array(gt[..., 1], np.float64)
    c1 = (gt_t > 2 / 3.0 * np.pi) & (ct < -2 / 3.0 * np.pi)
    ct = np.where(c1, ct + 2 * np.pi, ct)
    c2 = (gt_t < -2 / 3.0 * np.pi) & (ct > 2 / 3.0 * np.pi)
    gt_t = np.where(c2, gt_t + 2 * np.pi, gt_t)
    s = a * span_deg / 180.0 * np.pi
    g = span_deg / 180.0 * np.pi
    iw = np.minimum(ct + s / 2, gt_t + g / 2) - np.maximum(ct - s / 2, gt_t - g / 2)
    ih = np.minimum(cp + s / 2, gt_p + g / 2) - np.maximum(cp - s / 2, gt_p - g / 2)
    return np.where((iw > 0) & (ih > 0), iw * ih / (g * g), 0.0)


# --------------------------------------------------------------------------- #
# losses
# --------------------------------------------------------------------------- #


def mse(y_true, y_pred):
    """Keras 'mean_squared_error' reduced by fit(): mean over every element
    (mycode/FoV_seq2seq.py:103; mycode/cost.py:20-29 with add_xyz_sum1=False)."""
    return float(np.mean((y_pred - y_true) ** 2))


def gauss_nll(y_true, y_pred, running_length=10, fps=FPS, eps=1e-7):
    """mycode/cost.py:138-187 ``likelihood_loss``.  y_true (B,T,90) interleaved
    xyz frames, y_pred (B,T,6)=[ux,uy,uz,varx,vary,varz]."""
    total = 0.0
    for a in range(3):
        u = y_pred[:, :, a:a + 1]
        v = np.clip(np.abs(y_pred[:, :, 3 + a:4 + a]), 1e-4, 2.0)
        x = y_true[:, :, a::3]
        l = np.log(v + eps) + (x - u) ** 2 / (v + eps)
        l = np.clip(l, -2000.0, 2000.0)
        total = total + l
    loss = np.mean(np.sum(np.sum(total, axis=2), axis=1))
    return float(loss / running_length / fps)


def categorical_crossentropy(y_true, y_pred, eps=1e-7):
    """Keras 'categorical_crossentropy' on probabilities, last axis = classes
    (mycode/convlstm_heatmap.py:281): renormalise, clip to [eps,1-eps], -sum t*log p,
    mean over the remaining axes."""
    p = y_pred / np.sum(y_pred, axis=-1, keepdims=True)
    p = np.clip(p, eps, 1.0 - eps)
    return float(np.mean(-np.sum(y_true * np.log(p), axis=-1)))


# --------------------------------------------------------------------------- #
# optimisers (Keras forms)
# --------------------------------------------------------------------------- #


def adam_step(p, g, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras 2.2 Adam.get_updates: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    p -= lr_t * m / (sqrt(v) + eps)  (eps OUTSIDE the bias correction).  t is the
    1-based iteration.  ('Adam' string default, mycode/FoV_seq2seq.py:103)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


def rmsprop_step(p, g, a, lr=1e-3, rho=0.9, eps=1e-7):
    """Keras 2.2 RMSprop.get_updates ('RMSprop' string default,
    mycode/convlstm_seq2seq.py:287)."""
    a = rho * a + (1 - rho) * g * g
    p = p - lr * g / (np.sqrt(a) + eps)
    return p, a


# --------------------------------------------------------------------------- #
# initialisers (from-scratch loss curves only)
# --------------------------------------------------------------------------- #


def glorot_uniform(rng, shape, dtype=np.float32):
    """Keras glorot_uniform: fan_in/fan_out use the receptive field for conv kernels."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def orthogonal(rng, shape, dtype=np.float32):
    """Keras Orthogonal(gain=1): SVD of a normal matrix, flattened over leading dims."""
    rows = int(np.prod(shape[:-1]))
    cols = shape[-1]
    a = rng.normal(0.0, 1.0, (rows, cols))
    u, _, vt = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == (rows, cols) else vt
    return q.reshape(shape).astype(dtype)


def lstm_bias(units, dtype=np.float32):
    """zeros with the forget block = 1 (unit_forget_bias=True)."""
    b = np.zeros(4 * units, dtype)
    b[units:2 * units] = 1.0
    return b


def init_lstm(rng, in_dim, units, prefix, out):
    out[prefix + "/kernel"] = glorot_uniform(rng, (in_dim, 4 * units))
    out[prefix + "/recurrent_kernel"] = orthogonal(rng, (units, 4 * units))
    out[prefix + "/bias"] = lstm_bias(units)


def init_convlstm(rng, kh, kw, cin, filters, prefix, out):
    out[prefix + "/kernel"] = glorot_uniform(rng, (kh, kw, cin, 4 * filters))
    out[prefix + "/recurrent_kernel"] = orthogonal(rng, (kh, kw, filters, 4 * filters))
    out[prefix + "/bias"] = lstm_bias(filters)


def init_dense(rng, in_dim, out_dim, prefix, out):
    out[prefix + "/kernel"] = glorot_uniform(rng, (in_dim, out_dim))
    out[prefix + "/bias"] = np.zeros(out_dim, np.float32)


def init_conv(rng, kshape, prefix, out):
    out[prefix + "/kernel"] = glorot_uniform(rng, kshape)
    out[prefix + "/bias"] = np.zeros(kshape[-1], np.float32)


# --------------------------------------------------------------------------- #
# canonical model graphs (SURVEY.md 8a')
# --------------------------------------------------------------------------- #


def init_fov_seq2seq(seed=1, num_encoder_tokens=90, num_decoder_tokens=6, latent_dim=64):
    """Weights of M1 (num_encoder_tokens=90) / M2 (=6)."""
    rng = np.random.default_rng(seed)
    w = {}
    init_lstm(rng, num_encoder_tokens, latent_dim, "encoder", w)
    init_lstm(rng, num_decoder_tokens, latent_dim, "decoder", w)
    init_dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
    return w


def fov_seq2seq_forward(w, enc_in, dec_in, teacher_forcing=True, decoder_no_init_state=False,
                        recurrent_activation="hard_sigmoid", steps=None):
    """M1/M2.  Teacher forcing: mycode/FoV_seq2seq.py:82-97 (dec_in (B,T,6)).
    Autoregressive: mycode/FoV_seq2seq.py:154-178 and the in-graph form
    mycode/FoV_seq2seq_no_teac_forc.py:88-129 (dec_in (B,1,6), output re-fed).
    ``decoder_no_init_state`` is the ablation at FoV_seq2seq_no_teac_forc.py:29,98-101
    (zero decoder state at step 0)."""
    _, h, c = lstm(enc_in, w["encoder/kernel"], w["encoder/recurrent_kernel"], w["encoder/bias"],
                   recurrent_activation=recurrent_activation)
    if decoder_no_init_state:
        h, c = np.zeros_like(h), np.zeros_like(c)
    if teacher_forcing:
        seq, _, _ = lstm(dec_in, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], h, c, recurrent_activation)
        return dense(seq, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
    T = steps if steps is not None else 10
    x = dec_in[:, 0]
    outs = []
    for _ in range(T):
        h, c = lstm_step(x, h, c, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], recurrent_activation)
        y = dense(h, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        outs.append(y)
        x = y
    return np.stack(outs, axis=1)


def init_stacked_fov_seq2seq(seed=1, n_layers=2, num_encoder_tokens=6, num_decoder_tokens=6, latent_dim=64):
    """Weights of the 2- / 3-layer target-only models (mycode/Fov_seq2seq_2layers.py:232-272, mycode/3layers.py:223-275):
    n_layers encoder LSTMs and n_layers decoder LSTMs of latent_dim // 2 units, Dense(num_decoder_tokens, tanh)."""
    rng = np.random.default_rng(seed)
    units = latent_dim // 2
    w = {}
    for l in range(n_layers):
        init_lstm(rng, num_encoder_tokens if l == 0 else units, units, "encoder%d" % l, w)
    for l in range(n_layers):
        init_lstm(rng, num_decoder_tokens if l == 0 else units, units, "decoder%d" % l, w)
    init_dense(rng, units, num_decoder_tokens, "decoder_dense", w)
    return w


def stacked_fov_seq2seq_forward(w, enc_in, dec_in, n_layers=2, share_last_decoder=None,
                                recurrent_activation="hard_sigmoid"):
    """Teacher-forced forward of the stacked models: encoder layer l reads the hidden sequence of layer l-1, decoder
    layer l starts from the final state of encoder layer l and reads the hidden sequence of decoder layer l-1.
    ``share_last_decoder`` (default: n_layers == 3): mycode/3layers.py:266 calls ``decoder_lstm2`` again for the third
    decoder layer, so layers 2 and 3 of the decoder share one set of weights - it is what the graph computes."""
    if share_last_decoder is None:
        share_last_decoder = n_layers == 3
    xe, xd = enc_in, dec_in
    for l in range(n_layers):
        pe = "encoder%d" % l
        pd = "decoder%d" % (l - 1 if (share_last_decoder and l == n_layers - 1) else l)
        xe, h, c = lstm(xe, w[pe + "/kernel"], w[pe + "/recurrent_kernel"], w[pe + "/bias"],
                        recurrent_activation=recurrent_activation)
        xd, _, _ = lstm(xd, w[pd + "/kernel"], w[pd + "/recurrent_kernel"], w[pd + "/bias"], h, c,
                        recurrent_activation=recurrent_activation)
    return dense(xd, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")


def init_given_others_seq2seq(seed=1, num_user=34, latent_dim=32, num_encoder_tokens=6, num_decoder_tokens=6,
                              variant="mlp_mixing"):
    """Weights of mycode/given_others_gt_mean_var_seq2seq.py:97-166: two encoder and two decoder LSTMs of latent_dim
    (32) units, Dense(6, tanh), and per variant: 'mlp_mixing' (the script's default) a Dense(6, tanh) over
    [others' mean/var of the step (num_user-1, 6) ; decoder prediction (6)]; 'others_mlp' Dense(256, relu) ->
    Dense(latent_dim, relu) on the others slice, whose output is concatenated with the decoder state before
    decoder_dense; 'others_lstm' (:151-158, the branch the script's own flags select under model_others) two stacked
    Bidirectional(LSTM(latent_dim), merge_mode='concat') over the others' future, decoder_dense on
    [bi-LSTM output (2*latent_dim) ; decoder state]; 'conv_mixing' (:190-199) Conv2D(8,(1,3)) -> Conv2D(8,(1,3)) ->
    Conv2D(1,(1,3)), all relu, over the (1, 6, num_user) image of [others ; prediction]; 'target_only' nothing else."""
    rng = np.random.default_rng(seed)
    w = {}
    for l in range(2):
        init_lstm(rng, num_encoder_tokens if l == 0 else latent_dim, latent_dim, "encoder%d" % l, w)
    for l in range(2):
        init_lstm(rng, num_decoder_tokens if l == 0 else latent_dim, latent_dim, "decoder%d" % l, w)
    oth = (num_user - 1) * 6
    if variant == "others_lstm":
        for l in range(2):
            for d in ("fwd", "bwd"):
                init_lstm(rng, oth if l == 0 else 2 * latent_dim, latent_dim, "others_bilstm%d_%s" % (l, d), w)
        init_dense(rng, 3 * latent_dim, num_decoder_tokens, "decoder_dense", w)
    elif variant == "conv_mixing":
        init_dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
        for l, (ci, co) in enumerate([(num_user, 8), (8, 8), (8, 1)]):
            init_conv(rng, (1, 3, ci, co), "mixing_conv%d" % l, w)
    elif variant == "others_mlp":
        init_dense(rng, oth, 256, "others_dense1", w)
        init_dense(rng, 256, latent_dim, "others_dense2", w)
        init_dense(rng, 2 * latent_dim, num_decoder_tokens, "decoder_dense", w)
    else:
        init_dense(rng, latent_dim, num_decoder_tokens, "decoder_dense", w)
        if variant == "mlp_mixing":
            init_dense(rng, oth + num_decoder_tokens, num_decoder_tokens, "mixing", w)
    return w


def given_others_seq2seq_forward(w, enc_in, oth_in, dec_in, variant="mlp_mixing", teacher_forcing=False,
                                 recurrent_activation="hard_sigmoid"):
    """mycode/given_others_gt_mean_var_seq2seq.py:97-308 with cfg.input_mean_var / predict_mean_var: 2-layer fc-LSTM
    encoder-decoder; every step's output is re-fed as the next decoder input unless teacher_forcing
    (cfg.teacher_forcing, default False).  enc_in (B,10,6), oth_in (B,T,num_user-1,6) others' ground-truth mean/var
    of the future seconds, dec_in (B,1,6) (or (B,T,6) teacher-forced) -> (B,T,6).

    'others_lstm': Keras' Bidirectional runs the backward LSTM on the time-reversed input and reverses its output
    sequence back before the concat; called with return_state=True it returns [seq, fwd_h, fwd_c, bwd_h, bwd_c], and
    the script passes that LIST to the second Bidirectional (:154-155), which Keras reads as input + initial_state:
    the second pair starts from the first pair's final states.  In the teacher-forced graph the script takes
    get_dim1_layer(decoder2_outputs) = the decoder's FIRST step for every output step (:238); restated as written."""
    ra = recurrent_activation
    B, T = oth_in.shape[0], oth_in.shape[1]
    L = lambda n: (w[n + "/kernel"], w[n + "/recurrent_kernel"], w[n + "/bias"])
    e1, h1, c1 = lstm(enc_in, *L("encoder0"), recurrent_activation=ra)
    _, h2, c2 = lstm(e1, *L("encoder1"), recurrent_activation=ra)
    if variant == "others_lstm":
        y, st = oth_in.reshape(B, T, -1), {"fwd": (None, None), "bwd": (None, None)}
        for l in range(2):
            f, fh, fc = lstm(y, *L("others_bilstm%d_fwd" % l), *st["fwd"], recurrent_activation=ra)
            b, bh, bc = lstm(y[:, ::-1], *L("others_bilstm%d_bwd" % l), *st["bwd"], recurrent_activation=ra)
            y, st = np.concatenate([f, b[:, ::-1]], axis=-1), {"fwd": (fh, fc), "bwd": (bh, bc)}
        oth_seq = y
    if teacher_forcing:
        d1, _, _ = lstm(dec_in, *L("decoder0"), h1, c1, recurrent_activation=ra)
        d2, _, _ = lstm(d1, *L("decoder1"), h2, c2, recurrent_activation=ra)
    x = dec_in[:, 0]
    outs = []
    for t in range(T):
        if teacher_forcing:
            s2 = d2[:, t]
        else:
            h1, c1 = lstm_step(x, h1, c1, *L("decoder0"), recurrent_activation=ra)
            h2, c2 = lstm_step(h1, h2, c2, *L("decoder1"), recurrent_activation=ra)
            s2 = h2
        flat = oth_in[:, t].reshape(B, -1)
        if variant == "target_only":
            y = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        elif variant == "others_lstm":
            s = d2[:, 0] if teacher_forcing else s2
            y = dense(np.concatenate([oth_seq[:, t], s], axis=-1), w["decoder_dense/kernel"], w["decoder_dense/bias"],
                      "tanh")
        elif variant == "conv_mixing":
            pred = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
            img = np.concatenate([oth_in[:, t], pred[:, None]], axis=1).transpose(0, 2, 1)[:, None]   # (B,1,6,num_user)
            for l in range(3):
                img = conv2d(img, w["mixing_conv%d/kernel" % l], w["mixing_conv%d/bias" % l], "relu")
            y = img[:, 0, :, 0]
        elif variant == "others_mlp":
            o = dense(flat, w["others_dense1/kernel"], w["others_dense1/bias"], "relu")
            o = dense(o, w["others_dense2/kernel"], w["others_dense2/bias"], "relu")
            y = dense(np.concatenate([o, s2], axis=-1), w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
        else:
            pred = dense(s2, w["decoder_dense/kernel"], w["decoder_dense/bias"], "tanh")
            y = dense(np.concatenate([flat, pred], axis=-1), w["mixing/kernel"], w["mixing/bias"], "tanh")
        outs.append(y)
        x = y
    return np.stack(outs, axis=1)


def init_others_lstm_span_whole(seed=1, num_user=34, kernel_size=5, latent_dim=64,
                                oth_filters=(32, 16, 8), flat_dense=256):
    """Weights of M3, the canonical concat-state model (SURVEY.md hazard 2)."""
    rng = np.random.default_rng(seed)
    w = {}
    cin = 6
    for l, f in enumerate(oth_filters):
        init_convlstm(rng, 1, kernel_size, cin, f, "oth_convlstm%d" % l, w)
        cin = f
    flat = (num_user - 1) * sum(oth_filters)
    init_dense(rng, flat, (num_user - 1) * 6, "oth_recon_dense", w)
    init_dense(rng, flat, flat_dense, "oth_flat_dense", w)
    init_lstm(rng, 6, latent_dim, "encoder", w)
    init_lstm(rng, 6, latent_dim, "decoder", w)
    init_dense(rng, latent_dim, 6, "encoder_dense", w)
    init_dense(rng, latent_dim + flat_dense, 6, "decoder_dense", w)
    return w


def others_convlstm_stack(w, x, prefix="oth_convlstm", n_layers=3, h0c0=None, dilation=(1, 1),
                          recurrent_activation="hard_sigmoid"):
    """Three stacked ConvLSTM2D layers, channel-concat of their sequences
    (mycode/others_LSTM_span_whole.py:88-102).  Returns (concat_seq, [(h,c)]*3)."""
    seqs, states = [], []
    cur = x
    for l in range(n_layers):
        p = "%s%d" % (prefix, l)
        h0, c0 = (None, None) if h0c0 is None else h0c0[l]
        cur, h, c = convlstm2d(cur, w[p + "/kernel"], w[p + "/recurrent_kernel"], w[p + "/bias"],
                               h0, c0, dilation, recurrent_activation)
        seqs.append(cur)
        states.append((h, c))
    return np.concatenate(seqs, axis=-1), states


def others_lstm_span_whole_forward(w, enc_in, oth_in, dec_in, recurrent_activation="hard_sigmoid"):
    """M3 forward.  enc_in (B,10,6), oth_in (B,20,1,33,6), dec_in (B,1,6) ->
    [decoder_outputs (B,10,6), decoder_outputs_oth (B,20,198), encoder_reconstruct_tar (B,10,6)]
    (mycode/others_LSTM_span_whole.py:80-132,226-349 with use_fclstm_tar=True and the
    intended nesting of mycode/others_LSTM_span_whole_attention.py:233-263)."""
    B, Tenc = enc_in.shape[:2]
    Tall = oth_in.shape[1]
    Tdec = Tall - Tenc
    oth_seq, _ = others_convlstm_stack(w, oth_in, recurrent_activation=recurrent_activation)
    flat = oth_seq.reshape(B, Tall, -1)
    r_oth = dense(flat, w["oth_recon_dense/kernel"], w["oth_recon_dense/bias"])
    enc_seq, h, c = lstm(enc_in, w["encoder/kernel"], w["encoder/recurrent_kernel"],
                         w["encoder/bias"], recurrent_activation=recurrent_activation)
    r_tar = dense(enc_seq, w["encoder_dense/kernel"], w["encoder_dense/bias"], "tanh")
    x = dec_in[:, 0]
    outs = []
    for t in range(Tdec):
        h, c = lstm_step(x, h, c, w["decoder/kernel"], w["decoder/recurrent_kernel"],
                         w["decoder/bias"], recurrent_activation)
        s = dense(flat[:, Tenc + t], w["oth_flat_dense/kernel"], w["oth_flat_dense/bias"])
        y = dense(np.concatenate([h, s], axis=1), w["decoder_dense/kernel"],
                  w["decoder_dense/bias"])
        outs.append(y)
        x = y
    return [np.stack(outs, axis=1), r_oth, r_tar]


def init_others_convlstm_target(seed=1, num_user=34, kernel_size=5, oth_filters=(32, 16, 8), tar_filters=(8, 4, 2)):
    """Weights of the all-ConvLSTM form of mycode/others_LSTM_span_whole.py (use_fclstm_tar=False, :133-199, raw xyz
    layout: cfg.input_mean_var = cfg.predict_mean_var = False, the defaults of mycode/config.py:69-75): others' stack
    on (1,fps,(num_user-1)*3) images, a target encoder stack and a target decoder stack of latent_dim_target = 8 / 4 / 2
    filters (the decoder reads [its own last output ; others' state of that second], 3 + 56 channels), Dense((num_user-1)*3)
    on the others' state, Dense(3,tanh) on the target's past state, Dense(3) on the decoder state."""
    rng = np.random.default_rng(seed)
    w = {}
    cin = (num_user - 1) * 3
    for l, f in enumerate(oth_filters):
        init_convlstm(rng, 1, kernel_size, cin, f, "oth_convlstm%d" % l, w)
        cin = f
    cin = 3
    for l, f in enumerate(tar_filters):
        init_convlstm(rng, 1, kernel_size, cin, f, "tar_enc_convlstm%d" % l, w)
        cin = f
    cin = 3 + sum(oth_filters)
    for l, f in enumerate(tar_filters):
        init_convlstm(rng, 1, kernel_size, cin, f, "tar_dec_convlstm%d" % l, w)
        cin = f
    init_dense(rng, sum(oth_filters), (num_user - 1) * 3, "oth_recon_dense", w)
    init_dense(rng, sum(tar_filters), 3, "encoder_dense", w)
    init_dense(rng, sum(tar_filters), 3, "decoder_dense", w)
    return w


def others_convlstm_target_forward(w, enc_in, oth_in, dec_in, recurrent_activation="hard_sigmoid"):
    """mycode/others_LSTM_span_whole.py:80-102,133-199,226-349 with use_fclstm_tar=False in the raw layout:
    enc_in (B,10,1,fps,3), oth_in (B,20,1,fps,(U-1)*3), dec_in (B,1,1,fps,3) ->
    [decoder_outputs (B,10,1,fps,3), decoder_outputs_oth (B,20,1,fps,(U-1)*3), encoder_reconstruct_tar (B,10,1,fps,3)].
    Per future second: decoder input = channel concat [last output ; others' state] (:273-275), three one-step
    ConvLSTMs seeded by the encoder states, Dense(3) on their concatenated states (:296), output re-fed (:317)."""
    ra = recurrent_activation
    Tenc = enc_in.shape[1]
    Tdec = oth_in.shape[1] - Tenc
    oth_seq, _ = others_convlstm_stack(w, oth_in, recurrent_activation=ra)
    r_oth = dense(oth_seq, w["oth_recon_dense/kernel"], w["oth_recon_dense/bias"])
    pst, states = others_convlstm_stack(w, enc_in, "tar_enc_convlstm", recurrent_activation=ra)
    r_tar = dense(pst, w["encoder_dense/kernel"], w["encoder_dense/bias"], "tanh")
    x = dec_in
    outs = []
    for t in range(Tdec):
        cat_in = np.concatenate([x, oth_seq[:, Tenc + t:Tenc + t + 1]], axis=-1)
        dstate, states = others_convlstm_stack(w, cat_in, "tar_dec_convlstm", h0c0=states, recurrent_activation=ra)
        y = dense(dstate, w["decoder_dense/kernel"], w["decoder_dense/bias"])
        outs.append(y)
        x = y
    return [np.concatenate(outs, axis=1), r_oth, r_tar]


def init_convlstm_seq2seq(seed=1, in_ch=30, filters=(32, 16, 8), kernel_size=5,
                          head=(512, 1024, 30), head_kind="conv2d", head_kernel=None,
                          flat_dim=None):
    """Weights of M4 (heatmap form: head_kind='conv2d'; trajectory form: 'conv1d'
    with kernel 7, or 'dense')."""
    rng = np.random.default_rng(seed)
    w = {}
    for side in ("enc", "dec"):
        cin = in_ch
        for l, f in enumerate(filters):
            init_convlstm(rng, kernel_size, kernel_size, cin, f, "%s_convlstm%d" % (side, l), w)
            cin = f
    cat = sum(filters)
    if head_kind == "conv2d":
        cin = cat
        for l, f in enumerate(head):
            init_conv(rng, (kernel_size, kernel_size, cin, f), "head_conv%d" % l, w)
            cin = f
    elif head_kind == "conv1d":
        cin = cat
        k = head_kernel or 7
        for l, f in enumerate(head):
            init_conv(rng, (k, cin, f), "head_conv%d" % l, w)
            cin = f
    else:
        init_dense(rng, flat_dim, 6, "head_dense", w)
    return w


def philox4x32_10(counter, seed):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11 - the counter-based generator behind TF's
    random_normal, which K.random_normal at mycode/convlstm_seq2seq.py:57 calls): counter (n,) uint64 ->
    (n,4) uint32 words with counter words (lo, hi, 0, 0) and key (seed lo, seed hi).  Integer restatement the
    CUDA kernel fov_philox_normal is compared with bit for bit."""
    c = np.zeros((len(counter), 4), np.uint64)
    counter = np.asarray(counter, np.uint64)
    c[:, 0] = counter & np.uint64(0xFFFFFFFF)
    c[:, 1] = counter >> np.uint64(32)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c.astype(np.uint32)


def philox_normal(n, seed, offset=0):
    """n N(0,1) draws of the stream fov_philox_normal defines: element i = word i%4 of counter offset + i//4,
    Box-Muller on word pairs: u1 = (w0+1) 2^-32, u2 = w1 2^-32, (z0, z1) = sqrt(-2 ln u1) (cos, sin)(2 pi u2).
    Returns (words uint32 (n,), normals float64 (n,))."""
    nblk = (n + 3) // 4
    w = philox4x32_10(np.arange(nblk, dtype=np.uint64) + np.uint64(offset), seed)
    u1 = (w[:, 0::2].astype(np.float64) + 1.0) * 2.0 ** -32
    u2 = w[:, 1::2].astype(np.float64) * 2.0 ** -32
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], axis=-1)   # (nblk, 2 pairs, 2)
    return w.reshape(-1)[:n], z.reshape(-1)[:n]


def convlstm_seq2seq_forward(w, enc_in, dec_in, head_kind="conv2d", steps=10, dilation=(1, 1),
                             recurrent_activation="hard_sigmoid", noise=None, resample_mode="var_as_std"):
    """M4 forward (mycode/convlstm_seq2seq.py:100-126,146-165,209-281).
    heatmap form: enc_in (B,10,36,18,30), dec_in (B,1,36,18,30) -> (B,10,36,18,30);
    trajectory forms: enc_in (B,10,1,30,3), dec_in (B,1,1,30,3) ->
    conv1d head (B,10,1,30,3) (last Conv1D has softmax, :188-189) or dense head
    (B,10,6) with enc_in (B,10,1,1,6) (input_mean_var)."""
    B = enc_in.shape[0]
    _, states = others_convlstm_stack(w, enc_in, prefix="enc_convlstm", dilation=dilation,
                                      recurrent_activation=recurrent_activation)
    x = dec_in[:, 0]
    outs = []
    for _ in range(steps):
        hs = []
        cur = x
        new_states = []
        for l in range(3):
            p = "dec_convlstm%d" % l
            h, c = convlstm2d_step(cur, states[l][0], states[l][1], w[p + "/kernel"],
                                   w[p + "/recurrent_kernel"], w[p + "/bias"], dilation,
                                   recurrent_activation)
            new_states.append((h, c))
            hs.append(h)
            cur = h
        states = new_states
        d = np.concatenate(hs, axis=-1)                     # (B,H,W,56)
        if head_kind == "conv2d":
            y = conv2d(d, w["head_conv0/kernel"], w["head_conv0/bias"], "relu")
            y = conv2d(y, w["head_conv1/kernel"], w["head_conv1/bias"], "relu")
            y = conv2d(y, w["head_conv2/kernel"], w["head_conv2/bias"], "relu")
            y = softmax(y, -1)
            x = y
        elif head_kind == "conv1d":
            y = conv1d(d[:, 0], w["head_conv0/kernel"], w["head_conv0/bias"], "relu")
            y = conv1d(y, w["head_conv1/kernel"], w["head_conv1/bias"], "relu")
            y = conv1d(y, w["head_conv2/kernel"], w["head_conv2/bias"], "softmax")
            y = y[:, None]
            x = y
        else:
            y = dense(d[:, 0].reshape(B, -1), w["head_dense/kernel"], w["head_dense/bias"])
            if noise is not None:     # cfg.sample_and_refeed (mycode/convlstm_seq2seq.py:259-272)
                x = gaussian_resample(y[:, :3], y[:, 3:], noise[len(outs)], resample_mode)[:, None]
            else:
                x = y[:, None, None, :]
        outs.append(y)
    return np.stack(outs, axis=1)


def heatmap_batches_per_video(xyz, stride=10, running_length=10, bin_size=10):
    """The heatmap ConvLSTM's samples of ONE video, in the order mycode/data_generator_for_heatmap.py:103-216 yields
    them (target viewers outermost, windows in time order): per viewer the seconds of one-hot FoV-centre maps
    (:142-146) are cut into (past, future) windows by reshape2second_stacks(collapse_user=True) (:95), and
    _prepare_data (:17-75, cfg.input_mean_var = cfg.predict_mean_var = False) turns them into
    encoder input (N,10,36,18,30), decoder input = the last observed second (N,1,36,18,30) and target (N,10,36,18,30)
    (frames as channels: the transpose (0,1,4,3,2) of the stored (fps,18,36) stacks).
    xyz: (U, S, fps, 3) unit vectors -> (enc, dec0, target), N = U * windows.

    The shipped generator cannot run against the shipped utility.py (it hands 5-D stacks to the 3-D-only
    reshape2second_stacks, :94-95, and _reshape_others_data transposes six axes of a 4-D array, :23), so there is no
    reference output to record: this follows the statements on the target viewer's tensors, with the per-second stack
    flattened for the windowing; its two building blocks (one_hot_heatmaps, reshape2second_stacks) are pinned to the
    reference's own functions by golden vectors."""
    U, S = xyz.shape[:2]
    heat = one_hot_heatmaps(xyz, bin_size)                                   # (U,S,36,18,F)
    flat = heat.reshape(U, S, -1)
    enc, tgt = [], []
    for u in range(U):
        past, fut, _ = reshape2second_stacks(flat[u:u + 1], collapse_user=True, stride=stride,
                                             running_length=running_length)
        enc.append(past)
        tgt.append(fut)
    shp = (-1, running_length) + heat.shape[2:]
    enc, tgt = np.concatenate(enc).reshape(shp), np.concatenate(tgt).reshape(shp)
    return enc, enc[:, -1:], tgt
