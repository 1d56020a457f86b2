"""Synthetic head-orientation data of the reference's shapes: the gaze
datasets are not available offline.  Host-side NumPy; featurisation follows
mycode/utility.py:483-517 (mean / population variance per second), the xyz
convention follows mycode/dataIO.py:62-67 and the 10-degree one-hot binning
mycode/utility.py:536-571."""
from __future__ import annotations

import numpy as np

FPS = 30


def synth_trajectories(n_windows, n_viewers=34, seconds=20, seed=0):
    """Unit-sphere xyz trajectories, (n_windows, n_viewers, seconds*30, 3) float32.
    yaw0~U(-pi,pi), pitch0~N(0,0.3) clipped to +-1.4, AR(1)-smoothed N(0,0.02^2) rad
    per-frame increments, plus a slow drift shared by all viewers of a window."""
    rng = np.random.default_rng(seed)
    F = seconds * FPS
    yaw0 = rng.uniform(-np.pi, np.pi, (n_windows, n_viewers, 1))
    pitch0 = np.clip(rng.normal(0, 0.3, (n_windows, n_viewers, 1)), -1.4, 1.4)

    def ar1(shape):
        e = rng.normal(0, 0.02, shape).astype(np.float32)
        out = np.empty_like(e)
        acc = np.zeros(shape[:-1], np.float32)
        for f in range(shape[-1]):
            acc = 0.9 * acc + 0.1 * e[..., f]
            out[..., f] = acc
        return out

    common = ar1((n_windows, 1, F)) * 3.0
    dyaw = ar1((n_windows, n_viewers, F)) + common
    dpitch = ar1((n_windows, n_viewers, F)) * 0.5
    yaw = yaw0 + np.cumsum(dyaw, axis=-1)
    pitch = np.clip(pitch0 + np.cumsum(dpitch, axis=-1), -1.4, 1.4)
    x = np.sin(yaw) * np.cos(pitch)
    y = np.sin(pitch)
    z = np.cos(yaw) * np.cos(pitch)
    return np.clip(np.stack([x, y, z], axis=-1), -1.0, 1.0).astype(np.float32)


def mean_var(frames):
    """(..., 30, 3) -> (..., 6) = [mx,my,mz,vx,vy,vz] (np.var, ddof=0)."""
    return np.concatenate([frames.mean(axis=-2), frames.var(axis=-2)], axis=-1).astype(np.float32)


def one_hot_heatmaps(frames, bin_size=10):
    """(N,T,30,3) xyz -> (N,T,36,18,30): one one-hot FoV-centre map per frame, the 30 frames of a second stacked as
    channels (mycode/data_generator_for_heatmap.py:32,65-67).  Host-side twin of ``ops.one_hot_heatmaps`` with the
    reference's angles (mycode/dataIO.py:77-82) and binning (mycode/utility.py:533-539)."""
    x, y, z = (np.asarray(frames[..., i], np.float64) for i in range(3))
    theta = np.mod(np.arctan2(y, x), 2 * np.pi) - np.pi
    phi = np.mod(np.arctan2(z, np.sqrt(x ** 2 + y ** 2)) + np.pi / 2, np.pi)
    ti = np.floor((theta + np.pi) / np.pi * 180 / bin_size).astype(np.int64)
    ti[ti == 360 // bin_size] -= 1
    pi_ = np.floor(phi / np.pi * 180 / bin_size).astype(np.int64)
    pi_[pi_ == 180 // bin_size] -= 1
    N, T, Fr = ti.shape
    out = np.zeros((N, T, 360 // bin_size, 180 // bin_size, Fr), np.float32)
    n, t, f = np.meshgrid(np.arange(N), np.arange(T), np.arange(Fr), indexing="ij")
    out[n, t, ti, pi_, f] = 1.0
    return out


def make_m1_batch(n, seed=0, mean_var_encoder=False):
    """M1/M2 inputs: enc (n,10,90 | 6), dec_in (n,10,6), target (n,10,6), raw future (n,10,90)."""
    tr = synth_trajectories(n, 1, 20, seed)[:, 0].reshape(n, 20, FPS, 3)
    mv = mean_var(tr)
    enc = mv[:, :10] if mean_var_encoder else tr[:, :10].reshape(n, 10, 90)
    dec_in = mv[:, 9:19]
    return enc, dec_in, mv[:, 10:], tr[:, 10:].reshape(n, 10, 90)


def make_m3_batch(n, num_user=34, seed=0):
    """M3 inputs [enc (n,10,6), oth (n,20,1,num_user-1,6), dec0 (n,1,6)] and targets
    [fut (n,10,6), others (n,20,(num_user-1)*6), enc (n,10,6)]."""
    tr = synth_trajectories(n, num_user, 20, seed).reshape(n, num_user, 20, FPS, 3)
    mv = mean_var(tr)                                             # (n,U+1,20,6)
    tar, oth = mv[:, 0], mv[:, 1:].transpose(0, 2, 1, 3)          # (n,20,6), (n,20,U,6)
    enc = tar[:, :10]
    return ([enc, oth[:, :, None], tar[:, 9:10]],
            [tar[:, 10:], oth.reshape(n, 20, -1), enc])


def make_m4_batch(n, seed=0):
    """M4 heatmap inputs [enc (n,10,36,18,30), dec0 (n,1,36,18,30)] and target (n,10,36,18,30)."""
    tr = synth_trajectories(n, 1, 20, seed)[:, 0].reshape(n, 20, FPS, 3)
    hm = one_hot_heatmaps(tr)
    return [hm[:, :10], hm[:, 9:10]], [hm[:, 10:]]


def clip_xyz(per_video):
    """mycode/dataIO.py:16-26: clamp every video's 'x' / 'y' / 'z' arrays to [-1, 1] in place (the dataset pickles hold
    unit vectors with rounding spill); returns the same dict."""
    for v in per_video.values():
        for axis in ("x", "y", "z"):
            np.clip(v[axis], -1, 1, out=v[axis])
    return per_video


def rand_sample_ind(total_num_samples, num_testing_sample, batch_size, validation_ratio=0.1, rng=None):
    """mycode/utility.py:575-586: indices of a random subset of the first ``total - num_testing`` samples whose size
    makes both the training part and the ``validation_ratio`` part a whole number of batches (Keras' ``fit`` then never
    sees a ragged last batch).  ``rng``: a ``random.Random`` (default: the ``random`` module, as the reference)."""
    import random
    n = total_num_samples - num_testing_sample
    val_batches = int(n / batch_size * validation_ratio)
    n_train = int((1 - validation_ratio) / validation_ratio * val_batches * batch_size)
    n_val = int(val_batches * batch_size)
    return (rng or random).sample(range(n), n_train + n_val)


def rand_sample(data, sample_ind):
    """mycode/utility.py:589-591: the chosen samples in ascending index order."""
    return np.asarray(data)[np.sort(np.asarray(sample_ind, dtype=np.int64))]
