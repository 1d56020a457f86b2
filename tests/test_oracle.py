"""CPU tests of the oracle itself: pinned against the reference's own NumPy
functions (golden vectors), the two restatements against each other, torch.nn.LSTM
as an independent gate-algebra check, known-answer cases and finite differences."""
import os

import numpy as np
import pytest
import torch

from oracle import keras_numpy as kn
from oracle import keras_torch as kt


# ---------------- pinned: reference's own functions ------------------------ #

def test_featuriser_matches_reference_golden(golden):
    np.testing.assert_allclose(kn.get_gt_target_xyz(golden["xyz90_in"]), golden["xyz90_out"],
                               rtol=0, atol=1e-15)
    np.testing.assert_allclose(kn.get_gt_target_xyz(golden["xyz4_in"]), golden["xyz4_out"],
                               rtol=0, atol=1e-15)
    np.testing.assert_allclose(kn.get_gt_target_xyz_oth(golden["oth_in"]), golden["oth_out"],
                               rtol=0, atol=1e-15)


@pytest.mark.parametrize("stride", [10, 1, 2])
@pytest.mark.parametrize("collapse", [True, False])
def test_windowing_matches_reference_golden(golden, stride, collapse):
    a, b, c = kn.reshape2second_stacks(golden["stack_in"], collapse_user=collapse, stride=stride)
    tag = "stack_s%d_c%d" % (stride, int(collapse))
    assert np.array_equal(a, golden[tag + "_past"])
    assert np.array_equal(b, golden[tag + "_fut"])
    assert np.array_equal(c, golden[tag + "_futin"])


@pytest.mark.parametrize("stride,testing", [(10, True), (1, True), (5, False)])
@pytest.mark.parametrize("collapse", [True, False])
def test_windowing_purely_testing_and_raw_frames_match_reference_golden(golden, stride, testing, collapse):
    a, b, c = kn.reshape2second_stacks(golden["stack90_in"], collapse_user=collapse, stride=stride,
                                       purelly_testing=testing)
    tag = "stack90_s%d_t%d_c%d" % (stride, int(testing), int(collapse))
    assert np.array_equal(a, golden[tag + "_past"])
    assert np.array_equal(b, golden[tag + "_fut"])
    assert np.array_equal(c, golden[tag + "_futin"])


def test_whole_span_matches_reference_golden(golden):
    assert np.array_equal(kn.get_whole_span(golden["span_in"]), golden["span_out"])
    assert np.array_equal(kn.get_whole_span(golden["span5_in"]), golden["span5_out"])
    assert not golden["span_out"][-1].any()                     # "the last row is all zero!"


def test_one_hot_heatmaps_match_reference_golden(golden):
    ti, pj = kn.theta_phi_index(golden["onehot_in"])
    assert np.array_equal(ti, golden["onehot_theta"]) and np.array_equal(pj, golden["onehot_phi"])
    got = kn.one_hot_heatmaps(golden["onehot_in"])
    assert np.array_equal(got, golden["onehot_out"])
    assert got.sum() == golden["onehot_in"].shape[0] * golden["onehot_in"].shape[1] * 30   # one 1 per frame


def test_gaussian_fov_tiles_match_reference_golden():
    """crop_FoV_from_equirect / blur_head_direction_equirect restated (data_generator_gaussian_FoV.py:57-243):
    bit-exact against the outputs of the reference's own functions, incl. the poles, the theta seam and a call whose
    frames all wrap (the normaliser is then a wrapped frame's own peak), and the per-second distribution maps."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_gaussian_fov_golden.npz"))
    for kind in ("fov", "head"):
        assert np.array_equal(kn.gaussian_fov_per_video(g["pt"], kind), g[kind])
        assert np.array_equal(kn.gaussian_fov_per_video(g["pt_wrap"], kind), g[kind + "_wrap"])
        assert np.array_equal(kn.heatmap_sum(g[kind].copy()), g[kind + "_sum"])
    assert np.array_equal(kn.theta_phi_frames(g["xyz"]).reshape(3, 60, 2), g["xyz_phi_theta"])


def test_hit_rate_matches_reference_golden(golden):
    for a in (1.0, 0.75):
        got = kn.hit_rate(golden["hit_pred"], golden["hit_gt"], a=a)
        np.testing.assert_allclose(got, golden["hit_out_a%d" % int(a * 100)], rtol=0, atol=1e-14)


def test_resampler_matches_reference_golden(golden):
    mu, var, noise = golden["fake_mu"], golden["fake_var"], golden["fake_noise"]
    got = kn.gaussian_resample(mu[:, None], var[:, None], noise[:, :, None], "sqrt_floor")[..., 0]
    np.testing.assert_allclose(got, golden["fake_out"], rtol=0, atol=1e-14)


# ---------------- known-answer cases --------------------------------------- #

def test_hard_sigmoid_saturation():
    x = np.array([-3.0, -2.5, 0.0, 2.5, 3.0, 1.0])
    np.testing.assert_allclose(kn.hard_sigmoid(x), [0, 0, 0.5, 1, 1, 0.7])


def test_lstm_zero_weights_and_forget_bias():
    B, T, I, H = 2, 3, 4, 5
    x = np.ones((B, T, I))
    W, U = np.zeros((I, 4 * H)), np.zeros((H, 4 * H))
    seq, h, c = kn.lstm(x, W, U, np.zeros(4 * H))
    assert np.all(seq == 0) and np.all(c == 0)            # g = tanh(0) = 0
    b = np.zeros(4 * H)
    b[2 * H:3 * H] = 10.0                                   # g ~ 1, i=f=o=0.5
    seq, h, c = kn.lstm(x, W, U, b)
    g = np.tanh(10.0)
    c1 = 0.5 * g
    c2 = 0.5 * c1 + 0.5 * g
    np.testing.assert_allclose(seq[:, 1], 0.5 * np.tanh(c2), atol=1e-15)


def test_conv_same_delta_image():
    # 'same' padding on a delta image reproduces the (un-flipped) kernel, centred.
    k = np.arange(15, dtype=np.float64).reshape(3, 5, 1, 1)
    x = np.zeros((1, 7, 9, 1))
    x[0, 3, 4, 0] = 1.0
    y = kn.conv2d_same(x, k)[0, :, :, 0]
    # cross-correlation: y[p] = sum_d x[p+d-c] k[d] -> kernel appears flipped around the delta
    np.testing.assert_allclose(y[2:5, 2:7], k[::-1, ::-1, 0, 0])
    # even kernel: TF puts the extra pad at the end (low = 0 for k=2)
    k2 = np.array([1.0, 2.0]).reshape(1, 2, 1, 1)
    x2 = np.zeros((1, 1, 4, 1)); x2[0, 0, 0, 0] = 1.0
    np.testing.assert_allclose(kn.conv2d_same(x2, k2)[0, 0, :, 0], [1, 0, 0, 0])
    x2 = np.zeros((1, 1, 4, 1)); x2[0, 0, 3, 0] = 1.0
    np.testing.assert_allclose(kn.conv2d_same(x2, k2)[0, 0, :, 0], [0, 0, 2, 1])


def test_adam_first_step_known_answer():
    p, m, v = kn.adam_step(np.array([1.0]), np.array([0.5]), np.zeros(1), np.zeros(1), 1)
    # m=0.05, v=2.5e-4, lr_t = 1e-3*sqrt(1e-3)/0.1 ; step = lr_t*m/(sqrt(v)+1e-7)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(p, 1.0 - lr_t * 0.05 / (np.sqrt(2.5e-4) + 1e-7), rtol=1e-14)


def test_gauss_nll_hand_case():
    y_true = np.zeros((1, 10, 90))
    y_pred = np.zeros((1, 10, 6)); y_pred[..., 3:] = 1.0
    # var=1, x=u=0 -> log(1+1e-7) per frame per coord; sum over 30*10*3, /10/30
    np.testing.assert_allclose(kn.gauss_nll(y_true, y_pred), 3 * np.log(1 + 1e-7), rtol=1e-9)
    y_pred[..., 3:] = -5.0                                   # |var| clipped to 2
    np.testing.assert_allclose(kn.gauss_nll(y_true, y_pred), 3 * np.log(2 + 1e-7), rtol=1e-9)


# ---------------- independent restatements agree --------------------------- #

def test_lstm_matches_torch_nn_lstm_in_sigmoid_mode():
    rng = np.random.default_rng(0)
    B, T, I, H = 3, 6, 5, 8
    x = rng.normal(size=(B, T, I))
    W, U, b = rng.normal(size=(I, 4 * H)) * .3, rng.normal(size=(H, 4 * H)) * .3, rng.normal(size=4 * H) * .1
    seq, h, c = kn.lstm(x, W, U, b, recurrent_activation="sigmoid")
    ref = torch.nn.LSTM(I, H, batch_first=True).double()
    with torch.no_grad():
        ref.weight_ih_l0.copy_(torch.tensor(W.T)); ref.weight_hh_l0.copy_(torch.tensor(U.T))
        ref.bias_ih_l0.copy_(torch.tensor(b)); ref.bias_hh_l0.zero_()
        out, (hn, cn) = ref(torch.tensor(x))
    np.testing.assert_allclose(seq, out.numpy(), atol=1e-12)
    np.testing.assert_allclose(c, cn[0].numpy(), atol=1e-12)


def _rand_weights(init, seed=3, scale=1.0, **kw):
    w = init(seed=seed, **kw)
    rng = np.random.default_rng(seed + 100)
    return {k: (v.astype(np.float64) * scale + rng.normal(size=v.shape) * 0.05) for k, v in w.items()}


@pytest.mark.parametrize("tf", [True, False])
def test_m1_numpy_vs_torch(tf):
    rng = np.random.default_rng(1)
    w = _rand_weights(kn.init_fov_seq2seq)
    enc = rng.uniform(-1, 1, (4, 10, 90)); dec = rng.uniform(-1, 1, (4, 10 if tf else 1, 6))
    a = kn.fov_seq2seq_forward(w, enc, dec, teacher_forcing=tf)
    b = kt.fov_seq2seq_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(dec), teacher_forcing=tf)
    np.testing.assert_allclose(a, b.numpy(), atol=1e-10)


def test_m3_numpy_vs_torch():
    rng = np.random.default_rng(2)
    w = _rand_weights(kn.init_others_lstm_span_whole, num_user=6)
    enc = rng.uniform(-1, 1, (3, 10, 6)); oth = rng.uniform(-1, 1, (3, 20, 1, 5, 6)); dec = rng.uniform(-1, 1, (3, 1, 6))
    a = kn.others_lstm_span_whole_forward(w, enc, oth, dec)
    b = kt.others_lstm_span_whole_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(oth), torch.tensor(dec))
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y.numpy(), atol=1e-10)
    assert a[0].shape == (3, 10, 6) and a[1].shape == (3, 20, 30) and a[2].shape == (3, 10, 6)


@pytest.mark.parametrize("kind", ["conv2d", "conv1d", "dense"])
def test_m4_numpy_vs_torch(kind):
    rng = np.random.default_rng(4)
    if kind == "conv2d":
        w = _rand_weights(kn.init_convlstm_seq2seq, in_ch=5, filters=(4, 3, 2), kernel_size=3, head=(6, 7, 5))
        enc = rng.uniform(0, 1, (2, 3, 6, 4, 5)); dec = rng.uniform(0, 1, (2, 1, 6, 4, 5))
    elif kind == "conv1d":
        w = _rand_weights(kn.init_convlstm_seq2seq, in_ch=3, filters=(4, 3, 2), kernel_size=3, head=(6, 7, 3), head_kind="conv1d")
        enc = rng.uniform(-1, 1, (2, 3, 1, 8, 3)); dec = rng.uniform(-1, 1, (2, 1, 1, 8, 3))
    else:
        w = _rand_weights(kn.init_convlstm_seq2seq, in_ch=6, filters=(4, 3, 2), kernel_size=3, head_kind="dense", flat_dim=9)
        enc = rng.uniform(-1, 1, (2, 3, 1, 1, 6)); dec = rng.uniform(-1, 1, (2, 1, 1, 1, 6))
    a = kn.convlstm_seq2seq_forward(w, enc, dec, head_kind=kind, steps=3)
    b = kt.convlstm_seq2seq_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(dec), head_kind=kind, steps=3)
    np.testing.assert_allclose(a, b.numpy(), atol=1e-10)


def test_convlstm_dilation_and_dropout_masks_numpy_vs_torch():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 3, 5, 6, 3)); K = rng.normal(size=(3, 3, 3, 8)) * .3
    R = rng.normal(size=(3, 3, 2, 8)) * .3; b = rng.normal(size=8) * .1
    masks = [(rng.uniform(size=(2, 5, 6, 3)) > 0.3) / 0.7 for _ in range(4)]
    a, _, _ = kn.convlstm2d(x, K, R, b, dilation=(2, 2), dropout_masks=masks)
    t = lambda v: torch.tensor(v)
    bb, _, _ = kt.convlstm2d(t(x), t(K), t(R), t(b), dilation=(2, 2), dropout_masks=[t(m) for m in masks])
    np.testing.assert_allclose(a, bb.numpy(), atol=1e-10)


def test_losses_numpy_vs_torch():
    rng = np.random.default_rng(6)
    yt, yp = rng.uniform(-1, 1, (3, 10, 90)), rng.uniform(-1, 1, (3, 10, 6))
    np.testing.assert_allclose(kn.gauss_nll(yt, yp), kt.gauss_nll(torch.tensor(yt), torch.tensor(yp)).item(), rtol=1e-12)
    p = kn.softmax(rng.normal(size=(2, 4, 3, 5))); t1 = np.eye(5)[rng.integers(0, 5, (2, 4, 3))]
    np.testing.assert_allclose(kn.categorical_crossentropy(t1, p), kt.categorical_crossentropy(torch.tensor(t1), torch.tensor(p)).item(), rtol=1e-12)


def test_optimisers_numpy_vs_torch():
    rng = np.random.default_rng(7)
    p0 = rng.normal(size=9); w = {"p": torch.tensor(p0.copy())}
    adam = kt.KerasAdam(w); p, m, v = p0.copy(), np.zeros(9), np.zeros(9)
    for t in range(1, 4):
        g = rng.normal(size=9)
        p, m, v = kn.adam_step(p, g, m, v, t)
        adam.step({"p": torch.tensor(g)})
    np.testing.assert_allclose(p, w["p"].numpy(), atol=1e-14)
    w = {"p": torch.tensor(p0.copy())}; rms = kt.KerasRMSprop(w); p, a = p0.copy(), np.zeros(9)
    for t in range(3):
        g = rng.normal(size=9)
        p, a = kn.rmsprop_step(p, g, a); rms.step({"p": torch.tensor(g)})
    np.testing.assert_allclose(p, w["p"].numpy(), atol=1e-14)


# ---------------- finite differences pin the autograd gradients ------------ #

def test_m3_gradients_finite_difference():
    rng = np.random.default_rng(8)
    w = _rand_weights(kn.init_others_lstm_span_whole, num_user=4, scale=0.5)
    enc = rng.uniform(-1, 1, (2, 10, 6)); oth = rng.uniform(-1, 1, (2, 20, 1, 3, 6)); dec = rng.uniform(-1, 1, (2, 1, 6))
    tgt = [rng.uniform(-1, 1, (2, 10, 6)), rng.uniform(-1, 1, (2, 20, 18)), rng.uniform(-1, 1, (2, 10, 6))]

    def total(wd):
        o = kn.others_lstm_span_whole_forward(wd, enc, oth, dec)
        return sum(kn.mse(t, y) for t, y in zip(tgt, o))

    wt = kt.to_torch(w)
    _, _, grads = kt.loss_and_grads(kt.others_lstm_span_whole_forward, wt,
                                    [torch.tensor(enc), torch.tensor(oth), torch.tensor(dec)],
                                    [torch.tensor(t) for t in tgt], [kt.mse] * 3)
    eps = 1e-6
    for name in ["oth_convlstm0/recurrent_kernel", "oth_convlstm1/kernel", "decoder/recurrent_kernel",
                 "decoder_dense/kernel", "oth_flat_dense/kernel", "encoder/kernel", "encoder_dense/bias"]:
        flat_idx = rng.integers(0, w[name].size, 3)
        for fi in flat_idx:
            idx = np.unravel_index(fi, w[name].shape)
            wp = {k: v.copy() for k, v in w.items()}; wp[name][idx] += eps
            wm = {k: v.copy() for k, v in w.items()}; wm[name][idx] -= eps
            fd = (total(wp) - total(wm)) / (2 * eps)
            assert abs(fd - grads[name][idx].item()) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, grads[name][idx].item())


def test_philox_restatement_matches_random123_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: counter, key -> output):
    the pin for oracle.keras_numpy.philox4x32_10, which the CUDA stream fov_philox_normal must equal bit for bit."""
    w = kn.philox4x32_10(np.array([0], np.uint64), 0)[0]
    assert [int(v) for v in w] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    words, z = kn.philox_normal(200000, seed=99)
    assert words.dtype == np.uint32 and abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    # offset addressing: block k of the stream is counter k
    w2, _ = kn.philox_normal(8, seed=99, offset=5)
    assert np.array_equal(w2, words[20:28])


def test_resample_modes_and_torch_twin_agree():
    rng = np.random.default_rng(0)
    muvar = rng.uniform(0.05, 1, (7, 6)); noise = rng.normal(size=(7, 30, 3))
    muvar[2, 4] = -0.3
    for mode in ("sqrt_floor", "var_as_std"):
        a = kn.gaussian_resample(muvar[:, :3], muvar[:, 3:], noise, mode)
        b = kt.gaussian_resample(torch.tensor(muvar), torch.tensor(noise), mode).numpy()
        np.testing.assert_allclose(a, b, atol=1e-12)
    assert np.isnan(kn.gaussian_resample(muvar[:, :3], muvar[:, 3:], noise, "sqrt")[2, :, 1]).all()   # sqrt(negative), as written
    # M4 with sampling: NumPy and torch restatements agree
    w = kn.init_convlstm_seq2seq(seed=5, in_ch=3, filters=(4, 2, 2), kernel_size=3, head=None, head_kind="dense", flat_dim=8 * 30)
    enc = rng.uniform(-1, 1, (2, 3, 1, 30, 3)); dec = enc[:, -1:]
    nz = rng.normal(size=(3, 2, 30, 3))
    a = kn.convlstm_seq2seq_forward({k: v.astype(np.float64) for k, v in w.items()}, enc, dec, head_kind="dense", steps=3, noise=nz)
    b = kt.convlstm_seq2seq_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(dec), head_kind="dense", steps=3,
                                    noise=torch.tensor(nz)).numpy()
    np.testing.assert_allclose(a, b, atol=1e-10)


def test_get_data_restatement_matches_reference_golden():
    """oracle get_data (target / others split, duplicate padding, truncation, short-video skip) against the outputs of
    the reference's own get_data (tests/golden/make_get_data_golden.py), bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_get_data_golden.npz"))
    datadb = {k: {c: g["in_%s_%s" % (k, c)] for c in "xyz"} for k in ("v0", "v1", "v2")}
    for a, name in zip(kn.get_data(datadb, pick_user=False), ("all_past", "all_fut", "all_futin")):
        assert np.array_equal(a, g[name])
    for num_user in (4, 6):
        dups = list(g["u%d_dups" % num_user])
        out = kn.get_data(datadb, pick_user=True, num_user=num_user, draw=lambda n: int(dups.pop(0)))
        assert not dups
        for a, name in zip(out, ("tar_past", "tar_fut", "tar_futin", "oth_past", "oth_fut", "oth_futin")):
            assert np.array_equal(a, g["u%d_%s" % (num_user, name)]), name


def test_stacked_seq2seq_restatements_agree():
    """2- / 3-layer fc-LSTM stacks (Fov_seq2seq_2layers.py:232-272, 3layers.py:223-275): NumPy and torch restatements."""
    rng = np.random.default_rng(1)
    for L in (2, 3):
        w = kn.init_stacked_fov_seq2seq(seed=2, n_layers=L)
        enc = rng.uniform(-1, 1, (4, 10, 6)); dec = rng.uniform(-1, 1, (4, 10, 6))
        a = kn.stacked_fov_seq2seq_forward({k: v.astype(np.float64) for k, v in w.items()}, enc, dec, L)
        b = kt.stacked_fov_seq2seq_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(dec), L).numpy()
        np.testing.assert_allclose(a, b, atol=1e-10)
        assert a.shape == (4, 10, 6)
    # 3-layer graph: decoder layer 3 re-uses decoder1's weights (0-based), decoder2 is unused
    w3 = kn.init_stacked_fov_seq2seq(seed=2, n_layers=3)
    w3b = dict(w3); w3b["decoder2/kernel"] = w3["decoder2/kernel"] * 0
    x64 = {k: v.astype(np.float64) for k, v in w3.items()}; y64 = {k: v.astype(np.float64) for k, v in w3b.items()}
    assert np.array_equal(kn.stacked_fov_seq2seq_forward(x64, enc, dec, 3), kn.stacked_fov_seq2seq_forward(y64, enc, dec, 3))


def test_given_others_and_convlstm_target_restatements_agree():
    """given_others_gt_mean_var_seq2seq.py:97-308 (three variants, re-fed and teacher forced) and the all-ConvLSTM
    form of others_LSTM_span_whole.py (:133-199,273-317): NumPy and torch restatements agree; known structure checks
    (target_only ignores the others; the re-fed decoder really consumes its own output)."""
    rng = np.random.default_rng(3)
    enc = rng.uniform(-1, 1, (3, 10, 6)); oth = rng.uniform(-1, 1, (3, 10, 33, 6))
    for variant in ("mlp_mixing", "others_mlp", "target_only", "others_lstm", "conv_mixing"):
        w = kn.init_given_others_seq2seq(seed=4, variant=variant)
        w64 = {k: v.astype(np.float64) for k, v in w.items()}
        for tf in (False, True):
            dec = rng.uniform(-1, 1, (3, 10 if tf else 1, 6))
            a = kn.given_others_seq2seq_forward(w64, enc, oth, dec, variant, tf)
            b = kt.given_others_seq2seq_forward(kt.to_torch(w), torch.tensor(enc), torch.tensor(oth), torch.tensor(dec),
                                                variant, tf).numpy()
            np.testing.assert_allclose(a, b, atol=1e-10)
            assert a.shape == (3, 10, 6) and np.abs(a).max() <= 1.0          # tanh outputs
            if variant == "conv_mixing":
                assert a.min() >= 0.0                                        # the last mixing Conv2D is relu (:196-199)
            other = kn.given_others_seq2seq_forward(w64, enc, oth * 0.5, dec, variant, tf)
            assert np.array_equal(a, other) == (variant == "target_only")
        dec2 = dec[:, :1] + 0.1
        c = kn.given_others_seq2seq_forward(w64, enc, oth, dec2, variant, False)
        d = kn.given_others_seq2seq_forward(w64, enc, oth, dec[:, :1], variant, False)
        assert np.abs(c[:, -1] - d[:, -1]).max() > 0                          # the seed propagates through all 10 re-fed steps
    w = kn.init_others_convlstm_target(seed=5, num_user=5)
    x = [rng.uniform(-1, 1, (2, 10, 1, 30, 3)), rng.uniform(-1, 1, (2, 20, 1, 30, 12)), rng.uniform(-1, 1, (2, 1, 1, 30, 3))]
    a = kn.others_convlstm_target_forward({k: v.astype(np.float64) for k, v in w.items()}, *x)
    b = kt.others_convlstm_target_forward(kt.to_torch(w), *[torch.tensor(t) for t in x])
    for u, v in zip(a, b):
        np.testing.assert_allclose(u, v.numpy(), atol=1e-10)
    assert [u.shape for u in a] == [(2, 10, 1, 30, 3), (2, 20, 1, 30, 12), (2, 10, 1, 30, 3)]


def test_given_others_bilstm_structure():
    """others_lstm variant (given_others_gt_mean_var_seq2seq.py:151-158,236-241): the stacked Bidirectional LSTMs see
    the whole future of the others (step 0's output depends on the others' LAST second through the backward LSTM);
    the second pair starts from the first pair's final states (zeroing layer 0's recurrent kernels changes those
    states and so the result, even with layer 1's input sequence held fixed is not needed: checked directly against a
    hand-rolled Keras-semantics loop); the teacher-forced graph reads the decoder's first step throughout."""
    rng = np.random.default_rng(9)
    B, T, U, H = 2, 10, 5, 32
    w = {k: v.astype(np.float64) for k, v in kn.init_given_others_seq2seq(seed=2, num_user=U + 1, variant="others_lstm").items()}
    assert w["others_bilstm0_fwd/kernel"].shape == (U * 6, 4 * H) and w["others_bilstm1_bwd/kernel"].shape == (2 * H, 4 * H)
    assert w["decoder_dense/kernel"].shape == (3 * H, 6)
    enc = rng.uniform(-1, 1, (B, 10, 6)); oth = rng.uniform(-1, 1, (B, T, U, 6)); dec = rng.uniform(-1, 1, (B, 1, 6))
    a = kn.given_others_seq2seq_forward(w, enc, oth, dec, "others_lstm", False)
    oth2 = oth.copy(); oth2[:, -1] += 0.3
    b = kn.given_others_seq2seq_forward(w, enc, oth2, dec, "others_lstm", False)
    assert np.abs(a[:, 0] - b[:, 0]).max() > 1e-6
    # hand-rolled: per-step loops in Keras order, explicit reversal, explicit state hand-over
    L = lambda n: (w[n + "/kernel"], w[n + "/recurrent_kernel"], w[n + "/bias"])
    x = oth.reshape(B, T, -1)
    def run(xs, name, h, c):
        out = []
        for t in range(xs.shape[1]):
            h, c = kn.lstm_step(xs[:, t], h, c, *L(name))
            out.append(h)
        return np.stack(out, 1), h, c
    z = np.zeros((B, H))
    f0, fh, fc = run(x, "others_bilstm0_fwd", z, z)
    b0, bh, bc = run(x[:, ::-1], "others_bilstm0_bwd", z, z)
    y0 = np.concatenate([f0, b0[:, ::-1]], -1)
    f1, _, _ = run(y0, "others_bilstm1_fwd", fh, fc)
    b1, _, _ = run(y0[:, ::-1], "others_bilstm1_bwd", bh, bc)
    y1 = np.concatenate([f1, b1[:, ::-1]], -1)
    e1, h1, c1 = kn.lstm(enc, *L("encoder0")); _, h2, c2 = kn.lstm(e1, *L("encoder1"))
    xin, outs = dec[:, 0], []
    for t in range(T):
        h1, c1 = kn.lstm_step(xin, h1, c1, *L("decoder0"))
        h2, c2 = kn.lstm_step(h1, h2, c2, *L("decoder1"))
        xin = np.tanh(np.concatenate([y1[:, t], h2], -1) @ w["decoder_dense/kernel"] + w["decoder_dense/bias"])
        outs.append(xin)
    np.testing.assert_allclose(a, np.stack(outs, 1), atol=1e-12)
    # teacher forced: only the decoder's first step reaches the outputs
    dec_tf = rng.uniform(-1, 1, (B, T, 6)); dec_tf2 = dec_tf.copy(); dec_tf2[:, 1:] += 0.5
    c = kn.given_others_seq2seq_forward(w, enc, oth, dec_tf, "others_lstm", True)
    d = kn.given_others_seq2seq_forward(w, enc, oth, dec_tf2, "others_lstm", True)
    assert np.array_equal(c, d)


def test_host_utils_match_reference_golden():
    """data.rand_sample_ind / rand_sample (mycode/utility.py:575-591) and clip_xyz (mycode/dataIO.py:16-26) against
    outputs of the reference's own functions (tests/golden/make_host_utils_golden.py), same ``random`` seeds."""
    import random
    from longterm360fov_b200 import data
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_host_utils_golden.npz"))
    for i, (tot, ntest, bs, vr) in enumerate(g["cases"]):
        tot, ntest, bs = int(tot), int(ntest), int(bs)
        random.seed(100 + i)
        ind = data.rand_sample_ind(tot, ntest, bs, validation_ratio=float(vr))
        assert np.array_equal(np.array(ind), g["ind%d" % i])
        assert np.array_equal(np.array(data.rand_sample_ind(tot, ntest, bs, float(vr), rng=random.Random(100 + i))), g["ind%d" % i])
        n_val = int((tot - ntest) / bs * float(vr)) * bs
        assert n_val % bs == 0 and len(set(ind)) == len(ind) and max(ind) < tot - ntest   # whole validation batches
        src = np.random.default_rng(i).standard_normal((tot - ntest, 3)).astype(np.float32)
        assert np.array_equal(data.rand_sample(src, ind), g["picked%d" % i])
    keys = [(k, a) for k in ("v0", 3) for a in "xyz"]
    vids = {}
    for (k, a), arr in zip(keys, g["clip_in"]):
        vids.setdefault(k, {})[a] = arr.copy()
    out = data.clip_xyz(vids)
    assert out is vids
    for (k, a), want in zip(keys, g["clip_out"]):
        assert np.array_equal(out[k][a], want)


def test_heatmap_batches_per_video_structure():
    """kn.heatmap_batches_per_video (data_generator_for_heatmap.py:17-101 on the target viewer's tensors): sample
    (viewer u, window i) holds the one-hot maps of seconds i*stride .. +9 (past) and +10 .. +19 (future), viewers
    outermost; the decoder input is the last observed second; every frame-channel of every map holds exactly one 1."""
    from longterm360fov_b200 import data
    U, S, stride = 3, 27, 5
    xyz = data.synth_trajectories(1, n_viewers=U, seconds=S, seed=11)[0].reshape(U, S, 30, 3).astype(np.float64)
    enc, dec, tgt = kn.heatmap_batches_per_video(xyz, stride=stride)
    n = (S - 20) // stride + 1
    assert enc.shape == (U * n, 10, 36, 18, 30) and dec.shape == (U * n, 1, 36, 18, 30) and tgt.shape == enc.shape
    heat = kn.one_hot_heatmaps(xyz)
    for u in range(U):
        for i in range(n):
            assert np.array_equal(enc[u * n + i], heat[u, i * stride:i * stride + 10])
            assert np.array_equal(tgt[u * n + i], heat[u, i * stride + 10:i * stride + 20])
    assert np.array_equal(dec[:, 0], enc[:, -1])
    assert np.array_equal(enc.sum(axis=(2, 3)), np.ones((U * n, 10, 30)))


def test_conv_and_convlstm_against_scipy_correlate():
    """A third, library-backed statement of the convolutions: Keras 'same' Conv2D = cross-correlation with TF's
    low = floor(total / 2) padding, here through scipy.signal.correlate2d on explicitly padded planes,
    incl. an even kernel and a dilated one; then one ConvLSTM2D step assembled from those planes."""
    from scipy.signal import correlate2d
    rng = np.random.default_rng(21)

    def conv_ref(x, k, b, dil):
        B, H, W, Cin = x.shape
        kh, kw, _, Cout = k.shape
        kd = np.zeros(((kh - 1) * dil[0] + 1, (kw - 1) * dil[1] + 1, Cin, Cout))
        kd[::dil[0], ::dil[1]] = k                                            # dilation = zero-stuffed kernel
        th, tw = kd.shape[0] - 1, kd.shape[1] - 1
        xp = np.pad(x, ((0, 0), (th // 2, th - th // 2), (tw // 2, tw - tw // 2), (0, 0)))
        y = np.zeros((B, H, W, Cout))
        for n in range(B):
            for co in range(Cout):
                for ci in range(Cin):
                    y[n, :, :, co] += correlate2d(xp[n, :, :, ci], kd[:, :, ci, co], mode="valid")
        return y + b

    for (H, W, kh, kw, dil) in [(6, 5, 3, 3, (1, 1)), (5, 7, 4, 2, (1, 1)), (7, 6, 3, 3, (2, 1)), (1, 9, 1, 5, (1, 1))]:
        x = rng.standard_normal((2, H, W, 3)); k = rng.standard_normal((kh, kw, 3, 4)); b = rng.standard_normal(4)
        np.testing.assert_allclose(kn.conv2d(x, k, b, None, dil), conv_ref(x, k, b, dil), atol=1e-12)
    # one ConvLSTM2D step: z = x (*) K + b + h (*) R, gates i,f,c,o, hard sigmoid
    H, W, Cin, F = 4, 6, 3, 2
    x = rng.standard_normal((2, H, W, Cin)); h = rng.standard_normal((2, H, W, F)); c = rng.standard_normal((2, H, W, F))
    K = rng.standard_normal((3, 3, Cin, 4 * F)); R = rng.standard_normal((3, 3, F, 4 * F)); b = rng.standard_normal(4 * F)
    z = conv_ref(x, K, b, (2, 2)) + conv_ref(h, R, np.zeros(4 * F), (1, 1))      # the input kernel is dilated, R never
    hs = lambda v: np.clip(0.2 * v + 0.5, 0, 1)
    i, f, g, o = hs(z[..., :F]), hs(z[..., F:2 * F]), np.tanh(z[..., 2 * F:3 * F]), hs(z[..., 3 * F:])
    c1 = f * c + i * g
    h1 = o * np.tanh(c1)
    got_h, got_c = kn.convlstm2d_step(x, h, c, K, R, b, dilation=(2, 2))
    np.testing.assert_allclose(got_h, h1, atol=1e-12)
    np.testing.assert_allclose(got_c, c1, atol=1e-12)
