"""Red zones instead of compute-sanitizer (closed on this GPU pool): every buffer the autograd Functions hand to the C
ABI - outputs, saved activations, workspaces, scratch - is allocated with a poisoned guard band on both sides, whole
training steps and predictions of every model family run at awkward sizes, and afterwards every guard band must be
intact and every alignment gap of the flat parameter / gradient buckets still zero (an overrun of a gradient sink lands
in the gap or in the neighbouring parameter's gradient - the failure mode of the round-1 advisor finding)."""
import numpy as np
import pytest
import torch

from oracle import keras_numpy as kn

GUARD_BYTES = 4096


class GuardedTorch:
    """Stands in for the ``torch`` module inside longterm360fov_b200.ops: empty / zeros / *_like allocate
    [guard | body | guard] and return the body; everything else is the real module."""

    def __init__(self, guard_cpu=False):
        self._t = torch
        self.log = []
        self._guard_cpu = guard_cpu

    def __getattr__(self, name):
        return getattr(self._t, name)

    @staticmethod
    def _poison(dtype):
        if dtype.is_floating_point:
            return -12345.678
        return 0x5A if dtype in (torch.uint8, torch.int8) else 0x5A5A5A5A

    def _alloc(self, shape, dtype, device, zero):
        t = self._t
        dtype = dtype or t.float32
        device = t.device(device) if device is not None else t.device("cpu")
        if (device.type != "cuda" and not self._guard_cpu) or dtype == t.bool:
            return (t.zeros if zero else t.empty)(shape, dtype=dtype, device=device)
        n = 1
        for s in shape:
            n *= int(s)
        g = GUARD_BYTES // t.empty(0, dtype=dtype).element_size()
        flat = t.empty(n + 2 * g, dtype=dtype, device=device)
        p = self._poison(dtype)
        flat[:g] = p
        flat[g + n:] = p
        body = flat[g:g + n]
        if zero:
            body.zero_()
        self.log.append((flat, g, n, p))
        return body.view(tuple(int(s) for s in shape))

    @staticmethod
    def _shape(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(size[0])
        return tuple(size)

    def empty(self, *size, dtype=None, device=None, **kw):
        return self._alloc(self._shape(size), dtype, device, False)

    def zeros(self, *size, dtype=None, device=None, **kw):
        return self._alloc(self._shape(size), dtype, device, True)

    def empty_like(self, x, **kw):
        return self._alloc(tuple(x.shape), kw.get("dtype", x.dtype), kw.get("device", x.device), False)

    def zeros_like(self, x, **kw):
        return self._alloc(tuple(x.shape), kw.get("dtype", x.dtype), kw.get("device", x.device), True)

    def check(self):
        bad = []
        for i, (flat, g, n, p) in enumerate(self.log):
            want = torch.full((g,), p, dtype=flat.dtype, device=flat.device)
            if not torch.equal(flat[:g], want):
                bad.append((i, "head", n))
            if not torch.equal(flat[g + n:], want):
                bad.append((i, "tail", n))
        return bad


def test_guarded_allocator_itself():
    """The stand-in allocates what torch would (shape, dtype, zero fill) and notices a one-element overrun."""
    gt = GuardedTorch(guard_cpu=True)
    a = gt.empty(3, 5, device="cpu")
    b = gt.zeros((2, 7), dtype=torch.int32, device="cpu")
    c = gt.empty_like(torch.ones(4, dtype=torch.float64))
    d = gt.empty(100 + 256, dtype=torch.uint8, device="cpu")
    e = gt.zeros_like(a)
    assert a.shape == (3, 5) and a.dtype == torch.float32 and b.shape == (2, 7) and int(b.sum()) == 0
    assert c.dtype == torch.float64 and d.numel() == 356 and float(e.abs().sum()) == 0.0
    assert all(t.is_contiguous() and t.data_ptr() % 16 == 0 for t in (a, b, c, d, e))
    assert gt.check() == []
    flat, g, n, _ = gt.log[0]
    flat[g + n] = 0.0                                         # one float past the end of ``a``
    assert gt.check() == [(0, "tail", 15)]


def _gaps_clean(model):
    """Alignment gaps between the tensors of the flat parameter / gradient buckets (and the unused part of the 64
    reserved tail floats) still hold zeros."""
    spans = sorted((off, off + int(np.prod(shp))) for off, shp in model._offsets.values())
    mask = torch.ones(model.n_flat, dtype=torch.bool, device=model.device)
    for lo, hi in spans:
        mask[lo:hi] = False
    mask[-1] = False                                          # the data-parallel sample-count slot
    return bool((model.gflat[mask] == 0).all()) and bool((model.flat[mask] == 0).all())


def _cases():
    rng = np.random.default_rng(77)
    u = lambda *s: rng.uniform(-1, 1, s).astype(np.float32)
    out = []

    def m3(num_user, B):
        def build(fov):
            w = kn.init_others_lstm_span_whole(seed=3, num_user=num_user)
            m = fov.others_lstm_span_whole(num_user=num_user, weights=w).compile("Adam", ["mean_squared_error"] * 3)
            U = num_user - 1
            x = [u(B, 10, 6), u(B, 20, 1, U, 6), u(B, 1, 6)]
            y = [u(B, 10, 6), u(B, 20, U * 6), u(B, 10, 6)]
            return m, x, y
        return build
    out += [("m3_u34_b5", m3(34, 5)), ("m3_u6_b41", m3(6, 41)), ("m3_u34_b149", m3(34, 149))]

    def m1(B, tf):
        def build(fov):
            m = fov.fov_seq2seq(teacher_forcing=tf).compile("Adam", "mean_squared_error")
            return m, [u(B, 10, 90), u(B, 10 if tf else 1, 6)], [u(B, 10, 6)]
        return build
    out += [("m1_b7_tf", m1(7, True)), ("m1_b700_ar", m1(700, False)), ("m1_b6200_tf", m1(6200, True))]

    def m2(B):
        def build(fov):
            m = fov.fov_seq2seq_mu_var().compile("Adam", "mean_squared_error")
            return m, [u(B, 10, 6), u(B, 10, 6)], [u(B, 10, 6)]
        return build
    out += [("m2_b33", m2(33)), ("m2_b6150", m2(6150))]

    def m4(kind):
        def build(fov):
            from longterm360fov_b200.models import ConvLSTMSeq2Seq
            if kind == "conv2d":
                w = kn.init_convlstm_seq2seq(seed=5, in_ch=30, filters=(32, 16, 8), kernel_size=5, head=(24, 40, 30))
                x, y = [u(3, 4, 12, 6, 30), u(3, 1, 12, 6, 30)], [u(3, 3, 12, 6, 30)]
            elif kind == "conv1d":
                w = kn.init_convlstm_seq2seq(seed=5, in_ch=3, filters=(32, 16, 8), kernel_size=5, head=(24, 40, 3),
                                             head_kind="conv1d")
                x, y = [u(5, 4, 1, 30, 3), u(5, 1, 1, 30, 3)], [u(5, 3, 1, 30, 3)]
            else:
                w = kn.init_convlstm_seq2seq(seed=5, in_ch=6, filters=(4, 3, 2), kernel_size=3, head_kind="dense",
                                             flat_dim=9)
                x, y = [u(3, 4, 1, 1, 6), u(3, 1, 1, 1, 6)], [u(3, 3, 6)]
            m = ConvLSTMSeq2Seq(w, kind, max_decoder_seq_length=3).compile("RMSprop", "mean_squared_error")
            return m, x, y
        return build
    out += [("m4_" + k, m4(k)) for k in ("conv2d", "conv1d", "dense")]

    def given(variant, tf):
        def build(fov):
            m = fov.given_others_gt_mean_var_seq2seq(variant=variant, teacher_forcing=tf)
            m.compile("Adam", "mean_squared_error")
            return m, [u(37, 10, 6), u(37, 10, 33, 6), u(37, 10 if tf else 1, 6)], [u(37, 10, 6)]
        return build
    out += [("given_" + v + ("_tf" if tf else ""), given(v, tf)) for v, tf in
            (("mlp_mixing", False), ("others_lstm", False), ("conv_mixing", False), ("others_mlp", True))]

    def stacked(n, B):
        def build(fov):
            m = fov.stacked_fov_seq2seq(n_layers=n).compile("Adam", "mean_squared_error")
            return m, [u(B, 10, 6), u(B, 10, 6)], [u(B, 10, 6)]
        return build
    out += [("stacked3_b9", stacked(3, 9)), ("stacked2_b600", stacked(2, 600))]

    def allconv(fov):
        m = fov.others_convlstm_target(num_user=6).compile("Adam", ["mean_squared_error"] * 3)
        x = [u(5, 10, 1, 30, 3), u(5, 20, 1, 30, 15), u(5, 1, 1, 30, 3)]
        y = [np.asarray(o) for o in m.predict_on_batch(x)]
        return m, x, [t + 0.1 for t in y]
    out.append(("others_convlstm_target", allconv))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16x2", "fp32"])
@pytest.mark.parametrize("name,build", _cases(), ids=[c[0] for c in _cases()])
def test_no_write_outside_any_buffer(name, build, mode, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import ops
    model, x, y = build(fov)
    model.set_compute(mode)
    gt = GuardedTorch()
    monkeypatch.setattr(ops, "torch", gt)
    losses = [model.train_on_batch(x, y) for _ in range(2)]
    out = model.predict_on_batch(x)
    torch.cuda.synchronize()
    assert np.isfinite(losses).all() and all(np.isfinite(np.asarray(o)).all() for o in (out if isinstance(out, list) else [out]))
    assert len(gt.log) > 4, "the guarded allocator was not used"
    assert gt.check() == []
    assert _gaps_clean(model)


@pytest.mark.gpu
def test_red_zone_catches_a_real_overrun():
    """Negative control on the device: a kernel told to write one element more than its output buffer holds
    (fov_mse_fwd_bwd with numel + 1) trips the tail guard of exactly that buffer."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from longterm360fov_b200 import _lib, ops
    lib = _lib.load()
    n = 1000
    gt = GuardedTorch()
    y_pred = torch.rand(n + 1, device="cuda")
    y_true = torch.rand(n + 1, device="cuda")
    loss = gt.zeros(1, device="cuda")
    dy = gt.empty(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.fov_mse_fwd_bwd(n, ops.ptr(y_pred), ops.ptr(y_true), 1.0, ops.ptr(loss), ops.ptr(dy), st), "mse")
    torch.cuda.synchronize()
    assert gt.check() == []
    _lib.check(lib.fov_mse_fwd_bwd(n + 1, ops.ptr(y_pred), ops.ptr(y_true), 1.0, ops.ptr(loss), ops.ptr(dy), st), "mse")
    torch.cuda.synchronize()
    assert gt.check() == [(1, "tail", n)]


def _bench_cases():
    rng = np.random.default_rng(78)
    u = lambda *s: rng.uniform(-1, 1, s).astype(np.float32)

    def m3_bench(fov):
        from longterm360fov_b200 import data
        w = kn.init_others_lstm_span_whole(seed=3, num_user=34)
        m = fov.others_lstm_span_whole(num_user=34, weights=w).compile("Adam", ["mean_squared_error"] * 3)
        x, y = data.make_m3_batch(40, 34, seed=17)
        tile = lambda a: np.tile(a, (222,) + (1,) * (a.ndim - 1))                  # B = 8880, bench.py's per-GPU batch
        return m, "bf16x2", [tile(a) for a in x], [tile(a) for a in y]

    def m4_bench(fov):
        from longterm360fov_b200.models import ConvLSTMSeq2Seq
        w = kn.init_convlstm_seq2seq(seed=6, in_ch=30, filters=(32, 16, 8), kernel_size=5, head=(512, 1024, 30))
        m = ConvLSTMSeq2Seq(w, "conv2d", max_decoder_seq_length=3).compile("RMSprop", "mean_squared_error")
        B = 5                                                                      # 36 x 18 images, heads 512 / 1024 / 30
        return m, "bf16", [np.abs(u(B, 10, 36, 18, 30)), np.abs(u(B, 1, 36, 18, 30))], [np.abs(u(B, 3, 36, 18, 30))]

    def m2_bench(fov):
        m = fov.fov_seq2seq_mu_var(teacher_forcing=False).compile("Adam", "mean_squared_error")
        B = 37888                                                                  # config 3's training batch in bench.py
        return m, "bf16x2", [u(B, 10, 6), u(B, 1, 6)], [u(B, 10, 6)]
    return [("m3_b8880", m3_bench), ("m4_36x18_heads_512_1024", m4_bench), ("m2_b37888_ar", m2_bench)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,build", _bench_cases(), ids=[c[0] for c in _bench_cases()])
def test_no_write_outside_any_buffer_at_bench_shapes(name, build, monkeypatch):
    """The same red-zone check at the shapes bench.py times: config 2 at B = 8880 (bf16x2), config 5's full-size
    images with the 512 / 1024 heads (bf16, the TMA-fed weight gradients), config 3's model at B = 37 888."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import ops
    model, mode, x, y = build(fov)
    model.set_compute(mode)
    gt = GuardedTorch()
    monkeypatch.setattr(ops, "torch", gt)
    loss = model.train_on_batch(x, y)
    out = model.predict_on_batch(x)
    torch.cuda.synchronize()
    assert np.isfinite(loss) and all(np.isfinite(np.asarray(o)).all() for o in (out if isinstance(out, list) else [out]))
    assert len(gt.log) > 4
    assert gt.check() == []
    assert _gaps_clean(model)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16x2", "fp32"])
@pytest.mark.parametrize("name,build", _cases(), ids=[c[0] for c in _cases()])
def test_repeated_runs_agree(name, build, mode):
    """Racecheck by repetition (compute-sanitizer is closed on this pool): the forward pass of every model family has
    no atomics, so 8 repetitions on the same inputs must be BIT-identical - a shared-memory / TMEM / mbarrier race in
    a persistent kernel shows up as run-to-run differences; the gradients (weight gradients are summed with
    red.global.add, so only their rounding may differ) must agree to 1e-5 of their scale."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import longterm360fov_b200 as fov
    model, x, y = build(fov)
    model.set_compute(mode)
    as_list = lambda o: o if isinstance(o, list) else [o]
    first = [np.asarray(o).copy() for o in as_list(model.predict_on_batch(x))]
    for rep in range(7):
        again = as_list(model.predict_on_batch(x))
        for a, b in zip(first, again):
            assert np.array_equal(a, np.asarray(b)), (name, mode, rep)
    xs, ys = model._to_dev(x), model._to_dev(y)
    grads = []
    for rep in range(4):
        model.gflat.zero_()
        from longterm360fov_b200 import ops
        ops.set_math(mode)
        model._loss(model._forward(xs, True), ys).backward()
        ops.join_wgrad_stream()
        torch.cuda.synchronize()
        grads.append(model.gflat.clone())
    scale = float(grads[0].abs().max())
    assert scale > 0
    for g in grads[1:]:
        assert float((g - grads[0]).abs().max()) <= 1e-5 * scale, (name, mode)
