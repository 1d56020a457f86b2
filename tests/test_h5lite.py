"""h5lite (HDF5 subset reader / writer for Keras checkpoints and the reference's save2hdf5 / load_h5 caches,
mycode/utility.py:868-880, mycode/FoV_seq2seq.py:108): CPU-only host logic.

Pin: ``tests/golden/libhdf5_testdouble_7.4_GLNX86.mat`` is a file libhdf5 itself wrote (MATLAB 7.4's v7.3 MAT
format = HDF5 behind a 512-byte user block; the copy is SciPy's BSD-licensed test datum
``scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat``, the only HDF5 file in this image).  The reader must decode it,
and the writer's messages for the same content must be byte-identical to libhdf5's.
"""
import os
import struct

import numpy as np
import pytest

from longterm360fov_b200 import h5lite as h5

GOLD = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_testdouble_7.4_GLNX86.mat")


def test_reads_the_libhdf5_written_file():
    f = h5.File(GOLD)
    assert f.base == 512 and f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
    np.testing.assert_array_equal(np.array(d).ravel(), np.arange(9) * (np.pi / 4))
    assert d.attrs["MATLAB_class"] == b"double"
    assert f.get("absent") is None and "testdouble" in f and "absent" not in f


def test_writer_messages_equal_libhdf5_bytes(tmp_path):
    """Same content written by h5lite: dataspace / datatype / attribute / symbol-table messages, the local heap's
    name segment, the B-tree node header + keys and the symbol-node entry must equal what libhdf5 wrote."""
    ref = h5.File(GOLD)
    want = np.array(ref["testdouble"])
    p = str(tmp_path / "w.h5")
    with h5.Writer(p) as w:
        d = w.create_dataset("testdouble", data=want)
        d.attrs["MATLAB_class"] = np.bytes_(b"double")
    got = h5.File(p)
    rn, gn = ref["testdouble"]._node, got["testdouble"]._node
    for mtype in (0x01, 0x03):
        assert [b.rstrip(b"\0") for b in gn.find(mtype)] == [b.rstrip(b"\0") for b in rn.find(mtype)], hex(mtype)
    # attribute message: identical incl. padding, but for the string-padding nibble of its datatype (MATLAB wrote
    # H5T_STR_NULLTERM = 0, h5lite writes H5T_STR_NULLPAD = 1 as h5py does for NumPy 'S' data)
    want_attr = bytearray(rn.find(0x0C)[0])
    off = 8 + 16                                                     # 8-byte attribute header + padded name
    assert want_attr[off] == 0x13 and want_attr[off + 1] == 0x00
    want_attr[off + 1] = 0x01
    assert gn.find(0x0C)[0] == bytes(want_attr)
    # superblock: versions, offset / length sizes, group K values
    rb, gb = bytes(ref.buf), bytes(got.buf)
    assert gb[:8] == h5.SIG and gb[8:20] == rb[512 + 8:512 + 20]
    # root symbol-table message -> heap, B-tree, symbol node
    r_bt, r_hp = struct.unpack("<QQ", ref._node.find(0x11)[0])
    g_bt, g_hp = struct.unpack("<QQ", got._node.find(0x11)[0])
    assert ref._local_heap(r_hp)[:24] == got._local_heap(g_hp)[:24]  # "" at 0, "testdouble\0" padded at 8
    r_t, g_t = rb[512 + r_bt:512 + r_bt + 48], gb[g_bt:g_bt + 48]
    assert r_t[:24] == g_t[:24]                                      # TREE, type 0, level 0, 1 entry, no siblings
    assert r_t[24:32] == g_t[24:32] and r_t[40:48] == g_t[40:48]     # keys: heap offsets 0 and 8
    r_sn = 512 + struct.unpack("<Q", r_t[32:40])[0]
    g_sn = struct.unpack("<Q", g_t[32:40])[0]
    assert rb[r_sn:r_sn + 16] == gb[g_sn:g_sn + 16]                  # SNOD v1, 1 symbol, name offset 8
    assert rb[r_sn + 24:r_sn + 48] == gb[g_sn + 24:g_sn + 48]        # cache type 0, empty scratch pad
    np.testing.assert_array_equal(np.array(got["testdouble"]), want)


def test_round_trip_groups_dtypes_chunks(tmp_path):
    rng = np.random.default_rng(0)
    p = str(tmp_path / "a.h5")
    arrays = {
        "f32": rng.standard_normal((6, 256)).astype(np.float32),
        "f64": rng.standard_normal((3, 1, 5)),
        "f16": rng.standard_normal((7,)).astype(np.float16),
        "i16": rng.integers(-5, 5, (37, 13, 5)).astype(np.int16),
        "u8": rng.integers(0, 255, (1000,)).astype(np.uint8),
        "i64": np.arange(-3, 9, dtype=np.int64),
        "empty": np.zeros((0, 4), np.float32),
        "names": np.array([b"lstm_1", b"conv_lst_m2d_12"]),
    }
    with h5.Writer(p) as w:
        w.attrs["title"] = "unicode θ"
        w.attrs["vec"] = np.arange(5, dtype=np.int32)
        for k, a in arrays.items():
            w.create_dataset("plain/" + k, data=a)
        w.create_dataset("packed/i16", data=arrays["i16"], chunks=(8, 4, 5), compression="gzip", shuffle=True)
        w.create_dataset("packed/f32", data=arrays["f32"], chunks=(1, 3))            # 6 x 86 chunks: two-level tree
        w.create_dataset("scalar", data=np.float64(3.5))
        g = w.create_group("deep/er/still")
        g.attrs["depth"] = 3
        for i in range(40):                                                          # > 2K entries of a default SNOD
            w.create_dataset("many/d%02d" % i, data=np.arange(i, dtype=np.int64))
    f = h5.File(p)
    assert sorted(f.keys()) == ["deep", "many", "packed", "plain", "scalar"]
    assert f.attrs["title"].decode("utf8") == "unicode θ"
    np.testing.assert_array_equal(f.attrs["vec"], np.arange(5))
    for k, a in arrays.items():
        got = np.array(f["plain/" + k])
        assert got.dtype == a.dtype and got.shape == a.shape, k
        np.testing.assert_array_equal(got, a)
    np.testing.assert_array_equal(np.array(f["packed/i16"]), arrays["i16"])
    np.testing.assert_array_equal(np.array(f["/packed/f32"]), arrays["f32"])
    assert f["scalar"][()] == 3.5 and f["scalar"].shape == ()
    assert f["deep/er/still"].attrs["depth"] == 3
    assert f["many"].keys() == ["d%02d" % i for i in range(40)]
    np.testing.assert_array_equal(np.array(f["many/d07"]), np.arange(7))
    assert [k for k, _ in f.visit_datasets()][:3] == ["many/d00", "many/d01", "many/d02"]


def test_reference_cache_helpers(tmp_path):
    """save2hdf5 appends keys, load_h5 returns np.array(None) for a missing key (mycode/utility.py:868-880)."""
    p = str(tmp_path / "cache.h5")
    a = np.random.default_rng(1).standard_normal((4, 10, 6)).astype(np.float32)
    b = np.arange(12).reshape(3, 4)
    h5.save2hdf5(p, "_video_db_tar", a)
    h5.save2hdf5(p, "_video_db_oth", b)
    np.testing.assert_array_equal(h5.load_h5(p, "_video_db_tar"), a)
    np.testing.assert_array_equal(h5.load_h5(p, "_video_db_oth"), b)
    assert h5.load_h5(p, "nokey").shape == () and h5.load_h5(p, "nokey")[()] is None
    with pytest.raises(h5.H5Error):
        h5.save2hdf5(p, "_video_db_tar", a)


def test_not_hdf5_and_unsupported(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(h5.H5Error):
        h5.File(str(p))
    with pytest.raises(h5.H5Error):
        with h5.Writer(str(tmp_path / "y.h5")) as w:
            w.create_dataset("c", data=np.zeros(3, np.complex64))


class _FakeModel:
    """The weight-file half of models.Model without a GPU: same methods, host arrays instead of the flat bucket."""
    from longterm360fov_b200.models import Model as _M
    _keras_layers, _match_keras_layers = _M._keras_layers, _M._match_keras_layers
    save_weights, load_weights = _M.save_weights, _M.load_weights
    weight_order = ["encoder/kernel", "encoder/recurrent_kernel", "encoder/bias",
                    "decoder/kernel", "decoder/recurrent_kernel", "decoder/bias",
                    "decoder_dense/kernel", "decoder_dense/bias"]
    shapes = [(6, 256), (64, 256), (256,), (6, 256), (64, 256), (256,), (64, 6), (6,)]

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.w = [rng.standard_normal(s).astype(np.float32) for s in self.shapes]

    def get_weights_dict(self):
        return dict(zip(self.weight_order, self.w))

    def set_weights(self, arrays):
        self.w = [np.array(a) for a in arrays]


def test_model_h5_weights_by_name(tmp_path):
    p = str(tmp_path / "fov_s2s_withTfor_epoch03-0.0123.h5")
    a, b = _FakeModel(0), _FakeModel(1)
    a.save_weights(p)
    f = h5.File(p)
    assert [s.decode() for s in f.attrs["layer_names"]] == ["encoder", "decoder", "decoder_dense"]
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.2.4"
    assert [s.decode() for s in f["decoder"].attrs["weight_names"]] == \
        ["decoder/kernel:0", "decoder/recurrent_kernel:0", "decoder/bias:0"]
    assert f["decoder/decoder/kernel:0"].shape == (6, 256)
    b.load_weights(p)
    for x, y in zip(a.w, b.w):
        np.testing.assert_array_equal(x, y)


def test_model_h5_weights_from_a_keras_named_checkpoint(tmp_path):
    """A checkpoint as Keras writes it for the reference's unnamed layers (auto names, a weightless layer in the list,
    the ``model.save`` form under /model_weights): matched by order with shape checks, or by ``layer_map``."""
    a = _FakeModel(2)
    w = a.w
    layers = [("input_1", []),
              ("lstm_1", [("lstm_1/kernel:0", w[0]), ("lstm_1/recurrent_kernel:0", w[1]), ("lstm_1/bias:0", w[2])]),
              ("lstm_2", [("lstm_2/kernel:0", w[3]), ("lstm_2/recurrent_kernel:0", w[4]), ("lstm_2/bias:0", w[5])]),
              ("dense_1", [("dense_1/kernel:0", w[6]), ("dense_1/bias:0", w[7])])]
    p = str(tmp_path / "fov_s2s_tanh.h5")
    with h5.Writer(p) as wr:                                          # model.save layout
        g = wr.create_group("model_weights")
        g.attrs["layer_names"] = [n.encode() for n, _ in layers]
        for n, ws in layers:
            lg = g.create_group(n)
            lg.attrs["weight_names"] = np.array([k.encode() for k, _ in ws]) if ws else np.zeros((0,), "S1")
            for k, arr in ws:
                lg.create_dataset(k, data=arr)
        wr.attrs["model_config"] = '{"class_name": "Model"}'
    b = _FakeModel(3)
    b.load_weights(p)
    for x, y in zip(a.w, b.w):
        np.testing.assert_array_equal(x, y)
    c = _FakeModel(4)
    c.load_weights(p, layer_map={"lstm_2": "encoder", "lstm_1": "decoder", "dense_1": "decoder_dense"})
    np.testing.assert_array_equal(c.w[0], w[3])
    np.testing.assert_array_equal(c.w[3], w[0])
    # wrong architecture: one layer short, or a shape off
    p2 = str(tmp_path / "short.h5")
    h5.write_keras_weights(p2, layers[:3])
    with pytest.raises(ValueError, match="layer_map"):
        _FakeModel(5).load_weights(p2)
    bad = list(layers)
    bad[3] = ("dense_1", [("dense_1/kernel:0", np.zeros((64, 3), np.float32)), ("dense_1/bias:0", w[7])])
    p3 = str(tmp_path / "bad.h5")
    h5.write_keras_weights(p3, bad)
    with pytest.raises(ValueError, match="shape"):
        _FakeModel(6).load_weights(p3)


def test_model_save_with_optimizer_state(tmp_path):
    """model.save('x.h5') in Keras' model.save layout (model_weights / optimizer_weights / training_config) and
    load_weights + load_optimizer_weights restore everything (host logic; torch CPU tensors stand in for the bucket)."""
    import torch
    from longterm360fov_b200.models import Model, Adam, RMSprop

    class Fake(_FakeModel):
        save, load_optimizer_weights = Model.save, Model.load_optimizer_weights
        loss_kinds, loss_weights = ["mse"], [1.0]

        def __init__(self, seed, opt):
            super().__init__(seed)
            self._offsets, off = {}, 0
            for k, shp in zip(self.weight_order, self.shapes):
                self._offsets[k] = (off, shp)
                off += (int(np.prod(shp)) + 63) // 64 * 64
            self.n_flat = off + 64
            self.optimizer = opt
            opt.init(self.n_flat, "cpu")

    a = Fake(0, Adam(lr=2e-3))
    g = torch.Generator().manual_seed(1)
    for t in a.optimizer.state:
        t.copy_(torch.rand(a.n_flat, generator=g))
    a.optimizer.iterations = 17
    p = str(tmp_path / "fov_s2s_tanh.h5")
    a.save(p)
    f = h5.File(p)
    assert sorted(f.keys()) == ["model_weights", "optimizer_weights"]
    names = [n.decode() for n in f["optimizer_weights"].attrs["weight_names"]]
    assert names[0] == "Adam/iterations:0" and names[1] == "training/Adam/encoder/kernel/m:0" and len(names) == 17
    assert f["optimizer_weights/training/Adam/decoder_dense/bias/v:0"].shape == (6,)
    import json
    cfg = json.loads(f.attrs["training_config"].decode())
    assert cfg["optimizer_config"]["class_name"] == "Adam" and abs(cfg["optimizer_config"]["config"]["lr"] - 2e-3) < 1e-12
    b = Fake(1, Adam())
    b.load_weights(p)
    b.load_optimizer_weights(p)
    assert b.optimizer.iterations == 17 and abs(b.optimizer.lr - 2e-3) < 1e-12
    for k in a.weight_order:
        off, shp = a._offsets[k]
        n = int(np.prod(shp))
        for x, y in zip(a.optimizer.state, b.optimizer.state):
            assert torch.equal(x[off:off + n], y[off:off + n])
    for x, y in zip(a.w, b.w):
        np.testing.assert_array_equal(x, y)
    with pytest.raises(ValueError, match="RMSprop"):
        Fake(2, RMSprop()).load_optimizer_weights(p)
    # an uncompiled model's save holds weights only and still loads
    c = _FakeModel(3)
    c.optimizer = None
    c.save = Model.save.__get__(c)
    p2 = str(tmp_path / "w_only.h5")
    c.save(p2)
    assert h5.File(p2).keys() == ["model_weights"] and h5.read_keras_optimizer(p2) == (None, [])
    d = _FakeModel(4)
    d.load_weights(p2)
    np.testing.assert_array_equal(d.w[0], c.w[0])


def test_round_trip_property():
    """Property test (hypothesis): any tree of groups / datasets of the supported dtypes, any shapes incl. empty and
    scalar, contiguous or chunked (+ shuffle / deflate), with attributes, reads back exactly."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st, HealthCheck
    import tempfile

    dtypes = st.sampled_from(["<f4", "<f8", "<f2", "<i8", "<i4", "<i2", "|i1", "|u1", "<u2", "<u4", ">f4", ">i4"])
    shapes = st.lists(st.integers(0, 7), min_size=0, max_size=3).map(tuple)
    names = st.text(alphabet="abcXYZ019_:.- ", min_size=1, max_size=12).filter(lambda s: s.strip("/. ") != "")

    @st.composite
    def arrays(draw):
        dt = np.dtype(draw(dtypes))
        shape = draw(shapes)
        n = int(np.prod(shape)) if shape else 1
        seed = draw(st.integers(0, 2 ** 16))
        raw = np.random.default_rng(seed).integers(-100, 100, n)
        a = raw.astype(dt.newbyteorder("=")).reshape(shape)
        chunks = None
        if shape and n and draw(st.booleans()):
            chunks = tuple(draw(st.integers(1, max(1, s))) for s in shape)
        return a, chunks, draw(st.booleans()), draw(st.booleans())

    tree = st.dictionaries(names, st.one_of(arrays(), st.dictionaries(names, arrays(), max_size=3)), max_size=5)

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(tree, st.integers(0, 99))
    def run(t, attr_seed):
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "t.h5")
            with h5.Writer(p) as w:
                w.attrs["seed"] = attr_seed
                w.attrs["names"] = np.array([k.encode() for k in t]) if t else np.zeros((0,), "S1")

                def put(g, k, v):
                    a, chunks, comp, shuf = v
                    ds = g.create_dataset(k, data=a, chunks=chunks, compression="gzip" if (comp and chunks) else None,
                                          shuffle=bool(shuf and chunks))
                    ds.attrs["shape"] = np.array(a.shape, np.int64)
                for k, v in t.items():
                    if isinstance(v, dict):
                        g = w.create_group(k)
                        for k2, v2 in v.items():
                            put(g, k2, v2)
                    else:
                        put(w, k, v)
            with h5.File(p) as f:
                assert f.attrs["seed"] == attr_seed
                assert sorted(f.keys()) == sorted(t, key=lambda s: s.encode())
                assert [s.decode() for s in np.atleast_1d(f.attrs["names"])] == list(t)

                def check(node, v):
                    a = v[0]
                    got = np.array(node)
                    assert got.shape == a.shape and got.dtype.newbyteorder("=") == a.dtype.newbyteorder("=")
                    assert np.array_equal(got, a)
                    assert tuple(np.atleast_1d(node.attrs["shape"])) == a.shape
                for k, v in t.items():
                    if isinstance(v, dict):
                        assert sorted(f[k].keys()) == sorted(v, key=lambda s: s.encode())
                        for k2, v2 in v.items():
                            check(f[k][k2], v2)
                    else:
                        check(f[k], v)
    run()


def test_describe_lists_a_checkpoint(tmp_path):
    p = str(tmp_path / "ck.h5")
    a = _FakeModel(0)
    a.save_weights(p)
    lines = h5.describe(p)
    assert "@layer_names = |S13[3]" in lines and "encoder/" in lines
    assert any(ln.strip() == "kernel:0  float32[6, 256]" for ln in lines)
    assert any("@keras_version = '2.2.4'" in ln for ln in lines)
    assert h5.describe(GOLD) == ["testdouble  float64[9, 1]", "  @MATLAB_class = 'double'"]


def test_bidirectional_layers_are_one_keras_group(tmp_path):
    """The Bi-LSTM others branch: Keras stores a Bidirectional wrapper as ONE layer with six weights (forward kernel,
    recurrent kernel, bias, then backward); the model's two halves are saved and matched that way, by name for its
    own files and by order for a Keras-named checkpoint."""
    class Bi(_FakeModel):
        weight_order = ["others_bilstm0_fwd/kernel", "others_bilstm0_fwd/recurrent_kernel", "others_bilstm0_fwd/bias",
                        "others_bilstm0_bwd/kernel", "others_bilstm0_bwd/recurrent_kernel", "others_bilstm0_bwd/bias",
                        "decoder_dense/kernel", "decoder_dense/bias"]
        shapes = [(30, 128), (32, 128), (128,), (30, 128), (32, 128), (128,), (96, 6), (6,)]

    a, b, c = Bi(0), Bi(1), Bi(2)
    p = str(tmp_path / "bi.h5")
    a.save_weights(p)
    layers = h5.read_keras_weights(p)
    assert [n for n, _ in layers] == ["others_bilstm0", "decoder_dense"] and len(layers[0][1]) == 6
    b.load_weights(p)
    for x, y in zip(a.w, b.w):
        np.testing.assert_array_equal(x, y)
    w = a.w
    keras = [("bidirectional_1", [("bidirectional_1/forward_lstm_1/kernel:0", w[0]),
                                  ("bidirectional_1/forward_lstm_1/recurrent_kernel:0", w[1]),
                                  ("bidirectional_1/forward_lstm_1/bias:0", w[2]),
                                  ("bidirectional_1/backward_lstm_1/kernel:0", w[3]),
                                  ("bidirectional_1/backward_lstm_1/recurrent_kernel:0", w[4]),
                                  ("bidirectional_1/backward_lstm_1/bias:0", w[5])]),
             ("dense_1", [("dense_1/kernel:0", w[6]), ("dense_1/bias:0", w[7])])]
    p2 = str(tmp_path / "keras_bi.h5")
    h5.write_keras_weights(p2, keras)
    c.load_weights(p2)
    for x, y in zip(a.w, c.w):
        np.testing.assert_array_equal(x, y)
