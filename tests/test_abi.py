"""CPU tests: the C-ABI library builds, loads and exports every symbol include/fov360.h declares
(no compute calls without a GPU); host-side logic of the package."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="fov360.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fov_[a-z0-9_]+)\s*\(", src)))


def _exported_symbols(path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return sorted({l.split()[-1] for l in out.splitlines() if l.split() and l.split()[-1].startswith("fov_")})


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from longterm360fov_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
        assert name in _lib.SYMBOLS, "ctypes binding missing for %s" % name
    assert set(_lib.SYMBOLS) == set(declared)
    assert _lib.load().fov_version() >= 100
    # nothing undeclared leaves the library: every exported fov_* symbol is in fov360.h (the boundary) or in
    # fov_debug.h (diagnostic switches, not on the product path)
    debug = _declared_symbols("fov_debug.h")
    assert all(n.startswith("fov_debug_") for n in debug)
    exported = _exported_symbols(_lib.LIB_PATH)
    assert set(exported) == set(declared) | set(debug), sorted(set(exported) ^ (set(declared) | set(debug)))
    # the data-parallel entry points exist and validate their arguments without a GPU or NCCL
    lib = _lib.load()
    assert lib.fov_dp_unique_id_bytes() == 128 and lib.fov_dp_world() == 1 and lib.fov_dp_rank() == 0
    assert lib.fov_dp_init(None, 0, 1) == -1 and lib.fov_dp_destroy() == 0
    assert lib.fov_dp_allreduce(ctypes.c_void_p(16), 4, None) == -1 and b"fov_dp_init" in lib.fov_last_error()


def test_struct_layouts_match_the_c_header(tmp_path):
    """sizeof/offsetof of every struct, as gcc sees include/fov360.h, equal the ctypes mirror."""
    import subprocess
    from longterm360fov_b200 import _lib
    pairs = {"fov_lstm_cfg": _lib.LstmCfg, "fov_lstm_weights": _lib.LstmWeights, "fov_lstm_saved": _lib.LstmSaved,
             "fov_lstm_io": _lib.LstmIO, "fov_lstm_grads": _lib.LstmGrads, "fov_conv_cfg": _lib.ConvCfg,
             "fov_convlstm_cfg": _lib.ConvLstmCfg, "fov_convlstm_io": _lib.ConvLstmIO,
             "fov_convlstm_grads": _lib.ConvLstmGrads}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fov360.h"', 'int main(void){']
    for cname, cls in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('return 0;}')
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_host_side_entry_points_need_no_gpu():
    """Pure host functions of the C ABI and argument validation (every error path returns before any CUDA call)."""
    from longterm360fov_b200 import _lib
    from oracle import keras_numpy as kn
    lib = _lib.load()
    for S in (20, 21, 27, 33, 600):
        for stride in (1, 2, 3, 5, 10):
            for testing in (False, True):
                vid = np.zeros((1, S, 2))
                n_ref = kn.reshape2second_stacks(vid, collapse_user=True, stride=stride, purelly_testing=testing)[0].shape[0]
                assert lib.fov_window_count(S, 10, stride, int(testing)) == n_ref, (S, stride, testing)
    assert lib.fov_window_count(19, 10, 1, 0) == 0                       # fewer than 2 x running_length seconds
    cfg = _lib.LstmCfg(64, 10, 10, 90, 6, 64, 6, 1, 1, 0, 0, 1, 2)
    assert lib.fov_lstm_bwd_ws_floats(ctypes.byref(cfg)) == (156 + 72) * 256    # [h|x|0] rows padded to 4 floats
    one = ctypes.c_void_p(16)                                            # never dereferenced: validation fails first
    assert lib.fov_window_stacks(3, 19, 6, 10, 1, 0, 1, one, one, one, one, None) == -1
    assert b"running_length" in lib.fov_last_error()
    assert lib.fov_window_stacks(3, 30, 6, 10, 20, 0, 1, one, one, one, one, None) == -1   # stride > running_length
    assert lib.fov_onehot_heatmaps(4, 30, 7, one, one, None) == -1       # 7 does not divide 180
    assert lib.fov_hit_rate(0, one, one, 1.0, 1.0, 1.0, 1.0, one, None) == -1
    assert lib.fov_whole_span(0, 8, one, one, None) == -1


def test_models_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import longterm360fov_b200 as fov
    with pytest.raises(fov._lib.FovError):
        fov.fov_seq2seq()
    from longterm360fov_b200 import ops
    with pytest.raises(fov._lib.FovError):
        ops.mean_var_xyz(torch.zeros(2, 90))


def test_initialisers_match_oracle_shapes_and_keras_rules():
    from longterm360fov_b200 import initializers as ini
    from oracle import keras_numpy as kn
    for a, b in [(ini.init_fov_seq2seq(), kn.init_fov_seq2seq()),
                 (ini.init_others_lstm_span_whole(), kn.init_others_lstm_span_whole()),
                 (ini.init_convlstm_seq2seq(head=(8, 8, 30)), kn.init_convlstm_seq2seq(head=(8, 8, 30)))]:
        assert {k: v.shape for k, v in a.items()} == {k: v.shape for k, v in b.items()}
    w = ini.init_others_lstm_span_whole()
    assert sum(v.size for v in w.values()) == 921858            # SURVEY.md 8a' M3 parameter count
    assert sum(v.size for v in ini.init_fov_seq2seq().values()) == 58246
    b = w["encoder/bias"]
    assert np.all(b[64:128] == 1) and b.sum() == 64              # unit_forget_bias
    U = w["encoder/recurrent_kernel"]
    np.testing.assert_allclose(U @ U.T, np.eye(64), atol=1e-5)   # orthogonal rows
    w4 = ini.init_convlstm_seq2seq()
    assert sum(v.size for v in w4.values()) == 15182814          # SURVEY.md 8a' M4 parameter count


def test_synthetic_data_shapes_and_featuriser_consistency():
    from longterm360fov_b200 import data
    from oracle import keras_numpy as kn
    x, y = data.make_m3_batch(4, 34, seed=1)
    assert [a.shape for a in x] == [(4, 10, 6), (4, 20, 1, 33, 6), (4, 1, 6)]
    assert [a.shape for a in y] == [(4, 10, 6), (4, 20, 198), (4, 10, 6)]
    enc, dec_in, tgt, raw = data.make_m1_batch(4, seed=1)
    # data_sanity_check of the reference (others_LSTM_span_whole.py:380-386): the first decoder
    # input second is the last encoder second
    np.testing.assert_allclose(dec_in[:, 0], kn.get_gt_target_xyz(enc[:, -1:].astype(np.float64))[:, 0], atol=1e-6)
    np.testing.assert_allclose(tgt, kn.get_gt_target_xyz(raw.astype(np.float64)), atol=1e-6)
    xh, yh = data.make_m4_batch(2, seed=1)
    assert xh[0].shape == (2, 10, 36, 18, 30) and np.all(xh[0].sum(axis=(2, 3)) == 1)


def test_callbacks_host_logic():
    from longterm360fov_b200.callbacks import EarlyStopping, ReduceLROnPlateau

    class Opt: lr = 1e-3
    class M: optimizer = Opt(); stop_training = False
    m = M()
    r = ReduceLROnPlateau(factor=0.2, patience=3, min_lr=1e-6); r.set_model(m); r.on_train_begin()
    e = EarlyStopping(patience=7); e.set_model(m); e.on_train_begin()
    for ep, v in enumerate([1.0, 0.9, 0.95, 0.95, 0.95, 0.95, 0.95, 0.95]):
        r.on_epoch_end(ep, {"val_loss": v}); e.on_epoch_end(ep, {"val_loss": v})
    assert abs(m.optimizer.lr - 4e-5) < 1e-12 and m.stop_training is False
    e.on_epoch_end(8, {"val_loss": 0.95})
    assert m.stop_training is True


def test_flags_shim_maps_the_reference_cfg_onto_builder_arguments():
    """flags.cfg carries mycode/config.py's defaults of the flags the path reads; builder_kwargs() yields arguments
    every builder's signature accepts (no GPU needed: signatures only)."""
    import inspect
    import longterm360fov_b200 as fov
    from longterm360fov_b200 import flags
    c = flags.reference_defaults()
    assert (c.fps, c.running_length, c.predict_step, c.batch_size, c.data_chunk_stride) == (30, 10, 10, 32, 10)
    assert (c.input_mean_var, c.predict_mean_var, c.sample_and_refeed, c.teacher_forcing, c.use_one_hot) == \
        (False, False, True, False, False)
    for script in flags.scripts():
        name, kw = flags.builder_kwargs(script, c)
        inspect.signature(getattr(fov, name)).bind_partial(**kw)
    assert flags.builder_kwargs("FoV_seq2seq.py", c)[1]["num_encoder_tokens"] == 90
    h = c.copy()
    h.use_one_hot = True
    assert flags.builder_kwargs("convlstm_seq2seq", h)[1]["use_one_hot"] is True and c.use_one_hot is False
    t = c.copy()
    t.predict_mean_var = True                                    # raw xyz in, mean/var out: the re-sampling graph
    assert flags.builder_kwargs("convlstm_seq2seq", t)[1]["sample_and_refeed"] is True
    t.input_mean_var = True
    assert flags.builder_kwargs("convlstm_seq2seq", t)[1]["sample_and_refeed"] is False
    assert flags.builder_kwargs("FoV_seq2seq_no_teac_forc", t)[1]["num_encoder_tokens"] == 6
    import pytest
    with pytest.raises(KeyError):
        flags.builder_kwargs("lstm", c)
    with pytest.raises(AttributeError):
        c.no_such_flag
