"""The reference's scripts read one global ``cfg`` (mycode/config.py, an EasyDict mutated at import) while they build
their graphs; the builders of this package take explicit keyword arguments instead.  This module is the bridge a
script keeps its ``cfg.`` reads on: ``cfg`` holds the reference's defaults of the flags that reach the hot path (and
only those), ``builder_kwargs(script, cfg)`` turns them into the keyword arguments of the matching builder, and
``build(script, cfg, **overrides)`` calls it.  Host-side only.

    from longterm360fov_b200.flags import cfg, build
    cfg.use_one_hot = True                      # what the heatmap runs of convlstm_seq2seq.py set
    model = build("convlstm_seq2seq", cfg).compile("RMSprop", "mean_squared_error")
"""
from __future__ import annotations


class Flags(dict):
    """dict with attribute access (the EasyDict surface the scripts use: ``cfg.fps``, ``cfg.fps = 30``)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def copy(self):
        return Flags(self)


def reference_defaults():
    """Defaults of mycode/config.py for the flags the hot path reads (line numbers of that file)."""
    f = Flags()
    f.update(
        batch_size=32, fps=30,                                                     # :12-13
        predict_len=10, running_length=10, predict_step=10,                       # :19-22 (process_in_seconds)
        shuffle_data=False, stateful_across_batch=False,                          # :60-61
        dropout_rate=0.3, conv_kernel_size=5, recurrent_dropout_rate=0.3,         # :62-64
        predict_mean_var=False, sample_and_refeed=True, input_mean_var=False,     # :65-66,68
        teacher_forcing=False, use_one_hot=False,                                 # :69-70
        data_chunk_stride=10,                                                     # :79-82 (overlapping chunks, seconds)
        dilation_rate=1, use_saliency=False,                                      # :93-94
        cut_data_head=False, purelly_testing=False, time_shift=False,             # :98-99,104
        target_user_only=False,                                                   # :110
    )
    return f


cfg = reference_defaults()

# script (module name under mycode/) -> (builder name in longterm360fov_b200, kwargs from the flags)
_SCRIPTS = {
    # latent_dim = 64, num_encoder_tokens = 3 * fps, teacher forced (FoV_seq2seq.py:19-28,82-103)
    "FoV_seq2seq": ("fov_seq2seq", lambda c: dict(
        num_encoder_tokens=3 * c.fps, max_encoder_seq_length=c.running_length, max_decoder_seq_length=c.predict_step,
        teacher_forcing=True)),
    # mean / var in and out (FoV_seq2seq_mu_var.py:40-49,219-248)
    "FoV_seq2seq_mu_var": ("fov_seq2seq_mu_var", lambda c: dict(
        max_encoder_seq_length=c.running_length, max_decoder_seq_length=c.predict_step, teacher_forcing=True)),
    # in-graph autoregressive decoder (FoV_seq2seq_no_teac_forc.py:37-149)
    "FoV_seq2seq_no_teac_forc": ("fov_seq2seq", lambda c: dict(
        num_encoder_tokens=6 if c.input_mean_var else 3 * c.fps, max_encoder_seq_length=c.running_length,
        max_decoder_seq_length=c.predict_step, teacher_forcing=False)),
    # concat-state model (others_LSTM_span_whole.py:77-353; SURVEY.md hazard 2: mean / var in and out)
    "others_LSTM_span_whole": ("others_lstm_span_whole", lambda c: dict(
        kernel_size=c.conv_kernel_size, max_encoder_seq_length=c.running_length, max_decoder_seq_length=c.predict_step,
        dropout=c.dropout_rate)),
    # ConvLSTM encoder-decoder, heatmap or trajectory form (convlstm_seq2seq.py:73-287)
    "convlstm_seq2seq": ("convlstm_seq2seq", lambda c: dict(
        kernel_size=c.conv_kernel_size, dilation_rate=c.dilation_rate, use_one_hot=c.use_one_hot,
        input_mean_var=c.input_mean_var, predict_mean_var=c.predict_mean_var, max_decoder_seq_length=c.predict_step,
        fps=c.fps, dropout=c.dropout_rate,
        sample_and_refeed=bool(c.sample_and_refeed and c.predict_mean_var and not c.input_mean_var
                               and not c.use_one_hot))),
    # 2-layer decoder given the others' mean / var (given_others_gt_mean_var_seq2seq.py:97-308)
    "given_others_gt_mean_var_seq2seq": ("given_others_gt_mean_var_seq2seq", lambda c: dict(
        teacher_forcing=c.teacher_forcing, target_user_only=c.target_user_only)),
    "Fov_seq2seq_2layers": ("stacked_fov_seq2seq", lambda c: dict(n_layers=2)),    # :232-272
    "3layers": ("stacked_fov_seq2seq", lambda c: dict(n_layers=3)),                # :223-275
}


def scripts():
    return sorted(_SCRIPTS)


def builder_kwargs(script, flags=None):
    """(builder name, kwargs) for ``mycode/<script>.py`` under ``flags`` (default: the module's ``cfg``)."""
    if script.endswith(".py"):
        script = script[:-3]
    if script not in _SCRIPTS:
        raise KeyError("no builder for %r; known scripts: %s" % (script, ", ".join(scripts())))
    name, fn = _SCRIPTS[script]
    c = flags if flags is not None else cfg
    missing = [k for k in reference_defaults() if k not in c]
    if missing:
        raise KeyError("flags lack %s" % missing)
    return name, fn(c)


def build(script, flags=None, **overrides):
    """The model ``mycode/<script>.py`` builds under ``flags``; ``overrides`` go to the builder as they are
    (``num_user=``, ``weights=``, ``device=`` ...)."""
    import longterm360fov_b200 as fov
    name, kw = builder_kwargs(script, flags)
    kw.update(overrides)
    return getattr(fov, name)(**kw)
