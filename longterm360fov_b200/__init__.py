"""longterm360fov_b200 - B200-native (sm_100a) implementation of the LongTerm360FoV
sequence-prediction hot path: Keras-shaped model builders over hand-written CUDA
kernels behind a C ABI (include/fov360.h).  No CPU fallback: importing the models
without libfov360.so raises."""
from . import _lib, h5lite
from .h5lite import load_h5, save2hdf5
from .callbacks import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau
from .pipeline import M3VideoBatches, M4HeatmapBatches
from .models import (Adam, ConvLSTMSeq2Seq, FovSeq2Seq, Model, OthersLSTMSpanWhole, RMSprop,
                     convlstm_seq2seq, fov_seq2seq, fov_seq2seq_mu_var, others_lstm_span_whole,
                     StackedFovSeq2Seq, stacked_fov_seq2seq, GivenOthersSeq2Seq, given_others_gt_mean_var_seq2seq,
                     OthersConvLSTMTarget, others_convlstm_target)

__all__ = ["fov_seq2seq", "fov_seq2seq_mu_var", "others_lstm_span_whole", "convlstm_seq2seq",
           "Model", "FovSeq2Seq", "OthersLSTMSpanWhole", "ConvLSTMSeq2Seq", "Adam", "RMSprop",
           "ModelCheckpoint", "ReduceLROnPlateau", "EarlyStopping", "M3VideoBatches", "M4HeatmapBatches", "StackedFovSeq2Seq",
           "stacked_fov_seq2seq", "GivenOthersSeq2Seq", "given_others_gt_mean_var_seq2seq",
           "OthersConvLSTMTarget", "others_convlstm_target", "h5lite", "load_h5", "save2hdf5"]
