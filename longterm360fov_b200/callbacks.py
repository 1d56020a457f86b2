"""The three Keras callbacks the reference scripts pass to fit()
(mycode/FoV_seq2seq.py:108-111): ModelCheckpoint, ReduceLROnPlateau, EarlyStopping."""
from __future__ import annotations

import numpy as np


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self):
        pass

    def on_epoch_end(self, epoch, logs):
        pass


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", save_best_only=False, verbose=0, **_):
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.best = np.inf

    def on_epoch_end(self, epoch, logs):
        cur = logs.get(self.monitor)
        if self.save_best_only:
            if cur is None or not cur < self.best:
                return
            self.best = cur
        path = self.filepath.format(epoch=epoch + 1, **logs)
        self.model.save_weights(path)


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor="val_loss", factor=0.1, patience=10, min_lr=0.0, min_delta=1e-4, cooldown=0, **_):
        self.monitor, self.factor, self.patience, self.min_lr = monitor, factor, patience, min_lr
        self.min_delta, self.cooldown = min_delta, cooldown

    def on_train_begin(self):
        self.best, self.wait, self.cooldown_counter = np.inf, 0, 0

    def on_epoch_end(self, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = self.model.optimizer.lr
                if old > self.min_lr:
                    self.model.optimizer.lr = max(old * self.factor, self.min_lr)
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", **_):
        self.monitor, self.min_delta, self.patience = monitor, abs(min_delta), patience

    def on_train_begin(self):
        self.best, self.wait = np.inf, 0

    def on_epoch_end(self, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.model.stop_training = True
