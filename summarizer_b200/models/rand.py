"""Uniform-random scores baseline (models/rand.py); no parameters, no compute."""
import torch
import torch.nn as nn

from . import Trainer


class Random(nn.Module):
    def forward(self, x):
        """x: (seq_len, batch_size, input_size) -> uniform scores (seq_len, batch_size, 1)"""
        seq_len, batch_size, _ = x.shape
        return torch.rand((seq_len, batch_size, 1)).to(x.device)


class RandomTrainer(Trainer):
    def _init_model(self):
        return Random()

    def train(self, fold):
        return self._train_supervised(fold, optimizer_params=[])
