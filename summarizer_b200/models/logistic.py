"""Logistic-regression baseline (models/logistic.py): Linear(1024,1)+Sigmoid per frame.  One GEMV — memory
bound and outside the hot-path kernels — kept as a plain PyTorch module so that benchmark.py
and the model registry stay complete."""
import torch.nn as nn

from . import Trainer


class LogisticRegression(nn.Module):
    def __init__(self, input_size=1024):
        super().__init__()
        self.input_size = input_size
        self.perceptron = nn.Linear(input_size, 1)           # reference attribute names: state-dict keys perceptron.*
        self.sig = nn.Sigmoid()

    def forward(self, x):
        """x: (seq_len, batch_size, input_size) -> (seq_len, batch_size, 1)"""
        assert x.shape[2] == self.input_size, f"Input size of {self.input_size} expected, {x.shape[2]} given."
        return self.sig(self.perceptron(x))


class LogisticRegressionTrainer(Trainer):
    def _init_model(self):
        return LogisticRegression()

    def train(self, fold):
        return self._train_supervised(fold)
