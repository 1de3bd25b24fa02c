"""Transformer-encoder scorer (models/transformer.py:19-96) — PASSTHROUGH: plain torch.nn modules, no kernels of this
library (SURVEY.md §2.1 keeps the multi-head encoder variants outside the hot path; §8b lists the class surface as
part of the drop-in).  Same constructor arguments, attribute / state-dict names (``transformer_encoder_layer``,
``transformer_encoder``, the ONE ``layer_norm`` shared by the encoder's final norm and the regressor head, ``k1``,
``k2``, optional ``pos_embed``) and forward contract (T, B, 1024) -> (T, B, 1), so reference checkpoints load.
The trainer is the shared supervised loop (transformer.py:113-190 is the VASNet loop with another module)."""
import math

import torch
import torch.nn as nn

from . import Trainer


def sincos_table(max_length, width):
    """The fixed table of transformer.py:34-38: column i (even) sin(pos / 10000^(2i/width)), column i+1
    cos(pos / 10000^(2(i+1)/width)) — evaluated in float64 and rounded once, like the reference's numpy scalars."""
    pos = torch.arange(max_length, dtype=torch.float64).unsqueeze(1)
    i = torch.arange(0, width, 2, dtype=torch.float64).unsqueeze(0)
    table = torch.zeros(max_length, width)
    table[:, 0::2] = torch.sin(pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * i / width)).float()
    table[:, 1::2] = torch.cos(pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * (i + 1) / width)).float()
    return table


class Transformer(nn.Module):
    def __init__(self, input_size=1024, encoder_layers=6, attention_heads=8, more_residuals=False, max_length=None,
                 pos_embed="simple", epsilon=1e-5, weight_init=None):
        super().__init__()
        self.input_size = input_size
        self.max_length = max_length
        if self.max_length:
            self.pos_embed_type = pos_embed
            if pos_embed == "simple":
                self.pos_embed = nn.Embedding(self.max_length, input_size)
            elif pos_embed == "attention":
                self.pos_embed = sincos_table(self.max_length, input_size)     # plain tensor, not a buffer (as upstream)
            else:
                self.max_length = None
        self.more_residuals = more_residuals
        self.dropout = nn.Dropout(0.5)
        self.layer_norm = nn.LayerNorm(input_size, epsilon)
        self.transformer_encoder_layer = nn.TransformerEncoderLayer(d_model=input_size, nhead=attention_heads,
                                                                    dim_feedforward=input_size, dropout=0.1, activation="relu")
        self.transformer_encoder = nn.TransformerEncoder(self.transformer_encoder_layer, num_layers=encoder_layers,
                                                         norm=self.layer_norm)
        self.k1 = nn.Linear(input_size, input_size)
        self.k2 = nn.Linear(input_size, 1)
        self.sigmoid = nn.Sigmoid()
        self.relu = nn.ReLU()
        init = {"he": nn.init.kaiming_uniform_, "kaiming": nn.init.kaiming_uniform_,
                "xavier": nn.init.xavier_uniform_}.get(weight_init.lower()) if weight_init else None
        if init is not None:                                  # transformer.py:57-70: feed-forward + head matrices only
            for layer in self.transformer_encoder.layers:
                init(layer.linear1.weight)
                init(layer.linear2.weight)
            init(self.k1.weight)
            init(self.k2.weight)

    def forward(self, x):
        """x: (seq_len, batch_size, input_size) -> y: (seq_len, batch_size, 1)"""
        seq_len, batch_size, _ = x.shape
        if self.max_length is not None:
            assert self.max_length >= seq_len, "input sequence has higher length than max_length"
            if self.pos_embed_type == "simple":
                pe = self.pos_embed(torch.arange(seq_len, device=x.device))
            else:
                pe = self.pos_embed[:seq_len].to(x.device)
            x += pe.unsqueeze(1)                              # in place on the caller's tensor, as upstream (:84,:86)
        encoder_out = self.transformer_encoder(x)
        if self.more_residuals:
            encoder_out = encoder_out + x
        y = self.layer_norm(self.dropout(self.relu(self.k1(encoder_out))))
        return self.sigmoid(self.k2(y))


class TransformerTrainer(Trainer):
    def _init_model(self):
        ep = self.hps.extra_params or {}
        return Transformer(encoder_layers=int(ep.get("encoder_layers", 6)), attention_heads=int(ep.get("attention_heads", 8)),
                           more_residuals=ep.get("more_residuals", False),
                           max_length=int(ep["max_pos"]) if "max_pos" in ep else None,
                           pos_embed=ep.get("pos_embed", "simple"), epsilon=float(ep.get("epsilon", 1e-5)),
                           weight_init=ep.get("weight_init", None))

    def train(self, fold):
        return self._train_supervised(fold)


if __name__ == "__main__":
    model = Transformer()
    print("Trainable parameters in model:", sum(p.numel() for p in model.parameters() if p.requires_grad))
    y = model(torch.randn(10, 3, 1024))
    assert y.shape == (10, 3, 1)
    _ = math.pi
