"""SumGAN with attention (models/sumgan_att.py:20-420) — PASSTHROUGH for the transformer parts: the selector and the
encoder-decoder autoencoder are plain ``torch.nn`` transformer modules (outside the hot path); the
discriminator is the same ``GAN`` / ``cLSTM`` as SumGAN and therefore runs on this library's recurrence kernels.
Class names, constructor arguments and attribute / state-dict names follow the reference so checkpoints interchange.
The trainer reuses SumGANTrainer's phase machinery (per-phase Adam, clip over ALL parameters, data-parallel
all-reduce) with the Wasserstein losses and the autoencoder pre-training of sumgan_att.py:171-230."""
import torch
import torch.nn as nn

from .sumgan import GAN, SumGANTrainer


class Transformer(nn.Module):
    """Selector: transformer encoder + Linear/Sigmoid per frame (sumgan_att.py:20-46)."""

    def __init__(self, input_size=1024, encoder_layers=4, attention_heads=8, epsilon=1e-5):
        super().__init__()
        self.input_size = input_size
        self.layer_norm = nn.LayerNorm(input_size, epsilon)
        self.transformer_encoder_layer = nn.TransformerEncoderLayer(d_model=input_size, nhead=attention_heads,
                                                                    dim_feedforward=input_size)
        self.transformer_encoder = nn.TransformerEncoder(self.transformer_encoder_layer, num_layers=encoder_layers,
                                                         norm=self.layer_norm)
        self.out = nn.Sequential(nn.Linear(input_size, 1), nn.Sigmoid())

    def forward(self, x):
        """x: (seq_len, batch_size, input_size) -> scores: (seq_len, batch_size, 1)"""
        return self.out(self.transformer_encoder(x))


class AutoencoderTransformer(nn.Module):
    """Encoder-decoder transformer reconstructing its (score-weighted) input (sumgan_att.py:48-80)."""

    def __init__(self, input_size=1024, encoder_layers=4, attention_heads=8, epsilon=1e-5):
        super().__init__()
        self.input_size = input_size
        self.transformer_encoder_layer = nn.TransformerEncoderLayer(d_model=input_size, nhead=attention_heads,
                                                                    dim_feedforward=input_size)
        self.transformer_encoder = nn.TransformerEncoder(self.transformer_encoder_layer, num_layers=encoder_layers)
        self.transformer_decoder_layer = nn.TransformerDecoderLayer(d_model=input_size, nhead=attention_heads,
                                                                    dim_feedforward=input_size)
        self.transformer_decoder = nn.TransformerDecoder(self.transformer_decoder_layer, num_layers=encoder_layers)

    def forward(self, x):
        """x: (seq_len, batch_size, input_size) -> x_hat of the same shape (the decoder attends to the encoded x)"""
        return self.transformer_decoder(x, self.transformer_encoder(x))


class Summarizer(nn.Module):
    def __init__(self, input_size=1024, s_encoder_layers=2, s_attention_heads=4, ae_encoder_layers=2, ae_attention_heads=4):
        """Summarizer: Selector (Transformer) + Autoencoder Transformer."""
        super().__init__()
        self.selector = Transformer(input_size=input_size, encoder_layers=s_encoder_layers, attention_heads=s_attention_heads)
        self.ae = AutoencoderTransformer(input_size=input_size, encoder_layers=ae_encoder_layers,
                                         attention_heads=ae_attention_heads)

    def forward(self, x, uniform=False, p=0.3):
        """x: (seq_len, B, input_size) -> (x_hat (seq_len, B, input_size), scores (seq_len, B, 1))"""
        if uniform:
            scores = torch.rand((x.shape[0], x.shape[1], 1)).to(x.device)
        else:
            scores = self.selector(x)
        return self.ae(x * scores), scores


class SumGANAtt(nn.Module):
    def __init__(self, input_size=1024, s_encoder_layers=2, s_attention_heads=4, ae_encoder_layers=2, ae_attention_heads=4,
                 cLSTM_hidden_size=1024, cLSTM_num_layers=2):
        """SumGAN: Summarizer + GAN"""
        super().__init__()
        self.summarizer = Summarizer(input_size=input_size, s_encoder_layers=s_encoder_layers,
                                     s_attention_heads=s_attention_heads, ae_encoder_layers=ae_encoder_layers,
                                     ae_attention_heads=ae_attention_heads)
        self.gan = GAN(input_size=input_size, hidden_size=cLSTM_hidden_size, num_layers=cLSTM_num_layers)

    def forward(self, x):
        """x: (seq_len, B, input_size) -> scores: (seq_len, B, 1)"""
        return self.summarizer.selector(x)


class SumGANAttTrainer(SumGANTrainer):
    def _init_model(self):
        ep = self.hps.extra_params or {}
        self.input_size = int(ep.get("input_size", 1024))
        self.s_encoder_layers = int(ep.get("s_encoder_layers", 2))
        self.s_attention_heads = int(ep.get("s_attention_heads", 4))
        self.ae_encoder_layers = int(ep.get("ae_encoder_layers", 2))
        self.ae_attention_heads = int(ep.get("ae_attention_heads", 4))
        self.cLSTM_hidden_size = int(ep.get("cLSTM_hidden_size", 256))
        self.cLSTM_num_layers = int(ep.get("cLSTM_num_layers", 2))
        self.sup = bool(ep.get("sup", True))
        self.pretrain_ae = int(ep.get("pretrain_ae", 80))
        self.epoch_noise = int(ep.get("epoch_noise", 0.2 * self.hps.epochs))
        model = SumGANAtt(input_size=self.input_size, s_encoder_layers=self.s_encoder_layers,
                          s_attention_heads=self.s_attention_heads, ae_encoder_layers=self.ae_encoder_layers,
                          ae_attention_heads=self.ae_attention_heads, cLSTM_hidden_size=self.cLSTM_hidden_size,
                          cLSTM_num_layers=self.cLSTM_num_layers)
        self.log.debug("Generator params: {}".format(sum([_.numel() for _ in model.summarizer.parameters()])))
        self.log.debug("Discriminator params: {}".format(sum([_.numel() for _ in model.gan.parameters()])))
        return model

    # ---- losses (sumgan_att.py:171-193) ----------------------------------------------------------------------
    def loss_ae(self, x, x_hat):
        """minimize E[l2_norm(x - x_hat)]"""
        return torch.norm(x - x_hat, p=2)

    def loss_sparsity(self, scores, sigma=None):
        """upstream leaves the unsupervised sparsity term as a constant 0"""
        return torch.tensor(0)

    def loss_gan_generator(self, probs_fake, probs_uniform):
        """maximize 0.5 * (cLSTM(x_hat) + cLSTM(x_hat_p))"""
        return torch.mean(-0.5 * (probs_fake + probs_uniform))

    def loss_gan_discriminator(self, probs_real, probs_fake, probs_uniform):
        """maximize cLSTM(x) - 0.5 * (cLSTM(x_hat) + cLSTM(x_hat_p))"""
        return torch.mean(-probs_real + 0.5 * (probs_fake + probs_uniform))

    # ---- phases ------------------------------------------------------------------------------------------------
    def _pretrain_epochs(self):
        return self.pretrain_ae

    def pretrain(self, fold):
        """Pretrain autoencoder before learning the GAN (sumgan_att.py:195-230; learning rate x 10)"""
        train_keys, _ = self._get_train_test_keys(fold)
        ae = self.model.summarizer.ae
        opt = torch.optim.Adam(ae.parameters(), lr=self.hps.lr * 10.0, weight_decay=self.hps.weight_decay)
        dp, rank, world = self._dp()
        for epoch in range(self.pretrain_ae):
            losses = []
            for key, n_active in self._groups(train_keys, dp, rank, world):
                loss = None
                if key is not None:
                    x, _ = self._video_tensors(key)
                    loss = self.loss_ae(x, ae(x))
                    losses.append(loss.detach())
                self._update(opt, loss, dp, n_active)
            if epoch % 10 == 0 or epoch == self.pretrain_ae - 1:
                avg = float(torch.stack(losses).mean()) if losses else float("nan")
                self.log.info(f"Pretrain: {epoch+1:3}/{self.pretrain_ae:3}   Lae: {avg:.05f}")

    def _make_optimizers(self):
        s = self.model.summarizer
        self.s_e_optimizer = self._adam(list(s.selector.parameters()) + list(s.ae.transformer_encoder.parameters())
                                        + list(s.ae.transformer_encoder_layer.parameters()))
        self.d_optimizer = self._adam(list(s.ae.transformer_decoder.parameters())
                                      + list(s.ae.transformer_decoder_layer.parameters()))
        self.c_optimizer = self._adam(self.model.gan.c_lstm.parameters())

    def train_step(self, x, y, epoch, dp=None, n_active=1):
        """The three updates of one video (sumgan_att.py:289-351); None on an idle data-parallel replica."""
        m, idle = self.model, x is None
        loss_s_e = loss_d = loss_c = None
        if not idle:                                             # selector + encoder
            x_hat, scores = m.summarizer(x)
            _, h = m.gan(torch.cat([x, x_hat], 1))
            sparsity = self.loss_sparsity_sup(scores, y) if self.sup else self.loss_sparsity(scores)
            loss_s_e = self.loss_recons(h[0:1], h[1:2]) + sparsity
        self._update(self.s_e_optimizer, loss_s_e, dp, n_active)
        if not idle:                                             # decoder
            x_hat, _ = m.summarizer(x)
            x_hat_p, _ = m.summarizer(x, uniform=True)
            probs, h = m.gan(torch.cat([x, x_hat, x_hat_p], 1))
            loss_d = self.loss_recons(h[0:1], h[1:2]) + self.loss_gan_generator(probs[1:2], probs[2:3])
        self._update(self.d_optimizer, loss_d, dp, n_active)
        if not idle:                                             # discriminator
            x_hat, scores = m.summarizer(x)
            x_hat_p, _ = m.summarizer(x, uniform=True)
            x_in = x
            if epoch < self.epoch_noise:
                x_in = torch.randn_like(x) * x
                x_hat = x_hat * torch.randn_like(x_hat)
                x_hat_p = x_hat_p * torch.randn_like(x_hat_p)
            probs, _ = m.gan(torch.cat([x_in, x_hat, x_hat_p], 1))
            probs_real, probs_fake, probs_uniform = probs[0:1], probs[1:2], probs[2:3]
            loss_c = self.loss_gan_discriminator(probs_real, probs_fake, probs_uniform)
        self._update(self.c_optimizer, loss_c, dp, n_active)
        if idle:
            return None
        return dict(Lse=loss_s_e.detach(), Ld=loss_d.detach(), Lc=loss_c.detach(), D_x=probs_real.detach().mean(),
                    D_x_hat=probs_fake.detach().mean(), D_x_hat_p=probs_uniform.detach().mean(), scores=scores.detach())


if __name__ == "__main__":
    model = SumGANAtt()
    print("Trainable parameters in model:", sum(p.numel() for p in model.parameters() if p.requires_grad))
