"""Load golden fixtures written by oracle/make_golden.py (outputs of the
unmodified reference run under oracle/pyro_min)."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    # name: (model kind, constructor kwargs)
    "ivae_1d_t": ("ivae", dict(data_dim=(64,), latent_dim=2, invariances=["t"])),
    "ivae_28_rt": ("ivae", dict(data_dim=(28, 28), latent_dim=2, invariances=["r", "t"])),
    "ivae_28_rt_beta3": ("ivae", dict(data_dim=(28, 28), latent_dim=2, invariances=["r", "t"])),
    "ivae_12_rts_cond_gauss": ("ivae", dict(
        data_dim=(12, 12), latent_dim=3, invariances=["r", "t", "s"], c_dim=3,
        activation="relu", sampler_d="gaussian", sigmoid_d=False,
        sc_prior=0.2, dx_prior=0.15, dy_prior=0.05)),
    "ivae_12_vanilla": ("ivae", dict(data_dim=(12, 12), latent_dim=2, invariances=None)),
    "ivae_16_s_softplus": ("ivae", dict(
        data_dim=(16, 16), latent_dim=2, invariances=["s"], activation="softplus",
        hidden_dim_e=[64, 32], hidden_dim_d=[64, 64, 64])),
    "ivae_12_r_cbern": ("ivae", dict(data_dim=(12, 12), latent_dim=2, invariances=["r"],
                                     sampler_d="continuous_bernoulli")),
    "jivae_28_r": ("jivae", dict(data_dim=(28, 28), latent_dim=2, discrete_dim=3,
                                 invariances=["r"])),
    # beta[1] = 1: the enumerated-site weighting does not depend on how poutine.scale interacts
    # with the Dice weights
    "jivae_28_r_beta1": ("jivae", dict(data_dim=(28, 28), latent_dim=2, discrete_dim=3,
                                       invariances=["r"])),
    "ssivae_16_r_unsup": ("ssivae", dict(data_dim=(16, 16), latent_dim=2, num_classes=4,
                                         invariances=["r"])),
    "ssivae_16_r_sup": ("ssivae", dict(data_dim=(16, 16), latent_dim=2, num_classes=4,
                                       invariances=["r"])),
    "ssreg_16_rt_unsup": ("ssreg", dict(data_dim=(16, 16), latent_dim=2, reg_dim=2,
                                        invariances=["r", "t"])),
    "ssreg_16_rt_sup": ("ssreg", dict(data_dim=(16, 16), latent_dim=2, reg_dim=2,
                                      invariances=["r", "t"])),
    # VED (reduced channel counts keep the fixtures small)
    "ved_im2spec_32_64": ("ved", dict(
        input_dim=(32, 32), output_dim=(64,), latent_dim=2,
        hidden_dim_e=[(8,), (16, 16), (32, 32)], hidden_dim_d=[(32, 32), (16, 16), (8,)])),
    "ved_spec2im_32_16": ("ved", dict(
        input_dim=(32,), output_dim=(16, 16), latent_dim=3, activation="tanh",
        sampler_d="gaussian", sigmoid_d=False, decoder_sig=0.3,
        hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])),
    # volumetric input and output (Conv3d nets)
    "ved_vol_8": ("ved", dict(
        input_dim=(8, 8, 8), output_dim=(8, 8, 8), latent_dim=2,
        hidden_dim_e=[(4,), (8, 8)], hidden_dim_d=[(8, 8), (4,)])),
    # batchnorm=True: BatchNorm2d encoder + BatchNorm1d decoder, and the other way round
    "ved_bn_im2spec_16_32": ("ved", dict(
        input_dim=(16, 16), output_dim=(32,), latent_dim=2, batchnorm=True,
        hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])),
    "ved_bn_spec2im_32_16": ("ved", dict(
        input_dim=(32,), output_dim=(16, 16), latent_dim=2, batchnorm=True, activation="tanh",
        hidden_dim_e=[(8,), (16, 16)], hidden_dim_d=[(16, 16), (8,)])),
}


class Golden:
    def __init__(self, name):
        self.name = name
        self.kind, self.kwargs = CASES[name]
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.raw = {k: z[k] for k in z.files}
        self.loss = float(self.raw["loss"])
        self.loss_step = float(self.raw["loss_step"])

    def t(self, key, dtype=torch.float32):
        return torch.from_numpy(self.raw[key]).to(dtype)

    def has(self, key):
        return key in self.raw

    def group(self, prefix, dtype=torch.float32):
        out = OrderedDict()
        for k, v in self.raw.items():
            if k.startswith(prefix + "."):
                out[k[len(prefix) + 1:]] = torch.from_numpy(v).to(dtype)
        return out

    def args(self, dtype=torch.float32):
        x = self.t("arg0", dtype)
        y = self.t("arg1", dtype) if self.has("arg1") else None
        return x, y

    def eps(self, dtype=torch.float32):
        e = self.group("eps", dtype)
        if len(e) > 1:          # several reparameterised sites (ss_reg_iVAE: "y" and "z")
            return dict(e)
        return next(iter(e.values()))

    def kw(self):
        return {k[3:]: v for k, v in self.raw.items() if k.startswith("kw.")}
