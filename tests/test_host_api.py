"""CPU tests of the host-side mirror of the reference interface: constructor
bookkeeping, state_dict layout, error conventions, helpers.  No compute calls
(those need the GPU and fail loudly without one)."""
import pytest
import torch

import pyroved_b200 as pv
from pyroved_b200 import _lib
from golden_util import Golden


@pytest.mark.parametrize("inv, coord", [(None, 0), (['r'], 1), (['t'], 2), (['s'], 1),
                                        (['r', 's', 't'], 4)])
def test_coord_bookkeeping_2d(inv, coord):
    m = pv.models.iVAE((8, 8), 2, inv, device="cpu")
    assert m.coord == coord and m.z_dim == 2 + coord


@pytest.mark.parametrize("inv, coord", [(None, 0), (['t'], 1)])
def test_coord_bookkeeping_1d(inv, coord):
    m = pv.models.iVAE((8,), 2, inv, device="cpu")
    assert m.coord == coord


@pytest.mark.parametrize("inv", [['r'], ['s'], ['r', 't'], ['t', 'r']])
def test_1d_rejects_non_translation(inv):
    with pytest.raises(ValueError):
        pv.models.iVAE((8,), 2, inv, device="cpu")


def test_unknown_sampler_keyerror():
    with pytest.raises(KeyError):
        pv.models.iVAE((8, 8), 2, ['r'], sampler_d="poisson", device="cpu")


def test_bad_in_dim_valueerror():
    with pytest.raises(ValueError):
        pv.nets.fcEncoderNet((1, 2, 3, 4))


def test_grid_not_implemented_for_3d():
    with pytest.raises(NotImplementedError):
        pv.utils.generate_grid((2, 2, 2))


def test_to_onehot_assertion():
    with pytest.raises(AssertionError):
        pv.utils.to_onehot(torch.tensor([3]), 3)


@pytest.mark.parametrize("name", ["ivae_1d_t", "ivae_28_rt", "ivae_12_rts_cond_gauss",
                                  "ivae_12_vanilla", "ivae_16_s_softplus"])
def test_state_dict_matches_reference_layout_and_init(name):
    """Same keys, shapes AND initial values as the reference constructor
    (same seed, same nn.Linear construction order)."""
    g = Golden(name)
    kw = dict(g.kwargs)
    seeds = {"ivae_12_rts_cond_gauss": 2, "ivae_12_vanilla": 3, "ivae_16_s_softplus": 4}
    m = pv.models.iVAE(seed=seeds.get(name, 1), device="cpu", **kw)
    ref = g.group("w0")
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
        assert torch.equal(sd[k], ref[k]), k


def test_grid_matches_reference_formula():
    g = pv.utils.generate_grid((3, 4))
    assert g.shape == (12, 2)
    assert torch.allclose(g[0], torch.tensor([-1., 1.]))
    assert torch.allclose(g[-1], torch.tensor([1., -1.]))
    assert torch.allclose(g[5], torch.tensor([0., 1. - 2. / 3.]))
    g1 = pv.utils.generate_grid((5,))
    assert g1.shape == (5, 1) and g1[0, 0] == 1 and g1[-1, 0] == -1


def test_split_latent_shapes():
    m = pv.models.iVAE((8, 8), 2, ['r', 't', 's'], device="cpu")
    phi, dx, sc, z = m.split_latent(torch.randn(5, 6))
    assert phi.shape == (5,) and dx.shape == (5, 2) and sc.shape == (5,) and z.shape == (5, 2)
    m1 = pv.models.iVAE((8,), 2, ['t'], device="cpu")
    phi, dx, sc, z = m1.split_latent(torch.randn(5, 3))
    assert phi is None and sc is None and dx.shape == (5, 1) and z.shape == (5, 2)


def test_latent_grid_helpers():
    z, (gx, gy) = pv.utils.generate_latent_grid(4)
    assert z.shape == (16, 2)
    assert torch.allclose(z[0], torch.stack([gx[0], gy[0]]))
    assert torch.allclose(z[5], torch.stack([gx[1], gy[1]]))
    c, d = pv.utils.generate_latent_grid_traversal(4, 2, 3, 0, 0, 16)
    assert c.shape == (16, 2) and d.shape == (16, 3)
    assert torch.all(d.sum(1) == 1)


def test_no_cpu_fallback():
    """Compute entry points refuse CPU tensors / CPU devices loudly."""
    m = pv.models.iVAE((8, 8), 2, ['r'], device="cpu")
    with pytest.raises(RuntimeError):
        pv.trainers.SVItrainer(m, device="cpu")
    with pytest.raises(_lib.PvbError):
        m.encoder_z(torch.zeros(2, 64))


def test_dataloader_protocol():
    x = torch.randn(10, 4)
    y = torch.randn(10, 2)
    l1 = pv.utils.init_dataloader(x, batch_size=4, shuffle=False)
    assert [len(b) for b in l1] == [1, 1, 1] and len(l1.dataset) == 10
    l2 = pv.utils.init_dataloader(x, y, batch_size=5)
    assert all(len(b) == 2 for b in l2)


def test_ved_state_dict_keys_and_seeded_init_match_reference():
    """VED(seed) builds the reference's parameter names, shapes AND initial values."""
    seeds = {"ved_spec2im_32_16": 2, "ved_bn_im2spec_16_32": 3, "ved_bn_spec2im_32_16": 4}
    for name in ("ved_im2spec_32_64", "ved_spec2im_32_16", "ved_bn_im2spec_16_32",
                 "ved_bn_spec2im_32_16"):
        g = Golden(name)
        m = pv.models.VED(seed=seeds.get(name, 1), device="cpu", **g.kwargs)
        w0 = g.group("w0")
        sd = m.state_dict()
        assert list(sd.keys()) == list(w0.keys())
        for k in sd:
            assert torch.equal(sd[k].float(), w0[k]), k
    # reference tests/test_conv.py:12-18: one BatchNorm per convolution when batchnorm=True
    for hidden, bnorm, n in (([(8,)], True, 1), ([(8,)], False, 0), ([(8,), (16, 16)], True, 3)):
        fe = pv.nets.conv.FeatureExtractor(2, conv_filters=hidden, batchnorm=bnorm)
        assert len([k for k in fe.state_dict() if "running_mean" in k]) == n
    m = pv.models.VED((64, 64), (128,), device="cpu")
    keys = list(m.state_dict().keys())
    for i in (0, 3, 5, 8, 10):
        assert "encoder_z.feature_extractor.layers.{}.weight".format(i) in keys
    for i in (4, 9, 12):
        assert "decoder.upsampler.layers.{}.conv.weight".format(i) in keys
    assert m.state_dict()["encoder_z.features2latent.fc_latent.weight"].shape == (4, 32768)
    assert m.state_dict()["decoder.latent2features.fc.weight"].shape == (2048, 2)
    assert sum(p.numel() for p in m.parameters()) == 577893


def test_conv_layer_plan_batchnorm_and_volumetric():
    """Host-side description of the conv nets (no kernels): activation fused into the preceding
    convolution, batch norm / pool / upsample as their own steps, shapes for 1-D / 2-D / 3-D data
    (reference tests/test_conv.py: feature extractor / upsampler in every dimensionality)."""
    from pyroved_b200.nets.conv import FeatureExtractor, Upsampler, layer_plan, out_shape
    fe = FeatureExtractor(2, 1, [(8,), (16, 16)], batchnorm=True, activation="lrelu", pool_last=False)
    plan = layer_plan(fe.layers, "lrelu")
    assert [k for k, _, _ in plan] == ["conv", "bn", "pool", "conv", "bn", "conv", "bn"]
    assert all(a == "lrelu" for k, _, a in plan if k == "conv")
    shape = (1, 12, 10)
    for kind, mod, _ in plan:
        shape = out_shape(kind, mod, shape)
    assert shape == (16, 6, 5)
    for ndim, size in ((1, (8,)), (2, (8, 8)), (3, (8, 8, 8))):
        fe = FeatureExtractor(ndim, 1, [(8, 8)], pool_last=True)
        shape = (1, *size)
        for kind, mod, _ in layer_plan(fe.layers, "lrelu"):
            shape = out_shape(kind, mod, shape)
        assert shape == (8, *[s // 2 for s in size])
        up = Upsampler(ndim, 8, [(8,), (4,)], output_channels=2)
        kinds = [k for k, _, _ in layer_plan(up.layers, "lrelu")]
        assert kinds == ["conv", "up", "conv", "conv", "up", "conv", "conv"]
        shape = (8, *size)
        for kind, mod, _ in layer_plan(up.layers, "lrelu"):
            shape = out_shape(kind, mod, shape)
        assert shape == (2, *[4 * s for s in size])
        blocks = [mod for kind, mod, _ in layer_plan(up.layers, "lrelu") if kind == "up"]
        assert all(b.mode == ("bilinear" if ndim == 2 else "nearest") for b in blocks)
    with pytest.raises(NotImplementedError):
        FeatureExtractor(2, 1, [(8,)], kernel_size=5, padding=2)


def test_set_encoder_set_decoder_accept_package_nets_and_reject_foreign_modules():
    """reference models/base.py:173-181: the plug-in hooks.  The step is built from the nets'
    structure, so package nets are accepted (conv encoder for iVAE included), foreign modules and
    nets of the wrong width are refused with a clear error (no silent eager fallback)."""
    from pyroved_b200.nets import convEncoderNet, fcDecoderNet, fcEncoderNet, sDecoderNet
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], device="cpu")
    m.set_encoder(convEncoderNet((28, 28), latent_dim=m.z_dim))
    assert any(k.startswith("encoder_z.feature_extractor.layers.0") for k in m.state_dict())
    m.set_decoder(sDecoderNet((28, 28), 2, 0, [128, 128], "tanh"))
    with pytest.raises(TypeError):
        m.set_encoder(torch.nn.Linear(784, 10))
    with pytest.raises(ValueError):
        m.set_encoder(convEncoderNet((28, 28), latent_dim=2))      # must cover the transform latents
    with pytest.raises(TypeError):
        m.set_decoder(fcDecoderNet((28, 28), 2))                   # spatial model, non-spatial decoder
    m2 = pv.models.iVAE((28, 28), 2, ['r', 't'], device="cpu")
    m2.set_encoder(fcEncoderNet((28, 28), 5, 0, [64, 64, 64], "relu"))
    with pytest.raises(ValueError):
        m2.set_encoder(fcEncoderNet((28, 28), 4, 0))
    with pytest.raises(TypeError):
        pv.models.jiVAE((28, 28), 2, 3, ['r'], device="cpu").set_encoder(
            convEncoderNet((28, 28), latent_dim=3))
