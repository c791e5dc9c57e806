"""The drop-in boundary beyond the train step (VERDICT r01 "missing" 1-3): `hippie.backbones` (ResNet18Enc / ResNet18Dec as
modules of their own, the reference's own `test_decoder`), `MultiModalCVAE.encode` / `.decode`, the class-embedding
surgery of the stage-3 script, the device tables behind `EphysBatchLoader` -- each against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import cvae_oracle as O
import parity_util as U

pytestmark = pytest.mark.gpu
ATOL = 1e-4  # north_star: embeddings within 1e-4 absolute


def _backbone_state(cfg, prefix, seed=42):
    """The oracle's state of one backbone of the full model, with the module prefix stripped (= the stand-alone
    module's state_dict keys)."""
    st = U.perturbed_state(cfg, seed=seed)
    return {k[len(prefix) + 1:]: v for k, v in st.items() if k.startswith(prefix + ".")}, st


def test_reference_backbone_test_passes():
    """hippie/backbones.py:156-165 of the reference (`test_decoder`), run on the engine through the alias package."""
    from hippie.backbones import test_decoder
    test_decoder("cuda")


@pytest.mark.parametrize("L", [50, 100])
def test_resnet18enc_module_matches_oracle(L):
    from hippie.backbones import BasicBlockEnc, ResNet18Enc
    cfg = O.CVAEConfig(z_dim=10)
    prefix = "encoder_mod1" if L == 50 else "encoder_mod2"
    sd, full = _backbone_state(cfg, prefix)
    torch.manual_seed(0)
    enc = ResNet18Enc(z_dim=10, input_size=L, max_batch=64)
    assert list(enc.state_dict().keys()) == list(sd.keys())  # the reference's keys, in its order
    assert isinstance(enc.layer2[0], BasicBlockEnc) and enc.layer2[0].conv1.weight.shape == (128, 64, 3)
    enc.load_state_dict(sd)
    enc.to("cuda:0")
    x = torch.randn(40, 1, L, generator=torch.Generator().manual_seed(3))
    s64 = U.to_dtype(full, torch.float64)
    for train in (False, True):
        enc.train(train)
        got = enc(x)
        cx = O._Ctx(s64, train)
        ref = O._encoder(cx, prefix, x.double())
        assert got.shape == (40, 20)
        assert (got.cpu().double() - ref).abs().max().item() <= ATOL
        if train:  # running statistics moved like nn.BatchNorm1d's
            new = enc.state_dict()
            for k, v in cx.new_buffers.items():
                if k.startswith(prefix + ".") and k.endswith("running_var"):
                    assert (new[k[len(prefix) + 1:]].cpu().double() - v).abs().max().item() <= 1e-5
    # another input length / a larger batch rebuild the engine and keep the parameters (the reference module takes any)
    y = enc.eval()(torch.randn(70, 1, L))
    assert y.shape == (70, 20) and torch.isfinite(y).all()


@pytest.mark.parametrize("L", [50, 100])
def test_resnet18dec_module_matches_oracle(L):
    from hippie.backbones import BasicBlockDec, ResizeConv1d, ResNet18Dec
    cfg = O.CVAEConfig(z_dim=10)
    prefix = "decoder_mod1" if L == 50 else "decoder_mod2"
    sd, full = _backbone_state(cfg, prefix)
    dec = ResNet18Dec(output_size=L, z_dim=10, max_batch=64)
    assert list(dec.state_dict().keys()) == list(sd.keys())
    assert isinstance(dec.layer4[1], BasicBlockDec) and isinstance(dec.layer4[1].conv1, ResizeConv1d)
    dec.load_state_dict(sd)
    dec.to("cuda:0")
    d = torch.randn(24, 20, generator=torch.Generator().manual_seed(4))
    s64 = U.to_dtype(full, torch.float64)
    for train in (False, True):
        dec.train(train)
        got = dec(d)
        ref = O._decoder(O._Ctx(s64, train), prefix, d.double())
        assert got.shape == (24, 1, L)
        assert (got.cpu().double() - ref).abs().max().item() <= ATOL * max(1.0, ref.abs().max().item())
    with pytest.raises(RuntimeError):
        dec.layer1[0](d)  # blocks are parameter containers; the layer program runs whole backbones


@pytest.mark.parametrize("multimodal", [True, False])
def test_encode_decode_match_oracle(multimodal):
    """MultiModalCVAE.encode / .decode (reference hippie/model.py:402-422) and the unimodal pair (:50-61): embedding
    rows are inputs."""
    from hippie_b200 import model as M
    cfg = O.CVAEConfig(z_dim=10, num_classes=4) if multimodal else \
        O.CVAEConfig(z_dim=10, num_classes=4, multimodal=False, output_size_wave=50)
    st = U.perturbed_state(cfg)
    if multimodal:
        m = M.MultiModalCVAE(10, 50, 100, 5, cfg.num_sources, 4, max_batch=32)
    else:
        m = M.hippieUnimodalCVAE(10, 50, 5, cfg.num_sources, 4, max_batch=32)
    m.load_state_dict(st)
    m.to("cuda:0").eval()
    B = 20
    x1, x2, labels, eps = U.case_inputs(cfg, B, True, seed=8)
    cls, src = labels.unbind(1)
    s64 = U.to_dtype(st, torch.float64)
    semb, cemb = st["source_embedding.weight"][src], st["class_embedding.weight"][cls]
    with torch.no_grad():
        o64, _, _ = O.forward(s64, cfg, x1.double(), x2.double() if multimodal else None, src, cls, eps.double(), train=False)
    h, mu, lv = m.encode(x1, x2, semb, cemb) if multimodal else m.encode(x1, semb, cemb)
    for got, k in ((h, "enc"), (mu, "mu"), (lv, "logvar")):
        assert (got.cpu().double() - o64[k]).abs().max().item() <= ATOL, k
    z = (o64["mu"] + eps.double() * torch.exp(0.5 * o64["logvar"])).float()
    if multimodal:
        r1, r2 = m.decode(z, semb, cemb)
        assert r1.shape == (B, 1, 50) and r2.shape == (B, 1, 100)
        assert (r2.cpu().double() - o64["dec2"]).abs().max().item() <= ATOL
    else:
        r1 = m.decode(z, semb, cemb)
    assert (r1.cpu().double() - o64["dec1"]).abs().max().item() <= ATOL


def test_class_embedding_surgery_is_honoured():
    """`model.class_embedding = nn.Embedding(k, 5)` (reference scripts/train_model_with_multimodal.py:378-379): the new
    table is the one the engine reads and trains, everything else is carried over."""
    from hippie_b200 import model as M
    torch.manual_seed(42)
    m = M.MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=32).to("cuda:0")
    before = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    torch.manual_seed(7)
    emb = torch.nn.Embedding(7, 5)
    m.class_embedding = emb
    assert m.num_classes == 7 and m.engine.num_classes == 7
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert sd["class_embedding.weight"].shape == (7, 5)
    assert torch.equal(sd["class_embedding.weight"], emb.weight.detach())
    for k, v in before.items():
        if k != "class_embedding.weight":
            assert torch.equal(sd[k], v), k
    assert isinstance(m.class_embedding, torch.nn.Module) and m.class_embedding.weight.shape == (7, 5)
    # forward with the classes only the new table has, against the oracle on the same state
    cfg = O.CVAEConfig(z_dim=10, num_classes=7)
    x1, x2, labels, eps = U.case_inputs(cfg, 16, True, seed=3)
    labels[:, 0] = torch.arange(16) % 7
    cls, src = labels.unbind(1)
    m.eval()
    out = m(x1, x2, src, cls, eps=eps.cuda())
    with torch.no_grad():
        o64, _, _ = O.forward(U.to_dtype(sd, torch.float64), cfg, x1.double(), x2.double(), src, cls, eps.double(), train=False)
    assert (out[1].cpu().double() - o64["mu"]).abs().max().item() <= ATOL
    # a supervised step trains the new table
    tm = M.MultiModalCVAETrainModule(m, learning_rate=1e-3, weight_decay=0.01, beta=0.5)
    m.train()
    tm.training_step((x1, x2, labels), 0, eps=eps.cuda())
    tm.optimizer.step(max_norm=1.0)
    after = m.state_dict()["class_embedding.weight"].cpu()
    assert not torch.equal(after, emb.weight.detach())
    with pytest.raises(ValueError):
        m.class_embedding = torch.nn.Embedding(7, 3)  # wrong hidden size


def test_batch_loader_over_device_tables_equals_host_tables():
    """`EphysBatchLoader` over `DeviceTable`s (raw float64 tables resident on the GPU, one preprocessing launch per batch)
    yields the batches of the host-side `EphysTensorDataset` path: same index stream, waveforms bit for bit, ISI <= 2 ulp."""
    from hippie_b200.dataloading import DeviceTable, EphysBatchLoader, EphysTensorDataset
    rng = np.random.default_rng(5)
    tabs = [(rng.normal(size=(37, 47)), np.abs(rng.normal(size=(37, 100))), rng.integers(0, 4, 37)),
            (rng.normal(size=(20, 60)), np.abs(rng.normal(size=(20, 100))), rng.integers(0, 4, 20))]
    host = EphysTensorDataset.concat([EphysTensorDataset(w, i, l) for w, i, l in tabs])
    dev = DeviceTable.concat([DeviceTable(w, i, l) for w, i, l in tabs])
    assert len(host) == len(dev) == 57
    torch.manual_seed(123)
    hb = list(EphysBatchLoader(host, 16, shuffle=True))
    torch.manual_seed(123)
    db = list(EphysBatchLoader(dev, 16, shuffle=True))
    assert len(hb) == len(db) == 4
    for (h1, h2, hl), (d1, d2, dl) in zip(hb, db):
        assert d1.is_cuda and torch.equal(d1.cpu(), h1) and torch.equal(dl.cpu(), hl)
        assert (d2.cpu().view(torch.int32) - h2.view(torch.int32)).abs().max().item() <= 2
