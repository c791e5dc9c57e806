"""CPU-side tests: the C-ABI library loads and exports what include/hippie_b200.h declares, the layout and the
nn.Module mirror agree with the oracle (== the reference's construction order and initial values), host logic."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
from collections import OrderedDict

from oracle import cvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from hippie_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "hippie_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hippie_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes binding covers the whole header
    assert _lib.lib().hippie_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("cfg", [O.CVAEConfig(z_dim=10), O.CVAEConfig(z_dim=32, num_classes=4),
                                 O.CVAEConfig(z_dim=64), O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50),
                                 O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100)])
def test_layout_matches_reference_construction_order(cfg):
    from hippie_b200.engine import Engine, LAYOUT_CONV_OKI
    e = Engine(cfg.z_dim, cfg.output_size_wave, cfg.output_size_isi, cfg.class_hidden_dim, cfg.num_sources,
               cfg.num_classes, cfg.multimodal, 64)
    spec = {n: (k, s) for n, k, s in O.model_spec(cfg)}
    assert [p.name for p in e.params] == O.param_names(cfg)
    end = 0
    for p in e.params:
        kind, shp = spec[p.name]
        assert p.shape == tuple(O._torch_shape(kind, shp))
        assert p.offset % 4 == 0 and p.offset >= end
        assert (p.layout == LAYOUT_CONV_OKI) == (kind == "conv_w")
        end = p.offset + p.numel
    assert e.param_floats >= end and e.param_floats % 4 == 0
    bn = [n[:-len(".running_mean")] for n, k, _ in O.model_spec(cfg) if n.endswith(".running_mean")]
    assert [b.name for b in e.bns] == bn
    assert sum(p.numel for p in e.params) == sum(int(np.prod(O._torch_shape(k, s))) for n, k, s in O.model_spec(cfg)
                                                 if not O.is_buffer(k))


@pytest.mark.parametrize("multimodal", [True, False])
def test_gradient_exchange_ranges_cover_the_buffer_in_completion_order(multimodal):
    """hippie_grad_split / hippie_grad_bounds: the ranges the data-parallel step all-reduces after parts 0, 2 and 3 are
    disjoint, cover every parameter, start at parameter boundaries and follow the order the backward pass completes
    them (latent head + decoders, then layer3 / layer4 / Linear of every encoder, then its stem / layer1 / layer2)."""
    from hippie_b200.engine import Engine
    cfg = O.CVAEConfig(z_dim=10) if multimodal else O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50)
    e = Engine(cfg.z_dim, cfg.output_size_wave, cfg.output_size_isi, cfg.class_hidden_dim, cfg.num_sources,
               cfg.num_classes, cfg.multimodal, 64)
    names = {p.offset: p.name for p in e.params}
    split, bounds = e.grad_split, e.grad_bounds
    assert len(bounds) == (2 if multimodal else 1)
    pre = ["encoder_mod1", "encoder_mod2"] if multimodal else ["encoder"]
    cover = []
    for (b, d, end), name in zip(bounds, pre):
        assert names[b] == name + ".conv1.weight" and names[d] == name + ".layer3.0.conv1.weight"
        assert b < d < end
        deep = [p for p in e.params if d <= p.offset < end]
        assert all(p.name.startswith(name + ".layer3") or p.name.startswith(name + ".layer4")
                   or p.name.startswith(name + ".linear") for p in deep)
        assert sum(p.numel for p in deep) > 0.9 * sum(p.numel for p in e.params if b <= p.offset < end)
        cover += [(b, d), (d, end)]
    cover.append((split, e.param_floats))
    cover.sort()
    assert cover[0][0] == 0 and all(a[1] == b[0] for a, b in zip(cover, cover[1:])) and cover[-1][1] == e.param_floats
    assert names[split] == ("fusion_encoder.0.weight" if multimodal else "encoder_fc.0.weight")


def test_error_codes_without_gpu():
    from hippie_b200.engine import Engine
    with pytest.raises(ValueError):
        Engine(0, 50, 100, 5, 5, 5, True, 64)
    e = Engine(10, 50, 100, 5, 5, 5, True, 64)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            e.allocate("cuda:0")


@pytest.mark.parametrize("multimodal", [True, False])
def test_module_mirror_state_dict_bit_exact_with_reference_init(multimodal):
    from hippie_b200.model import MultiModalCVAE, MultiModalCVAETrainModule, hippieUnimodalCVAE, \
        hippieUnimodalEmbeddingModelCVAE
    torch.manual_seed(42)
    if multimodal:
        cfg = O.CVAEConfig(z_dim=10)
        m = MultiModalCVAE(10, 50, 100, class_hidden_dim=5, num_sources=5, num_classes=5, max_batch=8)
        tm = MultiModalCVAETrainModule(m, learning_rate=1e-3, weight_decay=0.01, beta=0.5)
    else:
        cfg = O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100, num_classes=4)
        m = hippieUnimodalCVAE(10, 100, 5, 5, 4, max_batch=8)
        tm = hippieUnimodalEmbeddingModelCVAE(m, learning_rate=1e-3)
    ref = O.init_state(cfg, seed=42)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        assert torch.equal(sd[k], ref[k]), k
    assert list(tm.state_dict().keys()) == ["model." + k for k in ref]
    assert len(list(m.parameters())) == len(O.param_names(cfg))
    # load_state_dict round trip through the flat buffer (conv weights are stored [Cout][k][Cin])
    new = {k: (torch.randn_like(v) if v.is_floating_point() else v + 5) for k, v in ref.items()}
    m.load_state_dict(new)
    for k in new:
        assert torch.equal(m.state_dict()[k], new[k]), k
    w = m.state_dict()["decoder.layer4.1.conv1.conv.weight" if not multimodal else "decoder_mod1.layer4.1.conv1.conv.weight"]
    assert w.shape == (256, 512, 3) and not w.is_contiguous()
    # strict=False partial load, as the reference's stage-3 does after popping class_embedding
    part = {k: v for k, v in ref.items() if k != "class_embedding.weight"}
    res = m.load_state_dict(part, strict=False)
    assert res.missing_keys == ["class_embedding.weight"]
    # no CPU path
    with pytest.raises(RuntimeError, match="no CPU path"):
        if multimodal:
            m(torch.zeros(2, 1, 50), torch.zeros(2, 1, 100), torch.zeros(2, dtype=torch.long))
        else:
            m(torch.zeros(2, 1, 100), torch.zeros(2, dtype=torch.long))
    # torch-format optimizer state
    osd = tm.optimizer.state_dict()
    assert osd["param_groups"][0]["lr"] == 1e-3 and osd["state"] == {}


def test_ckpt_roundtrip(tmp_path):
    from hippie_b200.model import MultiModalCVAE, MultiModalCVAETrainModule
    torch.manual_seed(1)
    tm = MultiModalCVAETrainModule(MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=8))
    path = tmp_path / "epoch=0-step=1.ckpt"
    torch.save({"state_dict": tm.state_dict(), "epoch": 0, "global_step": 1}, path)
    torch.manual_seed(2)
    tm2 = MultiModalCVAETrainModule(MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=8))
    sd = torch.load(path)["state_dict"]
    tm2.load_state_dict(sd)
    for (k, a), (_, b) in zip(tm.state_dict().items(), tm2.state_dict().items()):
        assert torch.equal(a, b), k


def test_dataset_items_match_oracle_and_fixture(golden_dir):
    from hippie_b200.dataloading import EphysDataset, EphysDatasetLabeled, EphysTensorDataset
    fx = np.load(os.path.join(golden_dir, "cellexplorer_raw48.npz"))
    ds = EphysDataset(fx["wf"], fx["isi"], mode="both", normalize=False)
    lab = np.arange(len(ds)) % 4
    dl = EphysDatasetLabeled(fx["wf"], fx["isi"], lab, mode="both", normalize=False)
    dt = EphysTensorDataset(fx["wf"], fx["isi"], lab)
    for i in range(len(ds)):
        a, b = ds[i]
        assert np.array_equal(a.numpy(), fx["x1"][i]) and np.array_equal(b.numpy(), fx["x2"][i])
        oa, ob = O.dataset_item(fx["wf"][i], fx["isi"][i])
        assert torch.equal(a, oa) and torch.equal(b, ob)
        w, t, l = dl[i]
        assert torch.equal(w, a) and torch.equal(t, b) and l.dtype == torch.int64 and int(l) == lab[i]
        assert torch.equal(dt[i][0], a) and torch.equal(dt[i][1], b)
    w, t, l = dt.batch([3, 1, 2])
    assert w.shape == (3, 1, 50) and t.shape == (3, 1, 100) and l.tolist() == [lab[3], lab[1], lab[2]]
    assert EphysDatasetLabeled(fx["wf"], fx["isi"], lab, mode="wave", normalize=False)[0][0].shape == (1, 50)
    # empty dataset
    assert len(EphysDataset(np.zeros((0, 47)), np.zeros((0, 100)), mode="both", normalize=False)) == 0


def test_balanced_sampler_round_robin_and_seed():
    from hippie_b200.dataloading import BalancedBatchSampler
    labels = torch.tensor([0] * 5 + [1] * 2 + [2] * 1)
    s = BalancedBatchSampler(range(len(labels)), labels, seed=0)
    idx = list(iter(s))
    assert len(idx) == len(s) == 15
    assert [int(labels[i]) for i in idx] == [0, 1, 2] * 5  # strict round robin over classes
    assert sorted(i for i in idx if labels[i] == 0) == [0, 1, 2, 3, 4]  # majority class: every item exactly once
    assert list(iter(BalancedBatchSampler(range(len(labels)), labels, seed=0))) == idx
    assert list(iter(s)) == idx  # re-iterable


def test_shard_helpers():
    from hippie_b200.parallel import shard_batch_indices, shard_units
    perm = list(range(100, 100 + 23))
    got = [shard_batch_indices(perm, s, r, 2, 4) for s in range(3) for r in range(2)]
    assert got[0] == perm[0:4] and got[1] == perm[4:8] and got[4] == perm[16:20] and got[5] == perm[20:23]
    assert sum(got, []) == perm  # bit-exact partition of the permutation, in order
    spans = [shard_units(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_units(0, 0, 2) == (0, 0)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hippie_b200.parallel import all_reduce_gradients, broadcast_state, shard_batch_indices, train_step_overlapped
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
g = torch.arange(12, dtype=torch.float32) * (r + 1)
scale = all_reduce_gradients(g)
assert scale == 0.5 and torch.equal(g * scale, torch.arange(12, dtype=torch.float32) * 1.5), g
p = torch.full((5,), float(r))
broadcast_state([p], src=0)
assert torch.equal(p, torch.zeros(5))
mine = shard_batch_indices(list(range(10)), 0, r, w, 3)
allv = [None, None]
dist.all_gather_object(allv, mine)
assert allv == [[0, 1, 2], [3, 4, 5]]

class StubEngine:  # the host logic of the overlapped exchange, without a GPU: every part fills its gradient ranges
    grad_split = 5
    grad_bounds = [(0, 1, 3), (3, 4, 5)]  # (begin, deep, end) per encoder
    def __init__(self, rank):
        self.flat_grads = torch.zeros(12)
        self.rank, self.calls = rank, []
    def train_fwd_bwd_part(self, part, *a, **k):
        self.calls.append(part)
        head = torch.arange(5, dtype=torch.float32) * (self.rank + 1)
        if part == 0:
            self.flat_grads.zero_()
            self.flat_grads[self.grad_split:] = torch.arange(7, dtype=torch.float32) + 10 * self.rank
        elif part == 2:
            for _, d, e in self.grad_bounds:
                self.flat_grads[d:e] += head[d:e]
        elif part == 3:
            for b, d, _ in self.grad_bounds:
                self.flat_grads[b:d] += head[b:d]
        else:
            raise AssertionError("the overlapped step uses parts 0, 2, 3")
        return None
# sharded embedding inference: contiguous shards in rank order, gathered back into unit order (ragged: 7 units on 2 ranks)
from hippie_b200.parallel import gather_rows, shard_units
lo, hi = shard_units(7, r, w)
rows = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 3)
allrows = gather_rows(rows, 7)
assert torch.equal(allrows, torch.arange(7, dtype=torch.float32).unsqueeze(1).repeat(1, 3)), allrows
# the loader gives every rank the same number of steps and at least two rows per step, whatever the tail
import numpy as np
from hippie_b200.dataloading import EphysBatchLoader, EphysTensorDataset
for n in (64, 65, 67, 70, 95):
    ds = EphysTensorDataset(np.zeros((n, 50)), np.ones((n, 100)), np.arange(n))
    torch.manual_seed(5)
    ld = EphysBatchLoader(ds, 16, shuffle=True, rank=r, world=w)
    mine = [b[2].tolist() for b in ld]
    assert len(mine) == len(ld) and all(len(b) >= 2 for b in mine), (n, [len(b) for b in mine])
    both = [None, None]
    dist.all_gather_object(both, mine)
    assert len(both[0]) == len(both[1]), (n, len(both[0]), len(both[1]))
    seen = sorted(i for part in both for b in part for i in b)
    assert len(seen) == len(set(seen)) and len(seen) >= n - 3, (n, len(seen))  # a tail of < 2 * world rows is dropped
eng = StubEngine(r)
scale = train_step_overlapped(eng, None, None, None, None, None, 0.5)
assert scale == 0.5 and eng.calls == [0, 2, 3]
assert torch.equal(eng.flat_grads[5:], 2 * torch.arange(7, dtype=torch.float32) + 10), eng.flat_grads
assert torch.equal(eng.flat_grads[:5], 3 * torch.arange(5, dtype=torch.float32)), eng.flat_grads
dist.destroy_process_group()
print("ok", r)
'''


def test_data_parallel_exchange_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script), ROOT],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_batch_loader_matches_torch_dataloader_bit_exact():
    """EphysBatchLoader reproduces DataLoader's index stream, batch boundaries and default-generator consumption
    (shuffle, Subset, sampler), and shards global batches for data-parallel ranks."""
    from hippie_b200.dataloading import BalancedBatchSampler, EphysBatchLoader, EphysDatasetLabeled, EphysTensorDataset
    rng = np.random.default_rng(0)
    wf, isi, lab = rng.normal(size=(37, 47)), np.abs(rng.normal(size=(37, 100))), np.arange(37) % 4
    ds = EphysDatasetLabeled(wf, isi, lab, mode="both", normalize=False)
    dt = EphysTensorDataset(wf, isi, lab)

    def same(a, b):
        return len(a) == len(b) and all(torch.equal(x, y) for p, q in zip(a, b) for x, y in zip(p, q))

    for shuffle in (False, True):
        torch.manual_seed(3)
        a = list(torch.utils.data.DataLoader(ds, batch_size=8, shuffle=shuffle))
        sa = torch.get_rng_state()
        torch.manual_seed(3)
        loader = EphysBatchLoader(dt, 8, shuffle=shuffle)
        b = list(loader)
        assert same(a, b) and torch.equal(sa, torch.get_rng_state()) and len(loader) == 5
    idx = [5, 1, 9, 30, 2, 7, 8]
    torch.manual_seed(4)
    a = list(torch.utils.data.DataLoader(torch.utils.data.Subset(ds, idx), batch_size=3, shuffle=True))
    sa = torch.get_rng_state()
    torch.manual_seed(4)
    assert same(a, list(EphysBatchLoader(dt, 3, shuffle=True, indices=idx))) and torch.equal(sa, torch.get_rng_state())
    s1, s2 = BalancedBatchSampler(ds, torch.as_tensor(lab), seed=1), BalancedBatchSampler(dt, torch.as_tensor(lab), seed=1)
    assert same(list(torch.utils.data.DataLoader(ds, batch_size=6, sampler=s1)), list(EphysBatchLoader(dt, 6, sampler=s2)))
    # data parallel: the two ranks' batches interleave to the single-process batches of twice the size
    whole = list(EphysBatchLoader(dt, 8, shuffle=False))
    r0, r1 = list(EphysBatchLoader(dt, 4, rank=0, world=2)), list(EphysBatchLoader(dt, 4, rank=1, world=2))
    assert len(r0) == len(whole) and torch.equal(torch.cat([r0[0][0], r1[0][0]]), whole[0][0])
    assert sum(b[0].shape[0] for b in r0 + r1) == 37
    cat = EphysTensorDataset.concat([dt, EphysTensorDataset(wf[:5, :40], isi[:5, :51], lab[:5])])
    assert len(cat) == 42 and cat.wave.shape == (42, 1, 50) and cat.isi.shape == (42, 1, 100)


def test_trainer_callbacks_and_limits(tmp_path):
    from hippie_b200.trainer import EarlyStopping, ModelCheckpoint, _limit

    class FakeTrainer:
        current_epoch, global_step, should_stop, is_global_zero = 0, 0, False, True
        log_dir = str(tmp_path)
        saved = []

        def save_checkpoint(self, path):
            self.saved.append(path)
            open(path, "w").write("x")

    tr, ck, es = FakeTrainer(), ModelCheckpoint("val_loss"), EarlyStopping("val_loss", patience=2)
    for epoch, v in enumerate([3.0, 2.0, 2.5, 2.2, 2.1]):
        tr.current_epoch, tr.global_step = epoch, 10 * (epoch + 1)
        for cb in (ck, es):
            cb.on_validation_end(tr, None, {"val_loss": v})
        if tr.should_stop:
            break
    assert ck.best_model_score == 2.0 and ck.best_model_path.endswith("epoch=1-step=20.ckpt")
    assert os.path.exists(ck.best_model_path) and not os.path.exists(tr.saved[0])  # top-1: the older file is removed
    assert tr.should_stop and es.stopped_epoch == 3 and es.wait_count == 2
    assert _limit(10, None) == 10 and _limit(10, 0.25) == 2 and _limit(10, 3) == 3 and _limit(None, 0.5) is None
    with pytest.raises(RuntimeError, match="CPU"):
        from hippie_b200.trainer import Trainer
        Trainer(accelerator="cpu")


def test_cli_flags_match_the_reference():
    """Every flag of the reference's two CLIs exists with the same default (scripts/train_model_with_multimodal.py:38-69,
    scripts/inference_from_trained_model.py:15-46)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):
        spec = importlib.util.spec_from_file_location("cli_" + name, os.path.join(root, "scripts", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.path.insert(0, os.path.join(root, "scripts"))
        spec.loader.exec_module(mod)
        return mod

    a = load("train_model_with_multimodal").parse_args([])
    want = dict(z_dim=5, weight_decay=0.01, learning_rate=0.001, beta=1, dataset="cellexplorer-celltype", upload_model=False,
                wandb_tag="no_curr_sup_pretrain_data", project="HIPPIE", finetune_without_labels=True, pretrain_max_epochs=1,
                finetune_max_epochs=1, supervised_max_epochs=1, batch_size=512, supervised_batch_size=64,
                early_stopping_patience=30, gradient_clip_val=1.0, train_val_split=0.8, finetune_split=0.1,
                limit_train_batches=None, limit_val_batches=None, model_type="unimodal", mod1_weight=1.0, mod2_weight=1.0)
    for k, v in want.items():
        assert getattr(a, k) == v, k
    b = load("inference_from_trained_model").parse_args(["--wave-checkpoint", "w.ckpt", "--time-checkpoint", "t.ckpt"])
    assert (b.z_dim, b.dataset, b.output_dir) == (64, "cellexplorer-celltype", "./embeddings")


def test_oracle_matches_the_installed_reference_classes():
    """The pinning recipe as a CI check: oracle/ref_loader.py imports the UNMODIFIED reference modules (baseline/_ref, the
    offline pip install made by __graft_entry__.build(), or /root/reference) next to this repository's `hippie/` alias
    package, and the oracle reproduces the reference's forward, loss and gradients (fp64: to rounding)."""
    from oracle.ref_loader import load_reference, reference_root
    if reference_root() is None:
        pytest.skip("no reference here (neither baseline/_ref nor /root/reference)")
    R = load_reference()
    import hippie.model as alias  # the alias package still resolves to this repository afterwards
    assert alias.__file__.startswith(ROOT) and not R.model.__file__.startswith(os.path.join(ROOT, "hippie"))
    cfg = O.CVAEConfig(z_dim=10, num_classes=4)
    torch.manual_seed(42)
    ref = R.model.MultiModalCVAE(10, 50, 100, 5, cfg.num_sources, 4).double()
    tm = R.model.MultiModalCVAETrainModule(ref, learning_rate=1e-3, weight_decay=0.01, beta=0.5)
    tm.train()
    st = OrderedDict((k, v.detach().clone()) for k, v in ref.state_dict().items())
    x1, x2, labels, g = O.synthetic_batch(12, seed=3, labelled=True)
    labels[:, 0] %= 4
    eps = torch.randn(12, 10, generator=g).double()
    ref.reparameterize = lambda mu, lv: mu + eps * torch.exp(0.5 * lv)
    loss = tm.training_step((x1.double(), x2.double(), labels), 0)
    loss.backward()
    _, _, info = O.train_step(st, O.new_opt_state(st, cfg), cfg, x1.double(), x2.double(), labels, eps, lr=1e-3,
                              weight_decay=0.01, beta=0.5, max_norm=1.0)
    assert abs(float(loss) - float(info["loss"])) <= 1e-13 * abs(float(loss))
    gn = float(torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters() if p.grad is not None)))
    for n, p in ref.named_parameters():
        if p.grad is not None:  # (a bias in front of a BatchNorm has an analytically zero gradient: absolute floor)
            assert (p.grad - info["grads_raw"][n]).norm() <= 1e-11 * p.grad.norm() + 1e-13 * gn, n
