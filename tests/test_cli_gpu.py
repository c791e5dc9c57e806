"""End-to-end runs of the command-line drop-ins (scripts/) on a tiny synthetic data root: three training stages, the
.ckpt interchange between them, the CSV outputs, and the inference CLI on the checkpoints the training CLI wrote."""
import importlib.util
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    path = os.path.join(ROOT, "scripts", name + ".py")
    spec = importlib.util.spec_from_file_location("cli_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    spec.loader.exec_module(mod)
    return mod


def _write_table(root, folder, n, wave_w, isi_w, rng, labels=None):
    d = os.path.join(root, folder)
    os.makedirs(d, exist_ok=True)
    t = np.linspace(0, 1, wave_w)
    wf = np.stack([np.sin(2 * np.pi * (1 + i % 3) * t + rng.normal()) * rng.uniform(0.5, 1.5) for i in range(n)])
    isi = np.abs(rng.normal(size=(n, isi_w)))
    isi /= isi.sum(axis=1, keepdims=True)
    pd.DataFrame(wf).to_csv(os.path.join(d, "waveforms.csv"))  # leading unnamed index column, like the shipped files
    pd.DataFrame(isi).to_csv(os.path.join(d, "isi_dist.csv"))
    if labels is not None:
        pd.DataFrame({"0": labels}).to_csv(os.path.join(d, "labels.csv"))  # no `label` header, like the shipped files


@pytest.fixture(scope="module")
def data_root(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("datasets"))
    rng = np.random.default_rng(0)
    _write_table(root, "extracellular-mouse-a1", 90, 40, 51, rng)
    _write_table(root, "neonatal-mouse-brain-slice", 110, 50, 100, rng)
    _write_table(root, "juxtacellular-mouse-s1-celltype", 70, 351, 100, rng)
    names = np.array(["PV", "SST", "Pyra", "VIP"])
    _write_table(root, "cellexplorer-celltype", 120, 46, 100, rng, labels=names[np.arange(120) % 4])
    return root


@pytest.mark.parametrize("model_type", ["multimodal", "unimodal"])
def test_train_cli_three_stages_and_inference(data_root, tmp_path, model_type):
    cli = _load("train_model_with_multimodal")
    out = str(tmp_path / model_type)
    res = cli.main(["--model-type", model_type, "--z_dim", "10", "--dataset", "cellexplorer-celltype", "--batch-size", "64",
                    "--supervised-batch-size", "32", "--pretrain-max-epochs", "2", "--finetune-max-epochs", "1",
                    "--supervised-max-epochs", "1", "--data-root", data_root, "--out-dir", out, "--no-wandb", "--beta", "0.5"])
    kinds = ["joint"] if model_type == "multimodal" else ["waveform", "isi", "joint"]
    z = {"joint": 10 if model_type == "multimodal" else 20, "waveform": 10, "isi": 10}
    for k in kinds:
        pre = pd.read_csv(os.path.join(out, f"pretraining_cellexplorer-celltype_{k}_embeddings.csv"))
        assert len(pre) == 120 - int(0.1 * 120) and "embeddings" in pre.columns
        knn = pd.read_csv(os.path.join(out, f"cellexplorer-celltype_{k}_knn.csv"))
        assert len(knn) == 120 - int(0.8 * 120) and set(knn["true"]) <= {"PV", "SST", "Pyra", "VIP"}
        emb = pd.read_csv(os.path.join(out, f"cellexplorer-celltype_{k}_embeddings.csv"), index_col=0)
        assert emb.shape == (120, z[k] + 1)
        e = emb.drop(columns="label").to_numpy()
        assert np.isfinite(e).all() and np.abs(e.mean(axis=1)).max() < 1e-4  # per-row z-scored
    for stage in ("pretrain_ckpt", "supervised_ckpt"):
        for path in res[stage].values():
            ck = torch.load(path)
            assert {"state_dict", "epoch", "global_step", "optimizer_states"} <= set(ck)
            assert all(k.startswith("model.") for k in ck["state_dict"])
    # the inference CLI consumes those checkpoints
    inf = _load("inference_from_trained_model")
    emb_dir = str(tmp_path / (model_type + "_emb"))
    if model_type == "multimodal":
        r = inf.main(["--z_dim", "10", "--dataset", "cellexplorer-celltype", "--joint-checkpoint",
                      res["supervised_ckpt"]["joint"], "--output-dir", emb_dir, "--data-root", data_root, "--no-umap"])
        assert r["joint"].shape == (120, 10)
    else:
        r = inf.main(["--z_dim", "10", "--dataset", "cellexplorer-celltype", "--wave-checkpoint",
                      res["supervised_ckpt"]["waveform"], "--time-checkpoint", res["supervised_ckpt"]["isi"],
                      "--output-dir", emb_dir, "--data-root", data_root, "--no-umap"])
        assert r["joint"].shape == (120, 20)
        df = pd.read_csv(os.path.join(emb_dir, "cellexplorer-celltype_joint_embeddings.csv"))
        assert df.shape == (120, 20 + 2) and (df["label_name"] == "unknown").all()
