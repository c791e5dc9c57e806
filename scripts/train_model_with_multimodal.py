#!/usr/bin/env python
"""Command-line drop-in for the reference's `scripts/train_model_with_multimodal.py` on the sm_100a engine.

Same flags, same three stages and the same output files as the reference (scripts/train_model_with_multimodal.py:36-975):

  1. pre-train on every source table except the target dataset (multimodal: one joint cVAE; unimodal: a waveform and an
     ISI cVAE), keep the checkpoint with the lowest epoch-mean `val_loss`;
  2. label-free fine-tune on `--finetune-split` of the target dataset at lr/10, write
     `pretraining_<dataset>_{joint|waveform|isi}_embeddings.csv`;
  3. supervised fine-tune (class labels, balanced batches of `--supervised-batch-size`), KNN (k = 5..19) on the
     z-scored `encoded` embeddings, write `<dataset>_*_knn.csv` and `<dataset>_*_embeddings.csv`.

What differs, deliberately: training runs through `hippie_b200.trainer.Trainer` (Lightning is
not needed), batches come from `EphysBatchLoader` (same index stream and RNG consumption as the reference's DataLoaders,
one gather per batch), the labelled `mode="both"` dataset exists, a missing source table is skipped with a warning instead
of crashing, `labels.csv` files without a `label` column fall back to their last column, wandb / matplotlib are optional,
and `--data-root` / `--out-dir` / `--max-batch` are extra flags.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import pandas as pd
import torch
from torch.utils.data import random_split

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from hippie_b200.dataloading import BalancedBatchSampler, EphysBatchLoader, EphysTensorDataset  # noqa: E402
from hippie_b200.model import (MultiModalCVAE, MultiModalCVAETrainModule, hippieUnimodalCVAE,  # noqa: E402
                               hippieUnimodalEmbeddingModelCVAE)
from hippie_b200.trainer import EarlyStopping, ModelCheckpoint, Trainer  # noqa: E402
from utils import get_embeddings, get_embeddings_multimodal, make_confmat  # noqa: E402

# folder -> recording-source id of the condition embedding (reference scripts/...:81-89)
SOURCE_ID = {"extracellular-mouse-a1": 1, "cellexplorer-celltype": 3, "cellexplorer-area": 3,
             "juxtacellular-mouse-s1-celltype": 4, "juxtacellular-mouse-s1-area": 4, "allenscope-neuropixel": 3,
             "neonatal-mouse-brain-slice": 2}


def str2bool(v):
    """The reference declares --finetune-without-labels with type=bool, so any non-empty string parses as True
    (SURVEY.md F4); here 'false' / '0' / 'no' really switch it off."""
    return str(v).strip().lower() not in ("0", "false", "no", "off", "")


def parse_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("--z_dim", type=int, default=5, help="Dimension of latent space")
    p.add_argument("--weight-decay", type=float, default=0.01)
    p.add_argument("--learning-rate", type=float, default=0.001)
    p.add_argument("--beta", type=float, default=1, help="Weight for KL divergence loss")
    p.add_argument("--dataset", type=str, default="cellexplorer-celltype")
    p.add_argument("--upload-model", action="store_true")
    p.add_argument("--wandb-tag", type=str, default="no_curr_sup_pretrain_data")
    p.add_argument("--project", type=str, default="HIPPIE")
    p.add_argument("--finetune-without-labels", type=str2bool, default=True)
    p.add_argument("--pretrain-max-epochs", type=int, default=1)
    p.add_argument("--finetune-max-epochs", type=int, default=1)
    p.add_argument("--supervised-max-epochs", type=int, default=1)
    p.add_argument("--batch-size", type=int, default=512)
    p.add_argument("--supervised-batch-size", type=int, default=64)
    p.add_argument("--early-stopping-patience", type=int, default=30)
    p.add_argument("--gradient-clip-val", type=float, default=1.0)
    p.add_argument("--train-val-split", type=float, default=0.8)
    p.add_argument("--finetune-split", type=float, default=0.1)
    p.add_argument("--limit-train-batches", type=float, default=None)
    p.add_argument("--limit-val-batches", type=float, default=None)
    p.add_argument("--model-type", type=str, choices=["unimodal", "multimodal"], default="unimodal",
                   help="Whether to use separate models for each modality or a joint model")
    p.add_argument("--mod1-weight", type=float, default=1.0, help="Weight for the waveform modality loss in multimodal model")
    p.add_argument("--mod2-weight", type=float, default=1.0, help="Weight for the ISI modality loss in multimodal model")
    # additions
    p.add_argument("--data-root", type=str, default="datasets", help="directory holding <dataset>/waveforms.csv etc.")
    p.add_argument("--out-dir", type=str, default=".", help="where the CSV outputs and lightning_logs/ go")
    p.add_argument("--max-batch", type=int, default=None, help="engine workspace size (default: max of the batch sizes)")
    p.add_argument("--no-wandb", action="store_true", help="do not import / initialise wandb")
    return p.parse_args(argv)


class NullRun:
    """Stand-in for wandb when it is disabled or not installed."""

    def log(self, *a, **k):
        pass

    def log_artifact(self, *a, **k):
        pass

    def update_config(self, *a, **k):
        pass


class WandbRun(NullRun):
    def __init__(self, project, name):
        import wandb
        self.wandb = wandb
        self.run = wandb.init(project=project, name=name, mode=os.environ.get("WANDB_MODE", "offline"))

    def log(self, d, **k):
        self.wandb.log(d, **k)

    def log_metrics(self, d, step=None):
        self.wandb.log(d, step=step)

    def log_artifact(self, path, name=None, type=None):
        try:
            self.run.log_artifact(path, name=name, type=type)
        except Exception as e:  # offline runs cannot always stage artifacts
            print(f"wandb artifact {path}: {e}")

    def update_config(self, args):
        self.wandb.config.update(vars(args), allow_val_change=True)


def read_table(data_root, folder, dropna=False):
    """(waveforms, isi) as float arrays.  pandas reads the unnamed index column as a feature, exactly like the reference
    (scripts/...:626-627) -- checkpoints are only interchangeable if the inputs are."""
    wf = pd.read_csv(os.path.join(data_root, folder, "waveforms.csv"))
    isi = pd.read_csv(os.path.join(data_root, folder, "isi_dist.csv"))
    if dropna:
        wf, isi = wf.dropna(axis=1), isi.dropna(axis=1)
    return wf.to_numpy(), isi.to_numpy()


class Projected:
    """Serves (wave, label) or (isi, label) batches from a loader of (wave, isi, label) batches."""

    def __init__(self, loader, kind):
        self.loader, self.kind = loader, kind

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for w, t, lab in self.loader:
            yield (w, lab) if self.kind == "waveform" else (t, lab)


class Task:
    """One model family of the run: the joint multimodal cVAE, or the waveform / ISI unimodal cVAEs."""

    def __init__(self, kind, args, num_sources, max_batch):
        self.kind, self.args, self.num_sources, self.max_batch = kind, args, num_sources, max_batch

    def new_module(self, num_classes, lr):
        a = self.args
        if self.kind == "joint":
            m = MultiModalCVAE(z_dim=a.z_dim, output_size_wave=50, output_size_isi=100, class_hidden_dim=5,
                               num_sources=self.num_sources, num_classes=num_classes, max_batch=self.max_batch)
            return MultiModalCVAETrainModule(m, learning_rate=lr, weight_decay=a.weight_decay, beta=a.beta,
                                             mod1_weight=a.mod1_weight, mod2_weight=a.mod2_weight)
        m = hippieUnimodalCVAE(z_dim=a.z_dim, output_size=50 if self.kind == "waveform" else 100, class_hidden_dim=5,
                               num_sources=self.num_sources, num_classes=num_classes, max_batch=self.max_batch)
        # the reference builds the unimodal modules without beta=, so --beta is ignored there (default 1)
        return hippieUnimodalEmbeddingModelCVAE(m, learning_rate=lr, weight_decay=a.weight_decay)

    def rewrap(self, module, lr):
        """A fresh train module (new optimizer state, new lr) around the SAME model (scripts/...:740-747)."""
        a = self.args
        if self.kind == "joint":
            return MultiModalCVAETrainModule(module.model, learning_rate=lr, weight_decay=a.weight_decay, beta=a.beta,
                                             mod1_weight=a.mod1_weight, mod2_weight=a.mod2_weight)
        return hippieUnimodalEmbeddingModelCVAE(module.model, learning_rate=lr, weight_decay=a.weight_decay)

    def view(self, loader):
        return loader if self.kind == "joint" else Projected(loader, self.kind)

    def clip(self):
        # the reference's unimodal waveform Trainer omits gradient_clip_val (scripts/...:200-207)
        return None if self.kind == "waveform" else self.args.gradient_clip_val


def fit(task, module, train_loader, val_loader, epochs, ckpt, stop, args, logger):
    tr = Trainer(max_epochs=epochs, logger=logger, callbacks=[ckpt, stop], limit_train_batches=args.limit_train_batches,
                 limit_val_batches=args.limit_val_batches, gradient_clip_val=task.clip(),
                 default_root_dir=os.path.join(args.out_dir, "lightning_logs"))
    tr.fit(module, task.view(train_loader), task.view(val_loader))
    return ckpt.best_model_path


def embeddings_of(tasks, modules, loader):
    """{"joint": [N, z]} for the multimodal model; {"waveform", "isi", "joint"} for the unimodal pair."""
    if "joint" in tasks:
        return {"joint": get_embeddings_multimodal(loader, modules["joint"])}
    for m in modules.values():
        m.eval()
    w, t, j = get_embeddings(Projected(loader, "waveform"), Projected(loader, "isi"), modules["waveform"], modules["isi"])
    return {"waveform": w, "isi": t, "joint": j}


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("hippie_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    os.makedirs(args.out_dir, exist_ok=True)
    out = lambda name: os.path.join(args.out_dir, name)
    torch.manual_seed(42)

    all_sources = dict(SOURCE_ID)
    num_sources = max(all_sources.values()) + 1
    pretrain_folders = dict(SOURCE_ID)
    if "justacellular" in args.dataset:  # (sic) the reference's typo: juxtacellular sets are never excluded
        pretrain_folders.pop("justacellular-mouse-s1-celltype", None)
        pretrain_folders.pop("juxtacellular-mouse-s1-area", None)
    if "cellexplorer" in args.dataset:
        pretrain_folders.pop("cellexplorer-celltype", None)
        pretrain_folders.pop("cellexplorer-area", None)

    max_batch = args.max_batch or max(args.batch_size, args.supervised_batch_size, 128)
    kinds = ["joint"] if args.model_type == "multimodal" else ["waveform", "isi"]
    tasks = {k: Task(k, args, num_sources, max_batch) for k in kinds}
    run_name = f"{args.wandb_tag}{args.dataset}_{'joint' if args.model_type == 'multimodal' else 'unimodal'}_model_{args.z_dim}"
    run = NullRun()
    if not args.no_wandb:
        try:
            run = WandbRun(args.project, run_name)
        except Exception as e:
            print(f"wandb disabled: {e}")
    logger = run if hasattr(run, "log_metrics") else None

    # ------------------------------------------------------------------ stage 1: pre-training on the other sources
    tables = []
    for folder, sid in pretrain_folders.items():
        if folder == args.dataset:
            continue
        try:
            wf, isi = read_table(args.data_root, folder)
        except FileNotFoundError as e:
            print(f"skipping {folder}: {e}")
            continue
        print(f"Folder {folder} has shapes {wf.shape} and {isi.shape}")
        tables.append(EphysTensorDataset(wf, isi, np.full((wf.shape[0],), sid)))
    if not tables:
        raise SystemExit(f"no pre-training tables under {args.data_root}")
    pre = EphysTensorDataset.concat(tables)  # ConcatDataset order: folder order
    n = len(pre)
    n_train = int(args.train_val_split * n)
    train_idx, val_idx = random_split(list(range(n)), [n_train, n - n_train])
    modules, ckpts, stops, best = {}, {}, {}, {}
    for k, task in tasks.items():  # model construction order = RNG consumption order of the reference
        modules[k] = task.new_module(num_classes=5, lr=args.learning_rate)
    for k, task in tasks.items():
        ckpts[k] = ModelCheckpoint(monitor="val_loss", save_top_k=1, mode="min")
        stops[k] = EarlyStopping(monitor="val_loss", patience=args.early_stopping_patience, mode="min")
        train_loader = EphysBatchLoader(pre, args.batch_size, shuffle=True, indices=list(train_idx), pin_memory=True)
        val_loader = EphysBatchLoader(pre, args.batch_size, shuffle=False, indices=list(val_idx), pin_memory=True)
        best[k] = fit(task, modules[k], train_loader, val_loader, args.pretrain_max_epochs, ckpts[k], stops[k], args, logger)
        modules[k].load_state_dict(torch.load(best[k])["state_dict"])

    # ------------------------------------------------------------------ stage 2: label-free fine-tuning on the target
    wf_ft, isi_ft = read_table(args.data_root, args.dataset, dropna=True)
    target = EphysTensorDataset(wf_ft, isi_ft, np.full((wf_ft.shape[0],), all_sources[args.dataset]))
    if args.finetune_without_labels:
        idx = list(range(len(target)))
        meta_path = os.path.join(args.data_root, args.dataset, "metadata.csv")
        if os.path.exists(meta_path) and "chip" in args.dataset:  # recordings of the first ten time points train
            meta = pd.read_csv(meta_path)
            meta["datetime"] = pd.to_datetime(meta["datetime"]).dt.time
            first = meta["datetime"].sort_values().unique()[:10]
            ft_train = meta[meta["datetime"].isin(first)].index.tolist()
            ft_test = meta[~meta["datetime"].isin(first)].index.tolist()
        else:
            k = int(args.finetune_split * len(idx))
            ft_train, ft_test = [list(s) for s in random_split(idx, [k, len(idx) - k])]
        for k, task in tasks.items():
            modules[k] = task.rewrap(modules[k], lr=0.1 * args.learning_rate)
            tl = EphysBatchLoader(target, args.batch_size, shuffle=False, indices=ft_train, pin_memory=True)
            vl = EphysBatchLoader(target, args.batch_size, shuffle=False, indices=ft_test, pin_memory=True)
            best[k] = fit(task, modules[k], tl, vl, args.finetune_max_epochs, ckpts[k], stops[k], args, logger)
            modules[k].load_state_dict(torch.load(best[k])["state_dict"])
        emb_loader = EphysBatchLoader(target, args.batch_size, shuffle=False, indices=ft_test)
    else:
        emb_loader = EphysBatchLoader(target, args.batch_size, shuffle=False)
    for name, e in embeddings_of(tasks, modules, emb_loader).items():
        path = out(f"pretraining_{args.dataset}_{name}_embeddings.csv")
        pd.DataFrame({"embeddings": list(e)}).to_csv(path)
        run.log_artifact(path, name=os.path.basename(path), type=os.path.basename(path))

    # ------------------------------------------------------------------ stage 3: supervised fine-tuning + KNN
    from sklearn.preprocessing import LabelEncoder

    from hippie_b200.knn import knn_sweep  # KNeighborsClassifier / balanced accuracy / confusion matrix on the GPU
    ds = args.dataset
    sup_wf, sup_isi = read_table(args.data_root, ds)
    labels_path = os.path.join(args.data_root, ds, "labels.csv")
    if os.path.exists(labels_path):
        lab_df = pd.read_csv(labels_path)
        col = "label" if "label" in lab_df.columns else lab_df.columns[-1]  # shipped files lack a `label` header (F4)
        raw_labels = lab_df[col].values
    else:
        print(f"No labels.csv found for {ds}")
        raw_labels = np.zeros(len(sup_wf))
    le = LabelEncoder().fit(raw_labels)
    y = le.transform(raw_labels)
    idx = list(range(len(sup_wf)))
    n_train = int(args.train_val_split * len(idx))
    tr_idx, va_idx = [list(s) for s in random_split(idx, [n_train, len(idx) - n_train])]
    y_train, y_val = y[tr_idx], y[va_idx]
    # rows of the class table: every class of the label encoder, not only those that reached the training split (a
    # class seen only in validation would index past the table; nn.Embedding raises IndexError there)
    num_classes = len(le.classes_)
    src = all_sources[ds]
    stack = lambda yy: np.vstack((yy, src * np.ones_like(yy))).T  # [class, source] (hippie/model.py:456-460)
    d_train = EphysTensorDataset(sup_wf[tr_idx], sup_isi[tr_idx], stack(y_train))
    d_val = EphysTensorDataset(sup_wf[va_idx], sup_isi[va_idx], stack(y_val))
    d_all = EphysTensorDataset(sup_wf, sup_isi, stack(y))
    sup_modules, sup_ckpt, sup_stop, sup_best = {}, {}, {}, {}
    for k, task in tasks.items():
        sup_modules[k] = task.new_module(num_classes=num_classes, lr=0.1 * args.learning_rate)
        state = torch.load(best[k])["state_dict"]
        state.pop("model.class_embedding.weight")  # the class table is re-initialised for the real classes
        sup_modules[k].load_state_dict(state, strict=False)
    for k, task in tasks.items():
        sampler = BalancedBatchSampler(d_train, torch.as_tensor(y_train))
        tl = EphysBatchLoader(d_train, args.supervised_batch_size, sampler=sampler, pin_memory=True)
        vl = EphysBatchLoader(d_val, args.supervised_batch_size, shuffle=False, pin_memory=True)
        sup_ckpt[k] = ModelCheckpoint(monitor="val_loss", save_top_k=1, mode="min")
        sup_stop[k] = EarlyStopping(monitor="val_loss", patience=args.early_stopping_patience, mode="min")
        sup_best[k] = fit(task, sup_modules[k], tl, vl, args.supervised_max_epochs, sup_ckpt[k], sup_stop[k], args, logger)
        run.log({f"best_epoch_{k}": sup_best[k]})
        sup_modules[k].load_state_dict(torch.load(sup_best[k])["state_dict"])
        sup_modules[k].eval()

    e_train = embeddings_of(tasks, sup_modules, EphysBatchLoader(d_train, 128))
    e_val = embeddings_of(tasks, sup_modules, EphysBatchLoader(d_val, args.supervised_batch_size))
    e_all = embeddings_of(tasks, sup_modules, EphysBatchLoader(d_all, 128))
    neighbor_options = list(range(5, 20))
    for name in e_train:
        for k in neighbor_options:
            print(f"KNN with {k} neighbors")
        sweep = knn_sweep(e_train[name], y_train, e_val[name], y_val, neighbor_options)  # one pass for all k
        acc, best_k, pred = sweep["balanced_accuracy"], sweep["best_neighbors"], sweep["pred"]
        knn_path = out(f"{ds}_{name}_knn.csv")
        pd.DataFrame({"pred": le.inverse_transform(pred.astype(int)),
                      "true": le.inverse_transform(y_val.astype(int))}).to_csv(knn_path)
        run.log_artifact(knn_path, name=os.path.basename(knn_path), type=os.path.basename(knn_path))
        emb_df = pd.DataFrame(e_all[name])
        emb_df["label"] = le.inverse_transform(y.astype(int))
        emb_path = out(f"{ds}_{name}_embeddings.csv")
        emb_df.to_csv(emb_path)
        run.log_artifact(emb_path, name=os.path.basename(emb_path), type=os.path.basename(emb_path))
        run.log({f"best_balanced_accuracy_{name}": float(np.max(acc))})
        print(f"{name}: best balanced accuracy {np.max(acc):.4f} with {best_k} neighbors")
        fig = make_confmat(sweep["confusion"], le.classes_, best_k)
        if fig is not None and hasattr(run, "wandb"):
            run.log({f"{ds}_confusion_matrix_{name}": run.wandb.Image(fig)})
    if args.upload_model:
        for k, path in sup_best.items():
            run.log_artifact(path, name=f"{k}_model_ft_d{ds}_z{args.z_dim}_lr{args.learning_rate}.pt", type="model")
    run.update_config(args)
    return {"pretrain_ckpt": best, "supervised_ckpt": sup_best}


if __name__ == "__main__":
    main()
