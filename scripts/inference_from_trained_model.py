#!/usr/bin/env python
"""Command-line drop-in for the reference's `scripts/inference_from_trained_model.py`: embeddings of a dataset from
trained checkpoints, written as `<dataset>_{waveform,isi,joint}_embeddings.csv` (+ optional UMAP figures).

Reference behaviour kept (scripts/inference_from_trained_model.py:60-163): both CSV tables read with pandas (index column
included), NaN columns dropped, labels from `metadata.csv:label` or dummy zeros, batches of 128, `num_sources = 5`, a
`model.class_embedding.weight` whose row count does not match is dropped from the checkpoint (`strict=False`), eval mode,
per-row z-score with the sample standard deviation, `label` / `label_name` columns in the CSVs.

Additions: `--joint-checkpoint` embeds with a multimodal checkpoint instead of the unimodal pair (population z-score, as
`get_embeddings_multimodal` does); `--data-root`; UMAP / matplotlib are optional imports (the figures are skipped when
they are missing); the embedding pass runs on the GPU through the encoder-only engine call (the reference's script can
only run on CPU-only hosts because of `torch.device("gpu")`).

Several GPUs (BASELINE.json config 4, "embedding inference sharded across 8 x B200"): launched under
`torchrun --nproc-per-node N scripts/inference_from_trained_model.py ...` every rank embeds one contiguous shard of the
units (hippie_b200.parallel.shard_units; eval-mode BatchNorm, so units are independent and nothing is exchanged during
the compute), the rows are gathered in rank order and rank 0 writes the same CSV files a single GPU writes, bit for bit.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from hippie_b200.dataloading import EphysBatchLoader, EphysTensorDataset  # noqa: E402
from hippie_b200.parallel import gather_rows, init_from_env, shard_units  # noqa: E402
from hippie_b200.model import (MultiModalCVAE, MultiModalCVAETrainModule, hippieUnimodalCVAE,  # noqa: E402
                               hippieUnimodalEmbeddingModelCVAE)
from utils import get_embeddings, get_embeddings_multimodal  # noqa: E402

NUM_SOURCES = 5
BATCH = 128


def parse_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("--z_dim", type=int, default=64, required=False, help="Dimensionality of the latent space")
    p.add_argument("--dataset", type=str, default="cellexplorer-celltype", help="Dataset to perform inference on")
    p.add_argument("--wave-checkpoint", type=str, default=None, help="Path to the waveform model checkpoint")
    p.add_argument("--time-checkpoint", type=str, default=None, help="Path to the time model checkpoint")
    p.add_argument("--joint-checkpoint", type=str, default=None, help="Path to a multimodal checkpoint (instead of the pair)")
    p.add_argument("--output-dir", type=str, default="./embeddings", help="Directory to save embeddings and visualizations")
    p.add_argument("--data-root", type=str, default="datasets")
    p.add_argument("--no-umap", action="store_true")
    a = p.parse_args(argv)
    if not a.joint_checkpoint and not (a.wave_checkpoint and a.time_checkpoint):
        p.error("give --wave-checkpoint and --time-checkpoint (reference usage) or --joint-checkpoint")
    return a


def load_into(module, path, num_classes):
    ckpt = torch.load(path, map_location="cpu")
    state = ckpt["state_dict"]
    key = "model.class_embedding.weight"
    if key in state and state[key].size(0) != num_classes:
        print(f"Warning: class embedding size mismatch in {os.path.basename(path)}; removing it from the checkpoint")
        state.pop(key)
    module.load_state_dict(state, strict=False)
    return module.to(f"cuda:{torch.cuda.current_device()}").eval()


def umap_figure(embeddings, labels, title, path):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        import umap
    except Exception as e:
        print(f"skipping UMAP figure ({e})")
        return
    xy = umap.UMAP(random_state=42).fit_transform(embeddings)
    fig, ax = plt.subplots(figsize=(10, 8))
    for lab in np.unique(labels):
        sel = labels == lab
        ax.scatter(xy[sel, 0], xy[sel, 1], s=5, label=str(lab))
    ax.set_title(title)
    ax.legend(markerscale=3)
    fig.savefig(path, dpi=300, bbox_inches="tight")
    plt.close(fig)


def main(argv=None):
    args = parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("hippie_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    rank, world = init_from_env()
    os.makedirs(args.output_dir, exist_ok=True)
    torch.manual_seed(42)
    print(f"Loading dataset: {args.dataset}")
    folder = os.path.join(args.data_root, args.dataset)
    wf = pd.read_csv(os.path.join(folder, "waveforms.csv")).dropna(axis=1).to_numpy()
    isi = pd.read_csv(os.path.join(folder, "isi_dist.csv")).dropna(axis=1).to_numpy()
    labels, label_names = None, None
    meta = os.path.join(folder, "metadata.csv")
    if os.path.exists(meta):
        md = pd.read_csv(meta)
        if "label" in md.columns:
            labels = md["label"].to_numpy()
            label_names = md["label"].unique()
            print(f"Found {len(label_names)} unique labels: {label_names}")
    if labels is None:
        labels, label_names = np.zeros(wf.shape[0], dtype=np.int64), ["unknown"]
        print("No labels found, using dummy labels")
    # the reference hands `labels` to the modules as a 1-D label tensor, i.e. as the SOURCE id (hippie/model.py:461-462)
    codes = labels.astype(np.int64) if np.issubdtype(labels.dtype, np.number) else pd.factorize(labels)[0]
    num_classes = len(np.unique(labels))
    data = EphysTensorDataset(wf, isi, codes)
    lo, hi = shard_units(len(data), rank, world)  # this rank's contiguous shard of the units (all of them on one GPU)
    loader = EphysBatchLoader(data, BATCH, shuffle=False, indices=range(lo, hi))

    print("Loading models from checkpoints...")
    results = {}
    if args.joint_checkpoint:
        m = MultiModalCVAETrainModule(MultiModalCVAE(args.z_dim, 50, 100, 5, NUM_SOURCES, num_classes, max_batch=BATCH))
        results["joint"] = get_embeddings_multimodal(loader, load_into(m, args.joint_checkpoint, num_classes))
    else:
        wave = hippieUnimodalEmbeddingModelCVAE(hippieUnimodalCVAE(args.z_dim, 50, 5, NUM_SOURCES, num_classes, max_batch=BATCH))
        time = hippieUnimodalEmbeddingModelCVAE(hippieUnimodalCVAE(args.z_dim, 100, 5, NUM_SOURCES, num_classes, max_batch=BATCH))
        wave, time = load_into(wave, args.wave_checkpoint, num_classes), load_into(time, args.time_checkpoint, num_classes)
        proj = lambda kind: ((w, lab) if kind == "w" else (t, lab) for w, t, lab in loader)
        print("Extracting embeddings...")
        w, t, j = get_embeddings(proj("w"), proj("t"), wave, time)
        results = {"waveform": w, "isi": t, "joint": j}

    if world > 1:  # rank order = unit order: rank 0 writes what a single GPU would have written
        dev = torch.device("cuda", torch.cuda.current_device())
        results = {k: gather_rows(torch.as_tensor(v).to(dev), len(data)).cpu().numpy() for k, v in results.items()}
        if rank != 0:
            return results
    print("Saving embeddings...")
    for name, emb in results.items():
        df = pd.DataFrame(emb)
        df["label"] = labels
        if label_names is not None:
            names = [label_names[int(c)] if np.issubdtype(labels.dtype, np.number) and int(c) < len(label_names) else c
                     for c in labels]
            df["label_name"] = pd.Categorical(names)
        path = os.path.join(args.output_dir, f"{args.dataset}_{name}_embeddings.csv")
        df.to_csv(path, index=False)
        print(f"Saved {name} embeddings to {path}")
        if not args.no_umap:
            umap_figure(emb, labels, f"{args.dataset} {name} embeddings",
                        os.path.join(args.output_dir, f"{args.dataset}_{name}_umap.png"))
    return results


if __name__ == "__main__":
    main()
