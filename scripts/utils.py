"""Helpers shared by the command-line drop-ins (reference scripts/utils.py): embedding extraction and the confusion-
matrix figure.  The arithmetic runs in the sm_100a engine; plotting libraries are optional imports."""
from __future__ import annotations

import numpy as np
import torch


def get_embeddings_multimodal(loader, module) -> np.ndarray:
    """Reference scripts/train_model_with_multimodal.py:22-34: eval mode, `model(sample)[0]` (= `encoded`), per-row z-score
    with the population standard deviation.  Here: the encoder-only embedding pass with the z-score fused into the head
    kernel (hippie_embed, ddof 0); the two decoders the reference runs and discards are skipped."""
    model = module.model
    was_training = module.training
    module.eval()
    out = []
    with torch.no_grad():
        for data1, data2, labels in loader:
            src, cls = module._split_labels(labels)
            out.append(model.embed(data1, data2, src, cls, zscore_ddof=0)["enc"])
    if was_training:
        module.train()
    return torch.cat(out).cpu().numpy() if out else np.zeros((0, model.z_dim), dtype=np.float32)


def get_embeddings(dataloader_wave, dataloader_time, wave_model, time_model):
    """Reference scripts/utils.py:75-101 (unimodal pair): `encoded` of each model, per-row z-score with the SAMPLE standard
    deviation (torch.std, ddof 1), joint = wave || isi.  Returns (waveform, isi, joint) as numpy arrays."""
    e_wave, e_time = [], []
    with torch.no_grad():
        for (wave, label_wave), (time, label_time) in zip(dataloader_wave, dataloader_time):
            assert (label_wave == label_time).all()
            sw, cw = wave_model._split_labels(label_wave)
            st, ct = time_model._split_labels(label_time)
            e_wave.append(wave_model.model.embed(wave, sw, cw, zscore_ddof=1)["enc"])
            e_time.append(time_model.model.embed(time, st, ct, zscore_ddof=1)["enc"])
    if not e_wave:  # an empty shard of a data-parallel inference run
        z = np.zeros((0, wave_model.model.z_dim), dtype=np.float32)
        return z, z, np.zeros((0, 2 * wave_model.model.z_dim), dtype=np.float32)
    w = torch.cat(e_wave).cpu().numpy()
    t = torch.cat(e_time).cpu().numpy()
    return w, t, np.concatenate([w, t], axis=1)


def make_confmat(cm, label_names, n_neighbors):
    """Row-normalised confusion matrix annotated with the raw counts (reference scripts/utils.py:10-40).  Returns a
    matplotlib figure, or None when matplotlib is not installed (it is optional here)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return None
    cm = np.asarray(cm, dtype=np.float64)
    norm = cm / np.maximum(cm.sum(axis=1, keepdims=True), 1.0)
    fig, ax = plt.subplots()
    im = ax.imshow(norm, cmap="Blues", vmin=0.0, vmax=1.0)
    fig.colorbar(im, ax=ax)
    for i in range(cm.shape[0]):
        for j in range(cm.shape[1]):
            ax.text(j, i, f"{norm[i, j]:.2f}\n({int(cm[i, j])})", ha="center", va="center")
    ax.set_xticks(range(len(label_names)), labels=[str(n) for n in label_names], rotation=45, ha="right")
    ax.set_yticks(range(len(label_names)), labels=[str(n) for n in label_names])
    ax.set_title(f"{n_neighbors} neighbors")
    plt.close(fig)
    return fig
