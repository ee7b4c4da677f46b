"""Table / dense checkpoint I/O.

The reference moves weights in and out through Keras: `weights=[ndarray]` on the embedding layers (features/group.py:283-291,
fed by `FeaturePool(pre_embd)`) and `model.save_weights / load_weights` (docs/source/Quickstart.md:250-262).  Here:

    save_weights(model, dir)      one `.npy` per weight; tables with more than `shard_rows` rows are written as row shards
                                  streamed from the device chunk by chunk (a 40 M x 16 table never exists on the host in one piece)
    load_weights(model, dir)      the inverse, shape-checked, streamed back to the device the same way
    save_sharded_tables(engine, dir) / load_sharded_tables(engine, dir)
                                  the row-sharded engine: rank r writes the rows it owns (row % N == r) of every sharded table as
                                  `<table>.rank<r>of<N>.npy`, rank 0 the replicated tables and the dense parameters; loading accepts
                                  the same N only (re-sharding = load on one GPU, save, shard again)
    export_pre_embd(model)        {feature name: ndarray} in the form `FeaturePool(pre_embd=...)` takes (group.py:283-284)

Weights are keyed by the position of their layer in the model's execution order + the layer class + the weight name, which is
stable for the same constructor call; `manifest.json` also records the layer names for humans.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterator, Tuple

import numpy as np
import torch

SHARD_ROWS = 1 << 22


def _named_weights(model) -> Iterator[Tuple[str, str, torch.nn.Parameter]]:
    seen = set()
    for i, layer in enumerate(model._all_layers()):
        for key, p in layer._weights.items():
            if id(p) in seen:
                continue
            seen.add(id(p))
            yield f"{i:03d}_{type(layer).__name__}.{key}", layer.name, p


def _save_tensor(t: torch.Tensor, base: str, shard_rows: int):
    files = []
    if t.dim() == 2 and t.shape[0] > shard_rows:
        for r0 in range(0, t.shape[0], shard_rows):
            r1 = min(t.shape[0], r0 + shard_rows)
            f = f"{base}.rows{r0:012d}-{r1:012d}.npy"
            np.save(f, t[r0:r1].detach().cpu().numpy())
            files.append(os.path.basename(f))
    else:
        np.save(base + ".npy", t.detach().cpu().numpy())
        files.append(os.path.basename(base) + ".npy")
    return files


def _load_tensor(dst: torch.Tensor, directory: str, files) -> None:
    if len(files) == 1 and ".rows" not in files[0]:
        a = np.load(os.path.join(directory, files[0]))
        if tuple(a.shape) != tuple(dst.shape):
            raise ValueError(f"{files[0]}: shape {a.shape} does not match the model's {tuple(dst.shape)}")
        dst.copy_(torch.from_numpy(a).to(dst.dtype))
        return
    done = 0
    for f in files:
        span = f.split(".rows")[1].split(".npy")[0]
        r0, r1 = (int(x) for x in span.split("-"))
        a = np.load(os.path.join(directory, f))
        if a.shape[0] != r1 - r0 or tuple(a.shape[1:]) != tuple(dst.shape[1:]) or r1 > dst.shape[0]:
            raise ValueError(f"{f}: shard does not fit a table of shape {tuple(dst.shape)}")
        dst[r0:r1].copy_(torch.from_numpy(a).to(dst.dtype))
        done += r1 - r0
    if done != dst.shape[0]:
        raise ValueError(f"row shards cover {done} of {dst.shape[0]} rows")


def save_weights(model, directory: str, shard_rows: int = SHARD_ROWS) -> Dict:
    model.sync()  # fused training keeps dense weights in the engine's flat buffer
    os.makedirs(directory, exist_ok=True)
    manifest = {"format": "handyrec_b200.weights.v1", "weights": {}}
    for key, lname, p in _named_weights(model):
        files = _save_tensor(p.data, os.path.join(directory, key), shard_rows)
        manifest["weights"][key] = {"layer": lname, "shape": list(p.shape), "files": files}
    with open(os.path.join(directory, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def load_weights(model, directory: str) -> None:
    manifest = json.load(open(os.path.join(directory, "manifest.json")))
    have = manifest["weights"]
    for key, _, p in _named_weights(model):
        if key not in have:
            raise KeyError(f"{directory}: no weight {key!r} (was the model built by the same constructor call?)")
        _load_tensor(p.data, directory, have[key]["files"])
    fused = getattr(model, "_fused", None)
    if fused is not None and fused.engine is not None:  # dense / FM weights also live in the engine: push the loaded values there
        eng = fused.engine
        for i, d in enumerate(fused._dense_layers()):
            eng.set_dense_weights(i, d.kernel.data.cpu(), d.bias.data.cpu())
        eng.fm_w.copy_(fused.spec.fm.linear.data.reshape(-1))
        eng.fm_w0.copy_(fused.spec.fm.w_0.data.reshape(-1))
        fused._dirty = False


def export_pre_embd(model) -> Dict[str, np.ndarray]:
    """{feature name: table} for `FeaturePool(pre_embd=...)`: embedding layers are named "embd_" + feature (group.py:286)."""
    from .layers import CustomEmbedding

    model.sync()
    return {l.name[len("embd_"):]: l.embeddings.data.detach().cpu().numpy() for l in model._all_layers()
            if isinstance(l, CustomEmbedding) and l.name.startswith("embd_")}


# ---- row-sharded engine -------------------------------------------------------------------------
def save_sharded_tables(engine, directory: str) -> None:
    rank, world = engine.comm.rank, engine.world
    os.makedirs(directory, exist_ok=True)
    for t, w in enumerate(engine.tables):
        if t in engine.replicated:
            if rank == 0:
                np.save(os.path.join(directory, f"table{t:03d}.replicated.npy"), w.detach().cpu().numpy())
        else:
            np.save(os.path.join(directory, f"table{t:03d}.rank{rank}of{world}.npy"), w.detach().cpu().numpy())
    if rank == 0:
        np.save(os.path.join(directory, "dense_params.npy"), engine.params[: engine.n_dense_params].detach().cpu().numpy())


def load_sharded_tables(engine, directory: str) -> None:
    rank, world = engine.comm.rank, engine.world
    for t, w in enumerate(engine.tables):
        f = f"table{t:03d}.replicated.npy" if t in engine.replicated else f"table{t:03d}.rank{rank}of{world}.npy"
        a = np.load(os.path.join(directory, f))
        if tuple(a.shape) != tuple(w.shape):
            raise ValueError(f"{f}: shape {a.shape} does not match this rank's {tuple(w.shape)} (saved with another world size?)")
        w.copy_(torch.from_numpy(a))
    engine.params[: engine.n_dense_params].copy_(torch.from_numpy(np.load(os.path.join(directory, "dense_params.npy"))))
    engine._refresh_wt()
