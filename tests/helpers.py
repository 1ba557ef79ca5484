"""Shared helpers for the parity tests: build the B200 model with the oracle's synthetic weights,
compare against the golden fixtures generated from the unmodified reference."""
import os

import numpy as np
import torch

from oracle import tag_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "cfg1_b4_2s": dict(batch=4, n_samples=64000, ragged=False, seed=1, data_seed=0),
    "ragged_b4_1s": dict(batch=4, n_samples=32000, ragged=True, seed=2, data_seed=3),
}
SHARPEN = 300.0


def sub(t, n=512):
    flat = t.detach().float().cpu().reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].numpy()


def load_case(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = O.synth_state_dict(seed=cfg["seed"], sharpen=SHARPEN, perturb_bn=True)
    batch = O.synth_batch(cfg["batch"], cfg["n_samples"], seed=cfg["data_seed"], ragged=cfg["ragged"])
    return g, sd, batch


def build_model(sd=None, dtype="fp32", device="cuda", vocab=O.VOCAB):
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder
    from texttoaudiogrounding_b200.models.match import DotProduct
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    model = BiEncoder(Cnn8Rnn(32000, compute_dtype=dtype), EmbeddingAgg(vocab, 512), DotProduct(), 512)
    if sd is not None:
        model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return model.to(device)


def nchw(t):
    """[B,H,W,C] kernel layout -> the reference's [B,C,H,W]"""
    return t.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = np.linalg.norm(a) * np.linalg.norm(b)
    return float((a * b).sum() / d) if d > 0 else 1.0
