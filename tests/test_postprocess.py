"""SURVEY.md §8f rank 4(iii) — frame probabilities -> event regions for all (sample, threshold) pairs.  Integer /
boolean work: BIT-EXACT against the fixture produced by the unmodified reference helpers (oracle/make_golden_post.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import GOLDEN

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
CASES = {"a": 0.04, "b": 0.02, "c": 0.04}


def synth_scores(seed, B, T):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, T // 5 + 2, generator=g).repeat_interleave(5, dim=1)[:, :T]
    return (0.7 * base + 0.3 * torch.rand(B, T, generator=g)).clamp(1e-7, 1.0)


def golden_rows(g, name):
    B, T, n_th, window, n_connect = (int(v) for v in g[f"{name}/cfg"])
    return (B, T, n_th, window, n_connect), [tuple(int(v) for v in r) for r in g[f"{name}/rows"]]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_regions_match_reference(name):
    g = np.load(os.path.join(GOLDEN, "post_regions.npz"))
    (B, T, n_th, window, n_connect), want = golden_rows(g, name)
    sim = synth_scores(17, B, T).numpy()
    thresholds = np.arange(1 / (n_th * 2), 1, 1 / n_th)
    got = [(b, k, s, e) for b in range(B) for k, th in enumerate(thresholds)
           for s, e in O.frame_regions(sim[b], th, window, n_connect)]
    assert got == want


def _kernel_rows(sim, thresholds, window, n_connect):
    from texttoaudiogrounding_b200.utils.eval_util import frame_regions
    regions, counts = frame_regions(sim.cuda(), thresholds, window, n_connect)
    regions, counts = regions.cpu().numpy(), counts.cpu().numpy()
    assert counts.max() <= regions.shape[2]
    return [(b, k, int(regions[b, k, r, 0]), int(regions[b, k, r, 1]))
            for b in range(sim.shape[0]) for k in range(len(thresholds)) for r in range(counts[b, k])]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_kernel_regions_bit_exact_with_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, "post_regions.npz"))
    (B, T, n_th, window, n_connect), want = golden_rows(g, name)
    thresholds = np.arange(1 / (n_th * 2), 1, 1 / n_th)
    assert _kernel_rows(synth_scores(17, B, T), thresholds, window, n_connect) == want


@pytest.mark.gpu
@pytest.mark.parametrize("window,n_connect", [(1, 0), (3, 2), (7, 13), (2, 1)])
def test_kernel_regions_match_oracle_edge_cases(window, n_connect):
    # all-off / all-on rows, values exactly AT a threshold (strict >), regions touching both ends
    T = 40
    sim = synth_scores(5, 6, T)
    sim[0] = 0.0
    sim[1] = 1.0
    sim[2, :3], sim[2, -2:] = 0.9, 0.9
    thresholds = np.array([0.05, 0.5, float(sim[3, 7]), 0.95])
    got = _kernel_rows(sim, thresholds, window, n_connect)
    want = [(b, k, s, e) for b in range(6) for k, th in enumerate(thresholds)
            for s, e in O.frame_regions(sim[b].numpy(), th, window, n_connect)]
    assert got == want


@pytest.mark.gpu
def test_predictions_schema_at_full_size():
    from texttoaudiogrounding_b200.utils.eval_util import predictions, threshold_grid
    sim = synth_scores(9, 64, 250)
    pred = predictions(sim.cuda(), [f"clip{b}_0" for b in range(64)], 50, 1, 0.04)
    ths = threshold_grid(50)
    assert list(pred.keys()) == list(ths) and len(ths) == 50
    # spot-check eight (sample, threshold) pairs of the full-size batch against the oracle
    for b, k in [(0, 0), (5, 10), (17, 25), (31, 26), (40, 33), (50, 40), (63, 49), (8, 24)]:
        want = O.frame_regions(sim[b].numpy(), ths[k], 1, 13)
        got = [(d["onset"], d["offset"]) for d in pred[ths[k]] if d["filename"] == f"clip{b}_0"]
        assert got == want
    # size-independent property: regions of one pair are sorted, disjoint and separated by more than n_connect frames
    for th, rows in pred.items():
        by = {}
        for d in rows:
            by.setdefault(d["filename"], []).append((d["onset"], d["offset"]))
        for regs in by.values():
            for (s0, e0), (s1, e1) in zip(regs, regs[1:]):
                assert s0 < e0 and s1 - e0 > 13
