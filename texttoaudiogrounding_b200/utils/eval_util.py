"""GPU post-processing of frame probabilities — the per-sample, per-threshold host loop of the reference's
``Runner.eval_inference`` (python_scripts/training/run_strong.py:222-247) and its helpers ``median_filter``,
``connect_clusters``, ``find_contiguous_regions`` (utils/eval_util.py:18-116) as ONE kernel launch over all
(sample, threshold) pairs (csrc/postprocess.cu); results are bit-exact with the reference."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

from ..ops import call


def threshold_grid(n_thresholds: int) -> np.ndarray:
    """run_strong.py:205-206"""
    return np.arange(1 / (n_thresholds * 2), 1, 1 / n_thresholds)


def frame_regions(frame_sim: torch.Tensor, thresholds: Sequence[float], window_size: int, n_connect: int):
    """frame_sim [B, T] (CUDA) -> (regions int32 [B, n_th, R, 2], counts int32 [B, n_th]): regions[b, k, :counts[b, k]] are
    the (onset, offset) frame pairs of sample b at thresholds[k] after median filtering and cluster connection."""
    if not frame_sim.is_cuda:
        raise RuntimeError("frame_regions (B200) needs CUDA tensors: there is no CPU fallback")
    if frame_sim.ndim != 2:
        raise ValueError("frame_sim must be [B, T]")
    sim = frame_sim.float()
    if sim.stride(1) != 1:
        sim = sim.contiguous()
    B, T = sim.shape
    th = torch.as_tensor(np.asarray(thresholds, dtype=np.float64)).to(sim.device)
    max_regions = T // 2 + 1
    regions = torch.zeros(B, th.numel(), max_regions, 2, device=sim.device, dtype=torch.int32)
    counts = torch.empty(B, th.numel(), device=sim.device, dtype=torch.int32)
    call("tag_frame_regions", sim, sim.stride(0), th, B, T, th.numel(), int(window_size), int(n_connect), max_regions,
         regions, counts)
    return regions, counts


def predictions(frame_sim: torch.Tensor, filenames: List[str], n_thresholds: int, window_size: int,
                time_resolution: float) -> Dict[float, List[dict]]:
    """The ``pred_buffer`` of Runner.eval_inference for one batch: {threshold: [{"filename", "event_label", "onset",
    "offset"}]} with onset / offset in frames (the reference scales them by ``time_resolution`` afterwards)."""
    thresholds = threshold_grid(n_thresholds)
    n_connect = math.ceil(0.5 / time_resolution)
    regions, counts = frame_regions(frame_sim, thresholds, window_size, n_connect)
    regions, counts = regions.cpu().numpy(), counts.cpu().numpy()
    out = {th: [] for th in thresholds}
    for b, fname in enumerate(filenames):
        for k, th in enumerate(thresholds):
            for r in range(counts[b, k]):
                out[th].append({"filename": fname, "event_label": "fake_event",
                                "onset": int(regions[b, k, r, 0]), "offset": int(regions[b, k, r, 1])})
    return out
