"""Generate tests/golden/post_regions.npz with the UNMODIFIED reference post-processing helpers (utils/eval_util.py:
median_filter, connect_clusters, find_contiguous_regions) driven exactly as Runner.eval_inference does
(python_scripts/training/run_strong.py:231-241).  eval_util imports plotting / metric packages that are absent here
(matplotlib, sed_eval, psds_eval, sed_scores_eval); they are stubbed — none is touched by these helpers.
Build container only:   python oracle/make_golden_post.py          TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import tag_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_eval_util():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for name in ("matplotlib", "matplotlib.pyplot", "sed_eval"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name)
    stub("psds_eval", PSDSEval=object, plot_psd_roc=None)
    stub("psds_eval.psds", WORLD=None, PSDSEvalError=Exception)
    stub("sed_scores_eval", intersection_based=None)
    stub("sed_scores_eval.utils")
    stub("sed_scores_eval.utils.auc", staircase_auc=None)
    sys.path.insert(0, "/root/reference")
    import utils.eval_util as eu
    return eu


def synth_scores(seed: int, B: int, T: int):
    """smooth-ish probabilities with plateaus, so regions, gaps and filter effects all occur"""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, T // 5 + 2, generator=g).repeat_interleave(5, dim=1)[:, :T]
    return (0.7 * base + 0.3 * torch.rand(B, T, generator=g)).clamp(1e-7, 1.0)


def main():
    eu = import_eval_util()
    out = {}
    for name, (B, T, n_th, window, res) in {"a": (5, 250, 50, 1, 0.04), "b": (4, 101, 10, 5, 0.02),
                                            "c": (3, 63, 7, 4, 0.04)}.items():
        sim = synth_scores(17, B, T)
        thresholds = np.arange(1 / (n_th * 2), 1, 1 / n_th)
        n_connect = int(np.ceil(0.5 / res))
        rows = []
        for b in range(B):
            for k, th in enumerate(thresholds):
                filtered = eu.median_filter(sim[b].unsqueeze(0).cpu(), window_size=window, threshold=th)[0]
                reg = eu.find_contiguous_regions(eu.connect_clusters(filtered, n_connect))
                for r in reg:
                    rows.append((b, k, int(r[0]), int(r[1])))
                # the oracle restatement agrees pair by pair
                assert [tuple(int(v) for v in r) for r in reg] == O.frame_regions(sim[b].numpy(), th, window, n_connect)
        out[f"{name}/rows"] = np.asarray(rows, dtype=np.int64)
        out[f"{name}/cfg"] = np.asarray([B, T, n_th, window, n_connect], dtype=np.int64)
        print(name, len(rows), "regions")
    np.savez_compressed(os.path.join(OUT, "post_regions.npz"), **out)


if __name__ == "__main__":
    main()
