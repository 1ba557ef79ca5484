#!/usr/bin/env python
"""Forward-only (inference) throughput of the three model graphs on one B200, 10 s @ 32 kHz clips + 8-token phrases:
  * cnn8rnn-w2vmean, bs = 64, fp32 (BASELINE.json configs[1]) and bf16
  * cnn8rnn + CLAP text tower behind the HF surface, bs = 32, bf16 (configs[4]; random-init full-size tower)
CUDA events around N eval-mode forwards after warm-up, inputs resident in HBM.  Usage: bench_infer.py [json-out]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn  # noqa: E402
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder  # noqa: E402
from texttoaudiogrounding_b200.models.hf_modeling_grounding import (Cnn8RnnLaionClapGroundingConfig,  # noqa: E402
                                                                    Cnn8RnnLaionClapGroundingModel)
from texttoaudiogrounding_b200.models.match import DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg  # noqa: E402

L = 320000
ROWS = []


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    torch.manual_seed(0)
    for prec, B in (("fp32", 64), ("bf16", 64), ("bf16", 32)):
        model = BiEncoder(Cnn8Rnn(32000, compute_dtype=prec), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().eval()
        d = {"waveform": 0.1 * torch.randn(B, L, device="cuda"), "waveform_len": [L] * B, "specaug": False,
             "text": torch.randint(2, 5221, (B, 8), device="cuda"), "text_len": torch.full((B,), 8, device="cuda")}
        with torch.no_grad():
            ms = timeit(lambda: model(d))
        ROWS.append({"graph": "cnn8rnn-w2vmean forward", "dtype": prec, "batch": B, "ms": round(ms, 3),
                     "clips_per_s": round(B / ms * 1e3, 1), "algorithmic_tflops": round(B * 33.83 / ms, 1)})
        print(ROWS[-1])
        del model
    from transformers import ClapTextConfig
    B = 32
    model = Cnn8RnnLaionClapGroundingModel(Cnn8RnnLaionClapGroundingConfig(text_encoder_name=ClapTextConfig()))
    model = model.cuda().eval()
    wav = 0.1 * torch.randn(B, L, device="cuda")
    ids = torch.randint(4, 50000, (B, 10), device="cuda")
    ids[:, 0], ids[:, -1] = 0, 2
    tokens = {"input_ids": ids, "attention_mask": torch.ones(B, 10, dtype=torch.long, device="cuda")}
    with torch.no_grad():
        ms = timeit(lambda: model(wav, [L] * B, tokens))
        ms_text = timeit(lambda: model.model.text_encoder(tokens))
    ROWS.append({"graph": "cnn8rnn + CLAP text tower (HF facade) forward", "dtype": "bf16", "batch": B,
                 "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 1), "text_tower_ms": round(ms_text, 3)})
    print(ROWS[-1])
    if len(sys.argv) > 1:
        json.dump({"device": torch.cuda.get_device_name(0), "rows": ROWS}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
