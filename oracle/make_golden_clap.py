"""Generate tests/golden/clap_b3_1s.npz by running the UNMODIFIED reference Hugging Face graph
(models/hf_modeling_grounding.py: Cnn8Rnn :66-180, LaionClapEncoder :183-199, DotProduct :202-226,
BiEncoder(add_proj=True) :229-302) with a RANDOMLY INITIALISED full-size CLAP text tower (transformers
ClapTextModel / ClapProjectionLayer built from ClapTextConfig() under a fixed seed: the pretrained
laion/clap-htsat-fused weights cannot be downloaded here; LaionClapEncoder.__init__ is bypassed, its forward is the
reference's).  Build container only:   python oracle/make_golden_clap.py
TEST INFRASTRUCTURE ONLY (see oracle/make_golden.py)."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import tag_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASE = dict(batch=3, n_samples=32000, n_tokens=7, seed=8, data_seed=12, tower_seed=21)


def main():
    ref_shim.import_reference()
    import models.hf_modeling_grounding as hf
    tower, proj, extra = O.clap_case_modules(CASE["tower_seed"])
    batch = O.synth_clap_batch(CASE["batch"], CASE["n_samples"], CASE["n_tokens"], CASE["data_seed"])
    enc = hf.LaionClapEncoder.__new__(hf.LaionClapEncoder)
    nn.Module.__init__(enc)
    enc.model, enc.projection, enc.embed_dim = tower, proj, 512
    m = hf.BiEncoder(hf.Cnn8Rnn(32000), enc, hf.DotProduct(), 512, add_proj=True)
    sd = O.synth_state_dict(seed=CASE["seed"], sharpen=1.0, perturb_bn=True)
    sd = {k: v for k, v in sd.items() if k.startswith("audio_encoder.")}
    sd.update(extra)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("text_encoder.") for k in missing), (missing[:5], unexpected[:5])
    m.eval()
    with torch.no_grad():
        d = {"waveform": batch["waveform"], "waveform_len": batch["waveform_len"],
             "input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"],
             "text_len": batch["attention_mask"].sum(-1)}
        t = enc(d)
        o = m(d)
    out = {"input_ids": batch["input_ids"].numpy(), "attention_mask": batch["attention_mask"].numpy(),
           "seq_emb": t["seq_emb"].numpy(), "token_emb": t["token_emb"].numpy(),
           "frame_sim": o["frame_sim"].numpy(), "length": np.asarray(o["length"])}
    fs = o["frame_sim"].double().clamp(1e-12, 1 - 1e-12)
    lg = torch.log(fs / (1 - fs))
    print("frame_sim", tuple(o["frame_sim"].shape), "logits", lg.min().item(), lg.max().item())
    print("token_emb std", t["token_emb"].std().item(), "seq_emb norm", t["seq_emb"].norm(dim=-1))
    np.savez_compressed(os.path.join(OUT, "clap_b3_1s.npz"), **out)


if __name__ == "__main__":
    main()
