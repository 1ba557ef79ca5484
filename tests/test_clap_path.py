"""BASELINE.json configs[4] — the CLAP graph behind the Hugging Face surface (reference
models/hf_modeling_grounding.py): Cnn8Rnn + LaionClapEncoder + DotProduct in BiEncoder(add_proj=True).
The fixture comes from the unmodified reference graph with a seeded, randomly initialised full-size text tower
(oracle/make_golden_clap.py).  CPU: the library modules rebuilt under the same seed reproduce the fixture.  GPU: the
row kernels against torch, the B200 tower (bf16 tensor-core GEMMs) and the whole façade against the fixture
(north_star bf16 bar: 1e-2 on frame_sim)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tag_oracle as O
from helpers import GOLDEN, rel_err

CASE = dict(batch=3, n_samples=32000, n_tokens=7, seed=8, data_seed=12, tower_seed=21)


def load():
    g = np.load(os.path.join(GOLDEN, "clap_b3_1s.npz"))
    batch = O.synth_clap_batch(CASE["batch"], CASE["n_samples"], CASE["n_tokens"], CASE["data_seed"])
    assert np.array_equal(batch["input_ids"].numpy(), g["input_ids"])
    return g, batch


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_clap_text_encoder_matches_reference():
    g, batch = load()
    tower, proj, _ = O.clap_case_modules(CASE["tower_seed"])
    with torch.no_grad():
        t = O.clap_text_encoder(tower, proj, batch["input_ids"], batch["attention_mask"])
    np.testing.assert_allclose(t["seq_emb"].numpy(), g["seq_emb"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(t["token_emb"].numpy(), g["token_emb"], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------- GPU
def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.gpu
@pytest.mark.parametrize("E", [512, 768])
def test_transformer_row_kernels_match_torch(E):
    from texttoaudiogrounding_b200.ops import call
    R = 37
    a, r = torch.randn(R, E, generator=_gen(1)), torch.randn(R, E, generator=_gen(2))
    gam, bet = torch.rand(E, generator=_gen(3)) + 0.5, torch.randn(E, generator=_gen(4))
    out = torch.empty(R, E, device="cuda")
    call("tag_add_layernorm", a.cuda(), r.cuda(), gam.cuda(), bet.cuda(), out, R, E, 1e-12)
    assert (out.cpu() - F.layer_norm(a + r, (E,), gam, bet, 1e-12)).abs().max().item() < 1e-5
    call("tag_add_layernorm", a.cuda(), None, gam.cuda(), bet.cuda(), out, R, E, 1e-5)
    assert (out.cpu() - F.layer_norm(a, (E,), gam, bet, 1e-5)).abs().max().item() < 1e-5
    call("tag_unary_f32", a.cuda(), out, a.numel(), 0)
    assert (out.cpu() - F.gelu(a)).abs().max().item() < 1e-6
    call("tag_unary_f32", a.cuda(), out, a.numel(), 1)
    assert (out.cpu() - torch.tanh(a)).abs().max().item() < 1e-6
    call("tag_l2_normalize", a.cuda(), out, R, E, 1e-12)
    assert (out.cpu() - F.normalize(a, dim=-1)).abs().max().item() < 1e-6
    # embeddings: position ids skip the pads (RoBERTa), LayerNorm on the sum
    B, L, V, P = 3, 40, 100, 64
    ids = torch.randint(2, V, (B, L), generator=_gen(5))
    ids[0, 30:] = 1
    ids[1, 5] = 1
    word, pos, typ = (torch.randn(n, E, generator=_gen(6 + i)) for i, n in enumerate((V, P, 1)))
    mask = ids.ne(1).int()
    pid = (torch.cumsum(mask, 1) * mask).long() + 1
    ref = F.layer_norm(word[ids] + typ[0] + pos[pid], (E,), gam, bet, 1e-12)
    out = torch.empty(B * L, E, device="cuda")
    call("tag_roberta_embed_ln", ids.cuda(), word.cuda(), pos.cuda(), typ.cuda(), gam.cuda(), bet.cuda(), out, B, L, E,
         V, P, 1, 1e-12)
    assert (out.cpu().view(B, L, E) - ref).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_clap_tower_and_facade_match_reference_golden():
    from transformers import ClapTextConfig
    from texttoaudiogrounding_b200.models.hf_modeling_grounding import (Cnn8RnnLaionClapGroundingConfig,
                                                                        Cnn8RnnLaionClapGroundingModel)
    g, batch = load()
    tower, proj, extra = O.clap_case_modules(CASE["tower_seed"])
    model = Cnn8RnnLaionClapGroundingModel(Cnn8RnnLaionClapGroundingConfig(text_encoder_name=ClapTextConfig()))
    sd = {k: v for k, v in O.synth_state_dict(seed=CASE["seed"], sharpen=1.0, perturb_bn=True).items()
          if k.startswith("audio_encoder.")}
    sd.update(extra)
    sd.update({"text_encoder.model." + k: v for k, v in tower.state_dict().items()})
    sd.update({"text_encoder.projection." + k: v for k, v in proj.state_dict().items()})
    model.model.load_state_dict(sd, strict=True)           # the reference graph's state-dict keys, all of them
    model = model.cuda().eval()
    tokens = {"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]}
    with torch.no_grad():
        t = model.model.text_encoder({k: v.cuda() for k, v in tokens.items()})
        sim = model(batch["waveform"], batch["waveform_len"], tokens)
    valid = torch.as_tensor(g["attention_mask"]).bool()
    assert rel_err(t["token_emb"].cpu()[valid], torch.as_tensor(g["token_emb"])[valid]) < 2e-2
    assert rel_err(t["seq_emb"].cpu(), g["seq_emb"]) < 2e-2
    assert (t["seq_emb"].norm(dim=-1).cpu() - 1).abs().max().item() < 1e-5
    assert tuple(sim.shape) == g["frame_sim"].shape
    assert np.abs(sim.cpu().numpy() - g["frame_sim"]).max() <= 1e-2
    # from the third call with the same (batch, length) the tower replays as one CUDA graph: same numbers, and new
    # token ids of that shape go through the static buffers
    enc = model.model.text_encoder
    dev_tokens = {k: v.cuda() for k, v in tokens.items()}
    for _ in range(3):
        again = enc(dev_tokens)
    assert tuple(tokens["input_ids"].shape) in enc._graphs
    assert torch.equal(again["seq_emb"], t["seq_emb"]) and torch.equal(again["token_emb"], t["token_emb"])
    other = {"input_ids": dev_tokens["input_ids"].flip(0).contiguous(),
             "attention_mask": dev_tokens["attention_mask"].flip(0).contiguous()}
    replayed = enc(other)
    enc.use_graph = False
    eager = enc(other)
    assert torch.equal(replayed["seq_emb"], eager["seq_emb"]) and torch.equal(replayed["token_emb"], eager["token_emb"])
