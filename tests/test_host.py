"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the module
classes mirror the reference's surface (state-dict keys/shapes, config instantiation, error
behaviour) — no kernel is launched here."""
import ctypes
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import tag_oracle as O


def test_library_builds_and_exports_every_header_symbol():
    from texttoaudiogrounding_b200 import _lib
    _lib.build()
    protos = _lib.parse_header()
    assert len(protos) >= 29
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), name
    assert _lib.lib().tag_version() == 100


def test_no_cpu_fallback():
    from helpers import build_model
    from texttoaudiogrounding_b200.losses import FrameBceLoss
    model = build_model(None, "fp32", device="cpu")
    batch = O.synth_batch(1, 32000)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model({"specaug": False, **batch})
    with pytest.raises(KeyError):
        model({k: v for k, v in batch.items()})          # 'specaug' is a required key (reference quirk)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FrameBceLoss()({"frame_sim": torch.rand(2, 5), "label": torch.rand(2, 5), "length": [5, 5]})


def test_state_dict_keys_and_shapes_match_reference():
    from helpers import build_model
    model = build_model(None, "fp32", device="cpu")
    sd = model.state_dict()
    spec = {k: tuple(shape) for k, shape, _ in O.state_dict_spec()}
    assert set(sd) == set(spec)
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k], k
    assert sum(p.numel() for p in model.parameters()) == 8804800
    # frontend buffers equal torchaudio's (restated in the oracle, pinned in test_oracle_golden)
    np.testing.assert_allclose(sd["audio_encoder.melspec_extractor.mel_scale.fb"].numpy(),
                               O.melscale_fbanks().numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(sd["audio_encoder.melspec_extractor.spectrogram.window"].numpy(),
                               O.hann_window().numpy(), atol=1e-7)
    # load a reference-layout (NCHW-contiguous) state dict, read it back unchanged
    ref_sd = O.synth_state_dict(seed=7)
    model.load_state_dict(ref_sd, strict=True)
    w = model.audio_encoder.conv_block2.conv1.weight
    assert torch.equal(w.detach(), ref_sd["audio_encoder.conv_block2.conv1.weight"])
    assert w.permute(0, 2, 3, 1).is_contiguous()       # kernels see [Cout][kh][kw][Cin]


def test_initialisation_matches_reference_for_same_seed():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "init_seed1.npz"))
    from helpers import build_model
    torch.manual_seed(1)
    model = build_model(None, "fp32", device="cpu")
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(v.double().sum().item(), g["sum/" + k].item(), rtol=1e-6, atol=1e-6, err_msg=k)
            np.testing.assert_allclose(v.double().abs().sum().item(), g["abs/" + k].item(), rtol=1e-6, err_msg=k)


def test_run_strong_dialect_config_instantiates_b200_classes(tmp_path):
    from texttoaudiogrounding_b200.utils import train_util
    train_util.install_as_reference_modules()
    cfg = {
        "model": {
            "audio_encoder": {"type": "models.audio_encoder.Cnn8Rnn", "args": {"sample_rate": 32000}},
            "text_encoder": {"type": "models.text_encoder.EmbeddingAgg",
                             "args": {"vocab_size": 50, "embed_dim": 512, "aggregation": "mean"}},
            "match_fn": {"type": "models.match.DotProduct", "args": {}},
            "type": "models.audio_text_model.BiEncoder", "args": {"shared_dim": 512},
        },
        "loss": {"type": "losses.FrameBceLoss", "args": {}},
        "optimizer": {"type": "torch.optim.Adam", "args": {"lr": 0.001}},
        "trainer": {"max_grad_norm": 1.0},
    }
    base = tmp_path / "base.yaml"
    base.write_text(yaml.dump(cfg))
    child = tmp_path / "child.yaml"
    child.write_text(yaml.dump({"inherit_from": "base.yaml", "trainer": {"max_grad_norm": 2.0}}))
    config = train_util.parse_config_or_kwargs(str(child), **{"optimizer.args.lr": 0.01})
    assert config["trainer"]["max_grad_norm"] == 2.0 and config["optimizer"]["args"]["lr"] == 0.01
    model = train_util.get_model(config)
    import texttoaudiogrounding_b200.models.audio_encoder as ae
    assert isinstance(model.audio_encoder, ae.Cnn8Rnn)
    assert model.audio_encoder.embed_dim == 512 and model.audio_encoder.downsample_ratio == 4
    loss = train_util.init_obj_from_str(config["loss"])
    opt = train_util.init_obj_from_str(config["optimizer"], params=model.parameters())
    assert type(loss).__name__ == "FrameBceLoss" and opt.defaults["lr"] == 0.01
    # no CPU fallback: the mirrored modules refuse host tensors instead of computing with torch ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        train_util.init_obj_from_str({"type": "models.match.DotProduct", "args": {"text_level": "token"}})(
            {"audio_emb": torch.zeros(1, 2, 4), "text_emb": {"token_emb": torch.zeros(1, 2, 4)}})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        train_util.init_obj_from_str({"type": "models.match.DotProduct", "args": {"l2norm": True}})(
            {"audio_emb": torch.zeros(1, 2, 4), "text_emb": {"seq_emb": torch.zeros(1, 4)}})
    # the class name the reference's eg_configs use for the audio encoder resolves as well
    import texttoaudiogrounding_b200.models.audio_encoder as ae2
    assert train_util.get_obj_from_str("models.audio_encoder.Cnn8_Rnn") is ae2.Cnn8Rnn
    # the later-config components resolve through the same registry
    for typ, args in (("models.text_encoder.SelfAttention", {"vocab_size": 50, "embed_dim": 512, "num_heads": 8}),
                      ("models.match.CrossAttention", {"embed_dim": 512, "num_heads": 8, "dropout": 0.2}),
                      ("models.cross_encoder.CrossAttentionGating", {"embed_dim": 512}),
                      ("models.align.DotProduct", {"l2norm": False, "scaled": False}),
                      ("models.match.ExpNegL2", {"text_level": "seq"}),
                      ("models.sim_pooling.AudioMeanTextMean", {}),
                      ("losses.MaxMarginRankingLoss", {"margin": 1})):
        obj = train_util.init_obj_from_str({"type": typ, "args": args})
        assert type(obj).__module__.startswith("texttoaudiogrounding_b200."), typ


def test_freeze_and_train_mode_semantics():
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    enc = Cnn8Rnn(32000, freeze_cnn=True, freeze_bn=True)
    assert all(p.requires_grad for p in enc.rnn.parameters())
    assert not enc.conv_block1.conv1.weight.requires_grad and not enc.fc1.weight.requires_grad
    enc.train()
    assert enc.training and not enc.bn0.training and not enc.conv_block3.bn2.training
    length = torch.div(torch.div(torch.as_tensor([64000, 32000]), 320, rounding_mode="floor") + 1, 4,
                       rounding_mode="floor")
    assert length.tolist() == [50, 25]


def test_dict_tokenizer_matches_reference_behaviour():
    """reference datasets/text_tokenizer.py:9-58: whitespace tokens, <unk> for unknown words, zero padding,
    List[List[str]] -> [B, text_num, N]."""
    from texttoaudiogrounding_b200.datasets.text_tokenizer import DictTokenizer
    vocab = {"<pad>": 0, "<unk>": 1, "a": 2, "dog": 3, "barks": 4, "rain": 5}
    tok = DictTokenizer(vocab)
    out = tok(["a dog barks", "rain", "a cat"])
    assert out["text"].tolist() == [[2, 3, 4], [5, 0, 0], [2, 1, 0]]
    assert out["text_len"].tolist() == [3, 1, 2]
    out = tok([["a dog", "rain"], ["barks", "a dog barks"]])
    assert tuple(out["text"].shape) == (2, 2, 3) and out["text_len"].tolist() == [[2, 1], [1, 3]]
    assert tok.inverse_transform(out["text"][1].tolist()) == ["barks", "a dog barks"]


def test_hf_style_facade_builds_and_refuses_cpu():
    from texttoaudiogrounding_b200.models.hf_modeling_grounding import (
        Cnn8RnnW2vMeanGroundingConfig, Cnn8RnnW2vMeanGroundingModel)
    cfg = Cnn8RnnW2vMeanGroundingConfig(vocabulary={"<pad>": 0, "<unk>": 1, "dog": 2})
    model = Cnn8RnnW2vMeanGroundingModel(cfg)
    assert model.model.text_encoder.embedding.core.weight.shape == (3, 512)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 32000), [32000], ["dog"])          # CPU tensors: no fallback


def test_checkpoint_interchange_key_and_shape_matching(tmp_path):
    """Reference checkpoint format {"model": state_dict, "epoch", ...} (run_strong.py:679-690) and the loaders'
    key-AND-shape matching (utils/train_util.py:219-297, models/base.py:9-47; PANNs layout of
    models/audio_encoder.py:153-160): matching tensors are taken, everything else is skipped silently."""
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder
    from texttoaudiogrounding_b200.models.match import DotProduct
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    sd = O.synth_state_dict(seed=11, vocab=50)
    path = tmp_path / "ckpt.pth"
    torch.save({"model": sd, "epoch": 3, "metric_monitor": 0.5, "not_improve_cnt": 0}, path)
    model = BiEncoder(Cnn8Rnn(32000), EmbeddingAgg(50, 512), DotProduct(), 512)
    msgs = []
    model.load_pretrained(str(path), msgs.append)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # a PANNs-style audio checkpoint: weights under ["model"] with the encoder's own keys, one tensor of another
    # shape (skipped) and one unknown key (skipped)
    enc_sd = {k[len("audio_encoder."):]: v + 1.0 if v.is_floating_point() else v for k, v in sd.items()
              if k.startswith("audio_encoder.")}
    enc_sd["fc1.weight"] = torch.zeros(7, 7)
    enc_sd["fc_audioset.weight"] = torch.zeros(447, 512)
    ppath = tmp_path / "panns.pth"
    torch.save({"model": enc_sd}, ppath)
    enc = Cnn8Rnn(32000, pretrained=str(ppath), output_fn=msgs.append)
    assert torch.equal(enc.state_dict()["bn0.weight"], enc_sd["bn0.weight"])
    assert torch.equal(enc.state_dict()["rnn.weight_hh_l0_reverse"], enc_sd["rnn.weight_hh_l0_reverse"])
    assert enc.state_dict()["fc1.weight"].shape == (512, 512)            # mismatched shape: model value kept
    assert "fc_audioset.weight" not in enc.state_dict()


def test_clap_facade_is_a_pretrained_model_and_round_trips(tmp_path):
    """reference models/hf_modeling_grounding.py:305-352: PretrainedConfig / PreTrainedModel with the same fields and
    parameter names, so save_pretrained / from_pretrained directories interchange (tiny random tower here)."""
    from transformers import ClapTextConfig, PreTrainedModel
    from texttoaudiogrounding_b200.models.hf_modeling_grounding import (Cnn8RnnLaionClapGroundingConfig,
                                                                        Cnn8RnnLaionClapGroundingModel)
    cfg = Cnn8RnnLaionClapGroundingConfig(
        text_encoder_name=ClapTextConfig(num_hidden_layers=1, vocab_size=120, max_position_embeddings=40))
    assert cfg.sample_rate == 32000 and cfg.shared_dim == 512 and cfg.text_encoder_name == "laion/clap-htsat-fused"
    model = Cnn8RnnLaionClapGroundingModel(cfg)
    assert isinstance(model, PreTrainedModel)
    keys = set(model.state_dict())
    for k in ("model.audio_encoder.conv_block1.conv1.weight", "model.audio_encoder.rnn.weight_hh_l0_reverse",
              "model.text_encoder.model.embeddings.word_embeddings.weight",
              "model.text_encoder.model.encoder.layer.0.attention.self.query.weight",
              "model.text_encoder.projection.linear2.bias", "model.audio_proj.weight", "model.text_proj.bias"):
        assert k in keys, k
    model.save_pretrained(tmp_path)
    assert (tmp_path / "config.json").exists()
    again = Cnn8RnnLaionClapGroundingModel.from_pretrained(tmp_path)
    sd1, sd2 = model.state_dict(), again.state_dict()
    assert set(sd1) == set(sd2) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    # the README call: AutoModel.from_pretrained(<checkpoint dir>) after registering the B200 classes
    from transformers import AutoModel
    from texttoaudiogrounding_b200.models.hf_modeling_grounding import register_auto_classes
    register_auto_classes()
    auto = AutoModel.from_pretrained(tmp_path)
    assert type(auto) is Cnn8RnnLaionClapGroundingModel and auto.config.sample_rate == 32000
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        again(torch.zeros(1, 32000), [32000], {"input_ids": torch.tensor([[0, 5, 2]]),
                                                 "attention_mask": torch.ones(1, 3, dtype=torch.long)})


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores) must print one JSON line with the same
    metric / unit / config as the B200 arm plus the cpu_baseline and e2e objects; it needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    assert line["impl"] == "reference" and line["metric"] == base["metric"] and line["unit"] == "clips/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert "configs[2]" in line["config"]["workload"]
    # the line describes what was run: exactly the requested steps, on a stated sample of the batch, on no GPU
    assert line["steps"] == 1 and line["warmup"] == 0 and line["gpus_used"] == 0 and line["clips_per_step"] == 8
    assert abs(line["value"] - line["clips_per_step"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    staged = os.path.isfile(os.path.join(root, "oracle", "_ref", "models", "audio_encoder.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged or os.path.isdir("/root/reference/models") else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_fused_step_refuses_configurations_it_does_not_compute():
    """FusedTrainStep hard-codes mean-embedding -> scaled dot -> sigmoid; any other head must raise instead of being
    trained as that.  Host logic only: CPU tensors, no kernels."""
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder, MultiTextBiEncoder
    from texttoaudiogrounding_b200.models.match import DotProduct, ExpNegL2
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    from texttoaudiogrounding_b200.train import FusedTrainStep, WeakFusedTrainStep

    def enc():
        return Cnn8Rnn(32000, compute_dtype="fp32")

    bad = [
        BiEncoder(enc(), EmbeddingAgg(50, 512), ExpNegL2(), 512),
        BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(l2norm=True), 512),
        BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(text_level="token"), 512),
        BiEncoder(enc(), EmbeddingAgg(50, 512, aggregation="attention"), DotProduct(), 512),
        BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(), 512, add_proj=True),
        BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(), 512, upsample=True),
    ]
    for m in bad:
        with pytest.raises(NotImplementedError):
            FusedTrainStep(m, use_graph=False)
    with pytest.raises(NotImplementedError):
        WeakFusedTrainStep(BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(), 512), use_graph=False)
    with pytest.raises(NotImplementedError):
        WeakFusedTrainStep(MultiTextBiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(), 512, ["text"], add_proj=True),
                           use_graph=False)
    FusedTrainStep(BiEncoder(enc(), EmbeddingAgg(50, 512), DotProduct(), 512), use_graph=False)


def test_fused_step_optimizer_state_round_trips_through_torch_adam():
    """optimizer_state_dict() is torch.optim.Adam's layout in model.parameters() order (what the reference stores
    with include_optim_in_ckpt, run_strong.py:679-690): torch.optim.Adam loads it, and what Adam saves loads back."""
    from helpers import build_model
    from texttoaudiogrounding_b200.train import FusedTrainStep
    torch.manual_seed(0)
    model = build_model(None, "fp32", device="cpu", vocab=40)
    ts = FusedTrainStep(model, lr=2e-3, use_graph=False)
    assert ts.optimizer_state_dict()["state"] == {}
    ts.flat_m.copy_(torch.randn(ts.n_params))
    ts.flat_v.copy_(torch.rand(ts.n_params))
    ts.step_dev.fill_(7)
    ts.set_lr(5e-4)
    assert ts.lr == 5e-4 and float(ts.lr_dev) == pytest.approx(5e-4)
    sd = ts.optimizer_state_dict()
    opt = torch.optim.Adam(model.parameters(), lr=1.0)
    opt.load_state_dict(sd)
    assert opt.param_groups[0]["lr"] == 5e-4
    w = model.audio_encoder.conv_block2.conv1.weight
    off, k = ts._views[[id(p) for p in ts._params].index(id(w))]
    assert torch.equal(opt.state[w]["exp_avg"].permute(0, 2, 3, 1).reshape(-1), ts.flat_m[off:off + k])
    assert int(opt.state[w]["step"]) == 7
    torch.manual_seed(0)
    model2 = build_model(None, "fp32", device="cpu", vocab=40)
    ts2 = FusedTrainStep(model2, use_graph=False)
    ts2.load_optimizer_state_dict(opt.state_dict())
    assert torch.equal(ts2.flat_m, ts.flat_m) and torch.equal(ts2.flat_v, ts.flat_v)
    assert int(ts2.step_dev) == 7 and ts2.lr == 5e-4
