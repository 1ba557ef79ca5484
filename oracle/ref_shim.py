"""Import the *unmodified* reference (read-only at /root/reference) in the build
container.  TEST INFRASTRUCTURE ONLY — used by oracle/make_golden.py to generate the
committed fixtures in tests/golden/.  /root/reference does not exist on the GPU box, so
nothing in tests/ (-m gpu), smoke() or bench.py may import this module at run time.

Five third-party modules the reference imports at module top level are absent from the
image and irrelevant to the hot path; they are stubbed in sys.modules:
torchlibrosa (SpecAugmentation is constructed, never called with specaug=False),
hydra, h5py, sentence_transformers, fire.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TAG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install_stubs():
    import torch.nn as nn

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "torchlibrosa" not in sys.modules:
        tl = mod("torchlibrosa")
        aug = mod("torchlibrosa.augmentation")

        class SpecAugmentation(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def forward(self, x):
                return x

        tl.SpecAugmentation = SpecAugmentation
        aug.SpecAugmentation = SpecAugmentation
        tl.augmentation = aug
    if "hydra" not in sys.modules:
        h = mod("hydra")
        hu = mod("hydra.utils")
        hu.instantiate = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        h.utils = hu
    if "h5py" not in sys.modules:
        mod("h5py")
    if "sentence_transformers" not in sys.modules:
        st = mod("sentence_transformers")
        st.SentenceTransformer = object
    if "fire" not in sys.modules:
        mod("fire")


def import_reference():
    """Returns a namespace with the reference classes on the cnn8rnn-w2vmean path."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our package also has a top-level module called `losses`-like names under its own
    # namespace only, so plain `models` / `losses` resolve to the reference here.
    import models.audio_encoder as audio_encoder
    import models.text_encoder as text_encoder
    import models.match as match
    import models.audio_text_model as audio_text_model
    import losses
    ns = types.SimpleNamespace(
        Cnn8Rnn=audio_encoder.Cnn8Rnn, EmbeddingAgg=text_encoder.EmbeddingAgg,
        DotProduct=match.DotProduct, BiEncoder=audio_text_model.BiEncoder,
        FrameBceLoss=losses.FrameBceLoss)
    return ns


def build_reference_model(ns, state_dict, vocab=5221):
    model = ns.BiEncoder(ns.Cnn8Rnn(32000), ns.EmbeddingAgg(vocab, 512), ns.DotProduct(), 512)
    missing, unexpected = model.load_state_dict(state_dict, strict=True)
    return model


def reference_runner_forward(model, batch, training=True):
    """Runner.forward of python_scripts/training/run_strong.py:92-120, restated around the
    imported reference model (run_strong.py itself needs fire/psds_eval, absent here)."""
    import torch
    b = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            b[k] = v.long() if k == "text" else v.float()
        else:
            b[k] = v
    input_dict = {"specaug": False}
    input_dict.update(b)
    output = model(input_dict)
    if training:
        label = b["label"]
        frame_sim = output["frame_sim"]
        trunc = min(frame_sim.size(1), label.size(1))
        output.update({
            "frame_sim": frame_sim[..., :trunc],
            "label": label[..., :trunc],
            "length": torch.clamp(output["length"], 1, trunc),
        })
    return output
