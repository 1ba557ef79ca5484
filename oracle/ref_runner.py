"""The reference's own train / eval step on its own modules (staged under oracle/_ref by oracle/build_ref.py, or read
from /root/reference in the build container), on "cpu" or "cuda": the CPU baseline of `bench.py --impl reference`
and the `torch.cuda` eager comparator (`gpu_baseline`) of BASELINE.json configs[1].  The loop is the body of
Runner.train_epoch (python_scripts/training/run_strong.py:139-147) around Runner.forward (:92-120), which cannot be
imported itself (needs fire / psds_eval).  TEST / MEASUREMENT INFRASTRUCTURE ONLY — never imported by the product."""
import os
import time

from . import ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def reference_root():
    """oracle/_ref when staged (the only place it can be on the GPU box), else the build container's /root/reference."""
    if os.path.isfile(os.path.join(STAGED, "models", "audio_encoder.py")):
        return STAGED
    if ref_shim.available():
        return ref_shim.REFERENCE_ROOT
    return None


def available() -> bool:
    return reference_root() is not None


class ReferenceStep:
    """BiEncoder(Cnn8Rnn(32000), EmbeddingAgg(5221, 512), DotProduct(), 512) + FrameBceLoss + clip_grad_norm_(1.0) +
    Adam(lr=1e-3): the cnn8rnn-w2vmean strong-supervision configuration, built from the reference's classes."""

    def __init__(self, device="cpu", vocab=5221, seed=1, cudnn_benchmark=False, tf32=False, autocast_bf16=False):
        import torch
        root = reference_root()
        if root is None:
            raise RuntimeError("reference modules not available (run oracle/build_ref.py in the build container)")
        ref_shim.REFERENCE_ROOT = root
        ns = ref_shim.import_reference()
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.backends.cudnn.benchmark = bool(cudnn_benchmark)
            torch.backends.cudnn.deterministic = not cudnn_benchmark     # set_seed default (utils/train_util.py:44-45)
            torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
            torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.manual_seed(seed)
        self.model = ns.BiEncoder(ns.Cnn8Rnn(32000), ns.EmbeddingAgg(vocab, 512), ns.DotProduct(), 512).to(self.device)
        self.loss_fn = ns.FrameBceLoss()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=1e-3)
        self.autocast = bool(autocast_bf16)

    def _to_device(self, batch):
        torch = self.torch
        b = {}
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                b[k] = (v.long() if k == "text" else v.float()).to(self.device)      # run_strong.py:94-99
            else:
                b[k] = v
        return b

    def forward(self, b, training):
        """Runner.forward, run_strong.py:92-120 (batch already on the device)."""
        torch = self.torch
        input_dict = {"specaug": False}
        input_dict.update(b)
        if self.autocast:
            with torch.autocast(self.device.type, dtype=torch.bfloat16):
                output = self.model(input_dict)
            output["frame_sim"] = output["frame_sim"].float()
        else:
            output = self.model(input_dict)
        if training:
            label, frame_sim = b["label"], output["frame_sim"]
            trunc = min(frame_sim.size(1), label.size(1))
            output.update({"frame_sim": frame_sim[..., :trunc], "label": label[..., :trunc],
                           "length": torch.clamp(output["length"], 1, trunc)})
        return output

    def train_step(self, b):
        """run_strong.py:139-147"""
        torch = self.torch
        self.model.train()
        self.optimizer.zero_grad()
        output = self.forward(b, training=True)
        loss = self.loss_fn(output)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
        self.optimizer.step()
        return loss.item()

    def eval_step(self, b):
        torch = self.torch
        self.model.eval()
        with torch.no_grad():
            out = self.forward(b, training=False)
        return out["frame_sim"]

    def time(self, batch, steps, warmup, mode="train"):
        """Seconds per step (list of per-step wall times after ``warmup``; CUDA: synchronised around every step —
        the reference reads loss.item() every iteration anyway)."""
        torch = self.torch
        b = self._to_device(batch)
        fn = self.train_step if mode == "train" else self.eval_step
        times = []
        for i in range(warmup + steps):
            if self.device.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(b)
            if self.device.type == "cuda":
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return times
