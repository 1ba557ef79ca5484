"""Training-step plumbing of the strong-supervision runner.

* ``runner_forward`` / ``train_step`` mirror Runner.forward and the body of Runner.train_epoch
  (reference python_scripts/training/run_strong.py:92-120 and :139-147) on top of the mirrored
  nn.Modules (autograd path; any torch optimizer works).
* ``FusedTrainStep`` is the production path: all parameters, gradients and Adam moments live
  in flat fp32 buffers, the whole step (forward, backward, clip_grad_norm_, Adam) is a fixed
  sequence of libtag_b200 kernels captured in CUDA graphs, and data parallelism is one NCCL
  all-reduce of the flat gradient bucket per step (batches shard by clip; BN statistics stay
  per rank, as in the reference's single-process semantics).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import engine, ops
from .losses import FrameBceLoss
from .ops import call


def runner_forward(model: nn.Module, batch: Dict, device, training: bool = True) -> Dict:
    """Runner.forward (run_strong.py:92-120): move tensors, inject specaug=False, run the model,
    truncate frame_sim/label to a common length and clamp length."""
    b = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            b[k] = v.long().to(device) if k == "text" else v.float().to(device)
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


def train_step(model: nn.Module, batch: Dict, optimizer: torch.optim.Optimizer, loss_fn=None,
               max_grad_norm: float = 1.0, device="cuda"):
    """One iteration of Runner.train_epoch (run_strong.py:139-147)."""
    loss_fn = loss_fn or FrameBceLoss()
    optimizer.zero_grad()
    output = runner_forward(model, batch, device, training=True)
    loss = loss_fn(output)
    loss.backward()
    total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    return loss, total_norm


class LossHandle:
    """Pinned host slot + the event that marks its device->host copy complete."""

    def __init__(self, slot: torch.Tensor, event: torch.cuda.Event):
        self.slot, self.event = slot, event

    def result(self) -> float:
        self.event.synchronize()
        return float(self.slot)


class FusedTrainStep:
    """Flat-buffer, CUDA-graph-captured train step for BiEncoder(Cnn8Rnn, EmbeddingAgg, DotProduct)."""

    MAX_GRAPHS = 8          # captured (shape -> graphs) entries kept, least recently used evicted first

    def __init__(self, model: nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 1.0, process_group=None, use_graph: bool = True,
                 base_seed: int = 1):
        self.model = model
        self.enc = model.audio_encoder
        self.txt = model.text_encoder
        self._check_model(model)
        self.betas, self.eps, self.max_grad_norm = betas, eps, max_grad_norm
        self.pg = process_group
        self.world, self.rank = 1, 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        self.use_graph = use_graph
        # every rank draws its own dropout masks (same base seed on all ranks would repeat the masks N times)
        self.base_seed = base_seed + 7919 * self.rank
        self.scale = 1.0 / (self.enc.embed_dim ** 0.5) if getattr(model.match_fn, "scale", True) else 1.0
        self.device = next(model.parameters()).device
        self._flatten()
        # the learning rate lives on the device: a scheduler (run_strong.py:136-137) acts on replayed graphs too
        self.lr_dev = torch.full((1,), float(lr), device=self.device, dtype=torch.float32)
        self._lr = float(lr)
        if self.world > 1:
            self._sync_replicas()
        self._graphs = OrderedDict()      # shape key -> {"fwd_bwd", "optim", "static"}
        self._seen = {}                   # shape key -> number of steps run with it
        self._graph_pool = None
        self._static = None
        self._staging = None
        self._prefetched = None
        self._calls = 0
        # BatchNorm's num_batches_tracked (an int64 bookkeeping buffer no kernel reads) is counted on the host and
        # written back by flush_counters() / before state_dict(): nine 1-element kernels per step otherwise
        self._nbt_pending = 0
        self._loss_slots = None
        model.register_state_dict_pre_hook(lambda *a, **k: self.flush_counters())
        # weight-gradient GEMMs on a second stream (TAG_B200_OVERLAP=1)
        self.side_stream = torch.cuda.Stream(self.device) if os.environ.get("TAG_B200_OVERLAP", "0") == "1" else None
        # TAG_B200_DETERMINISTIC_WGRAD=1: two-pass split-K (slabs + ordered reduce) instead of fp32 atomics for the
        # tensor-core weight gradients; the workspace is shared, so the weight gradients stay on the main stream
        if os.environ.get("TAG_B200_DETERMINISTIC_WGRAD", "0") == "1" and self.device.type == "cuda":
            ops.set_deterministic_wgrad(True, self.device)
            self.side_stream = None
        # Data parallel: ONE logical all-reduce of the flat gradient bucket per step, issued inside the captured forward +
        # backward graph (NCCL collectives are graph-capturable; no host round trip between the two graphs).
        # Default: a single call as the last node.  TAG_B200_AR_OVERLAP=1 issues it as two calls: the bucket is ordered
        # [bn0, block-1 BatchNorms and convolutions | everything else]; the tail (99.9 % of the bytes) is final once conv
        # block 2 has run its backward and is reduced on a communication stream forked inside the captured region while
        # block 1 (1.3 ms of kernels) runs.  The convolution kernels are persistent with one CTA per SM (all registers,
        # ~200 KB of shared memory) and NCCL's CTAs cannot co-reside with them, so for that window the persistent kernels
        # are launched on SM_RESERVE fewer SMs (tag_set_sm_reserve) and NCCL is expected to be capped to as many CTAs
        # (NCCL_MAX_CTAS, set by bench.py before the process group is created).
        self._ar_split = self._views[self._n_late][0]
        self._overlap_ar = self.world > 1 and os.environ.get("TAG_B200_AR_OVERLAP", "0") == "1"
        self._sm_reserve = int(os.environ.get("TAG_B200_AR_SM_RESERVE", "8"))
        self._ar_in_graph = True
        self._comm_stream = None
        self._ar_done = None

    # ------------------------------------------------------------------ what this step computes
    def _check_model(self, model):
        """The fused step hard-codes: masked-mean word embeddings -> scaled dot product -> sigmoid -> clamp.  Any
        other head must go through the autograd path (``train_step``), never silently through this one."""
        from .models.match import DotProduct
        from .models.text_encoder import EmbeddingAgg
        mf = model.match_fn
        problems = []
        if type(mf) is not DotProduct or mf.l2norm or mf.text_level != "seq":
            problems.append("match_fn must be models.match.DotProduct(l2norm=False, text_level='seq')")
        if not isinstance(self.txt, EmbeddingAgg) or self.txt.agg != "mean":
            problems.append("text_encoder must be EmbeddingAgg(aggregation='mean')")
        if getattr(model, "cross_encoder", None) is not None:
            problems.append("cross_encoder is not None")
        if hasattr(model, "audio_proj") or hasattr(model, "text_proj"):
            problems.append("audio_proj / text_proj (add_proj or mismatched embed dims)")
        if getattr(model, "upsample", False):
            problems.append("upsample=True")
        if problems:
            raise NotImplementedError(f"{type(self).__name__} does not implement this configuration ("
                                      + "; ".join(problems) + "): use train.train_step (autograd path)")

    # ------------------------------------------------------------------ learning rate / optimizer state
    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float) -> None:
        self.set_lr(value)

    def set_lr(self, value: float) -> None:
        """Takes effect at the next step, captured graphs included (the kernel reads the device scalar)."""
        self._lr = float(value)
        self.lr_dev.fill_(self._lr)

    def _param_views(self, flat: torch.Tensor):
        """Views of a flat buffer shaped like the parameters, in the order of ``_flatten``."""
        out = []
        for p, (off, k) in zip(self._params, self._views):
            if p.dim() == 4:
                co, ci, kh, kw = p.shape
                out.append(flat[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2))
            else:
                out.append(flat[off:off + k].view(p.shape))
        return out

    def optimizer_state_dict(self) -> Dict:
        """The optimizer state in ``torch.optim.Adam.state_dict()`` layout with parameters indexed in
        ``model.parameters()`` order, i.e. what the reference stores under ``"optimizer"`` with
        ``include_optim_in_ckpt`` (run_strong.py:679-690) — loadable by ``torch.optim.Adam(model.parameters())``."""
        m = {id(p): v for p, v in zip(self._params, self._param_views(self.flat_m))}
        v = {id(p): x for p, x in zip(self._params, self._param_views(self.flat_v))}
        step = self.step_dev.to(torch.float32).cpu().reshape(())
        state, idx = {}, []
        for i, p in enumerate(self.model.parameters()):
            idx.append(i)
            if int(step) > 0:
                state[i] = {"step": step.clone(), "exp_avg": m[id(p)].detach().clone().contiguous(),
                            "exp_avg_sq": v[id(p)].detach().clone().contiguous()}
        group = {"lr": self._lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": idx}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: Dict) -> None:
        group = sd["param_groups"][0]
        self.set_lr(group["lr"])
        self.betas, self.eps = tuple(group["betas"]), group["eps"]
        m = {id(p): x for p, x in zip(self._params, self._param_views(self.flat_m))}
        v = {id(p): x for p, x in zip(self._params, self._param_views(self.flat_v))}
        steps = set()
        for i, p in zip(group["params"], self.model.parameters()):
            st = sd["state"].get(i)
            if st is None:
                m[id(p)].zero_(); v[id(p)].zero_()
                continue
            m[id(p)].copy_(st["exp_avg"]); v[id(p)].copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise NotImplementedError("per-parameter Adam step counts differ; the fused step keeps one counter")
        self.step_dev.fill_(steps.pop() if steps else 0)
        if self._graphs:
            self._graphs.clear()          # betas / eps are baked into the captured optimizer graph

    def _sync_replicas(self) -> None:
        """Data-parallel replicas start identical: parameters, Adam moments and BatchNorm buffers of rank 0."""
        dist = torch.distributed
        for t in (self.flat_p, self.flat_m, self.flat_v, self.step_dev):
            dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        for bn in self.enc._bns():
            for buf in (bn.running_mean, bn.running_var, bn.num_batches_tracked):
                if buf is not None:
                    dist.broadcast(buf, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                                   group=self.pg)

    # ------------------------------------------------------------------ flat buffers
    def _flatten(self):
        params = self.enc._param_list() + [self.txt.embedding.core.weight]
        # bucket order: the parameters whose gradients are written last by the backward pass (bn0 and conv block 1: its
        # two BatchNorms and two convolutions) sit at the head, everything else — final once block 2 has run its
        # backward — forms one contiguous tail (see _early_allreduce)
        late = params[0:6] + params[18:20]
        late_ids = {id(p) for p in late}
        params = late + [p for p in params if id(p) not in late_ids]
        self._n_late = len(late)
        if not all(p.requires_grad for p in params):
            raise NotImplementedError("FusedTrainStep trains all parameters; use the autograd path "
                                      "(train_step) with frozen sub-modules")
        known = {id(p) for p in params}
        extra = [n for n, p in self.model.named_parameters() if id(p) not in known]
        if extra:
            raise NotImplementedError(f"parameters outside the fused path: {extra}")
        n = sum(p.numel() for p in params)
        dev = self.device
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.n_params = n
        self._params = params
        off = 0
        self._views = []
        for p in params:
            k = p.numel()
            if p.dim() == 4:      # conv weight: keep [Cout][kh][kw][Cin] memory, [Cout,Cin,3,3] shape
                co, ci, kh, kw = p.shape
                pv = self.flat_p[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
                gv = self.flat_g[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                pv = self.flat_p[off:off + k].view(p.shape)
                gv = self.flat_g[off:off + k].view(p.shape)
            pv.copy_(p.data)
            p.data = pv
            p.grad = gv
            self._views.append((off, k))
            off += k
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.norm_out = torch.zeros(1, device=dev, dtype=torch.float32)
        self.loss_out = torch.zeros((), device=dev, dtype=torch.float32)
        self.Wt = self.enc._weights()
        assert self.Wt.w_ih.data_ptr() == self.enc.rnn.weight_ih_l0.data_ptr()
        self.G = self._grad_bundle()
        self.prep = None
        if self.enc.compute_dtype == torch.bfloat16 and ops.USE_TC and ops.USE_HALO:
            # every bf16 GEMM operand of the step is regenerated from the fp32 masters by one launch
            wp = ops.WeightPrep(dev)
            prepared = {}
            for i in range(1, 8):
                co, ci = engine.CHANNELS[i // 2][1], (engine.CHANNELS[i // 2][0] if i % 2 == 0 else engine.CHANNELS[i // 2][1])
                prepared[("f", i)] = wp.add(self.Wt.conv[i], 1, co, ci, 9)
                prepared[("t", i)] = wp.add(self.Wt.conv[i], 2, co, ci, 9)
            prepared["fc"] = wp.add(self.Wt.fc_w, 0, 512, 512, 1).view(512, 512)
            prepared["ih"] = wp.add(self.Wt.w_ih, 0, 1536, 512, 1).view(1536, 512)
            prepared["fc_t"] = wp.add(self.Wt.fc_w, 3, 512, 512, 1)
            prepared["ih_t"] = wp.add(self.Wt.w_ih, 3, 1536, 512, 1)
            self.Wt.prepared = prepared
            self.prep = wp

    def _grad_bundle(self) -> engine.EncoderGrads:
        enc = self.enc
        r = enc.rnn
        from .models.audio_encoder import _cat_if_needed, _packed_conv
        return engine.EncoderGrads(
            bn=[(bn.weight.grad, bn.bias.grad) for bn in enc._bns()],
            conv=[_packed_conv(c.weight.grad) for c in enc._convs()],
            fc_w=enc.fc1.weight.grad, fc_b=enc.fc1.bias.grad,
            w_ih=_cat_if_needed(r.weight_ih_l0.grad, r.weight_ih_l0_reverse.grad),
            b_ih=_cat_if_needed(r.bias_ih_l0.grad, r.bias_ih_l0_reverse.grad),
            w_hh=_cat_if_needed(r.weight_hh_l0.grad, r.weight_hh_l0_reverse.grad).view(2, 768, 256),
            b_hh=_cat_if_needed(r.bias_hh_l0.grad, r.bias_hh_l0_reverse.grad))

    # ------------------------------------------------------------------ the step
    def _fwd_bwd(self, wav, text, text_len, label, length):
        """Forward + backward into flat_g; returns nothing (loss in self.loss_out)."""
        enc = self.enc
        self.flat_g.zero_()
        if self.prep is not None:
            self.prep.run()
        emb, ectx = engine.encoder_forward(
            self.Wt, wav, training=True, bn_training=enc.bn0.training, dropout=enc.dropout_enabled,
            seed=self.base_seed, dtype=enc.compute_dtype, save=True, seed_dev=self.step_dev)
        B, Tp, D = emb.shape
        ew = self.txt.embedding.core.weight
        V = ew.shape[0]
        N = text.shape[1]
        seq = torch.empty(B, D, device=emb.device, dtype=torch.float32)
        call("tag_embed_mean_fwd", text, text_len, ew.data, None, seq, B, N, D, V)
        sim = torch.empty(B, Tp, device=emb.device, dtype=torch.float32)
        call("tag_dot_sigmoid_fwd", emb, seq, sim, None, B, Tp, D, self.scale)
        trunc = min(Tp, label.shape[1])
        d_sim = torch.zeros(B, Tp, device=emb.device, dtype=torch.float32)
        call("tag_frame_bce", sim, Tp, label, label.stride(0), length, B, trunc, self.loss_out, d_sim, Tp, 1.0)
        d_emb = torch.empty_like(emb)
        d_seq = torch.empty_like(seq)
        ws = torch.empty(B, Tp, device=emb.device, dtype=torch.float32)
        call("tag_dot_sigmoid_bwd", d_sim, sim, emb, seq, d_emb, d_seq, ws, B, Tp, D, self.scale)
        call("tag_embed_mean_bwd", text, text_len, d_seq, ew.grad, B, N, D, V)
        overlap = self._overlap_ar and self.world > 1       # (bench.py profiles one rank alone with world set to 1)
        engine.encoder_backward(self.Wt, ectx, d_emb, self.G, side_stream=self.side_stream,
                                on_block_done=self._early_allreduce if overlap else None)
        if overlap:
            self._finish_allreduce()
        elif self._ar_in_graph and self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
        self.sim = sim

    def _early_allreduce(self, blk: int) -> None:
        """After conv block 2 (index 1): all-reduce flat_g[_ar_split:] on the communication stream; block 1's persistent
        kernels leave ``_sm_reserve`` SMs to NCCL."""
        if blk != 1:
            return
        tail = self.flat_g[self._ar_split:]
        if self.device.type != "cuda":
            torch.distributed.all_reduce(tail, group=self.pg)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(self.device)
        ready = torch.cuda.Event()
        ready.record()
        self._comm_stream.wait_event(ready)
        with torch.cuda.stream(self._comm_stream):
            torch.distributed.all_reduce(tail, group=self.pg)
            self._ar_done = torch.cuda.Event()
            self._ar_done.record()
        ops.set_sm_reserve(self._sm_reserve)

    def _finish_allreduce(self) -> None:
        """The head of the bucket (bn0 and block 1) after the last backward kernel; joins the tail."""
        if self.device.type == "cuda":
            ops.set_sm_reserve(0)
        torch.distributed.all_reduce(self.flat_g[:self._ar_split], group=self.pg)
        if self._ar_done is not None:
            torch.cuda.current_stream().wait_event(self._ar_done)
            self._ar_done = None

    def _optim(self):
        self.sumsq.zero_()
        call("tag_sumsq", self.flat_g, self.n_params, self.sumsq)
        call("tag_clip_adam", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_params,
             self.sumsq, self.step_dev, 1.0 / self.world, float(self.max_grad_norm), float(self._lr), self.lr_dev,
             float(self.betas[0]), float(self.betas[1]), float(self.eps), self.norm_out)

    def _allreduce(self):
        """The whole bucket in one call between the graphs — for the steps whose backward does not issue it itself."""
        if self.world > 1 and not (self._overlap_ar or self._ar_in_graph):
            torch.distributed.all_reduce(self.flat_g, group=self.pg)

    def _eager(self, s):
        self._run_fwd_bwd(s)
        self._allreduce()
        self._optim()

    def _host_views(self, batch: Dict):
        """(key, shapes) of the static buffers for ``batch`` + the host-side tensors to upload."""
        wav = batch["waveform"]
        B, L = wav.shape
        Tp = (L // engine.HOP + 1) // 4
        label = batch["label"]
        trunc = min(Tp, label.shape[1])
        length = torch.as_tensor(batch["waveform_len"]).to(torch.long)
        length = torch.clamp((length // engine.HOP + 1) // 4, 1, trunc)      # run_strong.py:107-118
        src = {"waveform": wav, "text": batch["text"],
               "text_len": torch.as_tensor(batch["text_len"]).to(torch.long), "label": label, "length": length}
        # float16 waveforms (the reference's h5 storage type) are uploaded as they are and widened by the frontend
        wdt = torch.float16 if wav.dtype == torch.float16 else torch.float32
        key = (B, L, batch["text"].shape[1], label.shape[1], wdt)
        return key, src

    def _alloc_inputs(self, key):
        B, L, N, Tl, wdt = key
        dev = self.device
        return {
            "key": key,
            "waveform": torch.empty(B, L, device=dev, dtype=wdt),
            "text": torch.empty(B, N, device=dev, dtype=torch.long),
            "text_len": torch.empty(B, device=dev, dtype=torch.long),
            "label": torch.empty(B, Tl, device=dev, dtype=torch.float32),
            "length": torch.empty(B, device=dev, dtype=torch.long),
        }

    _INPUT_KEYS = ("waveform", "text", "text_len", "label", "length")

    def prefetch(self, batch: Dict) -> None:
        """Start the host->device copy of ``batch`` (collate schema, ideally pinned host tensors) on a copy
        stream, so that it overlaps the step currently running; a later ``step(batch)`` with the same dict picks
        the staged inputs up with a device-to-device copy instead of waiting for PCIe."""
        key, src = self._host_views(batch)
        if self._staging is None or self._staging["key"] != key:
            self._staging = self._alloc_inputs(key)
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staging_free = torch.cuda.Event()
            self._staging_ready = torch.cuda.Event()
            self._staging_free.record()
        self._copy_stream.wait_event(self._staging_free)        # the previous consumer has drained the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for k in self._INPUT_KEYS:
                self._staging[k].copy_(src[k], non_blocking=True)
            self._staging_ready.record()
        self._prefetched = batch

    def _prepare_static(self, batch: Dict):
        """Device-resident static input buffers (graph inputs)."""
        staged = self._prefetched is batch and self._staging is not None
        if staged:
            key, src = self._staging["key"], self._staging
        else:
            key, src = self._host_views(batch)
        if self._static is None or self._static["key"] != key:
            entry = self._graphs.get(key)
            # a shape seen before brings its own static buffers (its graphs read them); a new one gets fresh ones
            self._static = entry["static"] if entry is not None else self._alloc_inputs(key)
        s = self._static
        if staged:
            torch.cuda.current_stream().wait_event(self._staging_ready)
        for k in self._INPUT_KEYS:
            s[k].copy_(src[k], non_blocking=True)
        if staged:
            self._staging_free.record()
            self._prefetched = None
        return s

    def close(self) -> None:
        """Drop the captured CUDA graphs.  In data-parallel runs they contain NCCL kernels, so call this (or delete the
        step object) before ``torch.distributed.destroy_process_group()``."""
        self._graphs.clear()
        self._seen.clear()
        self._static = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def flush_counters(self) -> None:
        """Write the host-side step count into the BatchNorm ``num_batches_tracked`` buffers."""
        if self._nbt_pending:
            for bn in self.enc._bns():
                if bn.training and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += self._nbt_pending
            self._nbt_pending = 0

    def step_async(self, batch: Optional[Dict]) -> "LossHandle":
        """``step`` + a non-blocking device->host copy of this step's loss into a pinned slot.  The returned handle's
        ``result()`` waits for that copy only, so a training loop can queue step i+1 before it reads the loss of
        step i (the reference reads ``loss.item()`` synchronously, run_strong.py:147)."""
        if self._loss_slots is None:
            self._loss_slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._loss_idx = 0
        loss = self.step(batch)
        slot = self._loss_slots[self._loss_idx % len(self._loss_slots)]
        self._loss_idx += 1
        slot.copy_(loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return LossHandle(slot, ev)

    def step(self, batch: Optional[Dict]) -> torch.Tensor:
        """One train step on ``batch`` (collate schema; host or device tensors).  Returns the
        device scalar holding the loss of this step.  ``batch=None`` re-uses the inputs already
        resident in the static device buffers (device-only timing)."""
        s = self._static if batch is None else self._prepare_static(batch)
        if self.enc.training:
            self._nbt_pending += 1
        self._calls += 1
        key = s["key"]
        seen = self._seen[key] = self._seen.get(key, 0) + 1
        if not self.use_graph or seen == 1:
            # the first step of a shape runs eagerly (sets kernel attributes, primes the allocator); a collate that
            # pads to the longest clip (datasets/collate_function.py:43-84) may never repeat a shape: all eager
            self._eager(s)
            return self.loss_out
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(s)
            self._graphs[key] = entry
            while len(self._graphs) > self.MAX_GRAPHS:
                old_key, _ = self._graphs.popitem(last=False)
                self._seen.pop(old_key, None)
        else:
            self._graphs.move_to_end(key)
        entry["fwd_bwd"].replay()
        self._allreduce()
        entry["optim"].replay()
        return self.loss_out

    def _capture(self, s) -> Dict:
        """Two CUDA graphs (forward+backward, optimizer) for the shape of ``s``; all shapes share one memory pool
        (replays are serialised on one stream, so their intermediates may alias)."""
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, pool=self._graph_pool):
            self._run_fwd_bwd(s)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=self._graph_pool):
            self._optim()
        return {"fwd_bwd": g1, "optim": g2, "static": s}      # capture executes nothing

    def _run_fwd_bwd(self, s):
        self._fwd_bwd(s["waveform"], s["text"], s["text_len"], s.get("label"), s["length"])


class WeakFusedTrainStep(FusedTrainStep):
    """The same flat-buffer / CUDA-graph step for the weakly supervised multi-phrase configuration
    MultiTextBiEncoder(Cnn8Rnn, EmbeddingAgg, DotProduct, pooling=...) + ClipBceLoss + Adam — the body of the weak
    runners' train loop (reference python_scripts/training/run_weak_phrase.py:39-96; most of the reference's
    eg_configs).  Batch schema: ``waveform`` [B, L], ``waveform_len`` [B], ``text`` [B, n, N], ``text_len`` [B, n],
    ``label`` [B, n] (clip-level 0/1 targets), n <= 64 phrases per clip."""

    _INPUT_KEYS = ("waveform", "text", "text_len", "label", "length", "strong_label")

    def __init__(self, model: nn.Module, frame_weight: Optional[float] = None, **kw):
        """``frame_weight`` = w selects ClipFrameBceLoss: (1 - w) * clip BCE + w * frame BCE on ``strong_label``
        [B, T', n] (losses.py:186-210); None = ClipBceLoss."""
        from .models.audio_text_model import MultiTextBiEncoder
        from .models.utils import POOL_MODES
        if not isinstance(model, MultiTextBiEncoder):
            raise NotImplementedError("WeakFusedTrainStep drives models.audio_text_model.MultiTextBiEncoder")
        if model.pooling not in POOL_MODES:
            raise Exception(f"Unsupported pooling {model.pooling}")
        super().__init__(model, **kw)
        self._overlap_ar = self._ar_in_graph = False       # this step reduces the bucket between its two graphs (_allreduce)
        self.pool_mode = POOL_MODES[model.pooling]
        self.frame_weight = frame_weight
        self.loss_clip = torch.zeros((), device=self.device, dtype=torch.float32)
        self.loss_frame = torch.zeros((), device=self.device, dtype=torch.float32)

    def _host_views(self, batch: Dict):
        wav = batch["waveform"]
        B, L = wav.shape
        Tp = (L // engine.HOP + 1) // 4
        text = torch.as_tensor(batch["text"])
        n, N = text.shape[1], text.shape[2]
        if n > 64:
            raise NotImplementedError("WeakFusedTrainStep: at most 64 phrases per clip")
        length = torch.as_tensor(batch["waveform_len"]).to(torch.long)
        length = torch.clamp((length // engine.HOP + 1) // 4, 1, Tp)
        label = torch.as_tensor(batch["label"]).to(torch.float32)
        src = {"waveform": wav, "text": text.reshape(B * n, N),
               "text_len": torch.as_tensor(batch["text_len"]).to(torch.long).reshape(B * n),
               "label": label.reshape(B, n), "length": length}
        if self.frame_weight is not None:
            strong = torch.as_tensor(batch["strong_label"]).to(torch.float32)
            if tuple(strong.shape) != (B, Tp, n):
                raise RuntimeError(f"strong_label must be [B, T', n] = {(B, Tp, n)}, got {tuple(strong.shape)}")
            src["strong_label"] = strong
        else:
            src["strong_label"] = torch.zeros(1)
        wdt = torch.float16 if wav.dtype == torch.float16 else torch.float32
        return (B, L, n, N, wdt), src

    def _alloc_inputs(self, key):
        B, L, n, N, wdt = key
        dev = self.device
        return {
            "key": key,
            "waveform": torch.empty(B, L, device=dev, dtype=wdt),
            "text": torch.empty(B * n, N, device=dev, dtype=torch.long),
            "text_len": torch.empty(B * n, device=dev, dtype=torch.long),
            "label": torch.empty(B, n, device=dev, dtype=torch.float32),
            "length": torch.empty(B, device=dev, dtype=torch.long),
            "strong_label": torch.empty((B, (L // engine.HOP + 1) // 4, n) if self.frame_weight is not None else (1,),
                                        device=dev, dtype=torch.float32),
        }

    def _run_fwd_bwd(self, s):
        self._strong = s["strong_label"]
        super()._run_fwd_bwd(s)

    def _fwd_bwd(self, wav, text, text_len, label, length):
        enc = self.enc
        self.flat_g.zero_()
        if self.prep is not None:
            self.prep.run()
        emb, ectx = engine.encoder_forward(
            self.Wt, wav, training=True, bn_training=enc.bn0.training, dropout=enc.dropout_enabled,
            seed=self.base_seed, dtype=enc.compute_dtype, save=True, seed_dev=self.step_dev)
        B, Tp, D = emb.shape
        n = label.shape[1]
        dev = emb.device
        ew = self.txt.embedding.core.weight
        V, N = ew.shape[0], text.shape[1]
        f32 = dict(device=dev, dtype=torch.float32)
        seq = torch.empty(B * n, D, **f32)
        call("tag_embed_mean_fwd", text, text_len, ew.data, None, seq, B * n, N, D, V)
        sim = torch.empty(B, Tp, n, **f32)
        call("tag_multi_dot_sigmoid_fwd", emb, seq, sim, B, Tp, n, D, self.scale)
        clip = torch.empty(B, n, **f32)
        call("tag_pool_with_lens_fwd", sim, length, self.pool_mode, clip, B, Tp, n)
        # ClipBceLoss = mean BCE over all B*n clip-level probabilities (losses.py:38-43)
        if getattr(self, "_full_n", None) is None or self._full_n.numel() != B or int(self._n_cached) != n:
            self._full_n = torch.full((B,), n, device=dev, dtype=torch.long)
            self._n_cached = n
        d_clip = torch.empty(B, n, **f32)
        fw = self.frame_weight
        call("tag_frame_bce", clip, n, label, label.stride(0), self._full_n, B, n,
             self.loss_out if fw is None else self.loss_clip, d_clip, n, 1.0 if fw is None else 1.0 - fw)
        d_sim = torch.empty_like(sim)
        call("tag_pool_with_lens_bwd", d_clip, sim, clip, length, self.pool_mode, d_sim, B, Tp, n)
        if fw is not None:
            # frame BCE over [B, T', n] with the length mask expanded over the phrases (losses.py:26-35): element
            # (b, t, j) is valid iff its flat index t*n + j < length[b]*n
            strong = self._strong
            d_frame = torch.empty_like(sim)
            call("tag_frame_bce", sim, Tp * n, strong, Tp * n, length * n, B, Tp * n, self.loss_frame, d_frame, Tp * n, fw)
            d_sim.add_(d_frame)
            torch.add(self.loss_clip * (1.0 - fw), self.loss_frame, alpha=fw, out=self.loss_out)
        d_emb = torch.empty_like(emb)
        d_seq = torch.empty_like(seq)
        ws = torch.empty_like(sim)
        call("tag_multi_dot_sigmoid_bwd", d_sim, sim, emb, seq, d_emb, d_seq, ws, B, Tp, n, D, self.scale)
        call("tag_embed_mean_bwd", text, text_len, d_seq, ew.grad, B * n, N, D, V)
        engine.encoder_backward(self.Wt, ectx, d_emb, self.G, side_stream=self.side_stream)
        self.sim, self.clip = sim, clip


class _CtxShim:
    """Stand-in for the autograd context so a fused step can drive an autograd Function's forward / backward bodies."""
    needs_input_grad = (True, True, False, False, False, False, False)

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


class AlignFusedTrainStep(FusedTrainStep):
    """Flat-buffer / CUDA-graph train step of the sentence-level alignment configurations: AudioTextAlignByWord or
    AudioTextAlignByPhrase (Cnn8Rnn + EmbeddingAgg + align.DotProduct + a sim_pooling.Audio*Text* module) with
    MaxMarginRankingLoss and Adam (reference eg_configs/weakly_supervised/audiocaps/sentence_level/*).  Batch schema:
    word level ``text`` [B, N], ``text_len`` [B]; phrase level ``phrases`` [txt_num, N], ``phrases_len`` [txt_num],
    ``phrases_num`` (B counts summing to txt_num); plus ``waveform`` / ``waveform_len``."""
    _INPUT_KEYS = ("waveform", "text", "text_len", "pool_len", "pad_index", "length")

    def __init__(self, model: nn.Module, margin: float = 1.0, fix_norm: bool = True, lamda1: float = 1.0, **kw):
        from .models.align import AUDIO_POOL, TEXT_POOL
        from .models.audio_text_model import AudioTextAlignByWord
        super().__init__(model, **kw)
        self._overlap_ar = self._ar_in_graph = False       # this step reduces the bucket between its two graphs (_allreduce)
        self.word_level = isinstance(model, AudioTextAlignByWord)
        self.a_mode = AUDIO_POOL[model.sim_pooling.audio_pool]
        self.t_mode = TEXT_POOL[model.sim_pooling.text_pool]
        self.scale = 1.0 / (self.enc.embed_dim ** 0.5) if getattr(model.match_fn, "scaled", False) else 1.0
        self.margin, self.fix_norm, self.lamda1 = float(margin), bool(fix_norm), float(lamda1)

    def _host_views(self, batch: Dict):
        wav = batch["waveform"]
        B, L = wav.shape
        Tp = (L // engine.HOP + 1) // 4
        length = torch.as_tensor(batch["waveform_len"]).to(torch.long)
        length = (length // engine.HOP + 1) // 4
        if self.word_level:
            text = torch.as_tensor(batch["text"]).to(torch.long)
            text_len = torch.as_tensor(batch["text_len"]).to(torch.long)
            pool_len, pad_index, width = text_len, torch.zeros(1, dtype=torch.long), text.shape[1]
        else:
            key = batch.get("text_key", "phrases")
            text = torch.as_tensor(batch[key]).to(torch.long)
            text_len = torch.as_tensor(batch[f"{key}_len"]).to(torch.long)
            num = [int(v) for v in batch[f"{key}_num"]]
            width = max(num)
            pad_index = torch.as_tensor([b * width + j for b, c in enumerate(num) for j in range(c)], dtype=torch.long)
            pool_len = torch.as_tensor(num, dtype=torch.long)
        src = {"waveform": wav, "text": text, "text_len": text_len, "pool_len": pool_len, "pad_index": pad_index,
               "length": length}
        wdt = torch.float16 if wav.dtype == torch.float16 else torch.float32
        return (B, L, text.shape[0], text.shape[1], width, wdt), src

    def _alloc_inputs(self, key):
        B, L, rows, N, width, wdt = key
        dev = self.device
        lng = dict(device=dev, dtype=torch.long)
        self._width = width
        return {"key": key, "waveform": torch.empty(B, L, device=dev, dtype=wdt), "text": torch.empty(rows, N, **lng),
                "text_len": torch.empty(rows, **lng), "pool_len": torch.empty(B, **lng),
                "pad_index": torch.empty(1 if self.word_level else rows, **lng), "length": torch.empty(B, **lng)}

    def _check_model(self, model):
        from .models import align, sim_pooling
        from .models.audio_text_model import AudioTextAlignByPhrase, AudioTextAlignByWord
        from .models.text_encoder import EmbeddingAgg
        problems = []
        if not isinstance(model, (AudioTextAlignByWord, AudioTextAlignByPhrase)):
            problems.append("model must be AudioTextAlignByWord / AudioTextAlignByPhrase")
        if type(model.match_fn) is not align.DotProduct:
            problems.append("match_fn must be models.align.DotProduct")
        if not isinstance(getattr(model, "sim_pooling", None), sim_pooling._AudioTextPooling):
            problems.append("sim_pooling must be one of models.sim_pooling.Audio*Text*")
        if not isinstance(model.text_encoder, EmbeddingAgg) or model.text_encoder.agg != "mean":
            problems.append("text_encoder must be EmbeddingAgg(aggregation='mean')")
        if problems:
            raise NotImplementedError("AlignFusedTrainStep does not implement this configuration ("
                                      + "; ".join(problems) + "): use the autograd path")

    def _run_fwd_bwd(self, s):
        self._extra = (s["pool_len"], s["pad_index"], s["key"][4])
        self._fwd_bwd(s["waveform"], s["text"], s["text_len"], None, s["length"])

    def _fwd_bwd(self, wav, text, text_len, label, length):
        from .models.align import _AlignPoolFunction
        enc = self.enc
        pool_len, pad_index, width = self._extra
        self.flat_g.zero_()
        if self.prep is not None:
            self.prep.run()
        emb, ectx = engine.encoder_forward(
            self.Wt, wav, training=True, bn_training=enc.bn0.training, dropout=enc.dropout_enabled,
            seed=self.base_seed, dtype=enc.compute_dtype, save=True, seed_dev=self.step_dev)
        B, Tp, D = emb.shape
        dev = emb.device
        ew = self.txt.embedding.core.weight
        V = ew.shape[0]
        rows, N = text.shape
        f32 = dict(device=dev, dtype=torch.float32)
        seq = torch.empty(rows, D, **f32)
        token_emb = torch.empty(rows, N, D, **f32) if self.word_level else None
        call("tag_embed_mean_fwd", text, text_len, ew.data, token_emb, seq, rows, N, D, V)
        if self.word_level:
            text3 = token_emb                                              # [B, N, D]
        else:
            text3 = torch.zeros(B * width, D, **f32).index_copy_(0, pad_index, seq).view(B, width, D)
        ctx = _CtxShim()
        sim = _AlignPoolFunction.forward(ctx, emb, text3, length, pool_len, self.a_mode, self.t_mode, self.scale)
        d_sim = torch.empty(B, B, **f32)
        call("tag_max_margin_rank", sim, B, self.margin, self.lamda1, int(self.fix_norm), self.loss_out, d_sim)
        d_emb, d_text3 = _AlignPoolFunction.backward(ctx, d_sim)[:2]
        if self.word_level:
            call("tag_embed_token_bwd", text, d_text3.contiguous(), ew.grad, rows, N, D, V)
        else:
            d_seq = d_text3.reshape(B * width, D).index_select(0, pad_index)
            call("tag_embed_mean_bwd", text, text_len, d_seq, ew.grad, rows, N, D, V)
        engine.encoder_backward(self.Wt, ectx, d_emb.contiguous(), self.G, side_stream=self.side_stream)
        self.sim = sim


class AutogradTrainStep:
    """The same production plumbing — flat fp32 parameter / gradient / Adam buffers, ONE all-reduce of the flat
    gradient bucket, fused clip_grad_norm_ + Adam, CUDA-graph replay — around ANY of the mirrored model graphs,
    with forward and backward driven by autograd (every node is one of this package's C-ABI autograd Functions).
    This is how the configurations FusedTrainStep does not hard-code train at full speed: the attention-type heads of
    BASELINE.json configs[3] (SelfAttention + CrossAttentionGating / CrossAttention), add_proj, upsample, ...

    ``forward_fn(model, batch) -> loss`` is called with the static device batch; the default is Runner.forward +
    ``loss_fn`` (run_strong.py:92-120, 139-141).  Every tensor of ``batch`` — lengths included — is kept on the device
    so that the captured region contains no host->device copy."""

    def __init__(self, model: nn.Module, loss_fn=None, forward_fn=None, lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_grad_norm: float = 1.0, process_group=None, use_graph: bool = True):
        self.model = model
        self.loss_fn = loss_fn or FrameBceLoss()
        self.forward_fn = forward_fn or (lambda m, b: self.loss_fn(runner_forward(m, b, self.device, training=True)))
        self.betas, self.eps, self.max_grad_norm = betas, eps, max_grad_norm
        self.pg = process_group
        self.world, self.rank = 1, 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        self.use_graph = use_graph
        self.device = next(model.parameters()).device
        params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in params)
        dev = self.device
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.n_params = n
        off = 0
        for p in params:
            k = p.numel()
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                co, ci, kh, kw = p.shape          # conv weight: memory stays [Cout][kh][kw][Cin]
                pv = self.flat_p[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
                gv = self.flat_g[off:off + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                pv = self.flat_p[off:off + k].view(p.shape)
                gv = self.flat_g[off:off + k].view(p.shape)
            pv.copy_(p.data)
            p.data = pv
            p.grad = gv                          # autograd accumulates in place into the bucket
            off += k
        self._params = params
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.norm_out = torch.zeros(1, device=dev, dtype=torch.float32)
        self.loss_out = torch.zeros((), device=dev, dtype=torch.float32)
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr = float(lr)
        if self.world > 1:
            src = torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0
            for t in [self.flat_p] + [b for b in model.buffers()]:
                torch.distributed.broadcast(t, src=src, group=self.pg)
        self._graphs = OrderedDict()
        self._seen = {}
        self._graph_pool = None
        self._static = None
        self._cap_stream = None

    MAX_GRAPHS = 8
    lr = FusedTrainStep.lr
    set_lr = FusedTrainStep.set_lr

    def _fwd_bwd(self, s):
        from .models import nn_ops
        self.flat_g.zero_()
        prev, nn_ops.SEED_DEV = nn_ops.SEED_DEV, self.step_dev
        try:
            loss = self.forward_fn(self.model, {k: v for k, v in s.items() if k != "key"})
            loss.backward()
        finally:
            nn_ops.SEED_DEV = prev
        for p in self._params:                    # a parameter the graph did not reach keeps its zero slice
            assert p.grad is not None and p.grad.data_ptr() >= self.flat_g.data_ptr()
        self.loss_out.copy_(loss.detach())

    def _optim(self):
        self.sumsq.zero_()
        call("tag_sumsq", self.flat_g, self.n_params, self.sumsq)
        call("tag_clip_adam", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_params,
             self.sumsq, self.step_dev, 1.0 / self.world, float(self.max_grad_norm), float(self._lr), self.lr_dev,
             float(self.betas[0]), float(self.betas[1]), float(self.eps), self.norm_out)

    def _allreduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)

    @staticmethod
    def _items(batch: Dict):
        items = {k: torch.as_tensor(v) for k, v in batch.items() if not isinstance(v, (str, bool, list)) or
                 (isinstance(v, list) and v and isinstance(v[0], (int, float)))}
        key = tuple((k, tuple(t.shape), t.dtype) for k, t in sorted(items.items()))
        return items, key

    def _alloc(self, items, key):
        d = {k: torch.empty(t.shape, device=self.device, dtype=t.dtype) for k, t in items.items()}
        d["key"] = key
        return d

    def prefetch(self, batch: Dict) -> None:
        """Start the host->device copy of ``batch`` on a copy stream (see FusedTrainStep.prefetch)."""
        items, key = self._items(batch)
        if getattr(self, "_staging", None) is None or self._staging["key"] != key:
            self._staging = self._alloc(items, key)
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staging_free = torch.cuda.Event()
            self._staging_ready = torch.cuda.Event()
            self._staging_free.record()
        self._copy_stream.wait_event(self._staging_free)
        with torch.cuda.stream(self._copy_stream):
            for k, t in items.items():
                self._staging[k].copy_(t, non_blocking=True)
            self._staging_ready.record()
        self._prefetched = batch

    def _prepare_static(self, batch: Dict):
        staged = getattr(self, "_prefetched", None) is batch and getattr(self, "_staging", None) is not None
        if staged:
            items, key = {k: v for k, v in self._staging.items() if k != "key"}, self._staging["key"]
        else:
            items, key = self._items(batch)
        if self._static is None or self._static["key"] != key:
            entry = self._graphs.get(key)
            self._static = entry["static"] if entry is not None else self._alloc(items, key)
        if staged:
            torch.cuda.current_stream().wait_event(self._staging_ready)
        for k, t in items.items():
            self._static[k].copy_(t, non_blocking=True)
        if staged:
            self._staging_free.record()
            self._prefetched = None
        return self._static

    _loss_slots = None
    step_async = FusedTrainStep.step_async

    def step(self, batch: Optional[Dict]) -> torch.Tensor:
        """One train step; returns the device scalar holding this step's loss.  ``batch=None`` re-uses the inputs
        already resident in the static device buffers."""
        s = self._static if batch is None else self._prepare_static(batch)
        key = s["key"]
        seen = self._seen[key] = self._seen.get(key, 0) + 1
        if not self.use_graph:
            self._fwd_bwd(s)
            self._allreduce()
            self._optim()
            return self.loss_out
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(self.device)
        if seen == 1:
            # autograd binds every AccumulateGrad node to the stream its parameter was first used on, and a node that
            # lives on the legacy default stream cannot run inside a capture: the eager first step of a shape already
            # runs on the stream the graphs are captured on
            cur = torch.cuda.current_stream()
            self._cap_stream.wait_stream(cur)
            with torch.cuda.stream(self._cap_stream):
                self._fwd_bwd(s)
            cur.wait_stream(self._cap_stream)
            self._allreduce()
            self._optim()
            return self.loss_out
        entry = self._graphs.get(key)
        if entry is None:
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, pool=self._graph_pool, stream=self._cap_stream):
                self._fwd_bwd(s)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=self._graph_pool):
                self._optim()
            entry = self._graphs[key] = {"fwd_bwd": g1, "optim": g2, "static": s}
            while len(self._graphs) > self.MAX_GRAPHS:
                old, _ = self._graphs.popitem(last=False)
                self._seen.pop(old, None)
        else:
            self._graphs.move_to_end(key)
        entry["fwd_bwd"].replay()
        self._allreduce()
        entry["optim"].replay()
        return self.loss_out
