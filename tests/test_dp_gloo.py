"""World-size-2 `gloo` test (CPU) of the data-parallel host logic of FusedTrainStep: the flat
parameter / gradient bucket layout, parameter views aliasing the bucket, and the single all-reduce
per step.  (The kernels themselves need a GPU; the NCCL path runs under `bench.py --gpus N`.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import build_model
        from texttoaudiogrounding_b200.train import FusedTrainStep
        torch.manual_seed(1 + rank)          # replicas start DIFFERENT; the constructor broadcasts rank 0's state
        model = build_model(None, "fp32", device="cpu", vocab=300)
        model.audio_encoder.bn0.running_mean.fill_(float(rank))
        ts = FusedTrainStep(model, use_graph=False, base_seed=5)
        assert ts.world == world and ts.rank == rank
        assert ts.base_seed == 5 + 7919 * rank          # every rank draws its own dropout masks
        assert float(model.audio_encoder.bn0.running_mean.abs().max()) == 0.0
        n = ts.n_params
        assert n == sum(p.numel() for p in model.parameters())
        # parameters and .grad are views of the flat buffers, in the bucket order: what the backward pass writes last
        # (bn0, conv block 1's BatchNorms and convolutions) first, then everything else in _param_list() order
        enc_params = model.audio_encoder._param_list()
        late = enc_params[0:6] + enc_params[18:20]
        assert late[6] is model.audio_encoder.conv_block1.conv1.weight and late[5] is model.audio_encoder.conv_block1.bn2.bias
        rest = [p for p in enc_params if not any(p is q for q in late)] + [model.text_encoder.embedding.core.weight]
        off = 0
        for p in late + rest:
            assert p.data_ptr() == ts.flat_p.data_ptr() + 4 * off
            assert p.grad.data_ptr() == ts.flat_g.data_ptr() + 4 * off
            off += p.numel()
        assert off == n
        # GRU operands exist without copies
        assert ts.Wt.w_ih.shape == (1536, 512) and ts.Wt.w_hh.shape == (2, 768, 256)
        assert ts.Wt.w_ih.data_ptr() == model.audio_encoder.rnn.weight_ih_l0.data_ptr()
        # conv weights keep the reference's shape while their memory is [Cout][kh][kw][Cin]
        w = model.audio_encoder.conv_block3.conv2.weight
        assert tuple(w.shape) == (256, 256, 3, 3) and w.permute(0, 2, 3, 1).is_contiguous()
        # identical replicas on every rank
        chk = ts.flat_p.double().sum().reshape(1)
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        # one all-reduce(sum) over the whole bucket; the mean's 1/world is applied by the Adam kernel
        ts.flat_g.copy_(torch.arange(n, dtype=torch.float32) % 7 + rank)
        assert ts._ar_in_graph and not ts._overlap_ar   # default: the step issues the one all-reduce itself (last graph node)
        ts._ar_in_graph = False
        ts._allreduce()
        ts._ar_in_graph = True
        expect = (torch.arange(n, dtype=torch.float32) % 7) * world + sum(range(world))
        assert torch.equal(ts.flat_g, expect)
        fcb = [id(p) for p in ts._params].index(id(model.audio_encoder.fc1.bias))
        assert torch.equal(model.audio_encoder.fc1.bias.grad, expect[ts._views[fcb][0]:ts._views[fcb][0] + 512])
        # the same collective as two calls (TAG_B200_AR_OVERLAP=1: under the backward pass on GPUs): the tail of the bucket —
        # everything but bn0 and conv block 1, final once block 2 has run its backward — and then the head
        first_early = model.audio_encoder.conv_block2.bn1.weight
        assert ts.flat_g.data_ptr() + 4 * ts._ar_split == first_early.grad.data_ptr()
        assert ts._ar_split == 3 * 2 * 64 + 64 * 9 + 64 * 64 * 9 and ts._ar_split < 0.01 * n      # > 99 % of the bytes travel early
        ts.flat_g.copy_(torch.arange(n, dtype=torch.float32) % 5 + 2 * rank)
        for blk in (3, 2, 1, 0):
            ts._early_allreduce(blk)
            if blk == 2:                                    # nothing is reduced before block 2 is done
                assert torch.equal(ts.flat_g, torch.arange(n, dtype=torch.float32) % 5 + 2 * rank)
        ts._finish_allreduce()
        assert torch.equal(ts.flat_g, (torch.arange(n, dtype=torch.float32) % 5) * world + 2 * sum(range(world)))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
