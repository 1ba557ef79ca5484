"""tcgen05 implicit-GEMM kernels (through the C ABI) against fp32 PyTorch convolutions of the same
bf16-rounded operands, on every (W, Cin, Cout) shape the model uses plus odd heights / batches
(TMA out-of-bounds halo, partial M tiles, multi-tile persistence, accumulator double buffering)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


def g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.fixture
def halo_mode(request):
    """halo kernel scheduling: True = default (CTA pairs with cta_group::2 MMAs where the weights stream),
    "single" = one CTA per tile everywhere, False = the plain (no halo re-use) kernel."""
    from texttoaudiogrounding_b200 import _lib
    mode = request.param
    _lib.lib().tag_conv_halo_set_pair_mode(0 if mode == "single" else 1)
    yield bool(mode)
    _lib.lib().tag_conv_halo_set_pair_mode(1)


def _bf(x):
    return x.bfloat16().float()


SHAPES = [
    # B, H, W, Cin, Cout
    (2, 13, 64, 64, 64),      # block 1 conv2 (TH = 2, odd H -> clipped tile)
    (3, 10, 32, 64, 128),     # block 2 conv1
    (2, 9, 32, 128, 128),     # block 2 conv2
    (2, 17, 16, 128, 256),    # block 3 conv1
    (1, 25, 16, 256, 256),    # block 3 conv2
    (3, 50, 8, 256, 512),     # block 4 conv1 (two N tiles)
    (2, 33, 8, 512, 512),     # block 4 conv2
    (40, 50, 8, 256, 512),    # > 148 tiles: persistent loop + both TMEM stages
    (1, 16, 8, 256, 256),     # ONE tile: the second CTA of a pair computes a padding tile
    (3, 40, 16, 128, 64),     # block-2 conv1 dgrad shape (N = 64, weights stream), odd tile count
    (3, 47, 16, 64, 64),      # weights-resident 64 -> 64 layer: several row tiles per column, the last one clipped
    (5, 151, 64, 64, 64),     # the same with > 148 tiles: both accumulators, both MMA issuers, ring wrap-around
]


@pytest.mark.parametrize("halo_mode", [True, "single", False], indirect=True)
@pytest.mark.parametrize("B,H,W,Cin,Cout", SHAPES)
def test_tc_conv3x3_fwd(B, H, W, Cin, Cout, halo_mode):
    halo = halo_mode
    from texttoaudiogrounding_b200 import ops
    x = _bf(torch.randn(B, Cin, H, W, generator=g(1)))
    w = _bf(torch.randn(Cout, Cin, 3, 3, generator=g(2)) * (1.0 / (3 * Cin ** 0.5)))
    ref = F.conv2d(x.cuda(), w.cuda(), padding=1).permute(0, 2, 3, 1).contiguous()   # fp32 NHWC
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    wp = w.permute(0, 2, 3, 1).contiguous().cuda()
    wp = ops.prep_weight(wp, torch.bfloat16, W) if halo else wp.bfloat16()
    # fp32 output + statistics
    y32 = torch.empty(B, H, W, Cout, device="cuda")
    stats = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    ops.conv_fwd(xn, wp, y32, None, False, stats, B, H, W, Cin, Cout, 9)
    torch.cuda.synchronize()
    assert rel_err(y32, ref) < 2e-3, rel_err(y32, ref)
    np.testing.assert_allclose(stats[:Cout].cpu().numpy(), y32.double().sum((0, 1, 2)).cpu().numpy(),
                               rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(stats[Cout:].cpu().numpy(), y32.double().pow(2).sum((0, 1, 2)).cpu().numpy(),
                               rtol=1e-3)
    # bf16 output
    yb = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(xn, wp, yb, None, False, None, B, H, W, Cin, Cout, 9)
    assert rel_err(yb.float(), ref) < 6e-3


@pytest.mark.parametrize("halo_mode", [True, "single", False], indirect=True)
@pytest.mark.parametrize("B,H,W,Cin,Cout", SHAPES[:7])
def test_tc_conv3x3_dgrad_and_wgrad(B, H, W, Cin, Cout, halo_mode):
    halo = halo_mode
    from texttoaudiogrounding_b200 import ops
    x = _bf(torch.randn(B, Cin, H, W, generator=g(3))).cuda().requires_grad_(True)
    w = _bf(torch.randn(Cout, Cin, 3, 3, generator=g(4)) * (1.0 / (3 * Cin ** 0.5))).cuda().requires_grad_(True)
    dy = _bf(torch.randn(B, Cout, H, W, generator=g(5))).cuda()
    F.conv2d(x, w, padding=1).backward(dy)
    xn = x.detach().permute(0, 2, 3, 1).contiguous().bfloat16()
    dyn = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    wp32 = w.detach().permute(0, 2, 3, 1).contiguous()
    wt = ops.prep_weight_t(wp32, Cout, Cin, 9, torch.bfloat16, W if halo else None)
    dx = torch.empty(B, H, W, Cin, device="cuda")
    ops.conv_fwd(dyn, wt, dx, None, False, None, B, H, W, Cout, Cin, 9)
    assert rel_err(dx.permute(0, 3, 1, 2), x.grad) < 2e-3, rel_err(dx.permute(0, 3, 1, 2), x.grad)
    dw = torch.zeros(Cout, 3, 3, Cin, device="cuda")
    ops.conv_wgrad(dyn, xn, dw, B, H, W, Cin, Cout, 9, 1)
    torch.cuda.synchronize()
    assert rel_err(dw.permute(0, 3, 1, 2), w.grad) < 2e-3, rel_err(dw.permute(0, 3, 1, 2), w.grad)


@pytest.mark.parametrize("M,K,N", [(300, 512, 512), (1000, 512, 1536), (16000, 1536, 512)])
def test_tc_linear_fwd_and_wgrad(M, K, N):
    from texttoaudiogrounding_b200 import ops
    x = _bf(torch.randn(M, K, generator=g(6))).cuda()
    w = _bf(torch.randn(N, K, generator=g(7)) * 0.03).cuda()
    b = torch.randn(N, generator=g(8)).cuda()
    ref = F.relu(F.linear(x, w, b))
    y = torch.empty(M, N, device="cuda")
    ops.conv_fwd(x.bfloat16(), w.bfloat16(), y, b, True, None, 1, M, 1, K, N, 1)
    assert rel_err(y, ref) < 2e-3, rel_err(y, ref)
    dy = _bf(torch.randn(M, N, generator=g(9))).cuda()
    dw = torch.zeros(N, K, device="cuda")
    ops.conv_wgrad(dy.bfloat16(), x.bfloat16(), dw, 1, M, 1, K, N, 1, 1)
    ref_dw = dy.t() @ x
    assert rel_err(dw, ref_dw) < 2e-3, rel_err(dw, ref_dw)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(40, 50, 8, 256, 512), (6, 100, 64, 64, 64), (3, 40, 16, 128, 64)])
def test_sm_reserve_leaves_results_unchanged(B, H, W, Cin, Cout):
    """tag_set_sm_reserve: the persistent conv kernels on 8 or 41 fewer SMs (odd: a CTA pair loses its partner's SM too)
    compute the same tiles — bit-identical outputs; the knob resets."""
    from texttoaudiogrounding_b200 import ops
    x = _bf(torch.randn(B, H, W, Cin, generator=g(41))).cuda().bfloat16()
    w = ops.prep_weight((torch.randn(Cout, 3, 3, Cin, generator=g(42)) * (1.0 / (3 * Cin ** 0.5))).cuda(), torch.bfloat16, W)
    outs = []
    try:
        for reserve in (0, 8, 41):
            ops.set_sm_reserve(reserve)
            y = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
            ops.conv_fwd(x, w, y, None, False, None, B, H, W, Cin, Cout, 9)
            outs.append(y)
        torch.cuda.synchronize()
    finally:
        ops.set_sm_reserve(0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    with pytest.raises(Exception):
        ops.set_sm_reserve(65)


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [
    (8, 250, 64, 64, 64, 9),       # wgrad64 (conv_block1.conv2), ~100 splits
    (8, 125, 32, 64, 128, 9),      # wgrad64, Cout 128
    (8, 125, 32, 128, 128, 9),     # one-CTA wgrad
    (8, 62, 16, 256, 256, 9),      # CTA-pair wgrad
    (1, 4000, 1, 512, 1536, 1),    # linear (GRU input projection)
])
def test_deterministic_splitk_wgrad_is_bit_reproducible(B, H, W, Cin, Cout, taps):
    """tag_set_splitk_workspace: slabs + ordered reduce give the same bits on every run and agree with the atomic path;
    the result is ADDED to dw like the atomic path's."""
    from texttoaudiogrounding_b200 import ops
    x = _bf(torch.randn(B, H, W, Cin, generator=g(31))).cuda().bfloat16()
    dy = _bf(torch.randn(B, H, W, Cout, generator=g(32))).cuda().bfloat16()
    P = B * H * W
    splits = ops.wgrad_splits(P, Cin, Cout, taps)
    shape = (Cout, 3, 3, Cin) if taps == 9 else (Cout, Cin)
    atomic = torch.zeros(shape, device="cuda")
    ops.conv_wgrad(dy, x, atomic, B, H, W, Cin, Cout, taps, splits)
    ops.set_deterministic_wgrad(True)
    try:
        runs = []
        for _ in range(3):
            dw = torch.full(shape, 0.5, device="cuda")
            ops.conv_wgrad(dy, x, dw, B, H, W, Cin, Cout, taps, splits)
            runs.append(dw)
        torch.cuda.synchronize()
    finally:
        ops.set_deterministic_wgrad(False)
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    assert rel_err(runs[0] - 0.5, atomic) < 1e-4, rel_err(runs[0] - 0.5, atomic)
    again = torch.zeros(shape, device="cuda")
    ops.conv_wgrad(dy, x, again, B, H, W, Cin, Cout, taps, splits)      # the switch back to atomics works
    assert rel_err(again, atomic) < 1e-5


def test_deterministic_splitk_workspace_too_small_is_an_error():
    from texttoaudiogrounding_b200 import _lib, ops
    x = torch.zeros(8, 62, 16, 256, device="cuda", dtype=torch.bfloat16)
    dy = torch.zeros(8, 62, 16, 256, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(256, 3, 3, 256, device="cuda")
    ops.set_deterministic_wgrad(True, megabytes=1)
    try:
        with pytest.raises(_lib.TagError):
            ops.conv_wgrad(dy, x, dw, 8, 62, 16, 256, 256, 9, 8)
    finally:
        ops.set_deterministic_wgrad(False)


def test_tc_matches_simt_on_model_shapes_end_to_end():
    """Same bf16 train step with the tensor-core kernels and with the SIMT kernels."""
    from oracle import tag_oracle as O
    from helpers import build_model
    from texttoaudiogrounding_b200 import ops
    from texttoaudiogrounding_b200.train import FusedTrainStep
    sd = O.synth_state_dict(seed=1, sharpen=300.0, perturb_bn=True)
    batch = O.synth_batch(4, 64000, seed=0)
    res = {}
    for use_tc in (True, False):
        ops.USE_TC = use_tc
        try:
            model = build_model(sd, "bf16")
            model.train()
            model.audio_encoder.dropout_enabled = False
            ts = FusedTrainStep(model, use_graph=False)
            loss = ts.step(batch).item()
            torch.cuda.synchronize()
            res[use_tc] = (loss, ts.norm_out.item(), ts.flat_g.clone())
        finally:
            ops.USE_TC = True
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=2e-2)
    np.testing.assert_allclose(res[True][1], res[False][1], rtol=5e-2)
    a, b = res[True][2], res[False][2]
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 0.98, cos


@pytest.mark.parametrize("B,T", [(3, 7), (8, 25), (11, 50)])
def test_bigru_bf16_tensor_core_variant_close_to_oracle(B, T):
    """mma.sync GRU (bf16 W_hh / exchanged state, fp32 cell) vs the fp32 oracle recurrence."""
    from oracle import tag_oracle as O
    from texttoaudiogrounding_b200 import ops
    sd = O.synth_state_dict(seed=5)
    pre = "audio_encoder.rnn."
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
             "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse"]
    ws = {k: sd[pre + k].clone().requires_grad_(True) for k in names}
    x = (torch.randn(B, T, 512, generator=g(18)) * 0.5).requires_grad_(True)
    ref = O.bigru(x, {pre + k: v for k, v in ws.items()}, pre)
    d_out = torch.randn(ref.shape, generator=g(19))
    ref.backward(d_out)
    w_ih = torch.cat([ws["weight_ih_l0"], ws["weight_ih_l0_reverse"]]).detach().cuda()
    b_ih = torch.cat([ws["bias_ih_l0"], ws["bias_ih_l0_reverse"]]).detach().cuda()
    w_hh = torch.stack([ws["weight_hh_l0"], ws["weight_hh_l0_reverse"]]).detach().cuda().contiguous()
    b_hh = torch.cat([ws["bias_hh_l0"], ws["bias_hh_l0_reverse"]]).detach().cuda()
    rows = B * T
    gi = torch.nn.functional.linear(x.detach().cuda().reshape(rows, 512), w_ih, b_ih).contiguous()
    out = torch.empty(B, T, 512, device="cuda")
    gates = torch.empty(B, T, 2, 4, 256, device="cuda")
    ops.call("tag_gru_fwd_bf16", gi, w_hh, b_hh, out, gates, B, T)
    torch.cuda.synchronize()
    assert (out.cpu() - ref.detach()).abs().max().item() < 1e-2
    bf = dict(device="cuda", dtype=torch.bfloat16)
    dgi = torch.empty(rows, 1536, **bf)
    dgh = torch.empty(2, rows, 768, **bf)
    hprev = torch.empty(2, rows, 256, **bf)
    ops.call("tag_gru_bwd_bf16", d_out.cuda(), out, gates, w_hh, dgi, dgh, hprev, B, T)
    dx = (dgi.float() @ w_ih).cpu().reshape(B, T, 512)
    assert rel_err(dx, x.grad) < 3e-2, rel_err(dx, x.grad)
    dw = dgh[0].float().t() @ hprev[0].float()
    assert rel_err(dw.cpu(), ws["weight_hh_l0"].grad) < 3e-2
    # hprev is the state each step consumed: forward direction = out shifted by one step
    hp = hprev[0].reshape(B, T, 256)
    assert torch.equal(hp[:, 1:], out[:, :-1, :256].bfloat16()) and float(hp[:, 0].float().abs().max()) == 0.0


@pytest.mark.parametrize("halo_mode", [True, "single"], indirect=True)
@pytest.mark.parametrize("B,H,W,C", [(2, 21, 16, 128), (3, 9, 64, 64), (2, 17, 8, 512), (1, 16, 8, 256), (4, 77, 64, 64)])
def test_halo_dgrad_with_fused_bn_relu_backward_reduce(B, H, W, C, halo_mode):
    """dgrad epilogue fusion: ReLU gate from the saved activation + (sum g, sum g * a) in the epilogue, converted by
    tag_bn_red_act_to_xhat  ==  the separate mode-0 pass over the BatchNorm input (dbeta, dgamma)."""
    from texttoaudiogrounding_b200 import ops
    dy = _bf(torch.randn(B, H, W, C, generator=g(30))).cuda().bfloat16()
    w32 = (torch.randn(C, 3, 3, C, generator=g(31)) * (1.0 / (3 * C ** 0.5))).cuda()
    y1 = _bf(torch.randn(B, H, W, C, generator=g(32))).cuda().bfloat16()
    gamma = (torch.rand(C, generator=g(33)) + 0.5).cuda()
    gamma[1] = -gamma[1]
    beta = (torch.randn(C, generator=g(34)) * 0.3).cuda()
    mean, invstd = (torch.randn(C, generator=g(35)) * 0.1).cuda(), (torch.rand(C, generator=g(36)) + 0.5).cuda()
    scale, shift = gamma * invstd, beta - mean * gamma * invstd
    a1 = torch.empty_like(y1)                          # the activation as the forward pass saves it
    ops.scale_shift_act(y1, a1, scale, shift, C, relu=True)
    wt = ops.prep_weight_t(w32, C, C, 9, torch.bfloat16, W)
    # reference: plain dgrad, then the stand-alone reduce pass on the BatchNorm input
    da = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(dy, wt, da, None, False, None, B, H, W, C, C, 9)
    red_ref = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_bn_relu_pool_bwd", 0, y1, da, None, 1, scale, shift, mean, invstd, red_ref, 1, B, H, W, C, 0, 0, 0.0,
             0, None)
    # fused
    g_f = torch.empty_like(da)
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.conv_fwd(dy, wt, g_f, None, False, red, B, H, W, C, C, 9, bn_fuse=a1)
    ops.call("tag_bn_red_act_to_xhat", red, gamma, beta, C, 1.0, None, None)
    torch.cuda.synchronize()
    mask = a1.float() > 0
    assert torch.equal(g_f.float(), torch.where(mask, da.float(), torch.zeros_like(da.float())))
    # the gates agree except where the pre-activation rounds to +0 in bf16; xhat through the bf16 activation carries its
    # rounding (2^-9 of |a| / |gamma|)
    np.testing.assert_allclose(red[:C].cpu().numpy(), red_ref[:C].cpu().numpy(), rtol=2e-3, atol=2e-2)
    np.testing.assert_allclose(red[C:].cpu().numpy(), red_ref[C:].cpu().numpy(), rtol=1e-2, atol=0.3)


@pytest.mark.parametrize("halo_mode", [True, "single"], indirect=True)
@pytest.mark.parametrize("B,H,W,Cin,Cout,ph,pw,pdrop", [(2, 21, 16, 128, 256, 2, 2, 0.2), (3, 9, 32, 64, 128, 2, 2, 0.0),
                                                        (2, 17, 8, 256, 512, 1, 2, 0.2), (1, 16, 8, 128, 128, 1, 2, 0.0)])
def test_halo_dgrad_with_fused_pooled_bn_backward_reduce(B, H, W, Cin, Cout, ph, pw, pdrop, halo_mode):
    """The bn2 backward reductions of a pooled block from (dout, pooled output, open-gate counts) in the epilogue of the
    NEXT block's conv1 dgrad  ==  the stand-alone reduce pass over the full-resolution BatchNorm input (mode 0).
    Geometry: previous block [B, Hp, Wp, Cin] at full resolution -> pooled [B, H, W, Cin] = input of this conv (Cin -> Cout)."""
    from texttoaudiogrounding_b200 import ops
    Hp, Wp = H * ph + (1 if ph == 2 else 0), W * pw            # an odd row that floor-mode pooling drops
    y2 = _bf(torch.randn(B, Hp, Wp, Cin, generator=g(40))).cuda().bfloat16()
    gamma = (torch.rand(Cin, generator=g(41)) + 0.5).cuda()
    gamma[2] = -gamma[2]
    beta = (torch.randn(Cin, generator=g(42)) * 0.3).cuda()
    mean, invstd = (torch.randn(Cin, generator=g(43)) * 0.1).cuda(), (torch.rand(Cin, generator=g(44)) + 0.5).cuda()
    scale, shift = gamma * invstd, beta - mean * gamma * invstd
    seed = 1234
    p = torch.empty(B, H, W, Cin, device="cuda", dtype=torch.bfloat16)
    cnt = torch.empty(B, H, W, Cin, device="cuda", dtype=torch.uint8)
    ops.call("tag_bn_relu_pool_fwd", y2, p, cnt, 1, scale, shift, B, Hp, Wp, Cin, ph, pw, pdrop, seed, None)
    assert int(cnt.max()) <= 8 and set(torch.unique(cnt).tolist()) <= ({0, 5, 6, 7, 8} if ph == 2 else {0, 6, 8})
    if pdrop > 0:
        dropped = (p.float() == 0) & (cnt == 0)
        assert 0.1 < dropped.float().mean().item() < 0.6        # dropped elements carry the code 0
    # the conv whose dgrad produces d(pooled output)
    dy = _bf(torch.randn(B, H, W, Cout, generator=g(45))).cuda().bfloat16()
    w32 = (torch.randn(Cout, 3, 3, Cin, generator=g(46)) * (1.0 / (3 * Cin ** 0.5))).cuda()
    wt = ops.prep_weight_t(w32, Cout, Cin, 9, torch.bfloat16, W)
    dp = torch.empty(B, H, W, Cin, device="cuda", dtype=torch.bfloat16)
    red = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    ops.conv_fwd(dy, wt, dp, None, False, red, B, H, W, Cout, Cin, 9, bn_fuse=(p, cnt))
    ops.call("tag_bn_red_act_to_xhat", red, gamma, beta, Cin, 0.25 / (1.0 - pdrop) if pdrop > 0 else 0.25, None, None)
    dp_plain = torch.empty_like(dp)
    ops.conv_fwd(dy, wt, dp_plain, None, False, None, B, H, W, Cout, Cin, 9)
    assert torch.equal(dp, dp_plain)                            # no gating in this mode
    red_ref = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    ops.call("tag_bn_relu_pool_bwd", 0, y2, dp, None, 1, scale, shift, mean, invstd, red_ref, 1, B, Hp, Wp, Cin, ph, pw,
             pdrop, seed, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(red[:Cin].cpu().numpy(), red_ref[:Cin].cpu().numpy(), rtol=2e-3, atol=2e-2)
    # dgamma through the bf16 pooled output: 2^-9 of |dout * p| / |gamma| per window, a random walk over the windows
    np.testing.assert_allclose(red[Cin:].cpu().numpy(), red_ref[Cin:].cpu().numpy(), rtol=1e-2, atol=1.0)


def test_batched_weight_prep_equals_single_kernels():
    """tag_weight_prep_batch (one launch for all operands) == the per-weight prep kernels, bit for bit."""
    from texttoaudiogrounding_b200 import ops
    gen = torch.Generator().manual_seed(11)
    wp = ops.WeightPrep("cuda")
    cases = []
    for co, ci in [(64, 64), (128, 64), (256, 128)]:
        w = torch.randn(co, 3, 3, ci, generator=gen).cuda()
        cases.append((wp.add(w, 1, co, ci, 9), ops.prep_weight(w, torch.bfloat16, 8)))
        cases.append((wp.add(w, 2, co, ci, 9), ops.prep_weight_t(w, co, ci, 9, torch.bfloat16, 8)))
    lin = torch.randn(192, 320, generator=gen).cuda()
    cases.append((wp.add(lin, 0, 192, 320, 1), ops.to_bf16(lin).view(-1)))
    cases.append((wp.add(lin, 3, 192, 320, 1), ops.prep_weight_t(lin, 192, 320, 1, torch.bfloat16)))
    for got, _ in cases:
        got.zero_()
    wp.run()
    for i, (got, ref) in enumerate(cases):
        assert torch.equal(got.view(-1), ref.view(-1)), i


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 21, 16, 64, 64), (1, 9, 8, 128, 256), (2, 5, 32, 64, 128)])
def test_split_bf16_conv_is_fp32_accurate(B, H, W, Cin, Cout):
    """fp32 activations x split-bf16 operands on the halo tcgen05 kernel (the fp32 inference route): error ~2^-16
    relative, far inside the 1e-3 bar and 30x tighter than a TF32 convolution."""
    from texttoaudiogrounding_b200 import ops
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) * 0.05
    ref = F.conv2d(x.double(), w.double(), padding=1).float()
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = w.permute(0, 2, 3, 1).contiguous().cuda()
    assert ops.x3_eligible(wp, W)
    y = torch.empty(B, H, W, Cout, device="cuda")
    ops.conv_fwd(xn, ops.prep_weight_x3(wp, W), y, None, False, None, B, H, W, Cin, Cout, 9)
    err = (y.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-5, err
    # linear form (taps = 1) with bias + ReLU
    R = 300
    xl = torch.randn(R, Cin, generator=gen)
    wl = torch.randn(Cout, Cin, generator=gen) * 0.1
    bl = torch.randn(Cout, generator=gen)
    refl = F.relu(F.linear(xl.double(), wl.double(), bl.double())).float()
    yl = torch.empty(R, Cout, device="cuda")
    ops.conv_fwd(xl.cuda(), ops.prep_weight_x3(wl.cuda()), yl, bl.cuda(), True, None, 1, R, 1, Cin, Cout, 1)
    assert (yl.cpu() - refl).abs().max().item() / refl.abs().max().item() < 5e-5
