"""Per-kernel parity through the C ABI against fp32 PyTorch / oracle restatements of the same op,
on odd and ragged shapes (tile edges, floor-mode pooling leftovers, batch not a multiple of 8)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tag_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from texttoaudiogrounding_b200 import ops
    return ops


def g(seed=0):
    return torch.Generator().manual_seed(seed)


def test_logmel_matches_oracle_incl_short_and_tonal():
    ops = _ops()
    from texttoaudiogrounding_b200 import engine
    for L, B in [(64000, 3), (32000, 2), (6401, 1), (320000, 1)]:
        batch = O.synth_batch(B, L, seed=L % 7)
        wav = batch["waveform"]
        window, fb = O.hann_window(), O.melscale_fbanks()
        ref = O.logmel_db(wav, window, fb).transpose(1, 2)          # [B,T0,64]
        T0 = L // 320 + 1
        out = torch.empty(B, T0, 64, device="cuda")
        stats = torch.zeros(128, device="cuda", dtype=torch.float64)
        fbc = fb.cuda()
        ops.call("tag_logmel_fwd", wav.cuda(), B, L, L, window.cuda(), fbc, engine.compute_mel_range(fbc),
                 out, stats)
        err = (out.cpu() - ref).abs().max().item()
        assert err < 5e-3, (L, err)
        np.testing.assert_allclose(stats[:64].cpu().numpy(), ref.double().sum((0, 1)).numpy(), rtol=1e-4)
        # dense filterbank path (mel_range = NULL) gives the same answer
        out2 = torch.empty_like(out)
        ops.call("tag_logmel_fwd", wav.cuda(), B, L, L, window.cuda(), fbc, None, out2, None)
        assert (out2 - out).abs().max().item() < 1e-3


@pytest.mark.parametrize("L,B", [(64000, 3), (32000, 2), (6401, 1), (640, 2), (320000, 1), (319999, 2)])
def test_logmel_warp_kernel_matches_oracle(L, B):
    """tag_logmel_fwd_v2 (one warp per frame pair, register FFT): even / odd frame counts, clips shorter than two
    windows (every frame reflected), fp32 and fp16 waveforms, bn0 statistics."""
    ops = _ops()
    from texttoaudiogrounding_b200 import engine
    batch = O.synth_batch(B, L, seed=L % 7)
    wav = batch["waveform"]
    window, fb = O.hann_window(), O.melscale_fbanks()
    ref = O.logmel_db(wav, window, fb).transpose(1, 2)          # [B,T0,64]
    T0 = L // 320 + 1
    fbc = fb.cuda()
    mr = engine.compute_mel_range(fbc)
    nnz = engine.mel_nnz(mr)
    assert 0 < nnz <= engine.LOGMEL_FB_CAP
    out = torch.full((B, T0, 64), float("nan"), device="cuda")
    stats = torch.zeros(128, device="cuda", dtype=torch.float64)
    ops.call("tag_logmel_fwd_v2", wav.cuda(), 0, B, L, L, window.cuda(), fbc, mr, nnz, out, stats)
    err = (out.cpu() - ref).abs().max().item()
    assert err < 5e-3, (L, err)
    np.testing.assert_allclose(stats[:64].cpu().numpy(), ref.double().sum((0, 1)).numpy(), rtol=1e-4)
    np.testing.assert_allclose(stats[64:].cpu().numpy(), ref.double().pow(2).sum((0, 1)).numpy(), rtol=1e-4)
    # agrees with the shared-memory Stockham kernel far below the oracle tolerance
    out1 = torch.empty_like(out)
    ops.call("tag_logmel_fwd", wav.cuda(), B, L, L, window.cuda(), fbc, mr, out1, None)
    assert (out1 - out).abs().max().item() < 2e-3
    # float16 waveform (the reference's h5 storage type): identical to widening on the host first
    wav16 = wav.half()
    ref16 = O.logmel_db(wav16.float(), window, fb).transpose(1, 2)
    out16 = torch.empty_like(out)
    ops.call("tag_logmel_fwd_v2", wav16.cuda(), 2, B, L, L, window.cuda(), fbc, mr, nnz, out16, None)
    assert (out16.cpu() - ref16).abs().max().item() < 5e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 13, 8, 64, 64), (1, 7, 16, 64, 128), (3, 5, 4, 128, 256)])
def test_conv3x3_fwd_dgrad_wgrad_fp32(B, H, W, Cin, Cout):
    ops = _ops()
    x = torch.randn(B, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, 3, 3, generator=g(2)) * 0.05
    x.requires_grad_(True)
    w.requires_grad_(True)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn(y.shape, generator=g(3))
    y.backward(dy)
    xn = x.detach().permute(0, 2, 3, 1).contiguous().cuda()
    wp = w.detach().permute(0, 2, 3, 1).contiguous().cuda()
    yn = torch.empty(B, H, W, Cout, device="cuda")
    stats = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    ops.conv_fwd(xn, wp, yn, None, False, stats, B, H, W, Cin, Cout, 9)
    assert rel_err(yn.permute(0, 3, 1, 2).cpu(), y.detach()) < 1e-5
    np.testing.assert_allclose(stats[:Cout].cpu().numpy(), y.detach().double().sum((0, 2, 3)).numpy(),
                               rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[Cout:].cpu().numpy(), y.detach().double().pow(2).sum((0, 2, 3)).numpy(),
                               rtol=1e-4)
    dyn = dy.permute(0, 2, 3, 1).contiguous().cuda()
    wt = torch.empty(Cin, 3, 3, Cout, device="cuda")
    ops.call("tag_weight_flip_transpose", wp, wt, Cout, Cin, 9)
    dxn = torch.empty(B, H, W, Cin, device="cuda")
    ops.conv_fwd(dyn, wt, dxn, None, False, None, B, H, W, Cout, Cin, 9)
    assert rel_err(dxn.permute(0, 3, 1, 2).cpu(), x.grad) < 1e-5
    dw = torch.zeros(Cout, 3, 3, Cin, device="cuda")
    ops.conv_wgrad(dyn, xn, dw, B, H, W, Cin, Cout, 9, 3)
    assert rel_err(dw.permute(0, 3, 1, 2).cpu(), w.grad) < 1e-5


def test_linear_as_one_tap_conv_with_bias_relu_and_bf16_io():
    ops = _ops()
    M, K, N = 300, 512, 1536
    x = torch.randn(M, K, generator=g(4))
    w = torch.randn(N, K, generator=g(5)) * 0.03
    b = torch.randn(N, generator=g(6))
    ref = F.relu(F.linear(x, w, b))
    y = torch.empty(M, N, device="cuda")
    ops.conv_fwd(x.cuda(), w.cuda(), y, b.cuda(), True, None, 1, M, 1, K, N, 1)
    assert rel_err(y.cpu(), ref) < 1e-5
    xb = x.cuda().bfloat16()
    yb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(xb, w.cuda(), yb, b.cuda(), True, None, 1, M, 1, K, N, 1)
    ref_b = F.relu(F.linear(xb.float().cpu(), w, b))
    assert rel_err(yb.float().cpu(), ref_b) < 5e-3


def test_conv_c1_fwd_bwd():
    ops = _ops()
    B, H, W = 2, 11, 64
    x = torch.randn(B, 1, H, W, generator=g(7), requires_grad=True)
    w = (torch.randn(64, 1, 3, 3, generator=g(8)) * 0.2).requires_grad_(True)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn(y.shape, generator=g(9))
    y.backward(dy)
    xn = x.detach().reshape(B, H, W).cuda()
    yn = torch.empty(B, H, W, 64, device="cuda")
    stats = torch.zeros(128, device="cuda", dtype=torch.float64)
    ops.call("tag_conv_c1_fwd", xn, w.detach().reshape(64, 9).cuda(), yn, 0, stats, B, H, W)
    assert rel_err(yn.permute(0, 3, 1, 2).cpu(), y.detach()) < 1e-5
    np.testing.assert_allclose(stats[:64].cpu().numpy(), y.detach().double().sum((0, 2, 3)).numpy(), rtol=1e-4, atol=1e-3)
    dw = torch.zeros(64, 9, device="cuda")
    dx = torch.empty(B, H, W, device="cuda")
    ops.call("tag_conv_c1_bwd", dy.permute(0, 2, 3, 1).contiguous().cuda(), xn, w.detach().reshape(64, 9).cuda(), 0,
             dw, dx, B, H, W)
    assert rel_err(dw.cpu().reshape(64, 1, 3, 3), w.grad) < 1e-5
    assert rel_err(dx.cpu().reshape(B, 1, H, W), x.grad) < 1e-5


@pytest.mark.parametrize("B,H", [(2, 11), (3, 37), (1, 16)])
def test_conv_c1_bwd_bf16_mma(B, H):
    """bf16 Cin=1 backward (mma.sync kernel): dy, x, w rounded to bf16, fp32 accumulation — compare with
    fp32 autograd on the same rounded dy / x and the fp32 weights (tiles of 16 rows: ragged last tile, halo rows)."""
    ops = _ops()
    W = 64
    bf = torch.bfloat16
    x = torch.randn(B, 1, H, W, generator=g(17)).to(bf).float().requires_grad_(True)
    w32 = torch.randn(64, 1, 3, 3, generator=g(18)) * 0.2
    w = w32.clone().requires_grad_(True)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn(y.shape, generator=g(19)).to(bf).float()
    y.backward(dy)
    dw = torch.zeros(64, 9, device="cuda")
    dx = torch.empty(B, H, W, device="cuda")
    ops.call("tag_conv_c1_bwd", dy.permute(0, 2, 3, 1).contiguous().cuda().to(bf), x.detach().reshape(B, H, W).cuda().to(bf),
             w32.reshape(64, 9).cuda(), 1, dw, dx, B, H, W)
    # dw uses the bf16 operands exactly (fp32 accumulate; independent of w); dx uses the fp32 weights as bf16 hi + lo
    # pairs (two MMAs), i.e. w to 16 mantissa bits
    assert rel_err(dw.cpu().reshape(64, 1, 3, 3), w.grad) < 1e-4
    assert rel_err(dx.cpu().reshape(B, 1, H, W), x.grad) < 1e-4


@pytest.mark.parametrize("B,H", [(2, 11), (3, 37), (1, 16)])
def test_c1_one_pass_layer_statistics_forward_and_backward(B, H):
    """conv_block1.conv1 + bn1 + relu without the raw convolution output in HBM (Cin = 1):
    statistics from the 54 input moments == statistics of the convolution output; the one-pass forward ==
    relu(bn(conv)); the fused backward (g gated by the ReLU, BN backward applied on the recomputed conv output) ==
    autograd of conv2d -> batch_norm -> relu, in train and eval BatchNorm mode."""
    ops = _ops()
    W, C = 64, 64
    bf = torch.bfloat16
    x = torch.randn(B, 1, H, W, generator=g(41)).to(bf).float()
    w = torch.randn(C, 1, 3, 3, generator=g(42)) * 0.3
    gamma = torch.rand(C, generator=g(43)) + 0.5
    gamma[3] = -gamma[3]                                        # a negative scale
    beta = torch.randn(C, generator=g(44)) * 0.3
    xd = x.reshape(B, H, W).cuda().to(bf)
    wd = w.reshape(C, 9).cuda()
    # ---- statistics from moments
    mom = torch.empty(54, device="cuda", dtype=torch.float64)
    ops.call("tag_c1_moments", xd, 1, B, H, W, mom)
    stats = torch.empty(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_c1_stats_from_moments", mom, wd, stats)
    y = F.conv2d(x.double(), w.double(), padding=1)
    np.testing.assert_allclose(stats[:C].cpu().numpy(), y.sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(stats[C:].cpu().numpy(), y.pow(2).sum((0, 2, 3)).numpy(), rtol=1e-5)
    for training in (True, False):
        xr = x.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        yr = F.conv2d(xr, wr, padding=1)
        rm, rv = torch.randn(C, generator=g(45)) * 0.2, torch.rand(C, generator=g(46)) + 0.5
        if training:
            mean, var = yr.detach().mean((0, 2, 3)), yr.detach().var((0, 2, 3), unbiased=False)
        else:
            mean, var = rm, rv
        invstd = (var + 1e-5).rsqrt()
        a = F.relu(F.batch_norm(yr, None if training else rm, None if training else rv, gamma, beta, training, 0.1, 1e-5))
        da = torch.randn(a.shape, generator=g(47)).to(bf).float()
        a.backward(da)
        scale, shift = gamma * invstd, beta - mean * gamma * invstd
        # ---- forward
        ad = torch.empty(B, H, W, C, device="cuda", dtype=bf)
        ops.call("tag_conv_c1_fwd_act", xd, wd, scale.cuda(), shift.cuda(), ad, 1, B, H, W)
        assert rel_err(ad.float().cpu().permute(0, 3, 1, 2), a.detach()) < 4e-3
        # ---- backward: the gated gradient and the two reductions as the fused dgrad epilogue delivers them
        gate = (a.detach() > 0).float()
        gg = (da * gate)
        xhat = (yr.detach() - mean.view(1, C, 1, 1)) * invstd.view(1, C, 1, 1)
        red = torch.cat([gg.double().sum((0, 2, 3)), (gg * xhat).double().sum((0, 2, 3))]).cuda()
        dw = torch.zeros(C, 9, device="cuda")
        dx = torch.empty(B, H, W, device="cuda")
        ops.call("tag_conv_c1_bwd_bn", gg.permute(0, 2, 3, 1).contiguous().cuda().to(bf), xd, wd, scale.cuda(),
                 mean.cuda(), invstd.cuda(), red, int(training), dw, dx, B, H, W)
        # dy1 is rounded to bf16 before the two products (as the unfused path stores it)
        assert rel_err(dw.cpu().reshape(C, 1, 3, 3), wr.grad) < 6e-3, training
        assert rel_err(dx.cpu().reshape(B, 1, H, W), xr.grad) < 6e-3, training
    # ---- activation-domain reductions -> dgamma: (sum g a - beta sum g) / gamma
    red = torch.tensor([3.0, -2.0, 5.0, 7.0], dtype=torch.float64).cuda()        # C = 2: [sum g | sum g * a]
    ops.call("tag_bn_red_act_to_xhat", red, torch.tensor([2.0, 0.0]).cuda(), torch.tensor([0.5, 1.0]).cuda(), 2, 1.0, None, None)
    assert red.cpu().tolist() == [3.0, -2.0, (5.0 - 0.5 * 3.0) / 2.0, 0.0]
    red = torch.tensor([3.0, -2.0, 5.0, 7.0], dtype=torch.float64).cuda()        # with a dropout keep scale on sum g
    ops.call("tag_bn_red_act_to_xhat", red, torch.tensor([2.0, 4.0]).cuda(), torch.tensor([0.5, 1.0]).cuda(), 2, 1.25, None, None)
    assert red.cpu().tolist() == [3.75, -2.5, (5.0 - 0.5 * 3.75) / 2.0, (7.0 + 2.5) / 4.0]


@pytest.mark.parametrize("pdrop", [0.0, 0.2])
def test_freq_mean_bwd_with_fused_pooled_bn_backward_sums(pdrop):
    """Block 4: the gradient of the pooled output comes from the frequency mean; its bn2 backward sums ride along in the
    activation domain (sum dx * cnt, sum dx * p) and must equal the stand-alone reduce pass over the BatchNorm input."""
    ops = _ops()
    B, T, Wp, C = 3, 11, 8, 512                       # pooled [B, T, 4, C] from [B, T, 8, C] with (1, 2) pooling
    bf = torch.bfloat16
    y2 = torch.randn(B, T, Wp, C, generator=g(61)).cuda().to(bf)
    gamma = (torch.rand(C, generator=g(62)) + 0.5).cuda()
    gamma[5] = -gamma[5]
    beta = (torch.randn(C, generator=g(63)) * 0.3).cuda()
    mean, invstd = (torch.randn(C, generator=g(64)) * 0.1).cuda(), (torch.rand(C, generator=g(65)) + 0.5).cuda()
    scale, shift = gamma * invstd, beta - mean * gamma * invstd
    p = torch.empty(B, T, Wp // 2, C, device="cuda", dtype=bf)
    cnt = torch.empty(B, T, Wp // 2, C, device="cuda", dtype=torch.uint8)
    ops.call("tag_bn_relu_pool_fwd", y2, p, cnt, 1, scale, shift, B, T, Wp, C, 1, 2, pdrop, 77, None)
    rows = B * T
    dm = torch.randn(rows, C, generator=g(66)).cuda().to(bf)
    dp = torch.empty_like(p)
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_freq_mean_bwd", dm, 1, dp, 1, rows, Wp // 2, C, 0.5, 99, None, p, cnt, red)
    dp_plain = torch.empty_like(p)
    ops.call("tag_freq_mean_bwd", dm, 1, dp_plain, 1, rows, Wp // 2, C, 0.5, 99, None, None, None, None)
    assert torch.equal(dp, dp_plain)
    ops.call("tag_bn_red_act_to_xhat", red, gamma, beta, C, 0.25 / (1.0 - pdrop) if pdrop > 0 else 0.25, None, None)
    red_ref = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_bn_relu_pool_bwd", 0, y2, dp, None, 1, scale, shift, mean, invstd, red_ref, 1, B, T, Wp, C, 1, 2, pdrop,
             77, None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(red[:C].cpu().numpy(), red_ref[:C].cpu().numpy(), rtol=2e-3, atol=2e-2)
    np.testing.assert_allclose(red[C:].cpu().numpy(), red_ref[C:].cpu().numpy(), rtol=1e-2, atol=0.5)


@pytest.mark.parametrize("ph,pw,H", [(2, 2, 9), (1, 2, 6), (2, 2, 8)])
def test_bn_relu_pool_fwd_bwd_matches_torch_autograd(ph, pw, H):
    ops = _ops()
    B, W, C = 3, 8, 64
    y = torch.randn(B, C, H, W, generator=g(10), requires_grad=True)
    gamma = (torch.rand(C, generator=g(11)) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g(12)) * 0.2).requires_grad_(True)
    a = F.relu(F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5))
    p = F.avg_pool2d(a, (ph, pw)) + F.max_pool2d(a, (ph, pw))
    dp = torch.randn(p.shape, generator=g(13))
    p.backward(dp)
    yn = y.detach().permute(0, 2, 3, 1).contiguous().cuda()
    stats = torch.stack([yn.double().sum((0, 1, 2)), yn.double().pow(2).sum((0, 1, 2))]).reshape(-1).contiguous()
    aux = [torch.empty(C, device="cuda") for _ in range(4)]
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    ops.bn_finalize(stats, B * H * W, C, gamma.detach().cuda(), beta.detach().cuda(), rm, rv, 0.1, 1e-5, True,
                    True, *aux)
    ref_rv = 0.9 + 0.1 * y.detach().transpose(0, 1).reshape(C, -1).var(1, unbiased=True)
    np.testing.assert_allclose(rv.cpu().numpy(), ref_rv.numpy(), rtol=1e-4)
    Ho, Wo = H // ph, W // pw
    pn = torch.empty(B, Ho, Wo, C, device="cuda")
    ops.call("tag_bn_relu_pool_fwd", yn, pn, None, 0, aux[0], aux[1], B, H, W, C, ph, pw, 0.0, 0, None)
    assert rel_err(pn.permute(0, 3, 1, 2).cpu(), p.detach()) < 1e-5
    dpn = dp.permute(0, 2, 3, 1).contiguous().cuda()
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_bn_relu_pool_bwd", 0, yn, dpn, None, 0, *aux, red, 1, B, H, W, C, ph, pw, 0.0, 0, None)
    np.testing.assert_allclose(red[:C].cpu().numpy(), beta.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(red[C:].cpu().numpy(), gamma.grad.numpy(), rtol=1e-4, atol=1e-5)
    dyn = torch.empty_like(yn)
    ops.call("tag_bn_relu_pool_bwd", 1, yn, dpn, dyn, 0, *aux, red, 1, B, H, W, C, ph, pw, 0.0, 0, None)
    assert rel_err(dyn.permute(0, 3, 1, 2).cpu(), y.grad) < 1e-4


def test_bn_relu_nopool_bwd_matches_torch_autograd():
    ops = _ops()
    B, H, W, C = 2, 5, 4, 128
    y = torch.randn(B, C, H, W, generator=g(14), requires_grad=True)
    gamma = (torch.rand(C, generator=g(15)) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g(16)) * 0.2).requires_grad_(True)
    a = F.relu(F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5))
    da = torch.randn(a.shape, generator=g(17))
    a.backward(da)
    yn = y.detach().permute(0, 2, 3, 1).contiguous().cuda()
    stats = torch.stack([yn.double().sum((0, 1, 2)), yn.double().pow(2).sum((0, 1, 2))]).reshape(-1).contiguous()
    aux = [torch.empty(C, device="cuda") for _ in range(4)]
    ops.bn_finalize(stats, B * H * W, C, gamma.detach().cuda(), beta.detach().cuda(), None, None, 0.1, 1e-5, True,
                    False, *aux)
    dan = da.permute(0, 2, 3, 1).contiguous().cuda()
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.call("tag_bn_relu_pool_bwd", 0, yn, dan, None, 0, *aux, red, 1, B, H, W, C, 0, 0, 0.0, 0, None)
    np.testing.assert_allclose(red[C:].cpu().numpy(), gamma.grad.numpy(), rtol=1e-4, atol=1e-5)
    dyn = torch.empty_like(yn)
    ops.call("tag_bn_relu_pool_bwd", 1, yn, dan, dyn, 0, *aux, red, 1, B, H, W, C, 0, 0, 0.0, 0, None)
    assert rel_err(dyn.permute(0, 3, 1, 2).cpu(), y.grad) < 1e-4


@pytest.mark.parametrize("B,T", [(3, 7), (8, 25), (11, 50)])
def test_bigru_fwd_bwd_matches_oracle(B, T):
    ops = _ops()
    sd = O.synth_state_dict(seed=5)
    pre = "audio_encoder.rnn."
    x = torch.randn(B, T, 512, generator=g(18)) * 0.5
    ws = {k: sd[pre + k].clone().requires_grad_(True) for k in
          ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
           "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse"]}
    x.requires_grad_(True)
    ref = O.bigru(x, {pre + k: v for k, v in ws.items()}, pre)
    d_out = torch.randn(ref.shape, generator=g(19))
    ref.backward(d_out)

    w_ih = torch.cat([ws["weight_ih_l0"], ws["weight_ih_l0_reverse"]]).detach().cuda()
    b_ih = torch.cat([ws["bias_ih_l0"], ws["bias_ih_l0_reverse"]]).detach().cuda()
    w_hh = torch.stack([ws["weight_hh_l0"], ws["weight_hh_l0_reverse"]]).detach().cuda().contiguous()
    b_hh = torch.cat([ws["bias_hh_l0"], ws["bias_hh_l0_reverse"]]).detach().cuda()
    rows = B * T
    xc = x.detach().cuda().reshape(rows, 512)
    gi = torch.empty(rows, 1536, device="cuda")
    ops.conv_fwd(xc, w_ih, gi, b_ih, False, None, 1, rows, 1, 512, 1536, 1)
    out = torch.empty(B, T, 512, device="cuda")
    gates = torch.empty(B, T, 2, 4, 256, device="cuda")
    ops.call("tag_gru_fwd", gi, w_hh, b_hh, out, gates, B, T)
    assert (out.cpu() - ref.detach()).abs().max().item() < 2e-5
    dgi = torch.empty(rows, 1536, device="cuda")
    dgh = torch.empty(2, rows, 768, device="cuda")
    hprev = torch.empty(2, rows, 256, device="cuda")
    ops.call("tag_gru_bwd", d_out.cuda(), out, gates, w_hh, dgi, dgh, hprev, B, T)
    # input gradient and parameter gradients from dgi / dgh
    dx = dgi @ w_ih
    assert rel_err(dx.cpu().reshape(B, T, 512), x.grad) < 1e-4
    dw_ih = torch.zeros(1536, 512, device="cuda")
    ops.conv_wgrad(dgi, xc, dw_ih, 1, rows, 1, 512, 1536, 1, 2)
    ref_dw_ih = torch.cat([ws["weight_ih_l0"].grad, ws["weight_ih_l0_reverse"].grad])
    assert rel_err(dw_ih.cpu(), ref_dw_ih) < 1e-4
    for d, k in enumerate(["weight_hh_l0", "weight_hh_l0_reverse"]):
        dw = torch.zeros(768, 256, device="cuda")
        ops.conv_wgrad(dgh[d], hprev[d], dw, 1, rows, 1, 256, 768, 1, 2)
        assert rel_err(dw.cpu(), ws[k].grad) < 1e-4, k
    db = torch.zeros(768, device="cuda")
    ops.call("tag_colsum", dgh[1], 0, rows, 768, db)
    assert rel_err(db.cpu(), ws["bias_hh_l0_reverse"].grad) < 1e-4


def test_head_embed_dot_bce_fwd_bwd():
    ops = _ops()
    from texttoaudiogrounding_b200.models.text_encoder import _EmbedMeanFunction
    from texttoaudiogrounding_b200.models.match import _DotSigmoidFunction
    from texttoaudiogrounding_b200.losses import frame_bce
    B, T, N, V, D = 5, 37, 8, 101, 512
    emb = (torch.randn(V, D, generator=g(20)) * 2).requires_grad_(True)
    audio = torch.randn(B, T, D, generator=g(21)).requires_grad_(True)
    text = torch.randint(0, V, (B, N), generator=g(22))
    tl = torch.tensor([8, 3, 5, 1, 8])
    label = (torch.rand(B, T + 1, generator=g(23)) > 0.5).float()
    length = torch.tensor([37, 20, 1, 37, 30])
    # oracle
    t = O.embedding_mean({"text_encoder.embedding.core.weight": emb}, text, tl)
    sim, _ = O.dot_product_match(audio, t["seq_emb"])
    ref_loss = O.frame_bce_loss({"frame_sim": sim[:, :T], "label": label[:, :T], "length": length})
    ref_loss.backward()
    # CUDA
    e2 = emb.detach().cuda().requires_grad_(True)
    a2 = audio.detach().cuda().requires_grad_(True)
    tok, seq = _EmbedMeanFunction.apply(e2, text.cuda(), tl.cuda())
    s2 = _DotSigmoidFunction.apply(a2, seq, 1.0 / math.sqrt(D))
    loss = frame_bce(s2[:, :T], label[:, :T].cuda(), length)
    loss.backward()
    assert (s2.detach().cpu() - sim.detach()).abs().max().item() < 1e-5
    np.testing.assert_allclose(loss.item(), ref_loss.item(), rtol=1e-5)
    assert rel_err(a2.grad.cpu(), audio.grad) < 1e-4
    assert rel_err(e2.grad.cpu(), emb.grad) < 1e-4
    assert torch.equal(tok.cpu(), F.embedding(text, emb.detach()))


def test_clip_adam_matches_torch():
    ops = _ops()
    n = 10007
    p = torch.randn(n, generator=g(24))
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    pc, m, v = p.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, device="cuda", dtype=torch.int64)
    norm = torch.zeros(1, device="cuda")
    for it in range(3):
        gr = torch.randn(n, generator=g(25 + it)) * (3.0 if it == 0 else 0.001)
        ref_p.grad = gr.clone()
        tn = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ss = torch.zeros(1, device="cuda", dtype=torch.float64)
        gc = gr.cuda()
        ops.call("tag_sumsq", gc, n, ss)
        ops.call("tag_clip_adam", pc, gc, m, v, n, ss, step, 1.0, 1.0, 1e-3, None, 0.9, 0.999, 1e-8, norm)
        np.testing.assert_allclose(norm.item(), float(tn), rtol=1e-5)
        assert (pc.cpu() - ref_p.detach()).abs().max().item() < 2e-6


@pytest.mark.parametrize("mode", ["cosine", "cosine_unscaled", "expnegl2", "expnegl2_raw"])
def test_normalised_match_heads_match_torch_autograd(mode):
    """DotProduct(l2norm=True) (cosine) and ExpNegL2 (models/match.py:10-60) against autograd of the oracle."""
    from texttoaudiogrounding_b200.models.match import DotProduct, ExpNegL2
    B, T, D = 3, 41, 512
    a = torch.randn(B, T, D, generator=g(1)).requires_grad_(True)
    s = (torch.randn(B, D, generator=g(2)) * (0.05 if mode == "expnegl2_raw" else 1.0)).requires_grad_(True)
    if mode == "expnegl2_raw":
        a.data.mul_(0.05)
    w = torch.randn(B, T, generator=g(3))
    if mode.startswith("cosine"):
        ref = O.dot_product_match_l2norm(a, s, scale=mode == "cosine")
        head = DotProduct(l2norm=True, scale=mode == "cosine")
    else:
        ref = O.exp_neg_l2_match(a, s, l2norm=mode == "expnegl2")
        head = ExpNegL2(l2norm=mode == "expnegl2")
    (ref * w).sum().backward()
    ac, sc = a.detach().cuda().requires_grad_(True), s.detach().cuda().requires_grad_(True)
    out = head({"audio_emb": ac, "text_emb": {"seq_emb": sc}})
    (out * w.cuda()).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() < 1e-6
    assert rel_err(ac.grad.cpu(), a.grad) < 1e-4
    assert rel_err(sc.grad.cpu(), s.grad) < 1e-4
