"""Backward kernels (tcgen05 backward-filter GEMM, backward-data epilogue, pool / bias / loss / Adam kernels)
against torch fp64 autograd of the same ops (GPU only).  Tolerance: 3-pass mode <= 5e-5 of max|ref|."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def _pad_from_dense(k, x):
    return k.pad_nhwc(x.contiguous(), precise=True)


@pytest.mark.parametrize("B,H,W,Cin,Cout,tap_window,passes", [
    (1, 9, 11, 64, 64, 1, 3), (1, 9, 11, 64, 64, 0, 3), (2, 16, 20, 64, 128, 1, 3), (1, 23, 30, 128, 256, 1, 3),
    (1, 23, 30, 128, 256, 0, 3), (1, 37, 41, 36, 64, 1, 3), (1, 40, 33, 3, 64, 1, 3), (1, 40, 33, 3, 64, 0, 3),
    (1, 12, 13, 512, 512, 1, 3), (1, 30, 21, 256, 128, 1, 1),
])
def test_conv_wgrad(B, H, W, Cin, Cout, tap_window, passes):
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(H * 131 + W + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=gen)
    g = torch.randn(B, H, W, Cout, device="cuda", generator=gen)
    precise = passes == 3
    xa, ga = k.pad_nhwc(x, precise=precise), k.pad_nhwc(g, precise=precise)
    dw = torch.zeros(3, 3, Cin, Cout, device="cuda")
    k.conv_wgrad(xa, ga, dw, precise=precise, accumulate=True, tap_window=bool(tap_window))
    torch.cuda.synchronize()
    xr, gr = (x.double(), g.double()) if precise else (x.bfloat16().double(), g.bfloat16().double())
    ref = torch.nn.grad.conv2d_weight(xr.permute(0, 3, 1, 2), (Cout, Cin, 3, 3), gr.permute(0, 3, 1, 2), padding=1)
    ref = ref.permute(2, 3, 1, 0)  # OIHW -> HWIO
    assert _relerr(dw, ref) < (5e-5 if precise else 1e-5)
    # accumulate semantics: a second call doubles the result
    k.conv_wgrad(xa, ga, dw, precise=precise, accumulate=True, tap_window=bool(tap_window))
    torch.cuda.synchronize()
    assert _relerr(dw, 2 * ref) < (5e-5 if precise else 1e-5)


def test_first_layer_wgrad_through_im2col():
    """Backward-filter of the first image layer as ONE taps = 1 GEMM over the im2col rows (K = 27 of 32, wgrad_kernel<32>)
    == the nine-tap form over the 16-channel PAD image == torch."""
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, Cout = 2, 37, 120, 64
    x = torch.randn(B, H, W, 3, device="cuda", generator=gen) * 40
    g = torch.randn(B, H, W, Cout, device="cuda", generator=gen)
    ga = k.pad_nhwc(g, precise=True)
    col = k.im2col3x3(x, precise=True)
    assert col.c_pad == 32
    dw = torch.zeros(3, 3, 3, Cout, device="cuda")
    k.conv_wgrad(col, ga, dw.view(1, 27, Cout), precise=True, accumulate=True, tap_window=False)
    dw9 = torch.zeros(3, 3, 3, Cout, device="cuda")
    k.conv_wgrad(k.pad_nhwc(x, precise=True), ga, dw9, precise=True, accumulate=True)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (Cout, 3, 3, 3), g.double().permute(0, 3, 1, 2), padding=1)
    ref = ref.permute(2, 3, 1, 0)
    assert _relerr(dw, ref) < 5e-5 and _relerr(dw9, ref) < 5e-5


@pytest.mark.parametrize("R,K,N,accumulate", [(100, 1024, 200, 1), (256, 3136, 2048, 0), (77, 128, 50, 1), (300, 4096, 50, 1)])
def test_linear_wgrad(R, K, N, accumulate):
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(R + K)
    x = torch.randn(R, K, device="cuda", generator=gen)
    g = torch.zeros(R, k.round_up(N, 64), device="cuda")
    g[:, :N] = torch.randn(R, N, device="cuda", generator=gen)
    x_hi, x_lo = _split(x)
    g_hi, g_lo = _split(g)
    dw = torch.zeros(K, N, device="cuda") if accumulate else torch.full((K, N), 7.0, device="cuda")
    k.linear_wgrad(x_hi, x_lo, g_hi, g_lo, dw, precise=True, accumulate=bool(accumulate))
    torch.cuda.synchronize()
    ref = x.double().t() @ g[:, :N].double()
    assert _relerr(dw, ref) < 5e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 9, 11, 64, 64), (2, 16, 20, 64, 128), (1, 23, 30, 256, 128), (1, 12, 13, 512, 512)])
def test_conv_dgrad_with_mask_and_addend(B, H, W, Cin, Cout):
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(H * 7 + W + Cout)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=gen) * (2.0 / (9 * Cin)) ** 0.5
    g = torch.randn(B, H, W, Cout, device="cuda", generator=gen)
    xact = torch.relu(torch.randn(B, H, W, Cin, device="cuda", generator=gen))   # forward activation (ReLU output)
    addend = torch.randn(B, H, W, Cin, device="cuda", generator=gen)
    ga = k.pad_nhwc(g, precise=True)
    xa = k.pad_nhwc(xact, precise=True)
    wd = k.pack_weights_dgrad(w)
    out, dense = k.conv(ga, wd, relu=False, precise=True, out_pad=True, out_f32_dense=True, mask=xa, addend=addend,
                        use_bias=False)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w.double().permute(3, 2, 0, 1), g.double().permute(0, 3, 1, 2),
                                     padding=1).permute(0, 2, 3, 1)
    ref = (ref + addend.double()) * (xact > 0)
    assert _relerr(dense, ref) < 5e-5
    assert _relerr(k.unpad_nhwc(out), ref) < 5e-5
    assert float(out.hi[:, :, 0, :].abs().max()) == 0 and float(out.hi[:, H, :, :].abs().max()) == 0


def test_linear_dgrad_mask():
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(5)
    R, K, N = 200, 2048, 2048
    w = torch.randn(K, N, device="cuda", generator=gen) * 0.02
    g = torch.randn(R, N, device="cuda", generator=gen)
    act = torch.relu(torch.randn(R, K, device="cuda", generator=gen))
    act_hi, _ = _split(act)
    g_hi, g_lo = _split(g)
    wd = k.pack_weights_dgrad(w, cout_pad=N)
    hi, lo, f32 = k.linear(g_hi, g_lo, wd, relu=False, precise=True, out_bf16=True, out_f32=True, mask_hi=act_hi,
                           mask_scale=2.0, use_bias=False)
    torch.cuda.synchronize()
    ref = (g.double() @ w.double().t()) * (act_hi.float() > 0) * 2.0
    assert _relerr(f32, ref) < 5e-5
    assert _relerr(hi.float() + lo.float(), ref) < 5e-5


@pytest.mark.parametrize("B,H,W,C", [(1, 8, 10, 64), (2, 13, 17, 128), (1, 75, 75, 64)])
def test_maxpool_bwd(B, H, W, C):
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(H + W)
    x = torch.relu(torch.randn(B, H, W, C, device="cuda", generator=gen))
    x[:, ::3, ::2] = 0  # plenty of all-zero / tied windows
    xa = k.pad_nhwc(x, precise=True)
    xq = k.unpad_nhwc(xa).double().requires_grad_(True)   # the values the kernels actually see
    y = torch.nn.functional.max_pool2d(xq.permute(0, 3, 1, 2), 2, 2)
    g = torch.randn(B, H // 2, W // 2, C, device="cuda", generator=gen)
    ga = k.pad_nhwc(g, precise=True)
    gq = k.unpad_nhwc(ga).double()
    y.backward(gq.permute(0, 3, 1, 2))
    ref = xq.grad * (xq > 0)   # ReLU gate of the producing conv
    out = k.maxpool2x2_bwd(xa, ga)
    torch.cuda.synchronize()
    got = k.unpad_nhwc(out).double()
    # ties between strictly positive values have probability ~0 with random data; zero windows are gated off
    assert torch.equal(got, ref)
    assert float(out.hi[:, :, 0, :].abs().max()) == 0 and float(out.hi[:, H, :, :].abs().max()) == 0


def test_bias_grad_and_masked_pad():
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(11)
    g = torch.randn(2, 19, 23, 512, device="cuda", generator=gen)
    act = torch.relu(torch.randn(2, 19, 23, 512, device="cuda", generator=gen))
    aa = k.pad_nhwc(act, precise=True)
    ga = k.pad_nhwc_masked(g, aa, precise=True)
    ref = g * (act.bfloat16().float() > 0)
    assert _relerr(k.unpad_nhwc(ga), ref) < 1e-5
    db = torch.zeros(512, device="cuda")
    k.bias_grad(ga.hi, ga.lo, 512, db)
    torch.cuda.synchronize()
    assert _relerr(db, k.unpad_nhwc(ga).double().sum(dim=(0, 1, 2))) < 1e-5
    g2 = torch.randn(300, 64, device="cuda", generator=gen)
    h2, l2 = _split(g2)
    db2 = torch.zeros(50, device="cuda")
    k.bias_grad(h2, l2, 50, db2)
    assert _relerr(db2, (h2.double() + l2.double())[:, :50].sum(0)) < 1e-5


def test_losses_match_autograd():
    from mv3d_tf_b200._lib import check, current_stream, lib, ptr

    gen = torch.Generator(device="cuda").manual_seed(3)
    B, Hf, Wf, A = 2, 11, 13, 4
    cls = torch.randn(B, Hf, Wf, 2 * A, device="cuda", generator=gen, dtype=torch.float32)
    bbox = torch.randn(B, Hf, Wf, 6 * A, device="cuda", generator=gen) * 0.3
    labels = torch.randint(-1, 2, (B, Hf, Wf, A), device="cuda", generator=gen).float()
    targets = torch.randn(B, Hf * Wf * A, 6, device="cuda", generator=gen) * 0.3
    counts = torch.stack([torch.stack([(labels[b] != -1).sum(), (labels[b] == 1).sum()]) for b in range(B)]).int()
    c_pad = 64
    gh = torch.full((B, Hf + 1, Wf + 1, c_pad), 9.0, device="cuda", dtype=torch.bfloat16)
    gl = torch.full_like(gh, 9.0)
    loss = torch.zeros(4, device="cuda")
    check(lib().mv3d_rpn_loss(ptr(cls), ptr(bbox), ptr(labels), ptr(targets), ptr(counts), B, Hf, Wf, A, c_pad, 3.0,
                              ptr(gh), ptr(gl), ptr(loss), current_stream()), "mv3d_rpn_loss")
    torch.cuda.synchronize()

    def sl1(d, sigma2=9.0):
        return torch.where(d.abs() < 1 / sigma2, 0.5 * sigma2 * d * d, d.abs() - 0.5 / sigma2)
    cr = cls.double().requires_grad_(True)
    br = bbox.double().requires_grad_(True)
    tot_c = tot_b = 0
    for b in range(B):
        lab = labels[b].reshape(-1)
        sc = cr[b].reshape(-1, 2)
        keep = lab != -1
        tot_c = tot_c + torch.nn.functional.cross_entropy(sc[keep], lab[keep].long()) / B
        pos = lab == 1
        tot_b = tot_b + sl1(br[b].reshape(-1, 6)[pos] - targets[b].double()[pos]).sum(1).mean() / B
    (tot_c + tot_b).backward()
    assert abs(float(loss[0]) - float(tot_c)) < 1e-5 * max(1, abs(float(tot_c)))
    assert abs(float(loss[1]) - float(tot_b)) < 1e-5 * max(1, abs(float(tot_b)))
    got = (gh.float() + gl.float())[:, :Hf, 1:, :]
    assert _relerr(got[..., :2 * A], cr.grad) < 1e-4
    assert _relerr(got[..., 2 * A:8 * A], br.grad) < 1e-4
    assert float(got[..., 8 * A:].abs().max()) == 0 and float(gh[:, Hf].abs().max()) == 0 and float(gh[:, :, 0].abs().max()) == 0

    R, nb = 200, 48
    cs = torch.randn(R, 50, device="cuda", generator=gen)
    lab = torch.randint(0, 2, (R,), device="cuda", generator=gen).int()
    tg = torch.randn(R, nb, device="cuda", generator=gen) * 0.2
    rois = torch.zeros(R, 5, device="cuda")
    rois[120:, 0] = 1
    fc = torch.tensor([120, 80], device="cuda", dtype=torch.int32)
    g2h = torch.empty(R, 64, device="cuda", dtype=torch.bfloat16)
    g2l = torch.empty_like(g2h)
    loss2 = torch.zeros(2, device="cuda")
    check(lib().mv3d_rcnn_loss(ptr(cs), 50, ptr(cs[:, 2:]), 50, ptr(lab), ptr(tg), nb, ptr(rois), ptr(fc), 2, R, 64, 3.0,
                               ptr(g2h), ptr(g2l), ptr(loss2), current_stream()), "mv3d_rcnn_loss")
    torch.cuda.synchronize()
    csr = cs.double().requires_grad_(True)
    lc = lb = 0
    for b, sl in enumerate((slice(0, 120), slice(120, 200))):
        lc = lc + torch.nn.functional.cross_entropy(csr[sl, :2], lab[sl].long()) / 2
        lb = lb + sl1(csr[sl, 2:] - tg.double()[sl]).sum(1).mean() / 2
    (lc + lb).backward()
    assert abs(float(loss2[0]) - float(lc)) < 1e-5 and abs(float(loss2[1]) - float(lb)) < 1e-5 * max(1, float(lb))
    assert _relerr((g2h.float() + g2l.float())[:, :50], csr.grad) < 1e-4


def test_adam_matches_tf_formula():
    """tf.train.AdamOptimizer (TF 1.0): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps) -- note the
    epsilon sits outside the bias correction, unlike torch.optim.Adam, so the reference here is the formula itself."""
    from mv3d_tf_b200._lib import check, current_stream, lib, ptr

    gen = torch.Generator(device="cuda").manual_seed(9)
    n = 100003
    theta = torch.randn(n, device="cuda", generator=gen)
    ref = theta.clone().double()
    rm, rv = torch.zeros_like(ref), torch.zeros_like(ref)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    for step in range(1, 4):
        g = torch.randn(n, device="cuda", generator=gen)
        check(lib().mv3d_adam(ptr(theta), ptr(g), ptr(m), ptr(v), n, lr, b1, b2, eps, step, 0.5,
                              current_stream()), "mv3d_adam")
        gd = g.double() * 0.5
        rm = b1 * rm + (1 - b1) * gd
        rv = b2 * rv + (1 - b2) * gd * gd
        lr_t = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
        ref = ref - lr_t * rm / (rv.sqrt() + eps)
    torch.cuda.synchronize()
    assert float((theta.double() - ref).abs().max()) < 1e-6
    # (1 - beta2) is formed in float32 like TF's fp32 variables do: 1 - 0.999f carries a 4.7e-5 relative error
    assert _relerr(m, rm) < 1e-5 and _relerr(v, rv) < 1e-4


@pytest.mark.parametrize("kh,cin,cout,off", [(3, 64, 64, 0), (3, 36, 128, 0), (1, 4096, 50, 0), (1, 512, 24, 0), (3, 128, 256, 2),
                                              (1, 200, 2048, 0), (3, 9, 64, 1), (1, 4096, 2, 0)])
def test_weight_packers_exact(kh, cin, cout, off):
    """pack_weights (forward, K-major transposed) and pack_weights_dgrad (flipped taps) against the index formulas of
    include/mv3d_b200.h, bit for bit -- every kernel variant (64x64 tiles, 32x32 tiles, scalar; vector / scalar dgrad),
    including a weight view that is not 16-byte aligned (`off` floats into a flat buffer, as in the trainer)."""
    from mv3d_tf_b200 import kernels as k

    gen = torch.Generator(device="cuda").manual_seed(kh * 1000 + cin + cout)
    flat = torch.randn(off + kh * kh * cin * cout, device="cuda", generator=gen)
    w = flat[off:].view(kh, kh, cin, cout)
    taps = kh * kh
    pw = k.pack_weights(w, None)
    cp = pw.cin_pad
    ref = torch.zeros(cout, taps, cp, device="cuda")
    ref[:, :, :cin] = w.reshape(taps, cin, cout).permute(2, 0, 1)
    ref = ref.reshape(cout, taps * cp)
    hi = ref.bfloat16()
    assert torch.equal(pw.hi, hi) and torch.equal(pw.lo, (ref - hi.float()).bfloat16())
    pd = k.pack_weights_dgrad(w)
    cop = pd.cin_pad                     # (PackedWeight of the dgrad operand: its K extent is the padded cout)
    refd = torch.zeros(cin, taps, cop, device="cuda")
    refd[:, :, :cout] = w.reshape(taps, cin, cout).flip(0).permute(1, 0, 2)
    refd = refd.reshape(cin, taps * cop)
    hid = refd.bfloat16()
    assert torch.equal(pd.hi, hid) and torch.equal(pd.lo, (refd - hid.float()).bfloat16())
