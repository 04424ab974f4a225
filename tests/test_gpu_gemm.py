"""tcgen05 implicit-GEMM conv / fc against a torch fp64 reference of the same op (GPU only).
Tolerances: 3-pass (hi/lo split) mode <= 2e-5 of max|ref|; single-pass bf16 <= 2e-2."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def _relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K,passes,split_k", [
    (128, 32, 64, 1, 1), (128, 64, 64, 3, 1), (300, 128, 256, 3, 1), (1000, 512, 1152, 3, 1), (1000, 512, 1152, 1, 1),
    (257, 50, 4096, 3, 1), (300, 2048, 3136, 3, 4), (77, 24, 48, 3, 1), (640, 256, 16, 1, 1), (130, 8, 512, 3, 1),
    # split-K fc over <= 512 rows with N % 256 == 0: the swapped CTA-pair kernel (weights on the M side)
    (128, 256, 512, 3, 2), (512, 512, 1024, 3, 3), (65, 256, 192, 3, 2), (300, 2048, 25088, 3, 3), (256, 2048, 2048, 3, 3),
])
def test_fc_gemm(M, N, K, passes, split_k):
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(K, N, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    pw = k.pack_weights(w, b, cin_pad=K)
    a_hi, a_lo = _split(a)
    if passes == 3 and split_k > 1 and N % 256 == 0 and 64 <= M <= 512 and K % 64 == 0:
        assert k.gemm_kernel_name(1, K, N, 3, split_k, M, 0) == "fc_swapped_pair_kernel"
    hi, lo, f32 = k.linear(a_hi, a_lo, pw, relu=True, precise=(passes == 3), out_bf16=True, out_f32=True, split_k=split_k)
    torch.cuda.synchronize()
    if passes == 3:
        ref = torch.relu(a.double() @ w.double() + b.double())
        tol = 3e-5
    else:
        ref = torch.relu(a_hi.double() @ pw.hi.double().t() + b.double())  # same rounded operands
        tol = 1e-5
    assert f32.shape == (M, N)
    assert _relerr(f32, ref) < tol
    got = hi[:, :N].float() + (lo[:, :N].float() if lo is not None else 0)
    assert _relerr(got, ref) < (tol if lo is not None else 1e-2)


def _f16e5_rows(x):
    """(R, K) float32 -> the f16e5 operand rows (fp16 plane, byte plane) in 16-bit containers, built with torch: per
    64-element chunk 64 x e5m2(h) then 64 x e5m2((x - h) * 4096) (csrc/common.cuh)."""
    R, Kd = x.shape
    h = x.half()
    h8 = h.float().to(torch.float8_e5m2).view(torch.uint8)
    l8 = ((x - h.float()) * 4096.0).to(torch.float8_e5m2).view(torch.uint8)
    planes = torch.stack((h8.view(R, Kd // 64, 64), l8.view(R, Kd // 64, 64)), dim=2).reshape(R, 2 * Kd)
    return h.view(torch.bfloat16).contiguous(), planes.contiguous().view(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(300, 2048, 25088), (64, 256, 128), (301, 512, 1024), (512, 256, 4096)])
def test_fc_gemm_f16e5_operands(M, N, K):
    """fc over a few hundred rows with f16e5 operands (mv3d_gemm_desc.passes = 2 on the swapped split-K kernel: one fp16
    pass + one e5m2 pass carrying both first-order correction terms) vs torch float64; error ~2^-14.5 per product."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.relu(torch.randn(M, K, device="cuda", generator=g))
    w = torch.randn(K, N, device="cuda", generator=g) * 0.02
    b = torch.randn(N, device="cuda", generator=g)
    pw = k.pack_weights(w, b, cin_pad=K, fmt=k.FMT_F16E5)
    a_hi, a_lo = _f16e5_rows(a)
    assert k.gemm_kernel_name(1, K, N, 2, 3, M, 0) == "fc_swapped_pair_kernel"
    hi, lo, f32 = k.linear(a_hi, a_lo, pw, relu=True, precise=True, out_bf16=True, out_f32=True, split_k=3)
    torch.cuda.synchronize()
    ref = torch.relu(a.double() @ w.double() + b.double())
    assert _relerr(f32, ref) < 1e-4
    assert _relerr(hi[:, :N].float() + lo[:, :N].float(), ref) < 1e-4
    # and against the 3-pass result of the same layer
    a_h3, a_l3 = _split(a)
    _, _, f3 = k.linear(a_h3, a_l3, k.pack_weights(w, b, cin_pad=K), relu=True, precise=True, out_bf16=False, out_f32=True, split_k=3)
    assert _relerr(f32, f3) < 1e-4


@pytest.mark.parametrize("B,H,W,Cin,Cout,passes", [
    (1, 9, 11, 3, 64, 3), (2, 16, 20, 64, 64, 3), (1, 37, 41, 36, 64, 3), (1, 23, 50, 128, 256, 3),
    (1, 23, 50, 128, 256, 1), (1, 12, 13, 512, 512, 3), (1, 75, 75, 64, 128, 3), (1, 40, 33, 9, 64, 1),
    (1, 21, 40, 20, 64, 3), (1, 21, 40, 33, 256, 3),
])
def test_conv3x3(B, H, W, Cin, Cout, passes):
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H * 31 + W)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    precise = passes == 3
    a = k.pad_nhwc(x, precise=precise)
    assert torch.equal(k.unpad_nhwc(a), x if precise else x.bfloat16().float()) or precise
    pw = k.pack_weights(w, b)
    out, dense = k.conv(a, pw, relu=True, precise=precise, out_pad=True, out_f32_dense=True)
    torch.cuda.synchronize()
    if precise:
        xr, wr, tol = x.double(), w.double(), 3e-5
    else:
        xr, wr, tol = x.bfloat16().double(), w.bfloat16().double(), 1e-5
    ref = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), b.double(), padding=1)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    assert _relerr(dense, ref) < tol
    got = k.unpad_nhwc(out)
    assert _relerr(got, ref) < (tol if precise else 1e-2)
    # halo pixels of the PAD output must be exactly zero (they are the next layer's padding)
    assert float(out.hi[:, :, 0, :].abs().max()) == 0 and float(out.hi[:, H, :, :].abs().max()) == 0
    # chained: pool then a second conv consumes the PAD output directly
    if H >= 4 and W >= 4 and Cout <= 128:
        pooled = k.maxpool2x2(out)
        pr = torch.nn.functional.max_pool2d(ref.permute(0, 3, 1, 2), 2, 2)
        assert _relerr(k.unpad_nhwc(pooled), pr.permute(0, 2, 3, 1)) < (tol if precise else 1e-2)
        w2 = torch.randn(3, 3, Cout, 32, device="cuda", generator=g) * 0.05
        pw2 = k.pack_weights(w2, None)
        _, d2 = k.conv(pooled, pw2, relu=False, precise=precise, out_pad=False, out_f32_dense=True)
        r2 = torch.nn.functional.conv2d(pr if precise else k.unpad_nhwc(pooled).double().permute(0, 3, 1, 2),
                                        (w2.double() if precise else w2.bfloat16().double()).permute(3, 2, 0, 1),
                                        None, padding=1).permute(0, 2, 3, 1)
        assert _relerr(d2, r2) < (5e-5 if precise else 1e-5)


def test_softmax_pairs():
    from mv3d_tf_b200 import kernels as k

    x = torch.randn(1, 13, 17, 8, device="cuda") * 3
    y = k.softmax_pairs(x, 4)
    ref = torch.softmax(x.view(-1, 4, 2).double(), dim=-1).view(x.shape)
    assert _relerr(y, ref) < 1e-6


@pytest.mark.parametrize("B,H,W,Cout,passes", [(1, 9, 11, 64, 3), (2, 37, 50, 64, 3), (1, 64, 96, 64, 1)])
def test_first_layer_im2col_conv(B, H, W, Cout, passes):
    """conv1_1 on a 3-channel input through the im2col + single K=32 GEMM path == the nine-tap path == torch."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H + W)
    x = torch.randn(B, H, W, 3, device="cuda", generator=g) * 50
    w = torch.randn(3, 3, 3, Cout, device="cuda", generator=g) * 0.2
    b = torch.randn(Cout, device="cuda", generator=g)
    precise = passes == 3
    col = k.im2col3x3(x, precise=precise)
    pw = k.pack_weights(w.reshape(1, 1, 27, Cout), b, cin_pad=32)
    out, dense = k.conv(col, pw, relu=True, precise=precise, out_pad=True, out_f32_dense=True)
    out9, dense9 = k.conv(k.pad_nhwc(x, precise=precise), k.pack_weights(w, b), relu=True, precise=precise,
                          out_pad=True, out_f32_dense=True)
    torch.cuda.synchronize()
    xr, wr = (x.double(), w.double()) if precise else (x.bfloat16().double(), w.bfloat16().double())
    ref = torch.relu(torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), b.double(), padding=1))
    ref = ref.permute(0, 2, 3, 1)
    assert _relerr(dense, ref) < (3e-5 if precise else 1e-5)
    assert _relerr(dense, dense9) < 3e-5
    assert _relerr(k.unpad_nhwc(out), ref) < (3e-5 if precise else 1e-2)
    assert float(out.hi[:, :, 0, :].abs().max()) == 0 and float(out.hi[:, H, :, :].abs().max()) == 0


@pytest.mark.parametrize("B,H,W,Cin,Cout,fmt", [(1, 9, 11, 3, 64, 0), (2, 37, 50, 3, 64, 1), (1, 64, 96, 3, 64, 1),
                                                 (1, 33, 40, 1, 64, 0), (1, 20, 31, 4, 128, 1), (1, 17, 300, 2, 64, 1),
                                                 (1, 375, 1242, 3, 64, 1), (1, 64, 512, 3, 128, 0)])
def test_first_layer_direct_kernel(B, H, W, Cin, Cout, fmt):
    """conv1_1 on an image-like input vs torch in float64: Cout = 64 and Cin <= 3 go through the tcgen05 form (operand rows
    built in shared memory by the CTA, bf16 hi/lo 3-pass MMAs, first_layer_tcgen05.cu), anything else through the direct
    fp32 FMA kernel; the PAD output carries the consumer's operand format (0 = bf16 hi/lo, 1 = f16e5) and zero halos."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H * W + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g) * 50
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * 0.2
    b = torch.randn(Cout, device="cuda", generator=g)
    out = k.conv3x3_small_cin(x, w, b, relu=True, precise=True, out_fmt=fmt)
    torch.cuda.synchronize()
    assert out.fmt == fmt and out.c_pad == Cout
    ref = torch.relu(torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(),
                                                padding=1)).permute(0, 2, 3, 1)
    # operand rendering keeps ~2^-15 (f16e5) / 2^-17 (bf16 hi/lo) of each value; the fp32 accumulation itself ~1e-7
    assert _relerr(k.unpad_nhwc(out), ref) < (8e-5 if fmt else 2e-5)
    # the rendering is exactly that of the generic path applied to the same float32 values
    dense = k.unpad_nhwc(out)
    again = k.pad_nhwc(dense, precise=True, fmt=fmt)
    assert torch.equal(k.unpad_nhwc(again), dense)
    hi = out.hi.view(torch.int16)
    assert int(hi[:, :, 0, :].abs().max()) == 0 and int(hi[:, H, :, :].abs().max()) == 0
    if out.lo is not None:
        lo = out.lo.view(torch.int16)
        assert int(lo[:, :, 0, :].abs().max()) == 0 and int(lo[:, H, :, :].abs().max()) == 0


@pytest.mark.parametrize("B,H,W,Cin,Cout,passes", [
    (1, 12, 13, 64, 64, 3), (1, 23, 50, 128, 256, 3), (1, 23, 50, 128, 256, 1), (1, 30, 33, 256, 512, 3),
    (2, 75, 75, 64, 128, 3), (1, 200, 300, 64, 64, 3), (1, 87, 100, 512, 512, 3),
])
def test_conv3x3_pair_kernel_equals_single_cta(B, H, W, Cin, Cout, passes):
    """The CTA-pair (cta_group::2, 256 x N tiles) kernel against the single-CTA tap-reuse kernel and torch: the two
    kernels issue the same K order into fp32 TMEM accumulators, so their outputs are expected to be identical."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H * 131 + W + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    precise = passes == 3
    a = k.pad_nhwc(x, precise=precise)
    pw = k.pack_weights(w, b)
    prev = k.set_pair_mode(True)
    try:
        assert k.gemm_kernel_name(9, a.hi.shape[-1], Cout, passes).startswith("conv3x3_pair_kernel")
        out_p, dense_p = k.conv(a, pw, relu=True, precise=precise, out_pad=True, out_f32_dense=True)
        k.set_pair_mode(False)
        out_s, dense_s = k.conv(a, pw, relu=True, precise=precise, out_pad=True, out_f32_dense=True)
        torch.cuda.synchronize()
    finally:
        k.set_pair_mode(prev)
    xr, wr = (x.double(), w.double()) if precise else (x.bfloat16().double(), w.bfloat16().double())
    ref = torch.relu(torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), b.double(), padding=1))
    ref = ref.permute(0, 2, 3, 1)
    assert _relerr(dense_p, ref) < (3e-5 if precise else 1e-5)
    assert torch.equal(dense_p, dense_s)
    assert torch.equal(out_p.hi, out_s.hi)
    if precise:
        assert torch.equal(out_p.lo, out_s.lo)


def _e5m2(x):
    return x.float().to(torch.float8_e5m2).double()


def _f16e5_conv_emulation(x, w, b):
    """fp64 evaluation of exactly the operands the f16e5 mode feeds the tensor cores (include/mv3d_b200.h):
    fp16(x) (x) fp16(4096 w) + [e5m2(x_h) (x) e5m2(residual_w) + e5m2(residual_x * 4096) (x) e5m2(w)], times 2^-12."""
    S = 4096.0
    xh = x.clamp(-65504, 65504).half().double()
    xh8, xl8 = _e5m2(xh), _e5m2((x.double() - xh) * S)
    wh = (w * S).half().double()
    wl8, wh8 = _e5m2(w.double() * S - wh), _e5m2(w)
    cv = lambda a, k: torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), k.permute(3, 2, 0, 1), None, padding=1)
    y = (cv(xh, wh) + cv(xh8, wl8) + cv(xl8, wh8)) / S + b.double().view(1, -1, 1, 1)
    return torch.relu(y).permute(0, 2, 3, 1)


@pytest.mark.parametrize("pair", [True, False])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (1, 12, 13, 64, 64), (1, 23, 50, 128, 256), (1, 30, 33, 256, 512), (2, 75, 75, 64, 128), (1, 87, 100, 512, 512),
    (1, 37, 41, 36, 64), (1, 20, 60, 17, 64), (1, 20, 60, 48, 128),   # zero-padded 64-channel chunk: k-steps of zeros skipped
])
def test_conv3x3_f16e5(B, H, W, Cin, Cout, pair):
    """passes=2: fp16 main pass + one e5m2 pass carrying both first-order correction terms.  Checked (a) against an fp64
    evaluation of the very same quantised operands (layout / pairing / scaling exact up to fp32 accumulation) and
    (b) against the true fp64 convolution: <= 1.5e-4 of max|ref| per layer (measured ~3e-5; bf16 3-pass ~1e-5,
    plain fp16 ~4e-4), then pooled and chained into a second f16e5 conv."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H * 17 + W + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).abs() * 3
    x = x * (torch.rand(B, H, W, Cin, device="cuda", generator=g) < 0.6)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    prev = k.set_pair_mode(pair)
    try:
        a = k.pad_nhwc(x, fmt=k.FMT_F16E5)
        # the format itself: x ~= h + l/4096 to ~2^-14.5
        assert _relerr(k.unpad_nhwc(a), x) < 6e-5
        pw = k.pack_weights(w, b, fmt=k.FMT_F16E5)
        out, dense = k.conv(a, pw, relu=True, out_pad=True, out_f32_dense=True, out_fmt=k.FMT_F16E5)
        torch.cuda.synchronize()
        emu = _f16e5_conv_emulation(x, w, b)
        ref = torch.relu(torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1),
                                                    b.double(), padding=1)).permute(0, 2, 3, 1)
        assert _relerr(dense, emu) < 3e-5
        assert _relerr(dense, ref) < 1.5e-4
        assert _relerr(k.unpad_nhwc(out), dense) < 6e-5
        assert float(out.hi.view(torch.int16)[:, :, 0, :].abs().max()) == 0
        assert float(out.hi.view(torch.int16)[:, H, :, :].abs().max()) == 0
        assert float(out.lo.view(torch.int16)[:, :, 0, :].abs().max()) == 0
        pooled = k.maxpool2x2(out)
        pr = torch.nn.functional.max_pool2d(k.unpad_nhwc(out).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        assert torch.equal(k.unpad_nhwc(pooled), pr.contiguous())
        w2 = torch.randn(3, 3, Cout, 64, device="cuda", generator=g) * (2.0 / (9 * Cout)) ** 0.5
        pw2 = k.pack_weights(w2, None, fmt=k.FMT_F16E5)
        o2, d2 = k.conv(pooled, pw2, relu=False, out_pad=True, out_f32_dense=True, out_fmt=k.FMT_BF16X2)
        r2 = torch.nn.functional.conv2d(pr.double().permute(0, 3, 1, 2), w2.double().permute(3, 2, 0, 1), None,
                                        padding=1).permute(0, 2, 3, 1)
        assert _relerr(d2, r2) < 1.5e-4
        assert _relerr(k.unpad_nhwc(o2), d2) < 1e-5   # bf16 hi/lo rendering of the output for bf16x3 consumers
    finally:
        k.set_pair_mode(prev)


@pytest.mark.parametrize("B,H,W,Cin,Cout,fmt", [(1, 40, 250, 64, 64, 1), (1, 41, 251, 64, 64, 1), (2, 25, 130, 64, 64, 0),
                                                 (1, 37, 621, 128, 128, 1), (1, 16, 128, 64, 128, 0), (1, 2, 2, 64, 64, 1),
                                                 (1, 187, 1242, 64, 64, 1)])
def test_conv3x3_fused_maxpool(B, H, W, Cin, Cout, fmt):
    """mv3d_gemm_desc.pool: conv3x3 + bias + ReLU + 2x2/2 VALID max-pool in one kernel (CTA pair = two image rows, row maxima
    exchanged through distributed shared memory) against the same conv followed by the stand-alone pool kernel, and against
    torch in float64.  Odd H / W drop the last row / column (VALID); the pooled PAD halos must be zero."""
    from mv3d_tf_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H * W + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(3, 3, Cin, Cout, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    a = k.pad_nhwc(x, precise=True, fmt=fmt)
    pw = k.pack_weights(w, b, fmt=fmt)
    saved = k.POOL_MAX_WASTE
    k.POOL_MAX_WASTE = 1e9       # exercise ragged widths too
    try:
        assert k.pool_fusable(a, pw)
        fused, _ = k.conv(a, pw, relu=True, precise=True, out_fmt=fmt, pool=True)
    finally:
        k.POOL_MAX_WASTE = saved
    plain, _ = k.conv(a, pw, relu=True, precise=True, out_fmt=fmt)
    want = k.maxpool2x2(plain)
    torch.cuda.synchronize()
    assert (fused.H, fused.W, fused.fmt) == (H // 2, W // 2, fmt)
    got_d, want_d = k.unpad_nhwc(fused), k.unpad_nhwc(want)
    # same accumulators, same rendering; the two orders (render then max / max then render) agree except where two window
    # elements render to neighbouring values -- bounded by one rendering step
    assert float((got_d - want_d).abs().max()) <= 2.0 ** -12 * float(want_d.abs().max())
    assert float((got_d != want_d).float().mean()) < 1e-3
    xq = k.unpad_nhwc(a).double()
    ref = torch.relu(torch.nn.functional.conv2d(xq.permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(), padding=1))
    ref = torch.nn.functional.max_pool2d(ref, 2, 2).permute(0, 2, 3, 1)
    assert _relerr(got_d, ref) < (3e-4 if fmt else 5e-5)
    hi = fused.hi.view(torch.int16)
    assert int(hi[:, :, 0, :].abs().max()) == 0 and int(hi[:, H // 2, :, :].abs().max()) == 0
    lo = fused.lo.view(torch.int16)
    assert int(lo[:, :, 0, :].abs().max()) == 0 and int(lo[:, H // 2, :, :].abs().max()) == 0

