"""Per-kernel parity through the C ABI single-op entry points (ttasr_op_*), against fp32 torch controls on the GPU."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from ttasr import _lib

    return _lib


def _stream(torch, dev):
    return int(torch.cuda.current_stream(dev).cuda_stream)


@pytest.mark.parametrize("rows,d", [(7, 128), (1500, 384), (3001, 1280)])
@pytest.mark.parametrize("out_f32", [0, 1])
def test_layernorm(cuda_device, rows, d, out_f32):
    import torch

    L = _lib()
    g = torch.Generator(device=cuda_device).manual_seed(rows + d)
    x = torch.randn((rows, d), generator=g, device=cuda_device) * 3 + 1.5
    w = torch.randn(d, generator=g, device=cuda_device)
    b = torch.randn(d, generator=g, device=cuda_device)
    y = torch.empty((rows, d), dtype=torch.float32 if out_f32 else torch.bfloat16, device=cuda_device)
    L.check(L.lib().ttasr_op_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, d, out_f32,
                                        _stream(torch, cuda_device)))
    ref = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
    tol = 2e-5 if out_f32 else 4e-2
    assert (y.float() - ref).abs().max().item() <= tol
    if not out_f32:
        assert torch.equal(y, ref.to(torch.bfloat16)) or (y.float() - ref).abs().mean().item() < 2e-3


GEMM_SHAPES = [(128, 128, 64), (256, 256, 128), (300, 384, 384), (1500, 1152, 384), (1500, 1536, 384),
               (3000, 1280, 1280), (777, 256, 5120)]


@pytest.mark.parametrize("mode", ["bf16", "bf16_gelu", "f32", "f32_add", "f32_gelu_add"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES, ids=[f"m{m}n{n}k{k}" for m, n, k in GEMM_SHAPES])
@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
def test_gemm(cuda_device, cta_group, M, N, K, mode):
    import torch

    L = _lib()
    g = torch.Generator(device=cuda_device).manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn((M, K), generator=g, device=cuda_device) * 0.5).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device=cuda_device) * (K ** -0.5)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device=cuda_device)
    act = 1 if "gelu" in mode else 0
    out_f32 = mode.startswith("f32")
    add = torch.randn((M, N), generator=g, device=cuda_device) if "add" in mode else None
    ref = a.float() @ w.float().t() + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    if add is not None:
        ref = ref + add
    out = add.clone() if add is not None else torch.full((M, N), float("nan"), device=cuda_device,
                                                         dtype=torch.float32 if out_f32 else torch.bfloat16)
    L.check(L.lib().ttasr_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                  out.data_ptr() if add is not None else None, out.data_ptr(), M, N, K, act,
                                  1 if out_f32 else 0, cta_group, _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    tol = 2e-3 if out_f32 else 3e-2
    assert err <= tol, f"max abs err {err}"


ATTN_SHAPES = [(1, 128, 1), (1, 256, 2), (2, 300, 2), (1, 1500, 6), (3, 1500, 20)]


@pytest.mark.parametrize("B,T,H", ATTN_SHAPES, ids=[f"b{b}t{t}h{h}" for b, t, h in ATTN_SHAPES])
def test_attention(cuda_device, B, T, H):
    import torch

    L = _lib()
    d = 64 * H
    g = torch.Generator(device=cuda_device).manual_seed(B * 100 + T + H)
    qkv = torch.randn((B, T, 3 * d), generator=g, device=cuda_device)
    qkv[..., :d] *= 0.4   # q already carries the head_dim^-0.5 scale
    qkv[..., d:2 * d] *= 1.2
    qkv = qkv.to(torch.bfloat16)
    out = torch.full((B, T, d), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    L.check(L.lib().ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    q, k, v = (t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=-1))
    ref = (torch.softmax(q @ k.transpose(2, 3), dim=-1) @ v).transpose(1, 2).reshape(B, T, d)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, f"max abs err {err}"


def test_attention_large_logits_trigger_rescale(cuda_device):
    """Keys are ordered so the running max keeps growing: exercises the lazy O rescale path."""
    import torch

    L = _lib()
    B, T, H, d = 1, 1500, 1, 64
    g = torch.Generator(device=cuda_device).manual_seed(99)
    q = torch.randn((B, T, d), generator=g, device=cuda_device)
    k = torch.randn((B, T, d), generator=g, device=cuda_device)
    v = torch.randn((B, T, d), generator=g, device=cuda_device)
    k = k * torch.linspace(0.2, 6.0, T, device=cuda_device).view(1, T, 1)
    qkv = torch.cat([q, k, v], dim=-1).to(torch.bfloat16)
    out = torch.empty((B, T, d), dtype=torch.bfloat16, device=cuda_device)
    L.check(L.lib().ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, _stream(torch, cuda_device)))
    qf, kf, vf = (t.float() for t in qkv.split(d, dim=-1))
    ref = torch.softmax(qf @ kf.transpose(1, 2), dim=-1) @ vf
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2


def test_attention_base_moves_at_every_point_of_the_sweep(cuda_device):
    """Spike keys placed so that the exponent base has to move (a) in the first key tile after quarter 0 has already been
    exponentiated, (b) in a later tile by a key outside quarter 0, (c) in a later tile by a key inside quarter 0, each by
    far more than the lazy-rescale threshold; rows whose q points the other way see the spikes as very negative scores."""
    import torch

    L = _lib()
    B, T, H, d = 2, 700, 2, 64
    g = torch.Generator(device=cuda_device).manual_seed(7)
    q = torch.randn((B, T, H, d), generator=g, device=cuda_device) * 0.5
    q[..., 0] += 1.0
    k = torch.randn((B, T, H, d), generator=g, device=cuda_device)
    v = torch.randn((B, T, H, d), generator=g, device=cuda_device)
    for idx, amp in ((40, 60.0), (2 * 128 + 100, 120.0), (4 * 128 + 5, 180.0)):
        k[:, idx] = 0.0
        k[:, idx, :, 0] = amp
    qkv = torch.cat([t.reshape(B, T, H * d) for t in (q, k, v)], dim=-1).to(torch.bfloat16)
    out = torch.full((B, T, H * d), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    L.check(L.lib().ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(B, T, H, d).transpose(1, 2) for t in qkv.split(H * d, dim=-1))
    ref = (torch.softmax(qf @ kf.transpose(2, 3), dim=-1) @ vf).transpose(1, 2).reshape(B, T, H * d)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2


def test_bad_arguments_surface_as_errors(cuda_device):
    import torch

    L = _lib()
    a = torch.zeros((128, 64), dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(L.TtasrError):
        L.check(L.lib().ttasr_op_gemm(a.data_ptr(), a.data_ptr(), None, None, a.data_ptr(), 128, 100, 64, 0, 0, 0, None))
    with pytest.raises(L.TtasrError):
        L.check(L.lib().ttasr_op_gemm(None, a.data_ptr(), None, None, a.data_ptr(), 128, 128, 64, 0, 0, 0, None))


# ------------------------------------------------------------------ split residual stream + folded LayerNorm
def _split(x):
    import torch

    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


SPLIT_SHAPES = [(128, 128, 64), (300, 384, 384), (1500, 1280, 1280), (777, 256, 512)]


@pytest.mark.parametrize("with_lo", [True, False], ids=["hi+lo", "hi"])
@pytest.mark.parametrize("M,N,K", SPLIT_SHAPES, ids=[f"m{m}n{n}k{k}" for m, n, k in SPLIT_SHAPES])
@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
def test_gemm_split_residual(cuda_device, cta_group, M, N, K, with_lo):
    """x' = a w^T + b + (hi + lo) -> (hi', lo') in place, plus the per-64-column (mean, M2) partials of hi'.  The
    stream carries a +100 offset (spread 1) so a naive E[x^2] - mean^2 would lose every digit of the variance."""
    import torch

    L = _lib()
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    a = (torch.randn((M, K), generator=g, device=cuda_device) * 0.5).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device=cuda_device) * (K ** -0.5)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device=cuda_device)
    x = torch.randn((M, N), generator=g, device=cuda_device) + 100.0
    hi, lo = _split(x)
    x_in = hi.float() + (lo.float() if with_lo else 0.0)
    ref = a.float() @ w.float().t() + bias + x_in
    stats = torch.full((M, N // 64, 2), float("nan"), device=cuda_device)
    L.check(L.lib().ttasr_op_gemm_split(a.data_ptr(), w.data_ptr(), bias.data_ptr(), hi.data_ptr(),
                                        lo.data_ptr() if with_lo else None, hi.data_ptr(),
                                        lo.data_ptr() if with_lo else None, stats.data_ptr(), M, N, K, 0, cta_group,
                                        _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    ref_hi = ref.to(torch.bfloat16)
    # hi' is the bf16 rounding of the fp32 sum (ties may fall either way on accumulation-order noise)
    assert (hi.float() - ref).abs().max().item() <= 0.51 * 0.5 + 2e-3     # bf16 spacing at ~100 is 0.5
    assert (hi != ref_hi).float().mean().item() < 0.02
    if with_lo:
        assert ((hi.float() + lo.float()) - ref).abs().max().item() <= 6e-3   # 16 mantissa bits at ~100
    parts = hi.float().view(M, N // 64, 64)
    mean_ref = parts.mean(dim=2)
    m2_ref = ((parts - mean_ref[..., None]) ** 2).sum(dim=2)
    assert (stats[..., 0] - mean_ref).abs().max().item() <= 1e-3
    assert ((stats[..., 1] - m2_ref).abs() / (m2_ref + 1.0)).max().item() <= 2e-3


@pytest.mark.parametrize("act", [0, 1], ids=["id", "gelu"])
@pytest.mark.parametrize("M,N,K", [(300, 384, 384), (1500, 3840, 1280), (200, 512, 128)],
                         ids=["m300n384k384", "m1500n3840k1280", "m200n512k128"])
def test_gemm_with_folded_layernorm(cuda_device, M, N, K, act):
    """LayerNorm(x) W^T + b computed on the un-normalised bf16 rows: gamma folded into W, mean / rstd applied in the
    epilogue from the (mean, M2) partials the split residual GEMM writes.  Control: fp32 torch LayerNorm + matmul."""
    import torch

    L = _lib()
    g = torch.Generator(device=cuda_device).manual_seed(M * 3 + N + K)
    x = (torch.randn((M, K), generator=g, device=cuda_device) * 2.0 + 3.0).to(torch.bfloat16)
    x[:, 5] = 40.0
    W = (torch.randn((N, K), generator=g, device=cuda_device) * (K ** -0.5)).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device=cuda_device)
    gamma = 0.5 + 1.5 * torch.rand(K, generator=g, device=cuda_device)
    beta = 0.3 * torch.randn(K, generator=g, device=cuda_device)
    wf = (W.float() * gamma).to(torch.bfloat16)
    c1 = wf.float().sum(dim=1).contiguous()
    c2 = (b + W.float() @ beta).contiguous()
    parts = K // 64
    xp = x.float().view(M, parts, 64)
    mean_p = xp.mean(dim=2)
    stats = torch.stack([mean_p, ((xp - mean_p[..., None]) ** 2).sum(dim=2)], dim=2).contiguous()
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_device)
    L.check(L.lib().ttasr_op_gemm_lnfold(x.data_ptr(), wf.data_ptr(), c1.data_ptr(), c2.data_ptr(), stats.data_ptr(),
                                         parts, out.data_ptr(), M, N, K, act, 1e-5, 0, _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    ln = torch.nn.functional.layer_norm(x.float(), (K,), None, None, 1e-5)
    ref = ln @ wf.float().t() + c2   # the same rounded W' = bf16(W * gamma) on both sides
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-2, f"max abs err {err}"
    assert (out.float() - ref).abs().mean().item() <= 4e-3


@pytest.mark.parametrize("B,T,d,n_mels", [(1, 128, 128, 80), (3, 1500, 384, 80), (2, 1500, 1280, 128)],
                         ids=["b1t128d128", "b3t1500d384", "b2t1500d1280"])
def test_conv_stem_against_conv1d(cuda_device, B, T, d, n_mels):
    """a6 on its own: GELU(conv1) -> GELU(conv2, stride 2) + positions as implicit GEMMs over shifted / parity-split
    TMA views, against F.conv1d in fp32 — with special attention to the rows at the chunk edges (t = 0 and t = T - 1),
    where the zero padding of one chunk must not pick up its batch neighbour's frames."""
    import torch
    import torch.nn.functional as F

    L = _lib()
    g = torch.Generator(device=cuda_device).manual_seed(B + T + d)
    ld = (n_mels + 7) // 8 * 8
    feats = torch.randn((B, n_mels, 2 * T), generator=g, device=cuda_device)
    feats[:, :, :2] += 3.0       # make the edge frames loud: a leak across the chunk boundary would show
    feats[:, :, -2:] -= 3.0
    tm = torch.zeros((B, 2 * T, ld), dtype=torch.bfloat16, device=cuda_device)
    tm[:, :, :n_mels] = feats.transpose(1, 2).to(torch.bfloat16)
    w1 = (torch.randn((d, n_mels, 3), generator=g, device=cuda_device) * (3 * n_mels) ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn((d, d, 3), generator=g, device=cuda_device) * (3 * d) ** -0.5).to(torch.bfloat16)
    b1 = 0.1 * torch.randn(d, generator=g, device=cuda_device)
    b2 = 0.1 * torch.randn(d, generator=g, device=cuda_device)
    pos = torch.randn((T, d), generator=g, device=cuda_device)
    scratch = torch.empty((B, 2 * T, d), dtype=torch.bfloat16, device=cuda_device)
    out = torch.full((B, T, d), float("nan"), device=cuda_device)
    L.check(L.lib().ttasr_op_conv_stem(tm.data_ptr(), ld, n_mels, B, T, d, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                       b2.data_ptr(), pos.data_ptr(), scratch.data_ptr(), out.data_ptr(),
                                       _stream(torch, cuda_device)))
    torch.cuda.synchronize()
    x = tm[:, :, :n_mels].float().transpose(1, 2)
    h1 = F.gelu(F.conv1d(x, w1.float(), b1, padding=1))
    assert (scratch.float() - h1.transpose(1, 2)).abs().max().item() <= 3e-2
    h1r = h1.to(torch.bfloat16).float()                      # conv2 reads conv1's bf16 output
    ref = F.gelu(F.conv1d(h1r, w2.float(), b2, stride=2, padding=1)).transpose(1, 2) + pos
    err = (out - ref).abs()
    assert err.max().item() <= 2e-2, f"max abs err {err.max().item()}"
    edge = torch.cat([err[:, :2], err[:, -2:]], dim=1)
    assert edge.max().item() <= 2e-2, "chunk-boundary rows differ"
