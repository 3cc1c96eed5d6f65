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


def test_bad_arguments_surface_as_errors(cuda_device):
    import torch

    L = _lib()
    a = torch.zeros((128, 64), dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(L.TtasrError):
        L.check(L.lib().ttasr_op_gemm(a.data_ptr(), a.data_ptr(), None, None, a.data_ptr(), 128, 100, 64, 0, 0, 0, None))
    with pytest.raises(L.TtasrError):
        L.check(L.lib().ttasr_op_gemm(None, a.data_ptr(), None, None, a.data_ptr(), 128, 128, 64, 0, 0, 0, None))
