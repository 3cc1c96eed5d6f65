"""fp32 restatement of the Whisper encoder forward (TEST INFRASTRUCTURE — see oracle/__init__.py).

A floating-point path, so the oracle is a plain torch-fp32 functional restatement (CPU), following the
Hugging Face module the reference runs through `Seq2SeqTrainer` / `generate`
(`train_asr.py:539-545,697-716`); HF = site-packages/transformers v5.5.0:

* sinusoidal positions      HF/models/whisper/modeling_whisper.py:55-64
* attention (q scaled by d_h^-0.5 after bias, k has no bias, softmax over all 1500 keys, no mask)
                            HF/models/whisper/modeling_whisper.py:215-238, :279-282, :310, :331-357
* pre-LN block              HF/models/whisper/modeling_whisper.py:392-408   (LayerNorm eps 1e-5 :372,378)
* stem + final LN           HF/models/whisper/modeling_whisper.py:613-626, :643
* exact (erf) GELU          HF/models/whisper/configuration_whisper.py:140

Weights are a flat dict keyed by the HF encoder state-dict names (`conv1.weight`,
`layers.0.self_attn.q_proj.weight`, ...), which is also what the shipped loader accepts.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class Arch:
    name: str
    d_model: int
    layers: int
    heads: int
    ffn: int
    n_mels: int
    n_ctx: int = 1500

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads

    def flops_per_chunk(self) -> int:
        """MACs x 2 of conv/GEMM/attention only."""
        d, f, L, T = self.d_model, self.ffn, self.layers, self.n_ctx
        return (2 * 2 * T * 3 * self.n_mels * d + 2 * T * 3 * d * d
                + L * (8 * T * d * d + 4 * T * T * d + 4 * T * d * f))


ARCHS = {
    "micro": Arch("micro", 128, 2, 2, 512, 80),  # test-only: smallest shape every kernel accepts
    "tiny": Arch("tiny", 384, 4, 6, 1536, 80),
    "base": Arch("base", 512, 6, 8, 2048, 80),
    "small": Arch("small", 768, 12, 12, 3072, 80),
    "medium": Arch("medium", 1024, 24, 16, 4096, 80),
    "large-v2": Arch("large-v2", 1280, 32, 20, 5120, 80),
    "large-v3": Arch("large-v3", 1280, 32, 20, 5120, 128),
}


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """modeling_whisper.py:55-64."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def init_weights(arch: Arch, seed: int = 0, std: float = 0.02, ln_jitter: float = 0.0) -> dict:
    """Deterministic random-init weights under HF names (normal(0, std) matrices, like HF `_init_weights`).

    `ln_jitter` > 0 perturbs LayerNorm gains/biases and linear biases so that a test exercises them
    (HF initialises them to 1 / 0, which would hide a dropped bias)."""
    g = torch.Generator().manual_seed(seed)

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    d, f = arch.d_model, arch.ffn
    w = {
        "conv1.weight": n(d, arch.n_mels, 3),
        "conv1.bias": n(d, s=ln_jitter or 0.0) if ln_jitter else torch.zeros(d),
        "conv2.weight": n(d, d, 3),
        "conv2.bias": n(d, s=ln_jitter) if ln_jitter else torch.zeros(d),
        "embed_positions.weight": sinusoids(arch.n_ctx, d),
    }

    def ln(prefix):
        w[prefix + ".weight"] = torch.ones(d) + (n(d, s=ln_jitter) if ln_jitter else 0.0)
        w[prefix + ".bias"] = n(d, s=ln_jitter) if ln_jitter else torch.zeros(d)

    def lin(prefix, out_f, in_f, bias=True):
        w[prefix + ".weight"] = n(out_f, in_f)
        if bias:
            w[prefix + ".bias"] = n(out_f, s=ln_jitter) if ln_jitter else torch.zeros(out_f)

    for i in range(arch.layers):
        p = f"layers.{i}."
        ln(p + "self_attn_layer_norm")
        lin(p + "self_attn.q_proj", d, d)
        lin(p + "self_attn.k_proj", d, d, bias=False)
        lin(p + "self_attn.v_proj", d, d)
        lin(p + "self_attn.out_proj", d, d)
        ln(p + "final_layer_norm")
        lin(p + "fc1", f, d)
        lin(p + "fc2", d, f)
    ln("layer_norm")
    return w


def round_weights_bf16(w: dict) -> dict:
    """bf16-round the tensors the GPU path stores in bf16 (GEMM/conv matrices), keep the rest fp32 —
    the parity protocol of SURVEY.md section 8d: same rounded weights on both sides."""
    out = {}
    for k, v in w.items():
        is_matrix = k.endswith(".weight") and v.dim() >= 2 and not k.startswith("embed_positions")
        out[k] = v.to(torch.bfloat16).to(torch.float32) if is_matrix else v.clone()
    return out


def attention(x: torch.Tensor, w: dict, p: str, heads: int) -> torch.Tensor:
    """WhisperAttention.forward self-attention branch, modeling_whisper.py:284-357."""
    B, T, d = x.shape
    dh = d // heads
    q = (F.linear(x, w[p + "q_proj.weight"], w[p + "q_proj.bias"]) * dh ** -0.5)
    k = F.linear(x, w[p + "k_proj.weight"])
    v = F.linear(x, w[p + "v_proj.weight"], w[p + "v_proj.bias"])
    q = q.view(B, T, heads, dh).transpose(1, 2)
    k = k.view(B, T, heads, dh).transpose(1, 2)
    v = v.view(B, T, heads, dh).transpose(1, 2)
    a = torch.softmax(torch.matmul(q, k.transpose(2, 3)), dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, d)
    return F.linear(o, w[p + "out_proj.weight"], w[p + "out_proj.bias"])


def encoder_layer(x: torch.Tensor, w: dict, i: int, heads: int) -> torch.Tensor:
    """WhisperEncoderLayer.forward, modeling_whisper.py:380-414."""
    p = f"layers.{i}."
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"], 1e-5)
    x = x + attention(h, w, p + "self_attn.", heads)
    h = F.layer_norm(x, (d,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], 1e-5)
    h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
    x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
    return x


def stem(feats: torch.Tensor, w: dict) -> torch.Tensor:
    """conv1+GELU, conv2(stride 2)+GELU, permute, + positions. modeling_whisper.py:613-626."""
    if feats.shape[-1] != 3000:
        raise ValueError(
            f"Whisper expects the mel input features to be of length 3000, but found {feats.shape[-1]}."
        )
    x = F.gelu(F.conv1d(feats, w["conv1.weight"], w["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["conv2.weight"], w["conv2.bias"], stride=2, padding=1))
    return x.permute(0, 2, 1) + w["embed_positions.weight"]


@torch.no_grad()
def encoder_forward(feats, w: dict, arch: Arch, n_layers: int | None = None, final_ln: bool = True):
    """[B, n_mels, 3000] fp32 -> last_hidden_state [B, 1500, d] fp32."""
    x = stem(torch.as_tensor(feats, dtype=torch.float32), w)
    for i in range(arch.layers if n_layers is None else n_layers):
        x = encoder_layer(x, w, i, arch.heads)
    if final_ln:
        x = F.layer_norm(x, (arch.d_model,), w["layer_norm.weight"], w["layer_norm.bias"], 1e-5)
    return x


def parity_stats(got, ref) -> dict:
    """max-abs, mean-abs and cosine as SURVEY.md section 8d defines the encoder tolerance."""
    got = torch.as_tensor(got, dtype=torch.float64).flatten()
    ref = torch.as_tensor(ref, dtype=torch.float64).flatten()
    diff = (got - ref).abs()
    cos = float(torch.dot(got, ref) / (got.norm() * ref.norm() + 1e-300))
    return {"max_abs": float(diff.max()), "mean_abs": float(diff.mean()), "cosine": cos}
