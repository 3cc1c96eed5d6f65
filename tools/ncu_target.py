"""Small, fixed workloads for `ncu --set full` captures of one kernel family at a time (large-v3 shapes, B chunks).
    python tools/ncu_target.py {gemm_fc1|gemm_fc2|gemm_qkv|gemm_out|attention|frontend|layernorm|...} [B]
The default encoder path (split residual stream) runs the *_ln (LayerNorm folded into the consumer) and *_split (residual
GEMM on the (hi, lo) pair, emitting the LayerNorm partials) flavours: gemm_qkv_ln, gemm_fc1_ln, gemm_out_split,
gemm_fc2_split, layernorm_split (the final LayerNorm)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402

from ttasr import _lib as L  # noqa: E402


def main():
    what = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda", 0)
    lib = L.lib()
    st = int(torch.cuda.current_stream().cuda_stream)
    M = 1500 * B
    reps = 4
    if what.startswith("gemm_") and not what.endswith(("_ln", "_split")):
        N, K, act, f32, add = {"gemm_fc1": (5120, 1280, 1, 0, 0), "gemm_fc2": (1280, 5120, 0, 1, 1),
                               "gemm_qkv": (3840, 1280, 0, 0, 0), "gemm_out": (1280, 1280, 0, 1, 1)}[what]
        a = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        o = torch.zeros((M, N), device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        for _ in range(reps):
            L.check(lib.ttasr_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), o.data_ptr() if add else None,
                                      o.data_ptr(), M, N, K, act, f32, 0, st))
    elif what in ("gemm_qkv_ln", "gemm_fc1_ln"):
        N, K, act = (3840, 1280, 0) if what == "gemm_qkv_ln" else (5120, 1280, 1)
        a = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
        c1, c2 = torch.randn(N, device=dev), torch.randn(N, device=dev)
        parts = K // 64
        stats = torch.stack([torch.zeros((M, parts), device=dev), torch.full((M, parts), 64.0, device=dev)], dim=2).contiguous()
        o = torch.zeros((M, N), device=dev, dtype=torch.bfloat16)
        for _ in range(reps):
            L.check(lib.ttasr_op_gemm_lnfold(a.data_ptr(), w.data_ptr(), c1.data_ptr(), c2.data_ptr(), stats.data_ptr(), parts,
                                             o.data_ptr(), M, N, K, act, 1e-5, 0, st))
    elif what in ("gemm_out_split", "gemm_fc2_split"):
        N, K = (1280, 1280) if what == "gemm_out_split" else (1280, 5120)
        a = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        xh = torch.randn((M, N), device=dev).to(torch.bfloat16)
        xl = (torch.randn((M, N), device=dev) * 1e-3).to(torch.bfloat16)
        stats = torch.empty((M, N // 64, 2), device=dev)
        for _ in range(reps):
            L.check(lib.ttasr_op_gemm_split(a.data_ptr(), w.data_ptr(), bias.data_ptr(), xh.data_ptr(), xl.data_ptr(),
                                            xh.data_ptr(), xl.data_ptr(), stats.data_ptr(), M, N, K, 0, 0, st))
    elif what == "layernorm_split":
        from ttasr import B200WhisperEncoder  # noqa: F401  (the split LayerNorm has no single-op entry: run a micro encoder)
        raise SystemExit("layernorm_split: capture it from the bench launch list (one launch per forward)")
    elif what == "attention":
        H, T = 20, 1500
        qkv = torch.randn((B, T, 3 * 64 * H), device=dev)
        qkv[..., : 64 * H] *= 0.125
        qkv = qkv.to(torch.bfloat16)
        o = torch.empty((B, T, 64 * H), device=dev, dtype=torch.bfloat16)
        for _ in range(reps):
            L.check(lib.ttasr_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, st))
    elif what == "frontend":
        from ttasr import B200WhisperFeatureExtractor
        fe = B200WhisperFeatureExtractor(feature_size=128)
        pcm = (0.1 * torch.randn((max(B, 64), 480000), device=dev)).clamp_(-1, 1)
        for _ in range(reps):
            fe.extract(pcm)          # the fp32-out variant the front-end roofline is quoted on
    elif what == "layernorm":
        x = torch.randn((M, 1280), device=dev)
        g = torch.ones(1280, device=dev)
        y = torch.empty((M, 1280), device=dev, dtype=torch.bfloat16)
        for _ in range(reps):
            L.check(lib.ttasr_op_layernorm(x.data_ptr(), g.data_ptr(), g.data_ptr(), y.data_ptr(), M, 1280, 0, st))
    else:
        raise SystemExit(f"unknown target {what}")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
