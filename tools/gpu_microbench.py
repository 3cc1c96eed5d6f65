"""Kernel-level timings on one B200 (CUDA events, warm-up, L2-sized rotation) -> gpurun_out/microbench.json.
Times the C-ABI single ops on the encoder's shapes next to cuBLAS / SDPA controls run in the same process."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402

from ttasr import _lib as L  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    dev = torch.device("cuda", 0)
    out = {"device": torch.cuda.get_device_name(0)}
    lib = L.lib()
    st = lambda: int(torch.cuda.current_stream().cuda_stream)
    B = int(os.environ.get("MB_BATCH", "32"))
    M = 1500 * B
    res = []
    for name, N, K, act, f32, add in [("qkv", 3840, 1280, 0, 0, 0), ("out_proj", 1280, 1280, 0, 1, 1),
                                      ("fc1", 5120, 1280, 1, 0, 0), ("fc2", 1280, 5120, 0, 1, 1),
                                      ("qkv_small", 2304, 768, 0, 0, 0), ("fc1_small", 3072, 768, 1, 0, 0)]:
        a = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        o = torch.zeros((M, N), device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        flops = 2.0 * M * N * K
        row = {"name": name, "M": M, "N": N, "K": K}
        for cg in (1, 2):
            try:
                def run():
                    L.check(lib.ttasr_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                              o.data_ptr() if add else None, o.data_ptr(), M, N, K, act, f32, cg, st()))
                med, best = timeit(run)
                row[f"cg{cg}_ms"] = med
                row[f"cg{cg}_tflops"] = flops / med / 1e9
            except Exception as e:  # keep going: a failing variant must not hide the others
                row[f"cg{cg}_err"] = str(e)[:200]
                break
        med, best = timeit(lambda: torch.matmul(a, w.t()))
        row["cublas_ms"] = med
        row["cublas_tflops"] = flops / med / 1e9
        res.append(row)
        print(row, flush=True)
    out["gemm"] = res
    # attention
    try:
        H, T = 20, 1500
        d = 64 * H
        qkv = torch.randn((B, T, 3 * d), device=dev)
        qkv[..., :d] *= 0.125  # q carries the head_dim^-0.5 scale, as in the encoder
        qkv = qkv.to(torch.bfloat16)
        o = torch.empty((B, T, d), device=dev, dtype=torch.bfloat16)
        med, best = timeit(lambda: L.check(lib.ttasr_op_attention(qkv.data_ptr(), o.data_ptr(), B, T, H, st())))
        fl = 4.0 * B * H * T * T * 64
        q, k, v = (t.view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=-1))
        med2, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0))
        out["attention"] = {"B": B, "ms": med, "tflops": fl / med / 1e9, "sdpa_ms": med2, "sdpa_tflops": fl / med2 / 1e9}
        print(out["attention"], flush=True)
    except Exception as e:
        out["attention"] = {"err": str(e)[:300]}
        print(out["attention"], flush=True)
    # layernorm
    try:
        x = torch.randn((M, 1280), device=dev)
        g = torch.ones(1280, device=dev)
        y = torch.empty((M, 1280), device=dev, dtype=torch.bfloat16)
        med, best = timeit(lambda: L.check(lib.ttasr_op_layernorm(x.data_ptr(), g.data_ptr(), g.data_ptr(), y.data_ptr(), M, 1280, 0, st())))
        out["layernorm"] = {"rows": M, "ms": med, "GBps": M * 1280 * 6 / med / 1e6}
        print(out["layernorm"], flush=True)
    except Exception as e:
        out["layernorm"] = {"err": str(e)[:300]}
    # front end
    try:
        from ttasr import B200WhisperFeatureExtractor
        for n_mels in (80, 128):
            fe = B200WhisperFeatureExtractor(feature_size=n_mels)
            FB = 128
            pcm = (0.1 * torch.randn((FB, 480000), device=dev)).clamp_(-1, 1)
            med, best = timeit(lambda: fe.extract(pcm), iters=10)
            by = FB * (480000 * 4 + n_mels * 3000 * 4)
            out[f"frontend{n_mels}"] = {"B": FB, "ms": med, "best_ms": best, "GBps_algorithmic": by / med / 1e6,
                                        "chunks_per_s": FB / med * 1e3}
            print(out[f"frontend{n_mels}"], flush=True)
    except Exception as e:
        out["frontend"] = {"err": str(e)[:300]}
        print(out["frontend"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
