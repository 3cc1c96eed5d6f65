"""A/B timing of ttasr_op_attention across several builds of the library in ONE process (interleaved rounds, so
clock / power drift hits every variant alike).   python tools/attn_ab.py name=path.so [name=path.so ...] [B]"""
import ctypes as C
import statistics
import sys

import torch


def main():
    libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
    B = next((int(a) for a in sys.argv[1:] if a.isdigit()), 32)
    H, T = 20, 1500
    d = 64 * H
    dev = torch.device("cuda", 0)
    qkv = torch.randn((B, T, 3 * d), device=dev)
    qkv[..., :d] *= 0.125
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty((B, T, d), device=dev, dtype=torch.bfloat16)
    q, k, v = (t.view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=-1))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0).transpose(1, 2).reshape(B, T, d)
    fns = {}
    for name, path in libs:
        lib = C.CDLL(path)
        lib.ttasr_op_attention.restype = C.c_int
        lib.ttasr_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        fns[name] = lib
    st = int(torch.cuda.current_stream().cuda_stream)
    times = {n: [] for n in fns}
    for rnd in range(7):
        for name, lib in fns.items():
            for _ in range(2):
                assert lib.ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, st) == 0
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                lib.ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, st)
            b.record()
            torch.cuda.synchronize()
            times[name].append(a.elapsed_time(b) / 10)
            if rnd == 0:
                err = (out.float() - ref.float()).abs().max().item()
                print(f"{name}: max abs err vs SDPA {err:.4f}", flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0)
    b.record()
    torch.cuda.synchronize()
    fl = 4.0 * B * H * T * T * 64
    print(f"sdpa: {a.elapsed_time(b) / 10:.4f} ms")
    for name, ts in times.items():
        med = statistics.median(ts)
        print(f"{name}: median {med:.4f} ms  min {min(ts):.4f}  ({fl / med / 1e9:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()
