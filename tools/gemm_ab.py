"""A/B timing of ttasr_op_gemm (fc1 shape, GELU epilogue) across builds of the library in ONE process (interleaved
rounds).   python tools/gemm_ab.py name=path.so [name=path.so ...] [B]"""
import ctypes as C
import statistics
import sys

import torch


def main():
    libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    B = next((int(a) for a in sys.argv[1:] if a.isdigit()), 32)
    shape = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--shape=")), "fc1")
    N, K, act, f32 = {"fc1": (5120, 1280, 1, 0), "fc2": (1280, 5120, 0, 1), "qkv": (3840, 1280, 0, 0), "fc1_noact": (5120, 1280, 0, 0),
                      "out_proj": (1280, 1280, 0, 1)}[shape]
    M = 1500 * B
    dev = torch.device("cuda", 0)
    a = torch.randn((M, K), device=dev).to(torch.bfloat16)
    w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    o = torch.zeros((M, N), device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    ref = a[:4096].float() @ w.float().t() + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    fns = {}
    for name, path in libs:
        lib = C.CDLL(path)
        lib.ttasr_op_gemm.restype = C.c_int
        lib.ttasr_op_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        fns[name] = lib
    st = int(torch.cuda.current_stream().cuda_stream)
    # residual shapes add the output to itself (addend == out), as the encoder does; values drift, timing does not
    run = lambda lib: lib.ttasr_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), o.data_ptr() if f32 else None,
                                        o.data_ptr(), M, N, K, act, f32, 0, st)
    times = {n: [] for n in fns}
    for rnd in range(7):
        for name, lib in fns.items():
            for _ in range(2):
                assert run(lib) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(lib)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / 10)
            if rnd == 0 and not f32:
                err = (o[:4096].float() - ref).abs().max().item()
                print(f"{name}: max abs err vs torch fp32 reference {err:.4f}", flush=True)
    fl = 2.0 * M * N * K
    for name, ts in times.items():
        med = statistics.median(ts)
        print(f"{name}: median {med:.4f} ms  min {min(ts):.4f}  ({fl / med / 1e9:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()
