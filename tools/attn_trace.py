"""Timeline of one attention CTA -> per-phase cycle statistics.  Needs a library built with -DTTASR_ATTN_TRACE=1
(python taiwan-tongues-asr-ce_b200/build.py -DTTASR_ATTN_TRACE=1 --out=...):  python tools/attn_trace.py lib.so [B]"""
import ctypes as C
import collections
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "taiwan-tongues-asr-ce_b200"))
import torch  # noqa: E402


NAMES = {10: "loop_top", 11: "s_full_ok", 12: "ldtm_done", 13: "max_done", 14: "odone_ok", 15: "pre_done", 16: "token_ok",
         17: "sweep_done", 18: "p_ready_sent", 30: "wait_p0", 31: "wait_p1", 32: "got_p0", 33: "got_p1", 40: "wait_sf0",
         41: "wait_sf1", 42: "got_sf0", 43: "got_sf1", 44: "wait_kv", 45: "got_kv", 50: "wait_kvfree", 51: "got_kvfree",
         19: "q0_done", 60: "pv_begin", 61: "pv_mma1", 62: "pv_mma4", 63: "pv_mma8", 64: "pv_commit", 65: "pv_end", 20: "epi_wait", 21: "epi_go", 22: "epi_done"}


def main():
    lib = C.CDLL(sys.argv[1])
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    H, T, cap = 20, 1500, 16384
    d = 64 * H
    dev = torch.device("cuda", 0)
    qkv = torch.randn((B, T, 3 * d), device=dev)
    qkv[..., :d] *= 0.125
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty((B, T, d), device=dev, dtype=torch.bfloat16)
    trace = torch.zeros((4, cap), dtype=torch.int64, device=dev)
    st = C.c_void_p(int(torch.cuda.current_stream().cuda_stream))
    lib.ttasr_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
    lib.ttasr_debug_attention_trace.argtypes = [C.c_void_p, C.c_int]
    for _ in range(3):
        assert lib.ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, st) == 0
    torch.cuda.synchronize()
    assert lib.ttasr_debug_attention_trace(trace.data_ptr(), cap) == 0
    assert lib.ttasr_op_attention(qkv.data_ptr(), out.data_ptr(), B, T, H, st) == 0
    torch.cuda.synchronize()
    assert lib.ttasr_debug_attention_trace(None, 0) == 0
    tr = trace.cpu().numpy()
    for region, name in ((1, "softmax tile 0"), (2, "softmax tile 1"), (0, "MMA thread"), (3, "TMA producer")):
        ev = [(int(tr[region, i]), int(tr[region, i + 1])) for i in range(0, cap, 2) if tr[region, i] != 0]
        if not ev:
            continue
        print(f"== {name}: {len(ev)} events, span {ev[-1][1] - ev[0][1]} cycles")
        deltas = collections.defaultdict(list)
        for (ta, ca), (tb, cb) in zip(ev, ev[1:]):
            deltas[(ta, tb)].append(cb - ca)
        for (ta, tb), ds in sorted(deltas.items(), key=lambda kv: -sum(kv[1])):
            if len(ds) < 20:
                continue
            print(f"  {NAMES.get(ta, ta):>13s} -> {NAMES.get(tb, tb):<13s} n={len(ds):4d} median {statistics.median(ds):7.0f} "
                  f"mean {statistics.mean(ds):7.0f}  p90 {sorted(ds)[int(0.9 * len(ds))]:7.0f}  total {sum(ds):9d}")
        if region == 1:  # steady-state period of one tile-pass
            tops = [c for t, c in ev if t == 10]
            per = [b - a for a, b in zip(tops, tops[1:])]
            print(f"  tile-pass period: median {statistics.median(per):.0f} cycles (n={len(per)})")


if __name__ == "__main__":
    main()
