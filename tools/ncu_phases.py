"""Per-phase view of an `ncu --set full --import-source on` capture of a barrier-phased kernel: the SASS listing is cut
at every block barrier / mbarrier instruction and, per segment, the share of executed warp-instructions, of the
warp-state samples, the shared-memory wavefronts and the top stall reasons are printed (markdown).
    python tools/ncu_phases.py capture.ncu-rep [label ...]      labels name, in order, the segments of >= 40 SASS instructions"""
import csv
import io
import subprocess
import sys


def main(path, labels):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, R = rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    new = lambda: {"n": 0, "samples": 0, "exec": 0, "ldw": 0, "stw": 0, "st": {}}  # noqa: E731
    segs, cur, tot_e, tot_s = [], new(), 0, 0
    for r in R:
        src = r[ix["Source"]]
        s, e = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
        w = int(r[ix["L1 Wavefronts Shared"]] or 0)
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        cur["n"] += 1; cur["samples"] += s; cur["exec"] += e
        for k in stalls:
            v = int(r[ix[k]] or 0)
            if v:
                cur["st"][k[6:]] = cur["st"].get(k[6:], 0) + v
        if op == "LDS":
            cur["ldw"] += w
        if op == "STS":
            cur["stw"] += w
        tot_e += e; tot_s += s
        if op == "BAR" or "SYNCS" in src.split()[0 if not src.startswith("@") else 1]:
            segs.append(cur); cur = new()
    segs.append(cur)
    print(f"`{path.split('/')[-1]}`: {tot_e} warp-instructions executed, {tot_s} warp-state samples\n")
    print("| segment (ends at a barrier) | SASS instrs | executed | samples | LDS wavefronts | STS wavefronts | top stalls (samples) |")
    print("|---|---:|---:|---:|---:|---:|---|")
    li = 0
    for s in segs:
        if s["n"] < 40:   # (a static property of the binary, so the labels stay aligned from capture to capture)
            continue
        name = labels[li] if li < len(labels) else f"#{li}"
        li += 1
        top = ", ".join(f"{k} {v}" for k, v in sorted(s["st"].items(), key=lambda kv: -kv[1])[:4])
        print(f"| {name} | {s['n']} | {s['exec'] / tot_e * 100:.1f} % | {s['samples'] / tot_s * 100:.1f} % | "
              f"{s['ldw'] / 1e6:.1f} M | {s['stw'] / 1e6:.1f} M | {top} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
