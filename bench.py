#!/usr/bin/env python
"""bench.py — audio-seconds/second of the hot path (16 kHz PCM -> log-mel -> Whisper encoder) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload large-v3|small|tiny] [--batch B] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic 30 s chunks per GPU (BASELINE.json configs[2] by
default: whisper-large-v3, 128-bin, 256 chunks; random-init weights of that architecture, synthetic audio).
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field):

  value        whole-job audio-s/s with the PCM already resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the public API with HOST buffers: per step a pinned-host -> device copy of the
               PCM and a device -> pinned-host read of the hidden states inside the timed region
  roofline     the dominant kernel's algorithmic FLOP/s (per-launch CUDA events recorded inside the timed steps by
               the library's stage profiler) against the measured peak in MEASURED_PEAKS.json
  cpu_baseline the reference's CPU implementation (HF numpy extractor + HF fp32 encoder; oracle port if transformers
               is missing) timed on this box's host cores on a bounded sample of the same workload

--impl reference times only that CPU implementation, with all host threads, as the driver's reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "audio-sec/sec (log-mel+large-v3 encoder)"   # BASELINE.json's metric (the default workload)


def metric_name(workload: str) -> str:
    """BASELINE.json's metric for the default workload; other workloads (--workload small = configs[1]) say so."""
    return METRIC if workload == "large-v3" else f"audio-sec/sec (log-mel+{workload} encoder)"
UNIT = "audio-s/s"
N_SAMPLES = 480000
CHUNK_SECONDS = 30.0
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


def arch_table(name):
    from ttasr import EncoderConfig

    return EncoderConfig.named(name)


# ------------------------------------------------------------------------------------------------ CPU reference
def make_cpu_weights(cfg, seed=0):
    """Random-init encoder weights (std 0.02, LN gain 1 / bias 0) under HF names, on the CPU."""
    import torch

    g = torch.Generator().manual_seed(seed)
    d, f = cfg.d_model, cfg.encoder_ffn_dim
    w = {}

    def n(*shape):
        return torch.randn(*shape, generator=g) * 0.02

    w["conv1.weight"], w["conv1.bias"] = n(d, cfg.num_mel_bins, 3), torch.zeros(d)
    w["conv2.weight"], w["conv2.bias"] = n(d, d, 3), torch.zeros(d)
    from oracle.encoder import sinusoids

    w["embed_positions.weight"] = sinusoids(cfg.max_source_positions, d)
    for i in range(cfg.encoder_layers):
        p = f"layers.{i}."
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            w[p + ln + ".weight"], w[p + ln + ".bias"] = torch.ones(d), torch.zeros(d)
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{proj}.weight"] = n(d, d)
            if proj != "k_proj":
                w[p + f"self_attn.{proj}.bias"] = torch.zeros(d)
        w[p + "fc1.weight"], w[p + "fc1.bias"] = n(f, d), torch.zeros(f)
        w[p + "fc2.weight"], w[p + "fc2.bias"] = n(d, f), torch.zeros(d)
    w["layer_norm.weight"], w["layer_norm.bias"] = torch.ones(d), torch.zeros(d)
    return w


class CpuReference:
    """The reference's CPU implementation of the path: HF numpy log-mel + HF fp32 WhisperEncoder on all host
    threads (the reference's CTranslate2 CPU encoder is not installable offline — BASELINE.md section 4); falls back to
    the oracle port when transformers cannot be imported."""

    def __init__(self, workload: str):
        import torch

        from oracle import encoder as OE

        self.cfg = arch_table(workload)
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.w = make_cpu_weights(self.cfg)
        self.arch = OE.Arch(workload, self.cfg.d_model, self.cfg.encoder_layers, self.cfg.encoder_attention_heads,
                            self.cfg.encoder_ffn_dim, self.cfg.num_mel_bins)
        self.kind = "port"
        self.hf_fe = self.hf_enc = None
        try:
            from transformers import WhisperFeatureExtractor

            from oracle.gen_golden import hf_encoder

            self.hf_fe = WhisperFeatureExtractor(feature_size=self.cfg.num_mel_bins)
            self.hf_enc = hf_encoder(self.arch, self.w)
            self.kind = "reference"
        except Exception:
            pass

    def step(self, pcm):
        """pcm: numpy [n, 480000] fp32 -> hidden states; returns seconds."""
        import numpy as np
        import torch

        from oracle import encoder as OE
        from oracle import frontend as OF

        t0 = time.perf_counter()
        if self.hf_fe is not None:
            feats = self.hf_fe._np_extract_fbank_features(pcm, "cpu")
            with torch.no_grad():
                out = self.hf_enc(torch.from_numpy(np.ascontiguousarray(feats))).last_hidden_state
        else:
            feats = OF.log_mel_batch(pcm, self.cfg.num_mel_bins)
            out = OE.encoder_forward(torch.from_numpy(feats), self.w, self.arch)
        assert out.shape[1:] == (self.cfg.max_source_positions, self.cfg.d_model)
        return time.perf_counter() - t0


def synth_pcm_cpu(n, seed=1234):
    import numpy as np

    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal((n, N_SAMPLES)), -1, 1).astype(np.float32)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(args.workload)
    # one warm-up pass pages in the weights / spins up the threads and tells how long a chunk takes here; the step is
    # then sized (2..8 chunks) so that the whole --steps run stays within ~2.5 minutes on this box's cores
    warm = synth_pcm_cpu(2)
    ref.step(warm[:1])
    t_chunk = ref.step(warm) / 2.0
    per_step = args.ref_chunks if args.ref_chunks > 0 else int(max(2, min(8, 150.0 / max(args.steps, 1) / max(t_chunk, 1e-3))))
    pcm = synth_pcm_cpu(per_step)
    times = [ref.step(pcm) for _ in range(args.steps)]
    total = sum(times)
    value = per_step * CHUNK_SECONDS * args.steps / total
    sample = (f"{per_step} x 30 s chunk(s) per step (a bounded sample of config.workload, NOT its "
              f"{args.batch} chunks per step), {args.steps} steps; HF numpy log-mel + "
              f"{'HF' if ref.kind == 'reference' else 'oracle-port'} fp32 encoder on {ref.cores} CPU threads "
              "(the reference's CTranslate2 int8 CPU encoder is not installable offline)")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_gpu_batch=args.batch),
        "reference_step": {"chunks_per_step": per_step, "same_units_as_config": True,
                           "note": "config names the workload both arms are quoted on; this arm times a bounded sample "
                                   "of it per step (audio-s/s is per chunk, so the ratio is unit-consistent)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def workload_config(args, per_gpu_batch):
    cfg = arch_table(args.workload)
    return {
        "workload": f"whisper-{args.workload} ({cfg.num_mel_bins}-bin) log-mel front end + {cfg.encoder_layers}-layer "
                    f"encoder, {per_gpu_batch} x 30 s synthetic chunks per GPU per step, random-init weights",
        "chunks_per_gpu_per_step": per_gpu_batch, "chunk_seconds": CHUNK_SECONDS,
        "parallelism": f"dp{args.gpus} (independent chunks, replicated weights, no collective on the data path)",
        "l2_policy": "inputs larger than L2 (PCM batch >= 0.49 GB, activations >= 1 GB per kernel)",
    }


def make_gpu_weights(cfg, device, seed=0):
    import torch

    from ttasr.encoder import EncoderConfig  # noqa: F401

    g = torch.Generator(device=device).manual_seed(seed)
    d, f = cfg.d_model, cfg.encoder_ffn_dim

    def n(*shape):
        return (torch.randn(*shape, generator=g, device=device) * 0.02).to(torch.bfloat16)

    def z(k):
        return torch.zeros(k, device=device)

    def one(k):
        return torch.ones(k, device=device)

    half = d // 2
    inc = torch.log(torch.tensor(10000.0)) / (half - 1)
    inv = torch.exp(-inc * torch.arange(half, device=device))
    t = torch.arange(cfg.max_source_positions, device=device).view(-1, 1) * inv.view(1, -1)
    w = {"conv1.weight": n(d, cfg.num_mel_bins, 3), "conv1.bias": z(d), "conv2.weight": n(d, d, 3), "conv2.bias": z(d),
         "embed_positions.weight": torch.cat([t.sin(), t.cos()], dim=1), "layer_norm.weight": one(d),
         "layer_norm.bias": z(d)}
    for i in range(cfg.encoder_layers):
        p = f"layers.{i}."
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            w[p + ln + ".weight"], w[p + ln + ".bias"] = one(d), z(d)
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{proj}.weight"] = n(d, d)
            if proj != "k_proj":
                w[p + f"self_attn.{proj}.bias"] = z(d)
        w[p + "fc1.weight"], w[p + "fc1.bias"] = n(f, d), z(f)
        w[p + "fc2.weight"], w[p + "fc2.bias"] = n(d, f), z(d)
    return w


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = os.path.join("/tmp", f"ttasr_clocks_{os.getpid()}.csv")
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(gpu_index)], stdout=self.fh, stderr=subprocess.DEVNULL)
            time.sleep(0.3)   # nvidia-smi needs a moment before its first sample
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def library_baseline(cfg, weights, pcm, dev, our_value, our_frontend):
    """Informational: the same work done by LIBRARY kernels on the same B200 in the same process (SURVEY.md 2d's honest
    baseline) — Hugging Face WhisperEncoder in bf16 with SDPA attention (cuBLAS GEMMs, cuDNN/flash attention, eager
    LayerNorm/GELU) and torch.stft + matmul for the log-mel front end (the HF torch extractor's arithmetic, a4).
    Not the reference arm and not a target; it only turns "hand-written beats the libraries by X" into a number."""
    import torch
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder

    from ttasr.mel import slaney_mel_filters

    hcfg = WhisperConfig(d_model=cfg.d_model, encoder_layers=cfg.encoder_layers,
                         encoder_attention_heads=cfg.encoder_attention_heads, encoder_ffn_dim=cfg.encoder_ffn_dim,
                         num_mel_bins=cfg.num_mel_bins, decoder_layers=1, decoder_attention_heads=cfg.encoder_attention_heads,
                         decoder_ffn_dim=64, max_source_positions=cfg.max_source_positions, attn_implementation="sdpa")
    with torch.device(dev):
        hf = WhisperEncoder(hcfg).eval().to(torch.bfloat16)
    hf.load_state_dict({k: v.to(torch.bfloat16) for k, v in weights.items()}, strict=True)
    Bl = min(64, pcm.shape[0])
    window = torch.hann_window(400, device=dev)
    fb = torch.from_numpy(slaney_mel_filters(cfg.num_mel_bins).astype("float32")).to(dev)   # [201, n_mels]

    def torch_logmel(x):
        st = torch.stft(x, 400, 160, window=window, return_complex=True)
        mel = fb.T @ (st[..., :-1].abs() ** 2)
        lg = torch.clamp(mel, min=1e-10).log10()
        lg = torch.maximum(lg, lg.amax(dim=(1, 2), keepdim=True) - 8.0)
        return (lg + 4.0) / 4.0

    def ev():
        return torch.cuda.Event(enable_timing=True)

    with torch.no_grad():
        for _ in range(2):
            f = torch_logmel(pcm[:Bl])
            hf(f.to(torch.bfloat16))
        torch.cuda.synchronize()
        a, b, c = ev(), ev(), ev()
        n = 3
        fe_ms = enc_ms = 0.0
        for _ in range(n):
            a.record()
            f = torch_logmel(pcm[:Bl])
            b.record()
            hf(f.to(torch.bfloat16))
            c.record()
            torch.cuda.synchronize()
            fe_ms += a.elapsed_time(b)
            enc_ms += b.elapsed_time(c)
    lib_value = Bl * CHUNK_SECONDS * n / ((fe_ms + enc_ms) / 1e3)
    fe_bytes = Bl * (N_SAMPLES * 4 + cfg.num_mel_bins * 3000 * 4)
    return {"what": "HF WhisperEncoder bf16 + SDPA (cuBLAS / cuDNN / eager elementwise) and torch.stft log-mel, same GPU, "
                    f"same weights, {Bl} chunks per step, {n} steps after 2 warm-ups",
            "value": lib_value, "unit": UNIT, "ours_over_library": our_value / lib_value,
            "encoder_ms_per_chunk": enc_ms / n / Bl, "frontend_ms_per_chunk": fe_ms / n / Bl,
            "frontend_gbs": fe_bytes / (fe_ms / n) / 1e6,
            "frontend_ours_over_library": our_frontend["achieved"] / (fe_bytes / (fe_ms / n) / 1e6)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import ttasr
    from ttasr.dp import all_max

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # control plane only (barrier + max-over-ranks of the timings): the data path has no collective, so a CPU
        # (gloo) group is all that is needed — and NCCL's banner would pollute the one-JSON-line stdout contract
        dist.init_process_group("gloo")
    n_gpus = world

    cfg = arch_table(args.workload)
    B = args.batch
    fe = ttasr.B200WhisperFeatureExtractor(feature_size=cfg.num_mel_bins)
    weights = make_gpu_weights(cfg, dev, seed=0)
    enc = ttasr.B200WhisperEncoder(cfg, weights, residual=args.residual)
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pcm = (0.1 * torch.randn((B, N_SAMPLES), generator=g, device=dev)).clamp_(-1, 1)
    host_pcm = torch.empty((B, N_SAMPLES), dtype=torch.float32).pin_memory()
    host_pcm.copy_(pcm)
    host_out = torch.empty((B, cfg.max_source_positions, cfg.d_model), dtype=torch.bfloat16).pin_memory()
    stage = torch.empty_like(pcm)

    fe_events = []

    def step_device(record_fe=False):
        if record_fe:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        # the production path (ttasr.B200LogMelEncoder.encode_device): only the bf16 time-major tensor the conv stem
        # reads is written; the fp32 `input_features` nobody would read are not
        _, tm = fe.extract(pcm, return_time_major=True, features=False)
        if record_fe:
            b.record()
            fe_events.append((a, b))
        return enc.encode(tm, time_major_ld=tm.shape[2])

    def step_host():
        stage.copy_(host_pcm, non_blocking=True)
        _, tm = fe.extract(stage, return_time_major=True, features=False)
        hidden = enc.encode(tm, time_major_ld=tm.shape[2])
        host_out.copy_(hidden, non_blocking=True)
        return hidden

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last_local_ms = [0.0]

    def timed(fn, steps, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn(**kw)
        e1.record()
        barrier()
        last_local_ms[0] = e0.elapsed_time(e1)
        return all_max(last_local_ms[0]), out

    # ---- cross-rank determinism probe: the same seeded chunks must give the same bits on every rank
    from ttasr.dp import gather_host, probe_digest

    probe_sha, probe_same = probe_digest(ttasr.B200LogMelEncoder(fe, enc))

    # ---- device-resident arm
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    enc.profile(True)
    enc.profile_read(reset=True)
    sampler = ClockSampler(local)
    ms_total, out = timed(step_device, args.steps, record_fe=True)
    clocks = sampler.stop()
    per_rank = gather_host({"rank": rank, "ms_per_step": last_local_ms[0] / args.steps, "sm_mhz": clocks.get("sm_mhz"),
                            "power_w_max": clocks.get("power_w_max"), "reasons": clocks.get("reasons")})
    stages = enc.profile_read(reset=True)
    enc.profile(False)
    fe_ms = [a.elapsed_time(b) for a, b in fe_events]
    # the same launch group timed on its own (no encoder around it: the SMs are not power-capped at ~1.4 GHz then)
    # front-end roofline, measured on the variant the reference's `fe(...)` call maps to (fp32 `input_features` out):
    # the launch group on its own (the SMs are not power-capped at ~1.4 GHz then) ...
    def time_fe(n, **kw):
        out_ms = []
        for i in range(n + 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fe.extract(pcm, **kw)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                out_ms.append(a.elapsed_time(b))
        return out_ms

    fe_alone = time_fe(10)
    fe_alone_tm = time_fe(10, return_time_major=True, features=False)
    # ... and inside a step (between two encoder forwards, at the step's clocks)
    fe_instep = []
    for i in range(4):
        step_device()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fe.extract(pcm)
        b.record()
        if i >= 1:
            fe_instep.append((a, b))
    torch.cuda.synchronize()
    fe_instep = [a.elapsed_time(b) for a, b in fe_instep]
    assert bool(torch.isfinite(out.float()).all()), "non-finite hidden states"
    value = n_gpus * B * CHUNK_SECONDS * args.steps / (ms_total / 1e3)

    if os.environ.get("TTASR_PROFILE_STEP"):  # one extra step between cudaProfilerStart/Stop for `ncu --profile-from-start off`
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---- end-to-end arm (host buffers in, host buffers out) through the public pipeline object: every step's PCM
    # is copied from pinned host memory and every step's hidden states are copied back, inside the timed region;
    # stream_host overlaps step k's kernels with the copy-in of step k+1 and the copy-out of step k-1.
    pipe = ttasr.B200LogMelEncoder(fe, enc)
    for _ in range(min(args.warmup, 3)):
        step_host()
    pipe.stream_host([host_pcm] * 2, [host_out] * 2)
    torch.cuda.synchronize()

    def run_stream():
        pipe.stream_host([host_pcm] * args.steps, [host_out] * args.steps)

    e2e_ms, _ = timed(run_stream, 1)
    e2e_value = n_gpus * B * CHUNK_SECONDS * args.steps / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (stage profiler events, per launch)
    peaks, peak_src = measured_peaks()
    d, f, T, Hh = cfg.d_model, cfg.encoder_ffn_dim, cfg.max_source_positions, cfg.encoder_attention_heads
    M = B * T
    flops_per_launch = {
        "conv1_gemm": 2.0 * B * 2 * T * 3 * cfg.num_mel_bins * d, "conv2_gemm": 2.0 * M * 3 * d * d,
        "qkv_gemm": 2.0 * M * d * 3 * d, "attention": 4.0 * B * Hh * T * T * 64, "out_proj_gemm": 2.0 * M * d * d,
        "fc1_gemm": 2.0 * M * d * f, "fc2_gemm": 2.0 * M * d * f,
    }
    kernels = {}
    for name, (ms, n) in stages.items():
        if n == 0:
            continue
        per = ms / n
        k = {"launches": n, "ms_per_launch": per, "share_of_step": ms / ms_total}
        if name in flops_per_launch:
            k["tflops"] = flops_per_launch[name] / per / 1e9
        elif name == "layernorm":
            k["gbs"] = M * d * 6.0 / per / 1e6
        kernels[name] = k
    dominant = max((n for n in kernels if n in flops_per_launch), key=lambda n: stages[n][0])
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{dominant}:{args.workload}:B{B}")
    roofline = {
        "kernel": dominant, "bound": "tensor", "achieved": kernels[dominant]["tflops"], "peak": tensor_peak,
        "unit": "TFLOP/s", "frac": kernels[dominant]["tflops"] / tensor_peak, "traffic": traffic,
        "peak_source": f"{peak_src} sustained bf16 matmul (kernel timed inside a long step)",
        "algorithmic_flops_per_launch": flops_per_launch[dominant],
    }
    # the encoder's GEMM kernels taken together (the north star's ">= 60 % of the bf16 tensor-pipe peak" is about them)
    gemm_names = [n for n in ("qkv_gemm", "out_proj_gemm", "fc1_gemm", "fc2_gemm", "conv1_gemm", "conv2_gemm") if n in kernels]
    gemm_flops = sum(flops_per_launch[n] * kernels[n]["launches"] for n in gemm_names)
    gemm_ms = sum(kernels[n]["ms_per_launch"] * kernels[n]["launches"] for n in gemm_names)
    gemm_tflops = gemm_flops / gemm_ms / 1e9 if gemm_ms > 0 else 0.0
    gemm_roofline = {"bound": "tensor", "kernels": gemm_names, "achieved": gemm_tflops, "unit": "TFLOP/s",
                     "peak": tensor_peak, "frac": gemm_tflops / tensor_peak,
                     "frac_of_burst_peak": gemm_tflops / float(peaks["bf16_tflops"]),
                     "share_of_step": sum(kernels[n]["share_of_step"] for n in gemm_names)}
    fe_bytes = B * (N_SAMPLES * 4 + cfg.num_mel_bins * 3000 * 4)
    fe_bytes_tm = B * (N_SAMPLES * 4 + cfg.num_mel_bins * 3000 * 2)
    hbm_peak = float(peaks["hbm_gbs"])

    def fe_entry(ms, nbytes, how):
        return {"ms_per_launch_group": ms, "achieved": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / hbm_peak, "how": how}

    fe_med = statistics.median(fe_instep)
    frontend = {"bound": "hbm", "achieved": fe_bytes / fe_med / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "frac": fe_bytes / fe_med / 1e6 / hbm_peak, "ms_per_launch_group": fe_med,
                "algorithmic_bytes_per_chunk": N_SAMPLES * 4 + cfg.num_mel_bins * 3000 * 4,
                "variant": "f32 PCM in -> fp32 input_features out (what the reference's feature extractor returns), timed "
                           "between two encoder forwards at the step's power-capped clocks",
                "alone": fe_entry(statistics.median(fe_alone), fe_bytes,
                                  "10 launch groups back to back without the encoder (SM clock not power-capped)"),
                "production_in_step": fe_entry(
                    statistics.median(fe_ms), fe_bytes_tm,
                    "the launch group the timed steps really run: bf16 time-major features only "
                    f"({N_SAMPLES * 4 + cfg.num_mel_bins * 3000 * 2} algorithmic bytes per chunk)"),
                "production_alone": fe_entry(statistics.median(fe_alone_tm), fe_bytes_tm, "the same, back to back"),
                "note": "memset + 2 kernels (frames, conditional clamp)"}
    enc_flops = cfg.flops_per_chunk() * B * n_gpus * args.steps
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(host_pcm.numel() * 4) * n_gpus,
                "d2h_bytes_per_step": int(host_out.numel() * 2) * n_gpus},
        "gpu_launches": args.steps * (2 + enc.launches_per_forward - 1),
        "clocks": clocks, "roofline": roofline, "gemm_roofline": gemm_roofline, "frontend_roofline": frontend,
        "encoder_tflops_whole_step": enc_flops / (ms_total / 1e3) / 1e12 / n_gpus,
        "kernels": kernels,
        "residual_stream": enc.residual or os.environ.get("TTASR_RESIDUAL") or "split (library default)",
        "rank_probe": {"sha256": probe_sha, "identical_on_all_ranks": bool(probe_same), "ranks": n_gpus,
                       "what": "3 seeded probe chunks encoded on every rank before the timed region (ttasr.dp.probe_digest)"},
        "per_rank": per_rank,
    }
    if n_gpus == 1 and not args.no_library_baseline:
        try:
            line["gpu_library_baseline"] = library_baseline(cfg, weights, pcm, dev, value, frontend)
        except Exception as exc:  # informational only: never fail the bench line over it
            line["gpu_library_baseline"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    if n_gpus == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.workload)
        n_ref = args.ref_chunks if args.ref_chunks > 0 else 2     # bounded sample: ~10-30 s of CPU work for large-v3
        sample_pcm = host_pcm[:n_ref].numpy()
        ref.step(sample_pcm[:1])
        secs = ref.step(sample_pcm)
        line["cpu_baseline"] = {
            "value": n_ref * CHUNK_SECONDS / secs, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
            "sample": f"first {n_ref} chunk(s) of the same batch, one pass after a 1-chunk warm-up; HF numpy "
                      f"log-mel + {'HF' if ref.kind == 'reference' else 'oracle-port'} fp32 encoder "
                      "(the reference's CTranslate2 CPU encoder is not installable offline)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large-v3", choices=["tiny", "base", "small", "medium", "large-v2", "large-v3"])
    ap.add_argument("--batch", type=int, default=256, help="30 s chunks per GPU per step")
    ap.add_argument("--ref-chunks", type=int, default=0,
                    help="chunks per CPU reference step (0 = sized from a warm-up pass: 2..8) / cpu_baseline sample")
    ap.add_argument("--residual", default=None, choices=["split", "f32", "bf16"],
                    help="residual-stream representation of the encoder (default: the library's, see ttasr_abi.h)")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the informational HF bf16 + SDPA / torch.stft comparison on the same GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
