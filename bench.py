#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 spectrum engine (contract: see DESIGN.md "Measurement").

Metric (BASELINE.json): IQ MSamples/s ingested @ 1024 demod clients.
Workload at N=1 (BASELINE.json configs[1] at the metric's client count): 35 MSPS complex-IQ
synthetic, 2^20-point FFT, 1024 clients mixed AM/USB/LSB, free-running.

A step = one pass of the hot path over one ring of synthetic input: `ring` hops already resident in
HBM (ring x 4 MiB > L2, so no step re-reads its input from cache) -> `ring` 50 %-overlapped frames,
each: fused window+forward FFT -> int8 waterfall pyramid -> gather/IFFT/demod/DC/AGC/int16 for every
client. `value` = IQ samples all ranks processed / device time (CUDA events, max over ranks).
`e2e` = the same frames through the reference-facing C-ABI calls with HOST buffers in their block form
(b200_submit_block / b200_wait_block = load_complex_input -> execute -> signal_loop for 64 frames per call, up to
four blocks in flight), every host<->device copy inside the timed region. Other keys: `roofline` (forward group
timed alone; `traffic` from profiles/r2_traffic*.json, which names its build), `cpu_baseline` (oracle port + MKL on
the host cores, with the forward / clients split and the one-thread-FFT variant), `breakdown` (stage-isolated
timings and the bare cuFFT transform), `clocks`, `gpu_launches`.

Options: --config cfg1|cfg2|cfg3 (BASELINE configs), --waterfall-skip 0 (the reference's send cadence), --pcm16,
--e2e-raw s16|u8 (raw ADC halves, extra key `e2e_raw`), --mgpu-mode spectrum|scatter|scatter-dma|scatter-pull,
--impl cufft (library yardstick), --impl reference (CPU arm).

N > 1 (torchrun): rank 0 ingests and computes each spectrum batch, one exchange step over NVLink delivers it
to every rank, every rank demodulates its own 1024 clients of the whole stream (weak scaling: per-GPU client
load fixed). `value` = ingest rate x client shards (= ingest x clients_total / 1024): the stream rate the job
sustains per 1024-client shard, summed over the N shards; `ingest_msps` is the rate of the ONE ingested
stream. The same run also times the other exchange modes and a strong-scaling leg (1024 clients in total), runs
the exchange parity check of tests/mgpu_worker.py in its own process group, and reports all of it under "mgpu".
--impl reference uses the same definition on the same client total.

--impl reference : the reference's CPU path (oracle port; FFT by MKL through torch.fft as the
FFTW substitute) on the host cores, same config/metric, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM  # noqa: E402
from phantomsdr_b200.synth import make_clients  # noqa: E402

METRIC = "IQ MSamples/s ingested @ 1024 demod clients"
VALUE_DEF = ("ingest rate (frames/s x new samples per frame) x clients_total / clients_per_gpu: at N = 1 the ingest rate; at N > 1 every "
             "rank demodulates the whole ingested stream for its own shard of clients_per_gpu clients, so the job sustains that "
             "stream rate once per shard (weak scaling of the client load); ingest_msps is the single stream's rate")
UNIT = "MS/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cufft"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3"],
                    help="BASELINE.json configs: cfg1 = [0] rtlsdr 3.2 MSPS u8 IQ, 2^17, 1 USB + waterfall; cfg2 = [1] 35 MSPS IQ, 2^20 "
                         "(the metric's configuration, at 1024 clients); cfg3 = [2] 70 MSPS real, 2^21, 256 FM + waterfall")
    ap.add_argument("--waterfall-skip", type=int, default=1,
                    help="waterfall cadence (SURVEY 8f N3): pyramids only every skip-th frame; 0 = the reference's skip_num")
    ap.add_argument("--pcm16", action="store_true", help="int16 PCM hand-off (SURVEY 8f N2) in the e2e path")
    ap.add_argument("--no-mgpu-extras", action="store_true", help="N>1: only the default exchange mode, no strong-scaling leg")
    ap.add_argument("--fft-log2", type=int, default=20)
    ap.add_argument("--sps", type=int, default=35_000_000)
    ap.add_argument("--real", action="store_true", help="r2c input (cfg 3 shape)")
    ap.add_argument("--clients", type=int, default=1024, help="demod clients per GPU")
    ap.add_argument("--ring", type=int, default=64, help="hops resident in HBM = frames per step")
    ap.add_argument("--batch", type=int, default=64, help="frames per kernel launch group")
    ap.add_argument("--banks", type=int, default=2, help="pipeline depth: clients of batch k overlap the FFT of batch k+1")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-frames", type=int, default=768)
    ap.add_argument("--e2e-raw", default="", choices=["", "s16", "u8"],
                    help="also time the e2e path with raw ADC samples handed to b200_load_raw_input's block form (SURVEY 8f N1: "
                         "sample conversion fused into FFT pass 1, 2x / 4x fewer H2D bytes); reported as e2e_raw, never as e2e")
    ap.add_argument("--mgpu-mode", default="scatter-dma", choices=["spectrum", "scatter", "scatter-dma", "scatter-pull"],
                    help="N>1 exchange step. spectrum: NCCL broadcast of every spectrum batch (north_star's wording). scatter: "
                         "clients partitioned in (l, r) order and FFT pass 2 on the ingest rank stores each rank's sub-band "
                         "straight into that rank's memory over NVLink (peer stores fused into the kernel, flags instead of a "
                         "collective). scatter-dma (default, fastest measured: 194 vs 130 vs 129 GS/s at N=8): same partition "
                         "and flags, the sub-bands pushed by the copy engines so the ingest rank's SMs never wait on NVLink. "
                         "scatter-pull: every client rank's own copy engine fetches its sub-band from the ingest rank's bank")
    args = ap.parse_args()
    explicit = {a.split("=")[0] for a in sys.argv[1:] if a.startswith("--")}
    if args.config == "cfg1":
        if "--fft-log2" not in explicit: args.fft_log2 = 17
        if "--sps" not in explicit: args.sps = 3_200_000
        if "--clients" not in explicit: args.clients = 1
        args.modes = (USB,)
    elif args.config == "cfg3":
        if "--fft-log2" not in explicit: args.fft_log2 = 21
        if "--sps" not in explicit: args.sps = 70_000_000
        if "--clients" not in explicit: args.clients = 256
        args.real = True
        args.modes = (FM,)
    else:
        args.modes = (AM, USB, LSB)
    return args


def make_cfg(args) -> SpectrumConfig:
    return SpectrumConfig(sps=args.sps, fft_size=1 << args.fft_log2, is_real=args.real)


def client_table(cfg, count, rank=0, modes=(AM, USB, LSB)):
    return make_clients(cfg, count, seed=0x5EED + 1 + 1000 * rank, modes=modes)


def algorithmic_bytes_per_frame(cfg) -> int:
    """B_fwd = input halves + spectrum (+ wrap tail) + int8 pyramid."""
    N, R, n, L = cfg.fft_size, cfg.fft_result_size, cfg.audio_fft_size, cfg.downsample_levels
    pyr = sum(R >> i for i in range(L))
    if cfg.is_real:
        return 4 * N + 8 * (N // 2 + 1) + pyr
    return 8 * N + 8 * R + pyr


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvml), DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x10: "sync_boost"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            try:
                self._once()
            except Exception:
                pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# synthetic input (torch on the device: noise + tones, seeded) - plumbing, not the product path
# ------------------------------------------------------------------------------------------------
def fill_ring(torch, ring_t, cfg, seed):
    """ring_t: float32 [nhops, hop_floats] on the device. White noise sigma 1e-3 + 32 CW/AM/FM tones
    (SURVEY 8d), phase-continuous across hops."""
    g = torch.Generator(device=ring_t.device)
    g.manual_seed(seed)
    nhops, hop_floats = ring_t.shape
    n = cfg.hop_samples
    rs = np.random.Generator(np.random.PCG64(seed))
    scale = min(1.0, float(np.sqrt(2.0 ** 20 / cfg.fft_size)))
    tones = [(rs.uniform(0.01, 0.49) if cfg.is_real else rs.uniform(-0.49, 0.49),
              float(np.exp(rs.uniform(np.log(1e-4), np.log(1.5e-3)))) * scale, i % 4, rs.uniform(0, 2 * np.pi))
             for i in range(32)]
    for hidx in range(nhops):
        t = torch.arange(hidx * n, (hidx + 1) * n, device=ring_t.device, dtype=torch.float64)
        if cfg.is_real:
            x = torch.randn(n, generator=g, device=ring_t.device, dtype=torch.float64) * (1e-3 * scale)
        else:
            x = torch.complex(torch.randn(n, generator=g, device=ring_t.device, dtype=torch.float64),
                              torch.randn(n, generator=g, device=ring_t.device, dtype=torch.float64)) * (1e-3 * scale)
        for f, a, kind, ph0 in tones:
            ph = 2 * np.pi * f * t + ph0
            env = 1.0
            if kind == 1:
                env = 1.0 + 0.5 * torch.cos(2 * np.pi * 1000.0 / cfg.sps * t)
            elif kind == 3:
                ph = ph + (2500.0 / 400.0) * torch.sin(2 * np.pi * 400.0 / cfg.sps * t)
            s = a * env * (torch.cos(ph) if cfg.is_real else torch.polar(torch.ones_like(ph), ph))
            x = x + s
        if cfg.is_real:
            ring_t[hidx].copy_(x.to(torch.float32))
        else:
            ring_t[hidx].copy_(torch.view_as_real(x.to(torch.complex64)).reshape(-1))


# ------------------------------------------------------------------------------------------------
# reference / CPU baseline: the oracle port of the reference FFTW path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, nclients, frames, warm=1, modes=(AM, USB, LSB), fft_threads=0):
    """Times `frames` frames of the reference CPU path. FFT provider: MKL through torch.fft (the
    FFTW substitute) on `fft_threads` threads (0 = all host cores; 1 = the reference's default,
    src/spectrumserver.cpp:39); window/quantiser/pyramid/clients: oracle/ (OpenMP over all host cores, like
    fft_impl.cpp:32,53 and the asio pool). Returns (frames/s, info dict with the forward / clients split)."""
    import torch

    import oracle

    oracle.build()
    cores = os.cpu_count() or 1
    N = cfg.fft_size
    orc = oracle.OracleFFT(N, cfg.downsample_levels, cfg.brightness_offset)
    n = cfg.audio_fft_size
    orc.set_output_additional_size(n)
    orc.plan_r2c() if cfg.is_real else orc.plan_c2c()
    specs = client_table(cfg, nclients, modes=modes)
    clients = []
    for c in specs:
        o = oracle.OracleClient(cfg.is_real, n, cfg.audio_sps, cfg.fft_result_size)
        o.set_audio_range(c.l, c.mid, c.r)
        o.set_audio_demodulation(c.mode)
        clients.append(o)
    rng = np.random.default_rng(0x5EED)
    hops = [(rng.standard_normal(cfg.hop_floats) * 1e-3).astype(np.float32) for _ in range(4)]
    out = orc.outbuf
    inb = orc.inbuf
    nfl = N + 2 if cfg.is_real else 2 * N
    t_fwd = [0.0]

    def one(frame):
        t0 = time.perf_counter()
        a1, a2 = hops[frame % 4], hops[(frame + 1) % 4]
        if cfg.is_real:
            orc.load_real_input(a1, a2)
        else:
            orc.load_complex_input(a1, a2)
        # the thread count is process-wide (torch and the oracle share the OpenMP runtime): restrict it for the FFT only
        if fft_threads:
            torch.set_num_threads(fft_threads)
        X = torch.fft.rfft(torch.from_numpy(inb)) if cfg.is_real else torch.fft.fft(torch.from_numpy(inb).view(torch.complex64))
        if fft_threads:
            torch.set_num_threads(cores)
        out[:nfl] = torch.view_as_real(X).reshape(-1).numpy()
        orc.quantize()
        orc.wrap_copy(n)
        t_fwd[0] += time.perf_counter() - t0
        oracle.clients_send_audio(clients, out, N, cfg.is_real, frame)

    for f in range(warm):
        one(f)
    t_fwd[0] = 0.0
    t0 = time.perf_counter()
    for f in range(frames):
        one(warm + f)
    dt = time.perf_counter() - t0
    return frames / dt, {"cores": cores, "kind": "port",
                         "sample": f"{frames} frames of the same workload ({nclients} clients); FFT = MKL via torch.fft "
                                   f"(FFTW substitute) on {fft_threads or cores} thread(s), rest = oracle/ C port with OpenMP on "
                                   f"{cores} threads",
                         "split_ms_per_frame": {"forward_and_waterfall": 1e3 * t_fwd[0] / frames,
                                                "clients": 1e3 * (dt - t_fwd[0]) / frames}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_cfg(args)
    total_clients = args.clients * max(1, args.gpus)  # the same client total as the b200 arm serves at this N
    # bounded sample per step so that steps+warmup finish within minutes
    probe_fps, info = cpu_reference_run(cfg, total_clients, frames=3, warm=1, modes=args.modes)
    budget_s = 120.0
    per_step = max(1, int(min(args.ring, budget_s * probe_fps / max(1, args.steps + args.warmup))))
    t_all = []
    for s in range(args.warmup + args.steps):
        fps, info = cpu_reference_run(cfg, total_clients, frames=per_step, warm=0, modes=args.modes)
        if s >= args.warmup:
            t_all.append(per_step / fps)
    ms = 1e3 * float(np.mean(t_all))
    ingest = per_step * cfg.hop_samples / (ms * 1e-3) / 1e6
    value = ingest * max(1, args.gpus)  # VALUE_DEF: ingest x clients_total / clients_per_gpu
    info["sample"] = f"each step = {per_step} frames; " + info["sample"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, **info},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ingest_msps": ingest,
    }
    print(json.dumps(line), flush=True)


def cufft_yardstick(torch, cfg, frames, reps=10):
    """Bare cuFFT (through torch.fft) on the same box: `frames` batched transforms of the config's size, device-resident,
    no window, no normalisation, no display shift, no waterfall. The yardstick the reference's own GPU backend is built on
    (src/fft_cuda.cu:42,58,133-142); never linked into the product. Returns microseconds per frame."""
    dev = torch.device("cuda", torch.cuda.current_device())
    if cfg.is_real:
        x = torch.randn(frames, cfg.fft_size, device=dev, dtype=torch.float32)
        fn = lambda: torch.fft.rfft(x)  # noqa: E731
    else:
        x = torch.randn(frames, cfg.fft_size, 2, device=dev, dtype=torch.float32)
        xc = torch.view_as_complex(x)
        fn = lambda: torch.fft.fft(xc)  # noqa: E731
    for _ in range(3):
        y = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        y = fn()
    b.record()
    torch.cuda.synchronize()
    del y
    return a.elapsed_time(b) * 1e3 / (reps * frames)


def run_cufft(args):
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    cfg = make_cfg(args)
    us = cufft_yardstick(torch, cfg, args.batch)
    bytes_frame = (4 * cfg.fft_size + 8 * (cfg.fft_size // 2 + 1)) if cfg.is_real else 16 * cfg.fft_size
    print(json.dumps({"impl": "cufft", "what": f"torch.fft ({'R2C' if cfg.is_real else 'C2C'} 2^{args.fft_log2}, batch {args.batch}, "
                      "out of place, device-resident): the bare library transform, no window / shift / normalisation / waterfall",
                      "us_per_frame": us, "input_plus_output_GBps": bytes_frame / us / 1e3,
                      "msps": cfg.hop_samples / us}), flush=True)


def workload_config(cfg, args, world):
    return {
        "workload": f"{cfg.sps / 1e6:g} MSPS {'real' if cfg.is_real else 'complex-IQ'} synthetic, 2^{args.fft_log2} FFT, "
                    f"{args.clients} clients/GPU {'/'.join({USB: 'USB', LSB: 'LSB', AM: 'AM', FM: 'FM'}[m] for m in args.modes)} "
                    + {"cfg1": "(BASELINE.json configs[0] shape)", "cfg2": "(BASELINE.json configs[1] at the metric's client count)",
                       "cfg3": "(BASELINE.json configs[2])"}[args.config],
        "value_definition": VALUE_DEF,
        "waterfall_every_n_frames": args.waterfall_skip if args.waterfall_skip > 0 else cfg.skip_num,
        "fft_size": cfg.fft_size, "audio_fft_size": cfg.audio_fft_size, "downsample_levels": cfg.downsample_levels,
        "clients_per_gpu": args.clients, "clients_total": args.clients * world,
        "frames_per_step": args.ring, "frames_per_launch": args.batch, "pipeline_banks": args.banks,
        "l2_policy": f"inputs larger than L2: {args.ring} hops x {cfg.hop_floats * 4 / 2**20:g} MiB resident ring, "
                     "each hop read by two consecutive frames only",
        "parallelism": "single GPU" if world == 1 else (
                       f"rank 0 ingests + forward FFT; clients of the whole job sorted by (l, r) and split into {world} "
                       f"contiguous blocks; FFT pass 2 stores each rank's sub-band of every frame straight into that rank's "
                       f"HBM over NVLink (peer stores inside the kernel, stream-ordered flags, no collective); "
                       f"{args.clients} clients demodulated per rank" if (args.mgpu_mode == "scatter" and not cfg.is_real) else
                       f"rank 0 ingests + forward FFT; clients of the whole job sorted by (l, r) and split into {world} "
                       f"contiguous blocks; copy engines push each rank's sub-band of every batch into that rank's HBM over "
                       f"NVLink (CUDA IPC, stream-ordered flags, no SM time, no collective); "
                       f"{args.clients} clients demodulated per rank" if (args.mgpu_mode == "scatter-dma" and not cfg.is_real) else
                       f"rank 0 ingests + forward FFT; clients of the whole job sorted by (l, r) and split into {world} "
                       f"contiguous blocks; every client rank's copy engine pulls its sub-band of every batch from the ingest "
                       f"rank's HBM over NVLink (CUDA IPC, stream-ordered flags, no SM time, no collective); "
                       f"{args.clients} clients demodulated per rank" if (args.mgpu_mode == "scatter-pull" and not cfg.is_real) else
                       f"rank 0 ingests + forward FFT; NCCL broadcast of each spectrum batch over NVLink on a "
                       f"communication stream (overlaps the next batch's FFT); "
                       f"{args.clients} clients demodulated per rank ({world} ranks)"),
    }


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import copy

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run_leg(args, torch, dist, world, rank, local, dev, full=True)
    cfg = make_cfg(args)
    if world > 1 and not args.no_mgpu_extras and not cfg.is_real:
        # the other exchange modes (weak scaling, same client load) and a strong-scaling leg (1024 clients in total)
        extras = {}
        for mode in ("spectrum", "scatter", "scatter-dma", "scatter-pull"):
            if mode == args.mgpu_mode:
                continue
            a2 = copy.copy(args)
            a2.mgpu_mode = mode
            r = run_leg(a2, torch, dist, world, rank, local, dev, full=False)
            if rank == 0:
                extras[mode] = {k: r[k] for k in ("value", "ingest_msps", "ms_per_step")}
        a3 = copy.copy(args)
        a3.clients = max(1, args.clients // world)
        r = run_leg(a3, torch, dist, world, rank, local, dev, full=False)
        if rank == 0:
            line["mgpu"] = {
                "default_mode": args.mgpu_mode,
                "modes_weak": {args.mgpu_mode: {k: line[k] for k in ("value", "ingest_msps", "ms_per_step")}, **extras},
                "strong": {"clients_total": a3.clients * world, "clients_per_gpu": a3.clients, "mode": args.mgpu_mode,
                           "ingest_msps": r["ingest_msps"], "ms_per_step": r["ms_per_step"],
                           "note": "fixed client total: the ingest rate rises with N until rank 0's forward group bounds it"},
            }
        # untimed: the exchange step's own parity check (tests/mgpu_worker.py) in THIS process group - every rank's PCM on the
        # spectrum delivered by each exchange mode against a local recomputation, c2c and r2c
        verdict = "ok"
        try:
            sys.path.insert(0, str(ROOT / "tests"))
            import mgpu_worker

            for pcfg in mgpu_worker.CASES():
                mgpu_worker.exchange_case(pcfg, local, dev, rank, world)
        except Exception as exc:  # reported in the line, not raised: the measurements above stand on their own
            verdict = f"FAILED: {exc!r}"[:300]
        flags = [None] * world
        dist.all_gather_object(flags, verdict)
        if rank == 0:
            bad = [f"rank {i}: {v}" for i, v in enumerate(flags) if v != "ok"]
            line["mgpu"]["exchange_parity"] = ("PCM of every rank bit-identical to a local recomputation for NCCL broadcast, "
                                               "fused peer stores, copy-engine push and copy-engine pull (c2c 2^20 and r2c 2^21)") if not bad else bad
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_leg(args, torch, dist, world, rank, local, dev, full=True):
    from phantomsdr_b200.backend import B200FFT, OPT_PCM16, OPT_STAGE_MASK

    cfg = make_cfg(args)
    F, H = args.batch, args.ring
    assert H % F == 0, "--ring must be a multiple of --batch"
    n = cfg.audio_fft_size
    wf_skip = args.waterfall_skip if args.waterfall_skip > 0 else cfg.skip_num

    eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, cfg.brightness_offset, device=local)
    eng.set_output_additional_size(n)
    eng.plan_r2c() if cfg.is_real else eng.plan_c2c()
    E2E_DEPTH = 4  # host blocks in flight in the e2e path: the hop ring keeps that many blocks of halves (+ the shared older half)
    HR = max(H, E2E_DEPTH * F + 2)
    eng.set_hop_ring(HR)
    eng.set_batch_frames(F)
    eng.set_pipeline(args.banks)
    if args.pcm16:
        eng.set_option(OPT_PCM16, 1)
    eng.set_waterfall_cadence(wf_skip)
    eng.clients_create(args.clients, n, cfg.audio_sps)
    scatter = world > 1 and args.mgpu_mode in ("scatter", "scatter-dma", "scatter-pull") and not cfg.is_real
    dma = args.mgpu_mode == "scatter-dma"
    pull = args.mgpu_mode == "scatter-pull"
    if scatter:
        # the (l, r)-sorted client list of the WHOLE job, split into contiguous equal blocks
        from phantomsdr_b200.parallel import partition_clients
        everyone = client_table(cfg, args.clients * world, 0, args.modes)
        parts = partition_clients([(c.l, c.r) for c in everyone], world)
        my_clients = [everyone[i] for i in parts[rank]]
    else:
        my_clients = client_table(cfg, args.clients, rank, args.modes)
    for i, c in enumerate(my_clients):
        eng.client_open(i, c.l, c.mid, c.r, c.mode)
    stream = torch.cuda.Stream(device=dev)  # a real (non-legacy) stream shared by torch/NCCL and the engine
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    ring_t = torch.as_tensor(eng.device_hop_ring(HR), device=dev)
    spec_banks = []
    for b in range(args.banks):
        eng.select_bank(b)
        spec_banks.append(torch.as_tensor(eng.device_spectrum(F), device=dev))
    eng.select_bank(0)
    if rank == 0:
        fill_ring(torch, ring_t, cfg, seed=0x5EED + 2)
    torch.cuda.synchronize()

    # ---- scatter mode: IPC-mapped peer banks + flags ----
    peer_ready, my_ready, r0_consumed, my_consumed = [], None, [], None
    remote_spec, my_band = None, None
    ipc_opened = []
    if scatter:
        def sub_band(block):
            """<= 2 half-open ranges of spectrum indices covering every slice of the block (src/websocket.cpp:182)."""
            iv = sorted((cfg.slice_offset(c.l), cfg.slice_offset(c.l) + (c.r - c.l)) for c in block)
            merged = []
            for a, b in iv:
                if merged and a <= merged[-1][1]:
                    merged[-1][1] = max(merged[-1][1], b)
                else:
                    merged.append([a, b])
            while len(merged) > 2:  # close the smallest gap (sending a few extra bins is harmless)
                gi = min(range(len(merged) - 1), key=lambda i: merged[i + 1][0] - merged[i][1])
                merged[gi][1] = merged[gi + 1][1]
                del merged[gi + 1]
            while len(merged) < 2:
                merged.append([0, 0])
            return merged

        flags = eng.flag_buffer
        mine = {"spec": eng.ipc_export(eng.spectrum_base), "flags": eng.ipc_export(flags)}
        table = [None] * world
        dist.all_gather_object(table, mine)
        if rank == 0:
            ptrs = []
            for g in range(1, world):
                fl = eng.ipc_open(table[g]["flags"])
                ipc_opened.append(fl)
                if not pull:
                    sp = eng.ipc_open(table[g]["spec"])
                    ipc_opened.append(sp)
                    ptrs.append(sp + eng.spectrum_offset)
                peer_ready.append(fl)                                       # flag 0 of rank g: "bank k has landed" / "is ready"
                r0_consumed.append(flags + 8 * g)                           # flag g of rank 0: "rank g is done with bank k"
            if not pull:
                eng.set_peer_spectra(ptrs)
                for g in range(1, world):
                    (a0, b0), (a1, b1) = sub_band([everyone[i] for i in parts[g]])
                    eng.set_peer_ranges(g - 1, a0, b0, a1, b1)
            if dma:
                from phantomsdr_b200.backend import OPT_PEER_STORES
                eng.set_option(OPT_PEER_STORES, 0)
        else:
            my_ready = flags
            f0 = eng.ipc_open(table[0]["flags"])
            ipc_opened.append(f0)
            my_consumed = f0 + 8 * rank
            if pull:
                s0 = eng.ipc_open(table[0]["spec"])
                ipc_opened.append(s0)
                remote_spec = s0 + eng.spectrum_offset
                my_band = sub_band([everyone[i] for i in parts[rank]])
        dist.barrier()

    state = {"frame_num": 0, "batch_no": 0}
    comm = torch.cuda.Stream(device=dev) if (world > 1 and not scatter) else None
    ev_ready = [torch.cuda.Event() for _ in range(args.banks)]
    ev_done = [torch.cuda.Event() for _ in range(args.banks)]

    def batch(hop0):
        """Forward group of hops hop0.. on the ingest rank, the exchange step, every rank's clients - one batch of F frames."""
        frame_num, batch_no = state["frame_num"], state["batch_no"]
        bank = batch_no % args.banks
        eng.select_bank(bank)
        if scatter:
            seq = batch_no + 1
            if rank == 0 and dma:
                eng.execute_device(hop0, F)
                if seq > args.banks:          # peers must have handed this bank back before the DMA overwrites it
                    eng.enqueue_wait(2, r0_consumed, seq - args.banks)
                eng.push_peers(F)             # copy engines: every rank's sub-band of the batch, off the SMs
                eng.enqueue_signal(2, peer_ready, seq)
                eng.clients_execute_device(frame_num, F)
            elif rank == 0:
                if seq > args.banks:          # peers must have handed this bank back
                    eng.enqueue_wait(False, r0_consumed, seq - args.banks)
                eng.execute_device(hop0, F)  # pass 2 also stores every rank's sub-band into that rank's bank
                eng.enqueue_signal(False, peer_ready, seq)
                eng.clients_execute_device(frame_num, F)
            elif pull:
                eng.enqueue_wait(True, [my_ready], seq)
                eng.pull_spectrum(remote_spec, F, my_band[0][0], my_band[0][1], my_band[1][0], my_band[1][1])
                eng.enqueue_signal(True, [my_consumed], seq)   # pulled: the ingest rank may reuse its bank
                eng.clients_execute_device(frame_num, F)
            else:
                eng.enqueue_wait(True, [my_ready], seq)
                eng.clients_execute_device(frame_num, F)
                eng.enqueue_signal(True, [my_consumed], seq)
        else:
            if rank == 0:
                eng.execute_device(hop0, F)      # waits for this bank's previous clients, then FFT + pyramid
            else:
                eng.bank_acquire()                # the broadcast may overwrite the bank once its clients are done
            if world > 1:
                # the exchange step runs on its own stream: broadcast k overlaps the forward FFT of batch k+1
                ev_ready[bank].record(stream)
                comm.wait_event(ev_ready[bank])
                with torch.cuda.stream(comm):
                    dist.broadcast(spec_banks[bank], src=0)
                    ev_done[bank].record(comm)
                eng.client_stream_wait_event(ev_done[bank].cuda_event)
            eng.clients_execute_device(frame_num, F)
        state["frame_num"] += F
        state["batch_no"] += 1

    def step():
        for g in range(H // F):
            batch(g * F)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    eng.join_streams()
    barrier()
    sampler = ClockSampler(local)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    eng.join_streams()
    ev1.record(stream)
    barrier()
    sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    samples_per_step = H * cfg.hop_samples
    ingest = samples_per_step / (ms_step * 1e-3) / 1e6
    value = world * ingest  # VALUE_DEF

    # ---- e2e at N > 1: rank 0 ingests from HOST buffers, the exchange step, every rank copies its results to the host ----
    e2e = None
    h = n // 2
    pcm_bytes = 2 if args.pcm16 else 4
    nblk = max(2, args.e2e_frames // F)
    if full and not args.no_e2e and world > 1:
        host_in = [torch.empty((F, cfg.hop_floats), dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None
        if rank == 0:
            for t_ in host_in:
                t_.normal_(0, 1e-3)
        outs = [dict(pcm=eng.pinned(pcm_bytes * F * args.clients * h, np.uint8), pwr=eng.pinned(4 * F * args.clients, np.float32),
                     valid=eng.pinned(F * args.clients, np.uint8)) for _ in range(2)]
        pyr_host = torch.empty((F, eng.pyramid_stride), dtype=torch.int8).pin_memory() if rank == 0 else None
        copy_s = torch.cuda.Stream(device=dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def e2e_multi(blocks):
            for k in range(blocks):
                half = k & 1
                hop0 = half * (F + 1)  # two disjoint regions of the hop ring (HR >= 2 F + 2), one per block in flight
                if rank == 0:
                    # H2D of the block's F new halves into the ring slots the forward group of this batch reads
                    with torch.cuda.stream(copy_s):
                        if k >= 2:
                            copy_s.wait_event(ev_free[half])  # the forward group of block k - 2 read this region
                        ring_t[hop0 + 1:hop0 + 1 + F].copy_(host_in[half], non_blocking=True)  # one copy: F contiguous halves
                        ev_in[half].record(copy_s)
                    stream.wait_event(ev_in[half])
                if k >= 2:
                    eng.clients_fetch_wait(half)  # the host buffers of block k - 2 are about to be reused
                batch(hop0)
                if rank == 0:
                    ev_free[half].record(stream)
                    q = torch.as_tensor(eng.device_quantized(F), device=dev)
                    pyr_host.copy_(q, non_blocking=True)  # (engine stream, behind the forward group)
                eng.clients_fetch_async(half, F, outs[half]["pcm"], outs[half]["pwr"], outs[half]["valid"])
            for half in range(min(2, blocks)):
                eng.clients_fetch_wait(half)
            eng.join_streams()
            torch.cuda.synchronize()

        e2e_multi(3)
        barrier()
        t0 = time.perf_counter()
        e2e_multi(nblk)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e_ingest = nblk * F * cfg.hop_samples / dt / 1e6
        e2e = {"value": world * e2e_ingest, "unit": UNIT, "ingest_msps": e2e_ingest,
               "h2d_bytes_per_step": cfg.hop_floats * 4 * H,
               "d2h_bytes_per_step": (eng.pyramid_bytes + world * args.clients * (h * pcm_bytes + 5)) * H,
               "frames_timed": nblk * F,
               "path": f"rank 0: {F} new halves per batch from pinned host memory (H2D on a copy stream) -> forward group -> exchange "
                       f"({args.mgpu_mode}) -> every rank's clients -> every rank's PCM/pwr/valid (b200_clients_fetch_async) and rank 0's "
                       "pyramid back to pinned host memory; two batches in flight; value = ingest x client shards (VALUE_DEF)"}

    if scatter:
        barrier()
        if rank == 0:
            eng.set_peer_spectra([])  # the single-rank measurements below must not touch the peers
        assert eng.flag_error == 0, "a flag wait timed out"
    # ---- roofline of the dominant kernel group (forward FFT + waterfall), timed alone on rank 0 ----
    roofline, breakdown = None, {}
    if rank == 0 and full:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_frame = algorithmic_bytes_per_frame(cfg)
        eng.set_waterfall_cadence(1)  # the roofline counts a pyramid per frame

        def time_fwd(mask, reps=20):
            eng.set_option(OPT_STAGE_MASK, mask)
            for g in range(H // F):
                eng.execute_device(g * F, F)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                for g in range(H // F):
                    eng.execute_device(g * F, F)
            b.record(stream)
            torch.cuda.synchronize()
            eng.set_option(OPT_STAGE_MASK, 7)
            return a.elapsed_time(b) * 1e-3 / (reps * H)  # seconds per frame

        t_fwd = time_fwd(7)
        achieved = bytes_frame / t_fwd / 1e9
        for name, mask in (("fft_pass1", 1), ("fft_pass2", 2), ("pyramid", 4)):
            breakdown[name + "_us_per_frame"] = round(time_fwd(mask) * 1e6, 3)
        eng.set_waterfall_cadence(wf_skip)
        # client kernels alone
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        eng.select_bank(0)
        a.record(stream)
        reps = 20
        for r_ in range(reps):
            eng.clients_execute_device(state["frame_num"], F)
            state["frame_num"] += F
        eng.join_streams()
        b.record(stream)
        torch.cuda.synchronize()
        breakdown["clients_us_per_frame"] = round(a.elapsed_time(b) * 1e3 / (reps * F), 3)
        breakdown["forward_us_per_frame"] = round(t_fwd * 1e6, 3)
        try:  # the library transform alone, on the same box (yardstick, not part of the product)
            breakdown["cufft_bare_transform_us_per_frame"] = round(cufft_yardstick(torch, cfg, min(F, 32)), 3)
        except Exception as exc:
            breakdown["cufft_bare_transform_us_per_frame"] = f"unavailable: {exc!r}"
        traffic, traffic_src = None, None
        try:
            # measured once per kernel change with ncu --set full (tools/gpu_profile_pack.sh, tools/ncu_summary.py traffic);
            # the file names the build it was taken from
            for name in ("r2_traffic.json", "r2_traffic_r2c.json"):
                tr = json.loads((ROOT / "profiles" / name).read_text())
                if tr.get("batch") == F and tr.get("is_real") == cfg.is_real and tr.get("fft_log2") == args.fft_log2:
                    traffic = tr["dram_bytes_per_launch_group"]
                    traffic_src = tr.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "forward FFT + waterfall (fft_pass1 + fft_pass2 + pyramid, one launch each per "
                                             f"{F} frames)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": which, "algorithmic_bytes_per_frame": bytes_frame,
                    "algorithmic_bytes_per_launch_group": bytes_frame * F, "us_per_frame": t_fwd * 1e6}

    # ---- e2e at N = 1: same frames through the reference-facing C-ABI with HOST buffers ----
    # b200_submit_block / b200_wait_block = load_*_input + execute + signal_loop for F frames per call, pipelined:
    # H2D of block k+1, kernels of block k, D2H of block k's results (pyramid + PCM/pwr/valid) on their own streams.
    e2e_raw = None
    if full and not args.no_e2e and world == 1:
        eng.join_streams()
        eng.sync()
        rs = np.random.default_rng(0x5EED + 3 + rank)
        sets = []
        for _ in range(E2E_DEPTH):  # blocks in flight -> as many sets of pinned host result buffers
            sets.append(dict(pcm=eng.pinned(pcm_bytes * F * args.clients * h, np.uint8),
                             pwr=eng.pinned(4 * F * args.clients, np.float32),
                             valid=eng.pinned(F * args.clients, np.uint8),
                             pyr=eng.pinned(F * eng.pyramid_bytes, np.int8)))
        # pyramids come back only for the send frames (frame_num % skip == 0) of the timed frames 3F .. 3F + nblk F
        n_send = sum(1 for f in range(3 * F, 3 * F + nblk * F) if f % wf_skip == 0)
        dbytes_frame = eng.pyramid_bytes * n_send / (nblk * F) + args.clients * (h * pcm_bytes + 5)

        def host_halves(dt):
            halves_sets = []
            for _ in range(2):
                halves = []
                blockbuf = eng.pinned(F * cfg.hop_floats * np.dtype(dt).itemsize, dt)  # a block's halves back to back
                for _k in range(F):
                    hb = blockbuf[_k * cfg.hop_floats:(_k + 1) * cfg.hop_floats]
                    if dt == np.float32:
                        hb[:] = (rs.standard_normal(cfg.hop_floats) * 1e-3).astype(np.float32)
                    elif dt == np.int16:
                        hb[:] = (rs.standard_normal(cfg.hop_floats) * 33.0).astype(np.int16)       # ~1e-3 full scale
                    else:
                        hb[:] = (rs.standard_normal(cfg.hop_floats) * 2.0 + 128.0).astype(np.uint8)
                    halves.append(hb)
                halves_sets.append(halves)
            return halves_sets

        def e2e_run(halves_sets, prime, blocks, f0):
            eng.stream_prime(prime)
            for k in range(blocks):
                st = sets[k % E2E_DEPTH]
                if k >= E2E_DEPTH:
                    eng.wait_block()  # block k - depth used this buffer set
                eng.submit_block(halves_sets[k & 1], f0 + k * F, st["pcm"], st["pwr"], st["valid"], st["pyr"])
            for _ in range(min(E2E_DEPTH, blocks)):
                eng.wait_block()

        def e2e_measure(dt):
            hs = host_halves(dt)
            e2e_run(hs, hs[0][0], 3, 0)  # warm-up
            t0 = time.perf_counter()
            e2e_run(hs, hs[0][0], nblk, 3 * F)
            return time.perf_counter() - t0

        dt = e2e_measure(np.float32)
        common = f"-> forward FFT + pyramid (every {wf_skip} frame(s)) -> clients -> D2H of the int8 pyramids and " \
                 f"{'int16' if args.pcm16 else 'int32'} PCM / pwr / valid of every frame; four blocks in flight"
        e2e = {"value": nblk * F * cfg.hop_samples / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": cfg.hop_floats * 4 * H, "d2h_bytes_per_step": int(dbytes_frame * H),
               "frames_timed": nblk * F,
               "path": f"b200_submit_block / b200_wait_block: {F} frames per call from pinned host float halves (H2D of every new half) "
                       + common}
        if args.e2e_raw:
            try:
                from phantomsdr_b200.backend import OPT_INPUT_FORMAT, _FMT_OF_DTYPE
                rdt = np.int16 if args.e2e_raw == "s16" else np.uint8
                eng.join_streams()
                eng.sync()
                eng.set_option(OPT_INPUT_FORMAT, _FMT_OF_DTYPE[np.dtype(rdt).name])
                dtm = e2e_measure(rdt)
                e2e_raw = {"value": nblk * F * cfg.hop_samples / dtm / 1e6, "unit": UNIT, "format": args.e2e_raw,
                           "h2d_bytes_per_step": cfg.hop_floats * np.dtype(rdt).itemsize * H, "d2h_bytes_per_step": int(dbytes_frame * H),
                           "frames_timed": nblk * F,
                           "path": "as e2e, but the halves are raw ADC samples: SampleConverter (src/samplereader.cpp:29-66) runs inside "
                                   "FFT pass 1 " + common}
            except Exception as exc:
                e2e_raw = {"value": None, "unit": UNIT, "format": args.e2e_raw, "error": repr(exc)}

    cpu_baseline = None
    if rank == 0 and world == 1 and full and not args.no_cpu_baseline:
        try:
            fps_probe, _ = cpu_reference_run(cfg, args.clients, frames=2, warm=1, modes=args.modes)
            frames = args.cpu_frames or int(max(4, min(2000, 12.0 * fps_probe)))
            fps, info = cpu_reference_run(cfg, args.clients, frames=frames, warm=1, modes=args.modes)
            cpu_baseline = {"value": fps * cfg.hop_samples / 1e6, "unit": UNIT, **info}
            # the reference's default is fft_threads = 1 (src/spectrumserver.cpp:39): the same sample with a one-thread FFT
            fps1, info1 = cpu_reference_run(cfg, args.clients, frames=max(4, frames // 3), warm=1, modes=args.modes, fft_threads=1)
            cpu_baseline["fft_threads_1"] = {"value": fps1 * cfg.hop_samples / 1e6, "unit": UNIT,
                                             "split_ms_per_frame": info1["split_ms_per_frame"]}
        except Exception as exc:  # the oracle is test infrastructure; its absence must not hide the GPU number
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"unavailable: {exc!r}"}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args, world),
            "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "breakdown": breakdown,
            **({"e2e_raw": e2e_raw} if e2e_raw is not None else {}),
            "ingest_msps": ingest,
            "realtime_margin": ingest / (cfg.sps / 1e6),
        }
    if world > 1:
        dist.barrier()
        if rank == 0 or scatter:
            for ptr in ipc_opened:
                eng.ipc_close(ptr)
        dist.barrier()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    eng.close()
    return line


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cufft":
        run_cufft(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
