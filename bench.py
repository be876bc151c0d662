#!/usr/bin/env python
"""
bench.py -- headline benchmark of the wav -> x-vector hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload frontend|wav2xvec|tdnn|plda]

N = 1 (default) measures BASELINE config 2, the configuration the metric is quoted on:
  MFCC(30 mfcc / 30 mel) + CMVN(window 200) front-end on a batch of 1024 x 10 s of synthetic
  16 kHz audio (float32, +-32767 scale).  One "step" = one pass of the hot path over one batch.
N > 1 (launched under torchrun, one rank per GPU): the same per-GPU batch on every rank -- utterances
  are sharded, there is NO data-path collective ("weak" scaling); value = all ranks' audio seconds /
  max-over-ranks device time.

One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every field):
  value      audio-seconds per second with the batch already resident in HBM (device-timed, CUDA events)
  e2e        the same metric through the public layer API with HOST (pinned) buffers: H2D of the wav
             batch and D2H of the features are inside the timed region every step
  roofline   front-end kernel: algorithmic bytes (wav in + features out) / its own device time vs the
             measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/ktf_oracle.py, an op-for-op NumPy port of the reference's layers;
             the reference itself needs TensorFlow 2.8 which is not installable here) on a bounded sample

`--impl reference` times that CPU port with all host threads on the same workload definition
(a bounded sample per step); it is the only other place allowed to execute oracle/.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

SR = 16000
UTT_SECONDS = 10
UTT_SAMPLES = SR * UTT_SECONDS
BATCH = 1024
FRAMES = 1 + (UTT_SAMPLES - 400) // 160          # 998
NUM_CEPS = 30
CMVN_WINDOW = 200
ALGO_BYTES_PER_UTT = UTT_SAMPLES * 4 + FRAMES * NUM_CEPS * 4      # wav once + features once


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clocks.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(clocks)}


# ----------------------------------------------------------------------------------------------
# CPU port (oracle) timing helpers -- cpu_baseline leg and --impl reference
# ----------------------------------------------------------------------------------------------

def synth_wav_np(n_utt, seed):
    rng = np.random.default_rng(seed)
    return np.clip(rng.standard_normal((n_utt, UTT_SAMPLES), dtype=np.float32) * 3000.0, -32767, 32767)


def oracle_frontend(wav, chunk=8):
    """The reference layers' arithmetic (NumPy port), a few utterances at a time to bound host memory."""
    from oracle import ktf_oracle as O
    outs = []
    for i in range(0, wav.shape[0], chunk):
        x = O.framing(wav[i:i + chunk], 25.0, 10.0, float(SR))
        x = O.mfcc(x, num_mfccs=NUM_CEPS, num_mels=30)
        outs.append(O.cmvn(x, window=CMVN_WINDOW))
    return outs


def time_oracle(n_utt, threads, seed=0):
    """Runs the NumPy port on n_utt utterances split over `threads` host threads; returns seconds."""
    wav = synth_wav_np(n_utt, seed)
    chunks = [c for c in np.array_split(np.arange(n_utt), threads) if len(c)]
    t0 = time.perf_counter()
    if len(chunks) == 1:
        oracle_frontend(wav)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(chunks)) as ex:
            list(ex.map(lambda idx: oracle_frontend(wav[idx]), chunks))
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the CPU port of the reference's layers on all host threads (TensorFlow 2.8, which the
    reference needs, is not installable in this image: `import tensorflow` fails, no network)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_utt = max(threads, 32)                         # bounded sample of the 1024-utterance batch per step
    for _ in range(args.warmup):
        time_oracle(n_utt, threads)
    t = 0.0
    for s in range(args.steps):
        t += time_oracle(n_utt, threads, seed=s)
    value = n_utt * UTT_SECONDS * args.steps / t
    line = {
        "impl": "reference", "metric": "audio_sec_per_sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps * (BATCH / n_utt), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{n_utt} of {BATCH} utterances x {UTT_SECONDS} s per step "
                                   f"(NumPy port of the reference layers; TensorFlow unavailable)"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "BASELINE config 2: Framing(25ms/10ms) -> MFCC(30 mfcc, 30 mel, povey, dither 0) -> "
                        "CMVN(window 200), batch 1024 x 10 s synthetic 16 kHz float32 audio per GPU",
            "batch_per_gpu": BATCH, "utt_seconds": UTT_SECONDS, "frames_per_utt": FRAMES,
            "parallelism": f"utterance-sharded x{n_gpus}, no collective",
            "l2_policy": "input batch is 655 MB per step, larger than the 126 MB L2 (no explicit flush needed)"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import kaldi_tflite_b200 as ktf
    from kaldi_tflite_b200 import _native, _tensor

    dev = torch.device("cuda", local_rank)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav = (torch.randn((BATCH, UTT_SAMPLES), generator=g, device=dev) * 3000.0).clamp_(-32767, 32767)

    framing = ktf.layers.Framing(25.0, 10.0, float(SR), dynamic_input_shape=True)
    mfcc = ktf.layers.MFCC(num_mfccs=NUM_CEPS, num_mels=30)
    cmvn = ktf.layers.CMVN(center=True, window=CMVN_WINDOW, norm_vars=False)

    def step(x):
        return cmvn(mfcc(framing(x)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step(wav)
    barrier()

    # ---- device-resident timing (value) + per-kernel timing of the front-end kernel (roofline) ----
    fe = mfcc.frontend(framing.frameWidth, framing.frameShift)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    n0 = ktf.launch_count()
    sampler.start()
    barrier()
    ev0.record()
    for s in range(args.steps):
        k_ev[s][0].record()
        feats, _ = fe.forward(wav)                         # the fused front-end kernel (one launch)
        k_ev[s][1].record()
        out = cmvn(feats)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = ktf.launch_count() - n0
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())

    # ---- end-to-end through the public API with host buffers ---------------------------------------
    host_in = torch.empty((BATCH, UTT_SAMPLES), dtype=torch.float32, pin_memory=True)
    host_in.copy_(wav)
    host_out = torch.empty((BATCH, FRAMES, NUM_CEPS), dtype=torch.float32, pin_memory=True)
    for _ in range(2):
        host_out.copy_(step(host_in.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x = host_in.to(dev, non_blocking=True)              # H2D of this step's inputs
        y = step(x)                                         # public layer API
        host_out.copy_(y, non_blocking=True)                # D2H of this step's result
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e_max = float(t2.item())

    if rank == 0:
        audio_s = BATCH * UTT_SECONDS * world * args.steps
        peak, peak_src = measured_peaks()
        achieved = ALGO_BYTES_PER_UTT * BATCH / (ms_kernel * 1e-3) / 1e9
        # bounded CPU-port sample on rank 0 (N = 1 only): 1 thread, 192 utterances ~ 10-20 s
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 192
            try:                                            # keep BLAS / FFT pools at one thread
                from threadpoolctl import threadpool_limits
                with threadpool_limits(limits=1):
                    tc = time_oracle(n_cpu, 1)
            except ImportError:
                tc = time_oracle(n_cpu, 1)
            cpu = {"value": n_cpu * UTT_SECONDS / tc, "unit": "audio-s/s", "cores": 1, "kind": "port",
                   "sample": f"{n_cpu} utterances x {UTT_SECONDS} s of the same workload, NumPy port of the "
                             f"reference layers, 1 thread ({tc:.1f} s)"}
        line = {
            "metric": "audio_sec_per_sec", "value": audio_s / (ms_total_max * 1e-3), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "frontend_kernel<32,25> (framing+window+FFT+mel+log+DCT)",
                         "kernel_ms": ms_kernel,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UTT * BATCH},
            "e2e": {"value": audio_s / (ms_e2e_max * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": BATCH * UTT_SAMPLES * 4,
                    "d2h_bytes_per_step": BATCH * FRAMES * NUM_CEPS * 4,
                    "ms_per_step": ms_e2e_max / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        traffic_file = os.path.join(ROOT, "profiles", "frontend_traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as f:
                line["roofline"]["traffic"] = json.load(f).get("dram_bytes_per_launch")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
