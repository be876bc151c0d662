#!/usr/bin/env python
"""
Device-time of each stage of the hot path at the BASELINE.json config shapes (development aid;
the judged numbers come from bench.py).  Prints one JSON object per stage.

    python scripts/bench_parts.py [tdnn] [wav2xvec] [plda] [frontend] [--iters N]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import kaldi_tflite_b200 as ktf


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ktf.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (ktf.launch_count() - n0) // iters


def sitw_model(precision):
    import yaml
    with open(os.path.join(ROOT, "data", "kaldi_models", "configs", "0008_sitw_v2_1a.yml")) as f:
        cfg = yaml.safe_load(f)
    return ktf.models.SequentialFromConfig(cfg["model_config"], None, "cmvn2xvec", precision=precision, seed=0)


def extractor_cfg():
    import yaml
    with open(os.path.join(ROOT, "data", "tflite_models", "0008_sitw_v2_1a.yml")) as f:
        cfg = yaml.safe_load(f)["extractor"]
    g = os.path.join(ROOT, "tests", "golden")
    cfg["mfcc"]["dither"] = 0.0
    cfg["xvec"]["model_config_path"] = os.path.join(ROOT, cfg["xvec"]["model_config_path"])
    cfg["xvec"]["model_path"] = None
    cfg["xvec"]["global_mean_path"] = os.path.join(g, "sitw_mean.vec")
    cfg["xvec"]["lda_matrix_path"] = os.path.join(g, "sitw_transform.mat")
    return cfg


def gated_noise(n_utt, n_samples, seed, dev):
    """Gaussian noise (sigma 3000) gated by a random on/off envelope, 'silence' sigma 30."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n_utt, n_samples), generator=g, device=dev)
    seg = 3200                                               # 0.2 s envelope granularity
    nseg = (n_samples + seg - 1) // seg
    on = (torch.rand((n_utt, nseg), generator=g, device=dev) < 0.7).float()
    env = on.repeat_interleave(seg, dim=1)[:, :n_samples]
    return x * (30.0 + 2970.0 * env)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    iters = 10
    if "--iters" in sys.argv:
        iters = int(sys.argv[sys.argv.index("--iters") + 1])
    what = set(args) or {"tdnn", "wav2xvec", "plda", "frontend"}
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)

    if "frontend" in what:
        g = torch.Generator(device=dev).manual_seed(1234)
        wav = (torch.randn((1024, 160000), generator=g, device=dev) * 3000.0).clamp_(-32767, 32767)
        fr = ktf.layers.Framing(25.0, 10.0, 16000.0, dynamic_input_shape=True)
        mf = ktf.layers.MFCC(num_mfccs=30, num_mels=30)
        cm = ktf.layers.CMVN(center=True, window=200, norm_vars=False)
        ms, nl = timed(lambda: cm(mf(fr(wav))), iters)
        print(json.dumps({"stage": "frontend cfg2 1024x10s MFCC+CMVN", "ms": ms, "launches": nl,
                          "GBps": 777994240 / ms / 1e6}))
        del wav

    if "tdnn" in what:
        for prec in ("bf16",):
            mdl = sitw_model(prec)
            g = torch.Generator(device=dev).manual_seed(1234)
            x = torch.randn((512, 300, 30), generator=g, device=dev)
            flat = x.reshape(-1, 30)
            offs = torch.arange(513, device=dev, dtype=torch.int64) * 300
            ms, nl = timed(lambda: mdl.forward_ragged(flat, offs), iters)
            flops = 8.248e11
            print(json.dumps({"stage": f"tdnn cfg3 512x300 {prec}", "ms": ms, "launches": nl,
                              "TFLOPs": flops / ms / 1e9}))

    if "wav2xvec" in what:
        ext = ktf.models.XvectorExtractor(extractor_cfg(), precision="bf16", seed=0, allow_random_init=True)
        B = 512
        wav = gated_noise(B, 160000, 7, dev)
        ms, nl = timed(lambda: ext(wav), iters)
        _, inter = ext(wav, return_intermediate=True)
        kept = float(inter["mask"].mean().item())
        print(json.dumps({"stage": f"wav2xvec cfg4 shard {B}x10s bf16", "ms": ms, "launches": nl,
                          "audio_s_per_s": B * 10 / ms * 1e3, "vad_keep": kept}))
        del wav

    if "plda" in what:
        from test_gpu_tdnn_plda import synthetic_plda
        dim, n = 128, 16384
        mean, Tm, psi = synthetic_plda(dim)
        g = torch.Generator(device=dev).manual_seed(1234)
        x = torch.randn((n, dim), generator=g, device=dev)
        x = x / x.norm(dim=1, keepdim=True) * dim ** 0.5
        layer = ktf.layers.PLDA(dim, mean, Tm, psi, dtype=np.float32, return_transformed=False)
        u = layer.transformVector(x)
        ms, nl = timed(lambda: layer.logLikelihoodRatio(u, u), iters)
        print(json.dumps({"stage": f"plda score {n}x{n} d{dim} f32", "ms": ms, "launches": nl,
                          "TFLOPs": 2.0 * n * n * dim / ms / 1e9, "out_GBps": n * n * 4 / ms / 1e6}))


if __name__ == "__main__":
    main()
