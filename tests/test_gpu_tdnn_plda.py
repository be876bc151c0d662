"""
GPU parity of the TDNN / BatchNorm / StatsPooling / LDA / PLDA kernels through the public layer
API (C-ABI underneath), mirroring the reference's tests and tolerances, plus oracle comparisons
on seeded random weights for the full-width SITW stack (its Kaldi weights are not vendored).
"""

import json

import numpy as np
import pytest

from conftest import load_golden, golden_path, rmse, read_wav_int16
import helpers
from oracle import ktf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ktf():
    import kaldi_tflite_b200 as k
    return k


def test_tdnn_single_layer_vs_kaldi(ktf):
    # layers/tdnn/tdnn_test.py:45-57 (tolerance :31)
    g = load_golden("tdnn.npz")
    cfg = json.loads(str(g["single_cfg"]))
    r = ktf.io.KaldiNnet3Reader(golden_path("tdnn_single_layer.final.raw"), True)
    layer = ktf.layers.TDNN.from_config({**cfg, "precision": "f32"})   # the reference's 1e-6 is an fp32 tolerance
    layer.build(g["single_in"].shape)
    layer.set_weights([r.components[0]["params"], r.components[0]["bias"]])
    got = layer(g["single_in"])
    assert got.shape == g["single_out"].shape
    assert rmse(g["single_out"], got) <= 1e-6


def test_tdnn_narrow_vs_kaldi(ktf):
    # layers/tdnn/tdnn_test.py:105-119 (tolerance :117), layer by layer AND through the fused plan
    g = load_golden("tdnn.npz")
    r = ktf.io.KaldiNnet3Reader(golden_path("tdnn_narrow.final.raw"), True)
    layers = []
    for name, dim, ctx, use_relu, use_bn in helpers.NARROW_LAYERS:
        l = ktf.layers.TDNN(dim, context=ctx, name=f"{name}.affine", precision="f32")
        layers.append(l)
        if use_relu:
            layers.append(ktf.layers.ReLU(name=f"{name}.relu"))
        if use_bn:
            layers.append(ktf.layers.BatchNorm(name=f"{name}.batchnorm"))
    mdl = ktf.models.Sequential(layers, input_shape=(None, None, 3), precision="f32")
    for l in mdl.layers:
        l.set_weights(r.getWeights(l.name))
    y = g["narrow_in"]
    for l in mdl.layers:
        y = l(y)
    assert y.shape == g["narrow_out"].shape
    assert rmse(g["narrow_out"], y) <= 5e-4
    fused = mdl(g["narrow_in"])
    assert rmse(g["narrow_out"], fused) <= 5e-4
    assert np.max(np.abs(fused - y)) < 1e-5


def test_tdnn_modes_vs_oracle(ktf):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 37, 20)).astype(np.float32)
    for ctx, sub, pad in [([-2, 0, 2], 1, "SAME"), ([-3, -1, 0, 1], 1, "VALID"), ([0], 3, "SAME"),
                          ([-1, 0, 1], 2, "VALID"), ([-2, -1, 0, 1, 2], 1, "SAME"), ([0, 3], 1, "VALID")]:
        l = ktf.layers.TDNN(24, context=ctx, subsampling_factor=sub, padding=pad, seed=1, precision="f32")
        got = l(x)
        kernel, bias = l.get_weights()
        want = O.tdnn(x, kernel, bias, ctx, sub, pad)
        assert got.shape == want.shape == tuple(l.compute_output_shape(x.shape)), (ctx, sub, pad)
        assert np.max(np.abs(got - want)) < 1e-4, (ctx, sub, pad)
        # the same modes on the tcgen05 engine (default precision): exact product of the
        # bf16-rounded operands, fp32 accumulation
        lt = ktf.layers.TDNN(24, context=ctx, subsampling_factor=sub, padding=pad, seed=1)
        assert lt.precision is None
        got_tc = lt(x)
        want_tc = O.tdnn(_bf16_round(x), _bf16_round(kernel), bias, ctx, sub, pad)
        assert got_tc.shape == want.shape, (ctx, sub, pad)
        assert np.max(np.abs(got_tc - want_tc)) < 1e-4, (ctx, sub, pad)
    with pytest.raises(ValueError):
        ktf.layers.TDNN(8, subsampling_factor=0)
    with pytest.raises(ValueError):
        ktf.layers.TDNN(8, padding="FULL")


def test_batchnorm_layer(ktf):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 9, 16)).astype(np.float32)
    bn = ktf.layers.BatchNorm()
    mean, var = rng.standard_normal(16).astype(np.float32), rng.random(16).astype(np.float32) + 0.5
    bn.set_weights([np.float32(0.7), mean, var])
    got = bn(x)
    want = O.batchnorm(x, 0.7 * np.ones(16, np.float32), mean, var, 0.001)
    assert np.max(np.abs(got - want)) < 1e-5


def test_stats_pooling_vs_kaldi(ktf):
    # layers/stats/stats_pooling_test.py:48-88 (tolerance :26)
    g = load_golden("stats.npz")
    cfg = helpers.stats_default_cfg()
    cfg["reduce_time_axis"] = True
    got = ktf.layers.StatsPooling(**cfg)(g["in_stats_mean_std"])
    assert rmse(g["out_stats_mean_std"][:, 0:1, :], got) <= 4e-6
    for name, over in helpers.STATS_CONFIGS.items():
        cfg = helpers.stats_default_cfg()
        cfg.update(over)
        layer = ktf.layers.StatsPooling(**cfg)
        got = layer(g[f"in_{name}"])
        want = g[f"out_{name}"]
        assert got.shape == want.shape, name
        assert rmse(want, got) <= 4e-6, name
        for pad in ("SAME", "VALID"):
            cfg["padding"] = pad
            x = np.random.default_rng(4).standard_normal((2, 23, 6)).astype(np.float32)
            a = ktf.layers.StatsPooling(**cfg)(x)
            b = O.stats_pooling(x, **cfg)
            assert a.shape == b.shape, (name, pad)
            assert np.max(np.abs(a - b)) < 1e-5, (name, pad)


def test_plda_vs_kaldi(ktf):
    # layers/plda/plda_test.py:45-62 (tolerance :30), float32 like the reference test and float64
    g = load_golden("plda.npz")
    rd = ktf.io.KaldiPldaReader(golden_path("plda.bin"), True)
    assert np.allclose(rd.mean, g["mean"], atol=1e-9) and np.allclose(rd.psi, g["psi"], atol=1e-9)
    for dt in (np.float32, np.float64):
        plda = ktf.layers.PLDA(int(g["dim"]), rd.mean, rd.transformMat, rd.psi, dtype=dt)
        scores, transformed = plda(g["plda_input"])
        assert transformed.shape == g["plda_transformed"].shape
        assert scores.shape == g["scores"].shape
        assert scores.dtype == dt
        assert rmse(g["plda_transformed"], transformed) <= 2e-4
        assert rmse(g["scores"], scores) <= 2e-4
    only = ktf.layers.PLDA(int(g["dim"]), rd.mean, rd.transformMat, rd.psi, dtype=np.float32,
                           return_transformed=False)(g["plda_input"][:, 0, :])
    assert only.shape == (29, 29)


def synthetic_plda(dim, seed=1234):
    rng = np.random.default_rng(seed)
    psi = np.exp(np.linspace(3, -4, dim))
    q, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    Tm = q * rng.uniform(0.5, 2.0, size=(1, dim))
    mean = rng.standard_normal(dim) * 0.05
    return mean, Tm, psi


def test_plda_num_examples_vs_oracle(ktf):
    # enrolled vectors that average n utterances (plda.py:163-182, 215-231), both dtypes and
    # both length normalisations, against the float64 oracle of the direct broadcast form
    import torch
    dim, n = 128, 300
    mean, Tm, psi = synthetic_plda(dim)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for ne in (1.0, 3.0, 10.0):
        for simple in (False, True):
            u64 = O.plda_transform(x, mean, Tm, psi, True, simple, np.float64, num_examples=ne)
            want = O.plda_llr(u64, psi, np.float64, num_examples=ne)
            for dt in (np.float32, np.float64):
                layer = ktf.layers.PLDA(dim, mean, Tm, psi, simple_length_norm=simple, dtype=dt)
                u = layer.transformVector(xd, num_examples=ne)
                got = layer.logLikelihoodRatio(u, num_examples=ne).cpu().numpy()
                assert np.max(np.abs(u.cpu().numpy() - u64)) < (1e-4 if dt == np.float32 else 1e-9), (ne, simple, dt)
                tol = 1e-3 if dt == np.float32 else 1e-9
                assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < tol, (ne, simple, dt)
    with pytest.raises(AssertionError):
        ktf.layers.PLDA(dim, mean, Tm, psi).transformVector(xd, num_examples=0)


def test_plda_random_vs_oracle(ktf):
    # SURVEY.md 8d cfg5 parity: |delta| <= 1e-3 * max(|s|, 1) against the float64 oracle
    dim, n = 128, 700
    mean, Tm, psi = synthetic_plda(dim)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, dim))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True) * np.sqrt(dim)).astype(np.float32)
    want, _ = O.plda(x, mean, Tm, psi, dtype=np.float64)
    for dt in (np.float32, np.float64):
        for simple in (False, True):
            layer = ktf.layers.PLDA(dim, mean, Tm, psi, dtype=dt, simple_length_norm=simple,
                                    return_transformed=False)
            got = layer(x)
            ref = want if not simple else O.plda(x, mean, Tm, psi, simple_length_norm=True)[0]
            assert np.all(np.abs(got - ref) <= 1e-3 * np.maximum(np.abs(ref), 1.0)), (dt, simple)


def sitw_layers_for_oracle(mdl):
    import kaldi_tflite_b200 as k
    out = []
    for l in mdl.layers:
        if isinstance(l, k.layers.TDNN):
            out.append({"type": "affine", "kernel": l.kernel, "bias": l.bias, "context": l.context})
        elif isinstance(l, k.layers.ReLU):
            out.append({"type": "relu"})
        elif isinstance(l, k.layers.BatchNorm):
            out.append({"type": "batchnorm", "gamma": l.gamma, "mean": l.moving_mean,
                        "var": l.moving_variance, "epsilon": l.epsilon})
        elif isinstance(l, k.layers.StatsPooling):
            cfg = l.get_config()
            cfg.pop("name"), cfg.pop("trainable")
            out.append({"type": "stats", **cfg})
    return out


def sitw_model(ktf, precision=None, seed=0):
    import yaml, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "data", "kaldi_models", "configs", "0008_sitw_v2_1a.yml")) as f:
        cfg = yaml.safe_load(f)
    mdl = ktf.models.SequentialFromConfig(cfg["model_config"], None, "cmvn2xvec", precision=precision,
                                          seed=seed)
    rng = np.random.default_rng(seed + 100)
    for l in mdl.layers:                     # non-trivial BatchNorm statistics
        if isinstance(l, ktf.layers.BatchNorm):
            d = l.gamma.shape[0]
            l.set_weights([np.float32(1.0), rng.standard_normal(d).astype(np.float32) * 0.1,
                           (rng.random(d).astype(np.float32) + 0.5)])
    return mdl


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def test_sitw_stack_fp32_vs_oracle(ktf):
    # Full 512-wide SITW TDNN with seeded random weights (Kaldi final.raw not vendored):
    # x-vector (tdnn6.affine) cosine >= 0.9999 vs the oracle (north star), here at fp32.
    g = load_golden("tdnn.npz")
    x = np.stack([g["sitw_chunk_mfcc"][0], g["sitw_chunk_mfcc"][0][::-1]]).astype(np.float32)
    mdl = sitw_model(ktf, precision="f32")
    got = mdl(x)
    want = O.sequential(x, sitw_layers_for_oracle(mdl))
    assert got.shape == want.shape == (2, 1, 512)
    for b in range(2):
        assert cosine(got[b], want[b]) >= 0.99999
    assert np.max(np.abs(got - want)) < 1e-3


def _bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("D,U,ctx", [(512, 512, [-2, 0, 2]), (512, 1500, [0]), (30, 512, [-2, -1, 0, 1, 2]),
                                     (64, 40, [-3, 0, 3]), (3000, 512, [0])])
def test_tc_layer_vs_exact_bf16_product(ktf, D, U, ctx):
    # tcgen05 engine, single layer: compare with an fp32 evaluation of the SAME bf16-rounded operands
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3, 157, D)).astype(np.float32)
    l = ktf.layers.TDNN(U, context=ctx, precision="bf16", seed=3)
    got = l(x)
    kernel, bias = l.get_weights()
    want = O.tdnn(_bf16_round(x), _bf16_round(kernel), bias, ctx)
    assert got.shape == want.shape
    err = np.max(np.abs(got - want))
    assert err < 2e-3 * max(1.0, float(np.max(np.abs(want)))), err
    # and it is close to the full-precision layer (bf16 operand rounding only)
    full = O.tdnn(x, kernel, bias, ctx)
    assert np.max(np.abs(got - full)) < 0.05 * max(1.0, float(np.max(np.abs(full))))


def test_sitw_stack_bf16_vs_oracle(ktf):
    # north star: x-vector cosine >= 0.9999 at bf16 TDNN precision (operands bf16, fp32 accumulation in
    # TMEM, bf16 inter-layer activations), full SITW widths, seeded random weights
    g = load_golden("tdnn.npz")
    base = g["sitw_chunk_mfcc"][0].astype(np.float32)
    x = [base, base[::-1].copy(), base[:97].copy(), np.concatenate([base, base[::2]])]
    mdl = sitw_model(ktf, precision="bf16")
    layers = sitw_layers_for_oracle(mdl)
    for xi in x:
        got = mdl(xi[None])
        want = O.sequential(xi[None], layers)
        assert got.shape == want.shape == (1, 1, 512)
        assert cosine(got, want) >= 0.9999, cosine(got, want)
    # ragged batch through the stack equals the per-utterance runs
    import torch
    flat = torch.from_numpy(np.concatenate(x)).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum([len(v) for v in x])]), dtype=torch.int64, device="cuda")
    y, _ = mdl.forward_ragged(flat, offs)
    y = y.cpu().numpy()
    for b, xi in enumerate(x):
        assert cosine(y[b], mdl(xi[None])) > 0.99999


def test_xvector_extractor_bf16_vs_oracle(ktf):
    wav = read_wav_int16(golden_path("librispeech_2.wav"))
    cfg = extractor_cfg()
    ext = ktf.models.XvectorExtractor(cfg, precision="bf16", seed=0, allow_random_init=True)
    got = ext(wav)
    want = O.xvector_extractor(wav, cfg, sitw_layers_for_oracle(ext.xvec), ext.xvecGlobalMean, ext.ldaTransform)
    assert cosine(got, want) >= 0.9999, cosine(got, want)


def extractor_cfg():
    import os, yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "data", "tflite_models", "0008_sitw_v2_1a.yml")) as f:
        cfg = yaml.safe_load(f)["extractor"]
    cfg["mfcc"]["dither"] = 0.0
    cfg["xvec"]["model_config_path"] = os.path.join(root, cfg["xvec"]["model_config_path"])
    cfg["xvec"]["model_path"] = None
    cfg["xvec"]["global_mean_path"] = golden_path("sitw_mean.vec")
    cfg["xvec"]["lda_matrix_path"] = golden_path("sitw_transform.mat")
    return cfg


def test_xvector_extractor_cfg1_vs_oracle(ktf):
    # BASELINE config 1: librispeech_2.wav, batch 1, dither 0, real LDA/mean, random TDNN (seed 0)
    wav = read_wav_int16(golden_path("librispeech_2.wav"))
    cfg = extractor_cfg()
    ext = ktf.models.XvectorExtractor(cfg, precision="f32", seed=0, allow_random_init=True)
    got, inter = ext(wav, return_intermediate=True)
    assert got.shape == (128,)
    want, ointer = O.xvector_extractor(wav, cfg, sitw_layers_for_oracle(ext.xvec),
                                       ext.xvecGlobalMean, ext.ldaTransform, return_intermediate=True)
    mfcc = inter["mfcc"].cpu().numpy()
    assert mfcc.shape == (2246, 30)
    truth = O.mfcc(O.framing(wav[None], 25, 10, 16000), precise=True, **cfg["mfcc"])[0]
    assert np.max(np.abs(mfcc - truth)) < 2e-3            # see tests/test_gpu_frontend.py on tolerances
    assert np.quantile(np.abs(mfcc - truth), 0.9999) < 1e-3
    assert np.max(np.abs(mfcc - ointer["mfcc"][0])) < 2.5e-3
    mask = inter["mask"].cpu().numpy()
    omask = np.zeros(2246, np.float32)
    omask[ointer["vad_idx"][:, 1]] = 1
    assert np.array_equal(mask, omask)                      # VAD mask bit-exact
    assert int(mask.sum()) == 1813
    assert cosine(got, want) >= 0.9999
    assert abs(np.linalg.norm(got) - np.sqrt(128)) < 1e-3


def test_xvector_extractor_ragged_batch(ktf):
    wav = read_wav_int16(golden_path("librispeech_2.wav"))
    parts = [wav[:80000], wav[80000:200000], wav[150000:], wav[:48000]]
    ext = ktf.models.XvectorExtractor(extractor_cfg(), precision="f32", seed=0, allow_random_init=True)
    batch = ext(parts)
    assert batch.shape == (4, 128)
    for b, p in enumerate(parts):
        solo = ext(p)
        assert cosine(solo, batch[b]) > 0.999999
    uniform = ext(np.stack([wav[:48000], wav[48000:96000]]))
    assert uniform.shape == (2, 128)
    assert cosine(uniform[0], batch[3]) > 0.999999


def test_xvector_extractor_from_config(ktf):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    yml = os.path.join(root, "data", "tflite_models", "0008_sitw_v2_1a.yml")
    mean = os.path.join(root, "data/kaldi_models/0008_sitw_v2_1a/exp/xvector_nnet_1a/"
                              "xvectors_train_combined_200k/mean.vec")
    if not os.path.exists(mean):
        pytest.skip("symlinked SITW back-end files did not travel")
    cwd = os.getcwd()
    os.chdir(root)
    try:
        with pytest.raises(FileNotFoundError):            # final.raw is not vendored: the drop-in call must not
            ktf.models.XvectorExtractorFromConfig(yml)      # silently return random-weight x-vectors
        with pytest.warns(UserWarning):
            ext = ktf.models.XvectorExtractorFromConfig(yml, precision="f32", allow_random_init=True)
        assert ext.randomInit
    finally:
        os.chdir(cwd)
    assert ext.mfcc.windowing.dither == 1.0                 # YAML default, like the reference
    wav = read_wav_int16(golden_path("librispeech_2.wav"))
    a = ext(wav)
    ref = ktf.models.XvectorExtractor(extractor_cfg(), precision="f32", seed=0, allow_random_init=True)(wav)
    # reference's own e2e tolerance with dither on (xvector_extractor_test.py:30)
    assert 1.0 - cosine(a, ref) <= 0.075


def test_stream_batches_equals_one_shot(ktf):
    # host batch processed in chunks with overlapped copies == the one-shot call (utterances are independent)
    import torch
    from kaldi_tflite_b200 import parallel
    rng = np.random.default_rng(9)
    host = torch.from_numpy((rng.standard_normal((10, 16000)) * 3000).astype(np.float32)).pin_memory()
    fr = ktf.layers.Framing(dynamic_input_shape=True)
    mf = ktf.layers.MFCC(num_mfccs=30, num_mels=30)
    fn = lambda x: mf(fr(x))
    want = fn(host.cuda())
    got = parallel.stream_batches(fn, host, 3)
    assert torch.equal(got, want)
    out = torch.empty(tuple(want.shape), dtype=torch.float32).pin_memory()
    parallel.stream_batches(fn, host, 4, out)
    torch.cuda.synchronize()
    assert torch.equal(out, want.cpu())


def test_plda_tensor_core_path_large_vs_oracle(ktf):
    # fp16 hi/lo split GEMM on the tcgen05 engine: ragged sizes (not multiples of the 128 x 256 tile),
    # enroll != test, |delta| <= 1e-3 * max(|s|, 1) against the float64 oracle
    import torch
    dim, nt, ne = 128, 1037, 700
    mean, Tm, psi = synthetic_plda(dim)
    rng = np.random.default_rng(21)
    x = rng.standard_normal((nt + ne, dim))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True) * np.sqrt(dim)).astype(np.float32)
    layer = ktf.layers.PLDA(dim, mean, Tm, psi, dtype=np.float32, return_transformed=False)
    u = layer.transformVector(torch.from_numpy(x).cuda())
    got = layer.logLikelihoodRatio(u[:nt], u[nt:]).cpu().numpy()
    uo = O.plda_transform(x, mean, Tm, psi, dtype=np.float64)
    want = O.plda_llr(uo, psi)[:nt, nt:]
    assert got.shape == want.shape == (nt, ne)
    assert np.all(np.abs(got - want) <= 1e-3 * np.maximum(np.abs(want), 1.0))
    assert np.max(np.abs(got - want)) < 2e-4          # fp32-equivalent: far inside the gate
