"""
Pins the CPU oracle (oracle/ktf_oracle.py) against the reference's own golden
vectors -- outputs of real Kaldi binaries -- at the reference's own tolerances
(each test cites the reference test it mirrors).
"""

import json

import numpy as np
import pytest

from conftest import load_golden, golden_path, rmse, read_wav_int16
import helpers
from oracle import ktf_oracle as O
from kaldi_tflite_b200.io import KaldiNnet3Reader, KaldiPldaReader, ReadKaldiArray


@pytest.fixture(scope="module")
def fe():
    return load_golden("frontend.npz")


def _frames(wav, cfg):
    fr = cfg["framing"]
    size, shift, _ = O.frame_params(fr["frame_length_ms"], fr["frame_shift_ms"], fr["sample_frequency"])
    x = wav.reshape(1, -1)
    if not cfg["snip_edges"]:
        x = O.pad_waveform(x, size, shift)        # mfcc_test.py:189-193
    return O.framing(x, **fr)


def test_framing_matches_strided_view():
    # layers/dsp/framing_test.py:40-73 -- exact equality with Kaldi-style frame extraction.
    x = np.arange(16000 * 10, dtype=np.float32)
    for (length, shift, sr) in [(25, 10, 16000), (32, 16, 16000), (20, 10, 8000),
                                (25, 10, 8000), (30, 15, 44100), (10, 10, 16000)]:
        m = int(sr * length / 1000.0)
        k = int(sr * shift / 1000.0)
        got = O.framing(x[None], length, shift, sr)[0]
        M = 1 + (len(x) - m) // k
        want = np.lib.stride_tricks.sliding_window_view(x, m)[::k][:M]
        assert got.shape[1] == 2 * (m // 2)
        assert np.array_equal(got, want[:, :got.shape[1]][:got.shape[0]])
        assert got.shape[0] in (M, M + 1) if m % 2 else got.shape[0] == M


def test_mfcc_vs_kaldi(fe):
    # layers/dsp/mfcc_test.py:168-203, tolerance :32
    wav = fe["wav_trimmed"].astype(np.float32)
    worst = 0.0
    n = 0
    for key in fe.files:
        if not key.startswith("mfcc_conf_"):
            continue
        idx = key.split("_")[-1]
        cfg = helpers.mfcc_conf_to_kwargs(str(fe[key]))
        got = O.mfcc(_frames(wav, cfg), **cfg["mfcc"])
        want = fe[f"mfcc_{idx}"]
        assert got.shape == want.shape, (idx, got.shape, want.shape)
        e = rmse(want, got)
        worst = max(worst, e)
        assert e < 2.25e-4, (idx, e)
        n += 1
    assert n == 54
    print("mfcc worst rmse", worst)


def test_fbank_vs_kaldi(fe):
    # layers/dsp/filterbank_test.py:157-195, tolerance :32
    wav = fe["wav_trimmed"].astype(np.float32)
    n = 0
    for key in fe.files:
        if not key.startswith("fbank_conf_"):
            continue
        idx = key.split("_")[-1]
        cfg = helpers.fbank_conf_to_kwargs(str(fe[key]))
        w = O.windowing(_frames(wav, cfg), return_energy=False, **cfg["windowing"])
        got = O.filterbank(w, **cfg["fbank"])
        want = fe[f"fbank_{idx}"]
        assert got.shape == want.shape
        assert rmse(want, got) < 2.25e-5, idx
        n += 1
    assert n >= 48      # the reference test walks 48; all 54 fixture dirs carry fbank goldens


def test_cmvn_vs_kaldi():
    # layers/normalization/cmvn_test.py:153-195, tolerance :31 (SAME and VALID)
    g = load_golden("cmvn.npz")
    n = 0
    for key in g.files:
        if not key.startswith("conf_"):
            continue
        idx = key.split("_")[-1]
        kw = helpers.cmvn_conf_to_kwargs(str(g[key]))
        x, want = g[f"in_{idx}"], g[f"out_{idx}"]
        got = O.cmvn(x, padding="SAME", **kw)
        assert rmse(want, got) < 1e-5, idx
        N, T = kw["window"], x.shape[1]
        got_v = O.cmvn(x, padding="VALID", **kw)
        if T > N:
            assert rmse(want[:, N // 2: T - (N - 1) // 2], got_v) < 1e-5, idx
        n += 1
    assert n == 8


def test_vad_vs_kaldi_exact():
    # layers/dsp/vad_test.py:132-152 -- bit-exact
    g = load_golden("vad.npz")
    n = 0
    for key in g.files:
        if not key.startswith("conf_"):
            continue
        idx = key.split("_")[-1]
        kw = helpers.vad_conf_to_kwargs(str(g[key]))
        got = O.vad(g[f"in_{idx}"], **kw)
        assert np.array_equal(got, g[f"out_{idx}"]), idx
        kw["return_indexes"] = True
        ind = O.vad(g[f"in_{idx}"], **kw)
        assert np.array_equal(ind, np.argwhere(g[f"out_{idx}"][..., 0] > 0))
        n += 1
    assert n == 46


def test_tdnn_single_layer_vs_kaldi():
    # layers/tdnn/tdnn_test.py:45-57, tolerance :31
    g = load_golden("tdnn.npz")
    cfg = json.loads(str(g["single_cfg"]))
    r = KaldiNnet3Reader(golden_path("tdnn_single_layer.final.raw"), True)
    W, b = r.components[0]["params"], r.components[0]["bias"]
    kernel = O.kaldi_to_kernel(W, cfg["units"], len(cfg["context"]))
    got = O.tdnn(g["single_in"], kernel, b, cfg["context"], activation=cfg["activation"])
    assert got.shape == g["single_out"].shape
    assert rmse(g["single_out"], got) <= 1e-6


def narrow_layers(reader):
    layers = []
    for name, dim, ctx, use_relu, use_bn in helpers.NARROW_LAYERS:
        W, b = reader.getWeights(f"{name}.affine")
        layers.append({"type": "affine", "kernel": O.kaldi_to_kernel(W, dim, len(ctx)),
                       "bias": b, "context": ctx})
        if use_relu:
            layers.append({"type": "relu"})
        if use_bn:
            rms, mean, var = reader.getWeights(f"{name}.batchnorm")
            layers.append({"type": "batchnorm", "gamma": rms * np.ones_like(mean),
                           "mean": mean, "var": var})
    return layers


def test_tdnn_narrow_vs_kaldi():
    # layers/tdnn/tdnn_test.py:105-119, tolerance :117
    g = load_golden("tdnn.npz")
    r = KaldiNnet3Reader(golden_path("tdnn_narrow.final.raw"), True)
    got = O.sequential(g["narrow_in"], narrow_layers(r))
    assert got.shape == g["narrow_out"].shape
    assert rmse(g["narrow_out"], got) <= 5e-4


def test_stats_pooling_vs_kaldi():
    # layers/stats/stats_pooling_test.py:48-88, tolerance :26
    g = load_golden("stats.npz")
    cfg = helpers.stats_default_cfg()
    cfg["reduce_time_axis"] = True
    got = O.stats_pooling(g["in_stats_mean_std"], **cfg)
    assert rmse(g["out_stats_mean_std"][:, 0:1, :], got) <= 4e-6
    for name, over in helpers.STATS_CONFIGS.items():
        cfg = helpers.stats_default_cfg()
        cfg.update(over)
        got = O.stats_pooling(g[f"in_{name}"], **cfg)
        want = g[f"out_{name}"]
        assert got.shape == want.shape, name
        assert rmse(want, got) <= 4e-6, name


def test_plda_vs_kaldi():
    # layers/plda/plda_test.py:45-62, tolerance :30 (float32 parameters)
    g = load_golden("plda.npz")
    for dt, tol in ((np.float32, 2e-4), (np.float64, 2e-4)):
        scores, transformed = O.plda(g["plda_input"], g["mean"], g["transform"], g["psi"], dtype=dt)
        assert transformed.shape == g["plda_transformed"].shape
        assert scores.shape == g["scores"].shape
        assert rmse(g["plda_transformed"], transformed) <= tol
        assert rmse(g["scores"], scores) <= tol


def test_plda_gemm_form_equals_direct_form():
    # score = A_i + B_j + sum_d u_i c_d u_j  (what the GPU kernel evaluates)
    g = load_golden("plda.npz")
    psi = g["psi"].astype(np.float64)
    u = O.plda_transform(g["plda_input"][:, 0, :], g["mean"], g["transform"], psi)
    direct = O.plda_llr(u, psi)
    v1 = 1.0 + psi / (psi + 1.0)
    v0 = 1.0 + psi
    r = psi / (psi + 1.0)
    c = r / v1
    A = -0.5 * np.sum(u * u / v1, 1) + 0.5 * np.sum(u * u / v0, 1) - 0.5 * (np.sum(np.log(v1)) - np.sum(np.log(v0)))
    Bj = -0.5 * np.sum((r * u) ** 2 / v1, 1)
    gemm = A[:, None] + Bj[None, :] + (u * c) @ u.T
    assert np.max(np.abs(gemm - direct)) < 1e-9


def test_lda_backend_shapes():
    # models/kaldi/xvector_extractor.py:125-134,174-181 on the real SITW mean.vec / transform.mat
    mean_b = ReadKaldiArray(golden_path("sitw_mean.vec"), True)
    mean_t = ReadKaldiArray(golden_path("sitw_mean.vec.txt"), False)
    assert mean_b.shape == (512,) and rmse(mean_b, mean_t) < 5e-8      # io/kaldi/array_reader_test.py
    mat = ReadKaldiArray(golden_path("sitw_transform.mat"), True)
    assert mat.shape == (128, 513)
    x = np.random.default_rng(0).standard_normal((3, 1, 512)).astype(np.float32)
    y = O.lda_length_norm(x, mean_b, mat)
    assert y.shape == (3, 1, 128)
    assert np.allclose(np.linalg.norm(y, axis=-1), np.sqrt(128.0), rtol=1e-5)


def test_xvector_extractor_cfg1_shapes():
    # librispeech_2.wav -> 2246 frames, 1813 voiced (dither 0)
    wav = read_wav_int16(golden_path("librispeech_2.wav"))
    assert wav.shape == (359665,)
    fr = O.framing(wav[None], 25, 10, 16000)
    assert fr.shape == (1, 2246, 400)
    feats = O.mfcc(fr, num_mfccs=30, num_mels=30, sample_frequency=16000.0,
                   high_freq_cutoff=7600.0, low_freq_cutoff=20.0, dither=0.0)
    idx = O.vad(feats, 0.5, 5.5, 2, 0.12, True, 0)
    assert idx.shape == (1813, 2)
