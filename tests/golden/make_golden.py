#!/usr/bin/env python
"""
Generates the golden fixtures under tests/golden/ from the reference's own test
data (/root/reference/kaldi_tflite/lib/testdata).  Those files are OUTPUTS OF
REAL KALDI BINARIES (compute-mfcc-feats, compute-fbank-feats,
apply-cmvn-sliding, compute-vad, nnet3-compute, ivector-plda-scoring; see the
generator scripts cited in SURVEY.md section 4) plus the Kaldi-format binary
model files the readers are tested on.  /root/reference does not exist on the
GPU box, so the vectors are re-packed here as compact .npz archives and the
binary model files are carried over byte-for-byte (they are data, not source).

Run from the repo root (only needed when the reference's fixtures change):

    python tests/golden/make_golden.py

No TensorFlow is needed: the reference's fixture modules that hold literals
(plda_model.py, plda_scores.py, xvectors.py, tdnn_narrow.py,
tdnn_single_layer.py) are numpy-only once `kaldi_tflite` / `kaldi_tflite.lib`
are registered as namespace stubs so that the TF-importing `__init__`s never run.
"""

import importlib
import json
import os
import shutil
import sys
import types
import wave

import numpy as np

REF = "/root/reference"
TD = os.path.join(REF, "kaldi_tflite", "lib", "testdata")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_ark(path, sep=None):
    """Kaldi text archive -> {utt_id: 2-D float32 array}."""
    ark, cur, rows = {}, None, []
    with open(path) as f:
        for line in f:
            toks = line.replace(",", " ").split() if sep is None else line.split(sep)
            if not toks:
                continue
            if "[" in toks and "]" in toks:
                if len(toks) > 3:
                    ark[toks[0]] = np.array([[float(t) for t in toks[2:-1]]], dtype=np.float32)
                continue
            if "[" in toks:
                cur, rows = toks[0], []
                continue
            last = "]" in toks
            vals = [float(t) for t in toks if t != "]"]
            if vals:
                rows.append(vals)
            if last:
                ark[cur] = np.array(rows, dtype=np.float32)
                cur, rows = None, []
    return ark


def stack_ark(path):
    return np.stack(list(load_ark(path).values()), axis=0)


def read_wav_int16(path):
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1
        sr = w.getframerate()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    return sr, data.copy()


def read_conf(path):
    conf = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            k, v = line.split("=")
            conf[k.lstrip("-")] = v
    return conf


def stub_reference_namespaces():
    for name, sub in (("kaldi_tflite", ["kaldi_tflite"]),
                      ("kaldi_tflite.lib", ["kaldi_tflite", "lib"]),
                      ("kaldi_tflite.lib.testdata", ["kaldi_tflite", "lib", "testdata"]),
                      ("kaldi_tflite.lib.testdata.tdnn", ["kaldi_tflite", "lib", "testdata", "tdnn"]),
                      ("kaldi_tflite.lib.testdata.plda", ["kaldi_tflite", "lib", "testdata", "plda"]),
                      ("kaldi_tflite.lib.testdata.xvectors", ["kaldi_tflite", "lib", "testdata", "xvectors"])):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *sub)]
        sys.modules[name] = m


def main():
    os.makedirs(OUT, exist_ok=True)

    # ---------------- front-end: MFCC x54, fbank x48 -------------------------
    fe = {}
    src = os.path.join(TD, "feats", "src", "fbank_mfcc")
    sr, wav = read_wav_int16(os.path.join(TD, "librispeech_2_trimmed.wav"))
    assert sr == 16000
    fe["wav_trimmed"] = wav
    names = sorted(os.listdir(src))
    for name in names:
        d = os.path.join(src, name)
        sr_i, wav_i = read_wav_int16(os.path.realpath(os.path.join(d, "audio.wav")))
        assert sr_i == sr and np.array_equal(wav_i, wav), name
        idx = name.split("_")[1]
        if os.path.exists(os.path.join(d, "mfcc.ark.txt")):
            fe[f"mfcc_{idx}"] = stack_ark(os.path.join(d, "mfcc.ark.txt"))
            fe[f"mfcc_conf_{idx}"] = json.dumps(read_conf(os.path.join(d, "mfcc.conf")))
        if os.path.exists(os.path.join(d, "fbank.ark.txt")):
            fe[f"fbank_{idx}"] = stack_ark(os.path.join(d, "fbank.ark.txt"))
            fe[f"fbank_conf_{idx}"] = json.dumps(read_conf(os.path.join(d, "fbank.conf")))
    np.savez_compressed(os.path.join(OUT, "frontend.npz"), **fe)

    # ---------------- CMVN x8, VAD x46 -----------------------------------------
    for kind, infile, outfile in (("cmvn", "mfcc.ark.txt", "cmvn.ark.txt"),
                                  ("vad", "mfcc.ark.txt", "vad.ark.txt")):
        pack = {}
        src = os.path.join(TD, "feats", "src", kind)
        for name in sorted(os.listdir(src)):
            d = os.path.join(src, name)
            idx = name.split("_")[2]
            pack[f"in_{idx}"] = stack_ark(os.path.join(d, infile))
            out = stack_ark(os.path.join(d, outfile))
            if kind == "vad":
                out = out.transpose([0, 2, 1])          # (1, T, 1) like RefVAD.getOutputs
            pack[f"out_{idx}"] = out
            pack[f"conf_{idx}"] = json.dumps(read_conf(os.path.join(d, f"{kind}.conf")))
        np.savez_compressed(os.path.join(OUT, f"{kind}.npz"), **pack)

    # ---------------- stats pooling x8 -------------------------------------------
    pack = {}
    src = os.path.join(TD, "stats", "src")
    for name in sorted(os.listdir(src)):
        d = os.path.join(src, name)
        if not os.path.isdir(d):
            continue
        pack[f"in_{name}"] = stack_ark(os.path.join(d, "feat.ark.txt"))
        pack[f"out_{name}"] = stack_ark(os.path.join(d, "output.ark.txt"))
    np.savez_compressed(os.path.join(OUT, "stats.npz"), **pack)

    # ---------------- TDNN single layer + narrow, PLDA (literal modules) --------
    stub_reference_namespaces()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)                        # tdnn_single_layer.py opens a CWD-relative final.raw
    try:
        single = importlib.import_module("kaldi_tflite.lib.testdata.tdnn.tdnn_single_layer").RefTdnnSingleLayer
        narrow = importlib.import_module("kaldi_tflite.lib.testdata.tdnn.tdnn_narrow").RefTdnnNarrow
        pmodel = importlib.import_module("kaldi_tflite.lib.testdata.plda.plda_model").RefPldaModel
        pscores = importlib.import_module("kaldi_tflite.lib.testdata.plda.plda_scores").RefPldaScores
        xvecs = importlib.import_module("kaldi_tflite.lib.testdata.xvectors.xvectors").RefXVectors
    finally:
        os.chdir(cwd)

    pack = {
        "single_cfg": json.dumps(single.cfg),
        "single_in": single.inputs, "single_out": single.outputs,
        "narrow_in": narrow.inputs, "narrow_out": narrow.outputs,
        "narrow_config": json.dumps(list(narrow.config)),
        "narrow_components": json.dumps([
            {k: (v if isinstance(v, (str, int, float, bool)) else None) for k, v in c.items()}
            for c in narrow.components]),
    }
    for i, c in enumerate(narrow.components):
        for k, v in c.items():
            if isinstance(v, np.ndarray):
                pack[f"narrow_c{i}_{k}"] = v
            elif isinstance(v, (np.floating, np.integer)):
                pack[f"narrow_c{i}_{k}"] = np.asarray(v)
    pack["sitw_chunk_mfcc"] = stack_ark(os.path.join(TD, "mfcc_chunk_30_16khz.ark.txt"))
    # Goldens that need the un-vendored final.raw (kept for the day the weights are available).
    pack["sitw_tdnn6_out"] = stack_ark(os.path.join(
        TD, "tdnn", "src", "0008_sitw_v2_1a_tdnn6.affine", "output.ark.txt"))
    pack["sitw_e2e_xvector"] = stack_ark(os.path.join(
        TD, "models", "src", "0008_sitw_v2_1a", "xvector.ark.txt"))
    # CALLHOME diarization model (8 kHz, 23-dim input, models/kaldi/sequential_test.py:78): its input chunk; the Kaldi
    # output needs the un-vendored final.raw as well
    pack["callhome_chunk_mfcc"] = stack_ark(os.path.join(
        TD, "tdnn", "src", "0006_callhome_diarization_v2_1a_tdnn6.affine", "feat.ark.txt"))
    pack["callhome_tdnn6_out"] = stack_ark(os.path.join(
        TD, "tdnn", "src", "0006_callhome_diarization_v2_1a_tdnn6.affine", "output.ark.txt"))
    np.savez_compressed(os.path.join(OUT, "tdnn.npz"), **pack)

    np.savez_compressed(
        os.path.join(OUT, "plda.npz"),
        dim=np.asarray(pmodel.dim), mean=pmodel.mean, transform=pmodel.transformMat, psi=pmodel.psi,
        plda_input=xvecs.pldaInput(), plda_transformed=xvecs.pldaTransformed(withoutPCA=True),
        scores=pscores.scores(withoutPCA=True))

    # ---------------- binary Kaldi files carried byte-for-byte --------------------
    for rel, dst in (
        ("tdnn/src/tdnn_single_layer/final.raw", "tdnn_single_layer.final.raw"),
        ("tdnn/src/tdnn_narrow/final.raw", "tdnn_narrow.final.raw"),
        ("plda/plda", "plda.bin"),
        ("plda/xvectors_train_combined_200k/mean.vec", "sitw_mean.vec"),
        ("plda/xvectors_train_combined_200k/mean.vec.txt", "sitw_mean.vec.txt"),
        ("plda/xvectors_train_combined_200k/transform.mat", "sitw_transform.mat"),
        ("librispeech_2.wav", "librispeech_2.wav"),
    ):
        shutil.copyfile(os.path.join(TD, rel), os.path.join(OUT, dst))

    for f in sorted(os.listdir(OUT)):
        print(f"{os.path.getsize(os.path.join(OUT, f)):>10d}  {f}")


if __name__ == "__main__":
    main()
