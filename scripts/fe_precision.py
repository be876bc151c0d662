"""MFCC error of the front-end kernel on librispeech_2.wav against the float64 evaluation of the oracle (scratch tool)."""
import os, sys, wave
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kaldi_tflite_b200 as ktf
from oracle import ktf_oracle as O

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with wave.open(os.path.join(root, "tests", "golden", "librispeech_2.wav"), "rb") as w:
    wav = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32)
kw = dict(num_mfccs=30, num_mels=30, low_freq_cutoff=20.0, high_freq_cutoff=7600.0)
fr = ktf.layers.Framing(dynamic_input_shape=True)
mf = ktf.layers.MFCC(**kw)
got = mf(fr(wav[None]))
got = (got.cpu().numpy() if hasattr(got, 'cpu') else np.asarray(got))[0]
frames = O.framing(wav[None], 25, 10, 16000)
truth = O.mfcc(frames, precise=True, **kw)[0]
f32 = O.mfcc(frames, **kw)[0]
for name, x in (("kernel", got), ("f32 oracle", f32)):
    d = np.abs(x - truth)
    print(f"{name:10s} max {d.max():.3e} p99.99 {np.quantile(d, 0.9999):.3e} p99.9 {np.quantile(d, 0.999):.3e} rms {np.sqrt((d**2).mean()):.3e}")
