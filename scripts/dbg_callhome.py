import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import kaldi_tflite_b200 as ktf
from oracle import ktf_oracle as O
from test_gpu_round2 import _wav_8k, CALLHOME_MFCC
wav = _wav_8k()[None]
for snip in (True, False):
    fr = ktf.layers.Framing(25.0, 10.0, 8000.0, dynamic_input_shape=True, snip_edges=snip)
    got = ktf.layers.MFCC(**CALLHOME_MFCC)(fr(wav))
    x = wav if snip else O.pad_waveform(wav, 200, 80)
    frames = O.framing(x, 25.0, 10.0, 8000.0)
    truth = O.mfcc(frames, precise=True, **CALLHOME_MFCC)
    ora = O.mfcc(frames, **CALLHOME_MFCC)
    e = np.abs(got - truth)[0]; f = np.abs(ora - truth)[0]
    idx = np.unravel_index(np.argsort(e.ravel())[-8:], e.shape)
    print("snip", snip, "shape", e.shape, "max", e.max(), "oracle max", f.max(), "p999", np.quantile(e, .999), np.quantile(f, .999), "rmse", np.sqrt((e**2).mean()), np.sqrt((f**2).mean()))
    for t, c in zip(*idx):
        print("   frame", t, "ceps", c, "err", e[t, c], "oracle err", f[t, c], "value", truth[0, t, c])
    lm_t = O.filterbank(O.windowing(frames, return_energy=False, precise=True), precise=True, num_bins=23, sample_frequency=8000.0, low_freq_cutoff=20.0, high_freq_cutoff=3700.0)[0]
    t = idx[0][-1]
    print("   logmel of worst frame:", np.round(lm_t[t], 2))
