"""Quick device timing of the front-end and CMVN kernels at BASELINE config 2 (scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kaldi_tflite_b200 as ktf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device="cuda").manual_seed(1234)
wav = (torch.randn((B, 160000), generator=g, device="cuda") * 3000).clamp(-32767, 32767)
fr = ktf.layers.Framing(dynamic_input_shape=True)
mf = ktf.layers.MFCC(num_mfccs=30, num_mels=30)
cm = ktf.layers.CMVN(window=200)
for _ in range(3):
    x = mf(fr(wav)); y = cm(x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fe_t, cm_t = [], []
for _ in range(n):
    ev[0].record(); x = mf(fr(wav)); ev[1].record(); y = cm(x); ev[2].record()
    torch.cuda.synchronize()
    fe_t.append(ev[0].elapsed_time(ev[1])); cm_t.append(ev[1].elapsed_time(ev[2]))
fe_t.sort(); cm_t.sort()
t_fe, t_cm = fe_t[n // 2], cm_t[n // 2]
bytes_fe = B * 160000 * 4 + B * 998 * 30 * 4
print(f"B={B} frontend median {t_fe*1e3:.1f} us min {fe_t[0]*1e3:.1f} max {fe_t[-1]*1e3:.1f} "
      f"({bytes_fe/(t_fe*1e-3)/1e9:.0f} GB/s)  cmvn median {t_cm*1e3:.1f} us")
