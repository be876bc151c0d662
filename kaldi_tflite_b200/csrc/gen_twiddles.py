"""Regenerates twiddles64.h (compile-time FFT twiddles used by frontend.cu)."""
import numpy as np

N = 64
j = np.arange(N)
c = np.cos(2 * np.pi * j / N)
s = np.sin(2 * np.pi * j / N)
c[np.abs(c) < 1e-12] = 0
s[np.abs(s) < 1e-12] = 0


def fmt(a):
    return ", ".join(f"{float(np.float32(v))!r}f" for v in a)


with open("twiddles64.h", "w") as f:
    f.write("// Generated: cos/sin(2*pi*j/64) rounded to float32 (see csrc/gen_twiddles.py).\n")
    f.write("// W_64^j = (KTF_COS64[j], -KTF_SIN64[j]);  W_N^i = W_64^(i*64/N) for N | 64.\n#pragma once\n")
    f.write("static __device__ constexpr float KTF_COS64[64] = {%s};\n" % fmt(c))
    f.write("static __device__ constexpr float KTF_SIN64[64] = {%s};\n" % fmt(s))
