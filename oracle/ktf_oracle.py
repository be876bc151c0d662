"""
CPU oracle for the wav -> x-vector hot path of shahruk10/kaldi-tflite.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product path (`kaldi_tflite_b200`) never does;
it fails loudly when the CUDA library is missing.

It is an op-for-op float32 NumPy restatement of the reference layers' `call()`
bodies (the reference's arithmetic lives in tensorflow==2.8.0, which is not in
/root/reference and not installable here).  Every
function cites the reference file:line it follows (paths relative to
/root/reference/kaldi_tflite/lib/).

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
here against the reference's own golden vectors (real Kaldi binaries' outputs,
committed under tests/golden/ by tests/golden/make_golden.py) at the
reference's own tolerances: MFCC x54, fbank x48, CMVN x8 (SAME and VALID),
VAD x46 (bit-exact), TDNN single/narrow, StatsPooling x8, PLDA 29x29.
Not pinned against Kaldi (weights not vendored, need network): the full
512-wide SITW stack -- there the GPU path is compared with this oracle on
seeded random weights.
"""

from __future__ import annotations

import numpy as np
import scipy.fft

F32 = np.float32


# --------------------------------------------------------------------------
# layers/dsp/framing.py
# --------------------------------------------------------------------------

def frame_params(frame_length_ms=25.0, frame_shift_ms=10.0, sample_frequency=16000.0):
    """framing.py:96-105 -- frame size / shift in samples and the half size."""
    if frame_length_ms <= 0 or frame_shift_ms <= 0 or sample_frequency <= 0:
        raise ValueError("frame_length, frame_shift and sample_frequency should be > 0")
    size = int(sample_frequency * frame_length_ms / 1000.0)
    shift = int(sample_frequency * frame_shift_ms / 1000.0)
    if size <= 0 or shift <= 0:
        raise ValueError("frame_length / frame_shift too small")
    return size, shift, size // 2


def frame_indexes(num_samples, size, shift):
    """framing.py:212-241 -- centres range(half, N-half+1, shift), offsets [-half, half)."""
    half = size // 2
    centres = np.arange(half, num_samples - half + 1, shift)
    offsets = np.arange(-half, half)
    return centres[:, None] + offsets[None, :]


def framing(wav, frame_length_ms=25.0, frame_shift_ms=10.0, sample_frequency=16000.0):
    """framing.py:243-265 -- (B, N) -> (B, T, 2*half); no edge padding."""
    wav = np.asarray(wav)
    size, shift, _ = frame_params(frame_length_ms, frame_shift_ms, sample_frequency)
    if wav.shape[-1] < size:
        raise ValueError(f"input sample size must be >= frame size ({size})")
    idx = frame_indexes(wav.shape[-1], size, shift)
    return wav[..., idx]


# --------------------------------------------------------------------------
# layers/dsp/windowing.py
# --------------------------------------------------------------------------

def window_function(window_type, M, blackman_coeff=0.42):
    """windowing.py:130-156 -- computed in float64, cast to float32."""
    n = np.arange(0, M)
    if M == 1:
        w = np.ones(1, float)
    elif window_type == "hamming":
        w = np.hamming(M)
    elif window_type == "hanning":
        w = np.hanning(M)
    elif window_type == "povey":
        w = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / (M - 1))) ** 0.85
    elif window_type == "rectangular":
        w = np.ones((M,))
    elif window_type == "sine":
        w = np.sin(np.pi * n / (M - 1))
    elif window_type == "blackman":
        w = np.blackman(M)
        if blackman_coeff != 0.42:
            w = w - 0.42 + blackman_coeff
    else:
        raise ValueError(f"window_type '{window_type}' is not recognized")
    return w.astype(F32)


def log_energy(x, energy_floor=0.0, epsilon=1e-7):
    """windowing.py:174-178 -- clip(log(relu(sum x^2) + eps), floor, max)."""
    dt = x.dtype.type
    e = np.sum(x * x, axis=-1, keepdims=True, dtype=dt)
    e = np.log(np.maximum(e, dt(0)) + dt(np.float32(epsilon))).astype(dt)
    return np.clip(e, dt(energy_floor), np.finfo(F32).max).astype(dt)


def windowing(frames, window_type="povey", blackman_coeff=0.42, dither=0.0,
              remove_dc_offset=True, preemphasis_coefficient=0.97,
              return_energy=True, raw_energy=True, energy_floor=0.0, epsilon=1e-7,
              rng=None, precise=False):
    """windowing.py:180-209.  `precise=True` evaluates the same formulas in float64 (the
    constant tables stay the float32 ones the reference stores): the yardstick that tells
    float32 rounding noise apart from real differences."""
    F32 = np.float64 if precise else np.float32
    x = np.asarray(frames, dtype=F32)
    if dither != 0.0:
        rng = rng or np.random.default_rng()
        x = x + rng.standard_normal(x.shape).astype(F32) * F32(dither)
    if remove_dc_offset:
        x = x - np.mean(x, axis=-1, keepdims=True, dtype=F32)
    energy = None
    if return_energy and raw_energy:
        energy = log_energy(x, energy_floor, epsilon)
    if preemphasis_coefficient > 0:
        c = F32(np.float32(preemphasis_coefficient))      # TF rounds the python scalar to float32
        first = x[..., :1] - c * x[..., :1]
        rest = x[..., 1:] - c * x[..., :-1]
        x = np.concatenate([first, rest], axis=-1)
    x = x * window_function(window_type, x.shape[-1], blackman_coeff).astype(F32)
    if return_energy:
        if not raw_energy:
            energy = log_energy(x, energy_floor, epsilon)
        return x.astype(F32), energy
    return x.astype(F32)


# --------------------------------------------------------------------------
# layers/dsp/filterbank.py
# --------------------------------------------------------------------------

def next_pow2(n):
    """filterbank.py:133-136."""
    if (n & (n - 1) == 0) and n != 0:
        return n
    return 2 ** (n - 1).bit_length()


def mel_scale(freq):
    """filterbank.py:138-139."""
    return 1127.0 * np.log(1.0 + freq / 700.0)


def mel_bank(window_size, num_bins=23, sample_frequency=16000.0,
             high_freq_cutoff=0.0, low_freq_cutoff=20.0):
    """filterbank.py:85-100 (cut-off validation) and :141-189 (triangles).

    Returns (fft_length, bank) with bank float32 of shape (fft_length/2+1, num_bins).
    Bin fft_length/2 (Nyquist) and bin 0 always carry zero weight.
    """
    nyquist = sample_frequency / 2.0
    if sample_frequency <= 0:
        raise ValueError("sample_frequency must be > 0")
    if low_freq_cutoff > nyquist or low_freq_cutoff < 0:
        raise ValueError("low_freq_cutoff must be > 0 and < Nyquist")
    upper = high_freq_cutoff
    if upper <= 0:
        upper += nyquist
    if low_freq_cutoff >= upper:
        raise ValueError("lower_freq_cutoff must be < higher_freq_cutoff")

    fft_length = next_pow2(window_size)
    fft_bins = fft_length // 2
    bin_width = sample_frequency / fft_length
    mel_low = mel_scale(low_freq_cutoff)
    mel_high = mel_scale(upper)
    mel_delta = (mel_high - mel_low) / (num_bins + 1)

    bank = np.zeros([num_bins, fft_bins + 1], dtype=F32)
    mels = mel_scale(bin_width * np.arange(fft_bins))
    for i in range(num_bins):
        left = mel_low + (i * mel_delta)
        center = left + mel_delta
        right = center + mel_delta
        for j in range(fft_bins):
            mel = mels[j]
            if left < mel < right:
                if mel <= center:
                    bank[i, j] = (mel - left) / (center - left)
                else:
                    bank[i, j] = (right - mel) / (right - center)
    return fft_length, np.ascontiguousarray(bank.T)


def filterbank(frames, num_bins=23, sample_frequency=16000.0, high_freq_cutoff=0.0,
               low_freq_cutoff=20.0, use_log_fbank=True, use_power=True, epsilon=1e-7,
               precise=False):
    """filterbank.py:225-242 -- pad -> rfft -> abs -> pow2 -> @melBank -> log(relu+eps)."""
    F32 = np.float64 if precise else np.float32
    x = np.asarray(frames, dtype=F32)
    fft_length, bank = mel_bank(x.shape[-1], num_bins, sample_frequency,
                                high_freq_cutoff, low_freq_cutoff)
    spec = scipy.fft.rfft(x, n=fft_length, axis=-1)          # complex64
    spec = np.abs(spec).astype(F32)
    if use_power:
        spec = spec * spec
    feats = np.matmul(spec, bank.astype(F32)).astype(F32)
    if use_log_fbank:
        feats = np.log(np.maximum(feats, F32(0)) + F32(np.float32(epsilon))).astype(F32)
    return feats


# --------------------------------------------------------------------------
# layers/dsp/dct.py, layers/dsp/mfcc.py
# --------------------------------------------------------------------------

def dct_matrix(input_length, length):
    """dct.py:98-143 -- ortho DCT-II (N x K); column 0 overwritten with sqrt(1/N)."""
    if length <= 0:
        raise ValueError("DCT length must be > 0")
    if input_length < length:
        raise ValueError("input feature length must be >= DCT length")
    N = float(input_length)
    n = np.arange(input_length)
    k = np.expand_dims(np.arange(length, dtype=np.float64), 1)
    dct = np.cos((np.pi / N) * (n + 0.5) * k)
    dct[0] *= 1.0 / np.sqrt(2.0)
    dct *= np.sqrt(2.0 / N)
    dct = dct.T
    dct[:, 0] = np.sqrt(1.0 / N)
    return dct.astype(F32)


def lifter_coeffs(num_mfccs, cepstral_lifter):
    """mfcc.py:146-159."""
    n = np.arange(0, num_mfccs)
    q = cepstral_lifter
    return (1 + 0.5 * np.sin(np.pi * n / q) * q).astype(F32)


def mfcc(frames, num_mfccs=23, num_mels=23, cepstral_lifter=22, use_energy=True,
         sample_frequency=16000.0, high_freq_cutoff=0.0, low_freq_cutoff=20.0,
         use_log_fbank=True, use_power=True, window_type="povey", dither=0.0,
         remove_dc_offset=True, preemphasis_coefficient=0.97, raw_energy=True,
         energy_floor=0.0, epsilon=1e-7, precise=False):
    """mfcc.py:197-244 -- windowing -> filterbank -> DCT -> lifter -> C0 <- log-energy."""
    if num_mfccs > num_mels:
        raise ValueError("num_mfccs must be <= num_mels")
    F32 = np.float64 if precise else np.float32
    w = windowing(frames, window_type=window_type, dither=dither,
                  remove_dc_offset=remove_dc_offset,
                  preemphasis_coefficient=preemphasis_coefficient,
                  return_energy=use_energy, raw_energy=raw_energy,
                  energy_floor=energy_floor, epsilon=epsilon, precise=precise)
    if use_energy:
        w, energy = w
    fb = filterbank(w, num_bins=num_mels, sample_frequency=sample_frequency,
                    high_freq_cutoff=high_freq_cutoff, low_freq_cutoff=low_freq_cutoff,
                    use_log_fbank=use_log_fbank, use_power=use_power, epsilon=epsilon,
                    precise=precise)
    out = np.matmul(fb, dct_matrix(num_mels, num_mfccs).astype(F32)).astype(F32)
    if cepstral_lifter > 1 and num_mfccs > 1:
        out = out * lifter_coeffs(num_mfccs, cepstral_lifter).astype(F32)
    if use_energy:
        out = out.copy()
        out[..., 0] = energy[..., 0]
    return out.astype(F32)


# --------------------------------------------------------------------------
# layers/dsp/vad.py
# --------------------------------------------------------------------------

def vad(feats, energy_mean_scale=0.5, energy_threshold=5.0, frames_context=0,
        proportion_threshold=0.6, return_indexes=True, energy_coeff=0):
    """vad.py:156-203.  feats (B, T, D) -> float mask (B, T, 1) or int64 (n, 2) indexes."""
    if energy_mean_scale < 0:
        raise ValueError("`energy_mean_scale` must be >= 0")
    if frames_context < 0:
        raise ValueError("`frames_context` must be >= 0")
    if proportion_threshold <= 0 or proportion_threshold >= 1:
        raise ValueError("`proportion_threshold` must be between 0 and 1 (exlcusive)")
    x = np.asarray(feats, dtype=F32)
    e = x[..., energy_coeff:energy_coeff + 1]
    T = e.shape[-2]
    thr = F32(energy_threshold)
    if energy_mean_scale > 0:
        thr = thr + F32(energy_mean_scale) * np.mean(e, axis=-2, keepdims=True, dtype=F32)
    dec = e > thr
    c = frames_context
    if c > 0:
        d = dec.astype(F32)
        pad = np.pad(d, [(0, 0)] * (d.ndim - 2) + [(c, c), (0, 0)])
        counts = np.zeros_like(d)
        for k in range(2 * c + 1):                       # conv1d with ones kernel, SAME
            counts += pad[..., k:k + T, :]
        N = 2 * c + 1
        sizes = np.full((T,), N, dtype=F32)
        edge_sizes = list(range(N // 2 + 1, N, 1)) + list(range(N - 1, N // 2, -1))
        edge_idx = list(range(0, N // 2)) + list(range(-N // 2 + 1, 0))
        for i, s in zip(edge_idx, edge_sizes):           # scatter update, later wins
            sizes[(i + T) % T] = s
        prop = (counts / sizes.reshape((1,) * (d.ndim - 2) + (T, 1))).astype(F32)
        dec = prop >= F32(proportion_threshold)
    if return_indexes:
        return np.argwhere(dec[..., 0]).astype(np.int64)
    return dec.astype(F32)


# --------------------------------------------------------------------------
# layers/normalization/cmvn.py
# --------------------------------------------------------------------------

def _windowed_sums(x_padded, N, padding):
    """cmvn.py:146-184 -- cumsum difference, edges repeated for SAME."""
    cs = np.cumsum(x_padded, axis=-2, dtype=F32)
    s = cs[..., N:, :] - cs[..., :-N, :]
    if padding == "SAME":
        s = np.concatenate([np.repeat(s[..., :1, :], N // 2, axis=-2), s,
                            np.repeat(s[..., -1:, :], (N - 1) // 2, axis=-2)], axis=-2)
    return s


def cmvn(feats, center=True, norm_vars=False, window=600, min_window=100, padding="SAME"):
    """cmvn.py:186-250."""
    if not center:
        raise NotImplementedError("CMVN with center=False not supported yet")
    if window <= 0 or min_window <= 0:
        raise ValueError("`window` and `min_window` must be > 0")
    padding = padding.upper()
    if padding not in ("SAME", "VALID"):
        raise ValueError("`padding` should be either 'SAME' or 'VALID'")
    x = np.asarray(feats, dtype=F32)
    T = x.shape[-2]
    N = window
    std = None
    if T <= N:
        mean = np.sum(x, axis=-2, keepdims=True, dtype=F32) / F32(T)
        if norm_vars:
            x2 = np.sum(x * x, axis=-2, keepdims=True, dtype=F32) / F32(T)
            std = np.sqrt(x2 - mean * mean)
    else:
        xp = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(1, 0), (0, 0)])
        mean = _windowed_sums(xp, N, padding) / F32(N)
        if norm_vars:
            x2 = _windowed_sums(xp * xp, N, padding) / F32(N)
            std = np.sqrt(x2 - mean * mean)
    if padding == "VALID":
        a = N // 2
        b = T - (N - 1) // 2
        x = x[..., a:b, :]
    out = x - mean
    if norm_vars:
        out = out / std
    return out.astype(F32)


# --------------------------------------------------------------------------
# layers/tdnn/tdnn.py, layers/tdnn/utils.py, layers/normalization/batchnorm.py
# --------------------------------------------------------------------------

def kaldi_to_kernel(weights, units, kernel_width):
    """utils.py:22-28 -- Kaldi (U, K*D) -> (1, K, D, U): kernel[0,k,d,u] = W[u, k*D+d]."""
    w = np.asarray(weights)
    return w.flatten().reshape((1, -1, kernel_width, units), order="F").transpose([0, 2, 1, 3])


def tdnn_indices(T, context, subsampling_factor=1, padding="SAME"):
    """tdnn.py:224-249."""
    context = sorted(context)
    start, end = 0, T
    if padding == "VALID":
        if context[0] < 0:
            start = -context[0]
        if context[-1] > 0:
            end = T - context[-1]
    t = np.arange(start, end, subsampling_factor)
    idx = t[:, None] + np.asarray(context)[None, :]
    if padding == "SAME":
        idx = np.clip(idx, 0, T - 1)
    return idx


def tdnn(x, kernel, bias=None, context=(0,), subsampling_factor=1, padding="SAME",
         activation=None):
    """tdnn.py:251-280 -- gather (edge clamp) -> 1xK conv -> +bias -> activation.

    kernel is the TF-layout (1, K, D, U) tensor.
    """
    x = np.asarray(x, dtype=F32)
    idx = tdnn_indices(x.shape[1], list(context), subsampling_factor, padding.upper())
    g = x[:, idx, :]                                       # (B, T', K, D)
    B, Te, K, D = g.shape
    y = np.matmul(g.reshape(B, Te, K * D), np.asarray(kernel, dtype=F32).reshape(K * D, -1))
    if bias is not None:
        y = y + np.asarray(bias, dtype=F32)
    if activation == "relu":
        y = np.maximum(y, F32(0))
    elif activation is not None:
        raise NotImplementedError(activation)
    return y.astype(F32)


def relu(x):
    """keras ReLU (sequential.py:71-72)."""
    return np.maximum(np.asarray(x, dtype=F32), F32(0))


def batchnorm(x, gamma, mean, var, epsilon=0.001):
    """batchnorm.py:81-88 -- keras BN inference, center=False: gamma*(x-mean)/sqrt(var+eps)."""
    x = np.asarray(x, dtype=F32)
    inv = (np.asarray(gamma, dtype=F32) / np.sqrt(np.asarray(var, dtype=F32) + F32(epsilon))).astype(F32)
    return (x * inv + (-np.asarray(mean, dtype=F32) * inv)).astype(F32)


# --------------------------------------------------------------------------
# layers/stats/stats_pooling.py
# --------------------------------------------------------------------------

def _stats_all(x, input_period, include_std, epsilon):
    """stats_pooling.py:211-240."""
    if input_period > 1:
        x = x[:, ::input_period, :]
    mean = np.mean(x, axis=1, keepdims=True, dtype=F32)
    if not include_std:
        return mean
    x2 = np.mean(x * x, axis=1, keepdims=True, dtype=F32)
    var = x2 - mean * mean
    std = np.sqrt(np.maximum(var, F32(0)) + F32(epsilon))
    return np.concatenate([mean, std], axis=-1).astype(F32)


def _stats_eval_indices(T, left, right, input_period, output_period, padding):
    """stats_pooling.py:157-209."""
    # getStartEndSteps (:157-177): SAME returns (0, T) early; VALID returns (start, end + 1).
    if padding == "SAME":
        t = np.arange(0, T, output_period)
    else:
        start, end = 0, T
        if left < 0:
            start = -left
        if right > 0 and (right - left + 1) < T:
            end = T - right
        t = np.arange(start, end + 1, output_period)
    rc = right + 1
    if rc > T:
        rc = T
    offs = np.arange(left, rc, input_period)
    idx = t[:, None] + offs[None, :]
    mask = (idx >= 0) & (idx < T)
    return np.clip(idx, 0, T - 1), mask


def _stats_windows(x, left, right, input_period, output_period, include_std, epsilon, padding):
    """stats_pooling.py:242-295."""
    T = x.shape[1]
    idx, mask = _stats_eval_indices(T, left, right, input_period, output_period, padding)
    m = mask.astype(F32)[None, :, :, None]
    n = np.sum(m, axis=2)
    g = x[:, idx, :]
    mean = np.sum(g * m, axis=2, dtype=F32) / n
    if not include_std:
        return mean.astype(F32)
    g2 = (x * x)[:, idx, :]
    var = np.sum(g2 * m, axis=2, dtype=F32) / n - mean * mean
    std = np.sqrt(np.maximum(var, F32(0)) + F32(epsilon))
    return np.concatenate([mean, std], axis=-1).astype(F32)


def stats_pooling(x, left_context, right_context, input_period=1, output_period=1,
                  include_std=True, padding="SAME", epsilon=1e-10, reduce_time_axis=False):
    """stats_pooling.py:297-316."""
    if left_context > 0 or right_context < 0:
        raise ValueError("'left_context' must be <= 0 and 'right_context' must be >= 0")
    if input_period <= 0 or output_period <= 0:
        raise ValueError("'input_period' and 'output_period' must be > 0")
    if output_period % input_period != 0 and not reduce_time_axis:
        raise ValueError("'output_period' must be a multiple of 'input_period'")
    padding = padding.upper()
    if padding not in ("VALID", "SAME"):
        raise ValueError("padding should be either 'VALID' or 'SAME'")
    x = np.asarray(x, dtype=F32)
    if reduce_time_axis:
        return _stats_all(x, input_period, include_std, epsilon)
    T = x.shape[1]
    if padding == "SAME":
        s = _stats_windows(x, left_context, right_context, input_period, output_period,
                           include_std, epsilon, padding)
        if output_period > 1:
            s = np.repeat(s, output_period, axis=1)
        return s
    if T > (right_context - left_context + 1):
        return _stats_windows(x, left_context, right_context, input_period, output_period,
                              include_std, epsilon, padding)
    return _stats_all(x, input_period, include_std, epsilon)


# --------------------------------------------------------------------------
# layers/plda/plda.py
# --------------------------------------------------------------------------

LOG2PI = 1.8378770664093454835606594728112


def plda_transform(x, mean, transform, psi, normalize_length=True, simple_length_norm=False,
                   dtype=np.float64, num_examples=1.0):
    """plda.py:163-196 -- u = T x - T m; optional length normalisation.  x (B, dim) -> (B, dim)."""
    dt = np.dtype(dtype).type
    x = np.asarray(x).astype(dt)
    Tm = np.asarray(transform).astype(dt)
    m = np.asarray(mean).astype(dt).reshape(-1, 1)
    ps = np.asarray(psi).astype(dt)
    dim = dt(x.shape[-1])
    offset = dt(-1.0) * np.matmul(Tm, m)                   # (dim, 1)
    u = (offset + np.matmul(Tm, x.T)).T.astype(dt)         # (B, dim)
    if normalize_length:
        if simple_length_norm:
            nf = np.sqrt(dim) / np.linalg.norm(u, axis=1, keepdims=True)
        else:
            inv_covar = dt(1.0) / (ps + dt(1.0 / num_examples))
            dot = np.sum(inv_covar[None, :] * u * u, axis=1, keepdims=True)
            nf = np.sqrt(dim / dot)
        u = u * nf.astype(dt)
    return u.astype(dt)


def plda_llr(u, psi, dtype=np.float64, num_examples=1.0):
    """plda.py:198-245 -- direct (B, dim, B) broadcast form; score[i, j]: i = test row, j = enrolled col."""
    dt = np.dtype(dtype).type
    u = np.asarray(u).astype(dt)
    ps = np.asarray(psi).astype(dt)
    n = dt(num_examples)
    one = dt(1.0)
    dim = dt(u.shape[1])

    def loglike(inputs, mean, var):                         # inputs (B, dim, 1), mean (1, dim, B)
        logdet = np.sum(np.log(var))
        sq = (inputs - mean) ** 2
        dot = np.sum(sq * (one / var)[None, :, None], axis=1)
        return dt(-0.5) * (logdet + dt(LOG2PI) * dim + dot)

    inputs = u[:, :, None]
    mean = (n * ps[None, :] * u) / (n * ps + one)[None, :]  # (B, dim)
    mean = mean.T[None, :, :]                               # (1, dim, B)
    var = one + ps / (n * ps + one)
    given = loglike(inputs, mean, var)
    without = loglike(inputs, np.zeros_like(mean), one + ps)
    return (given - without).astype(dt)


def plda(x, mean, transform, psi, normalize_length=True, simple_length_norm=False,
         dtype=np.float64):
    """plda.py:247-263 -- returns (scores (B, B), transformed (B, dim, 1))."""
    x = np.asarray(x)
    if x.ndim == 3:
        x = x[:, 0, :]
    u = plda_transform(x, mean, transform, psi, normalize_length, simple_length_norm, dtype)
    return plda_llr(u, psi, dtype), u[:, :, None]


# --------------------------------------------------------------------------
# models/kaldi/sequential.py, models/kaldi/xvector_extractor.py
# --------------------------------------------------------------------------

def sequential(x, layers):
    """sequential.py:29-83 -- run a list of layer dicts produced by the host config parser.

    Each entry: {"type": "affine", "kernel": (1,K,D,U), "bias": (U,), "context": [...]}
              | {"type": "relu"} | {"type": "batchnorm", "gamma","mean","var","epsilon"}
              | {"type": "stats", **StatsPooling kwargs}
    """
    for l in layers:
        t = l["type"]
        if t == "affine":
            x = tdnn(x, l["kernel"], l.get("bias"), l.get("context", [0]),
                     l.get("subsampling_factor", 1), l.get("padding", "SAME"),
                     l.get("activation"))
        elif t == "relu":
            x = relu(x)
        elif t == "batchnorm":
            x = batchnorm(x, l["gamma"], l["mean"], l["var"], l.get("epsilon", 0.001))
        elif t == "stats":
            kw = {k: v for k, v in l.items() if k != "type"}
            x = stats_pooling(x, **kw)
        else:
            raise ValueError(f"unsupported layer type '{t}'")
    return x


def lda_length_norm(x, global_mean, lda_mat_kaldi):
    """xvector_extractor.py:129-134, 174-181 -- transform.mat is (lda_dim, dim+1) = [L | o]."""
    x = np.asarray(x, dtype=F32)
    mat = np.asarray(lda_mat_kaldi, dtype=F32)
    off = mat[..., -1:].T
    L = mat[..., :-1].T
    y = np.matmul(x - np.asarray(global_mean, dtype=F32), L) + off
    norm = np.linalg.norm(y, axis=-1, keepdims=True)
    ratio = norm / np.sqrt(F32(y.shape[-1]))
    return (y / ratio).astype(F32)


def xvector_extractor(wav, cfg, layers, global_mean, lda_mat_kaldi, return_intermediate=False):
    """xvector_extractor.py:137-186 -- batch-1 wav -> (lda_dim,) x-vector.

    cfg: the `extractor` section of data/tflite_models/*.yml (framing/mfcc/vad/cmvn dicts).
    """
    wav = np.asarray(wav, dtype=F32).reshape(1, -1)
    fr = {k: v for k, v in cfg["framing"].items() if k != "dynamic_input_shape"}
    x = framing(wav, **fr)
    feats = mfcc(x, **cfg["mfcc"])
    vkw = dict(cfg["vad"])
    vkw["return_indexes"] = True
    idx = vad(feats, **vkw)
    x = feats[idx[:, 0], idx[:, 1]][None]
    x = cmvn(x, **cfg["cmvn"])
    emb = sequential(x, layers)
    out = lda_length_norm(emb, global_mean, lda_mat_kaldi)
    out = np.squeeze(out)
    if return_intermediate:
        return out, {"mfcc": feats, "vad_idx": idx, "embedding": emb}
    return out


# --------------------------------------------------------------------------
# kaldi_numpy/frame_extraction.py (host-side helper for Kaldi snip-edges=false)
# --------------------------------------------------------------------------

def pad_waveform(x, frame_size, frame_shift):
    """frame_extraction.py:28-89 -- mirror padding so that Framing matches snip-edges=false."""
    x = np.asarray(x)
    N = x.shape[-1]
    M = (N + (frame_shift // 2)) // frame_shift
    Nv = (M - 1) * frame_shift + frame_size
    left_over = abs(N - Nv)
    left = (frame_size - frame_shift) // 2
    right = left_over - left
    lp = np.flip(x[..., :left], axis=-1)
    rp = np.flip(x[..., -right:], axis=-1)
    return np.concatenate([lp, x, rp], axis=-1)
