"""Shared helpers: translate the Kaldi .conf dictionaries stored with the golden
vectors into the reference's layer kwargs (same mapping as the reference's
fixture loaders, testdata/feats/feats.py:40-212)."""

import json


def _b(v):
    return v == "true"


def mfcc_conf_to_kwargs(conf_json):
    conf = json.loads(conf_json) if isinstance(conf_json, str) else conf_json
    cfg = {"snip_edges": False, "framing": {}, "mfcc": {}}
    for k, v in conf.items():
        if k == "sample-frequency":
            cfg["framing"]["sample_frequency"] = float(v)
            cfg["mfcc"]["sample_frequency"] = float(v)
        elif k == "frame-length":
            cfg["framing"]["frame_length_ms"] = float(v)
        elif k == "frame-shift":
            cfg["framing"]["frame_shift_ms"] = float(v)
        elif k == "use-energy":
            cfg["mfcc"]["use_energy"] = _b(v)
        elif k == "raw-energy":
            cfg["mfcc"]["raw_energy"] = _b(v)
        elif k == "dither":
            cfg["mfcc"]["dither"] = float(v)
        elif k == "low-freq":
            cfg["mfcc"]["low_freq_cutoff"] = float(v)
        elif k == "high-freq":
            cfg["mfcc"]["high_freq_cutoff"] = float(v)
        elif k == "num-mel-bins":
            cfg["mfcc"]["num_mels"] = int(v)
        elif k == "num-ceps":
            cfg["mfcc"]["num_mfccs"] = int(v)
        elif k == "snip-edges":
            cfg["snip_edges"] = _b(v)
        else:
            raise ValueError(k)
    return cfg


def fbank_conf_to_kwargs(conf_json):
    conf = json.loads(conf_json) if isinstance(conf_json, str) else conf_json
    cfg = {"snip_edges": False, "framing": {}, "windowing": {}, "fbank": {}}
    for k, v in conf.items():
        if k == "sample-frequency":
            cfg["framing"]["sample_frequency"] = float(v)
            cfg["fbank"]["sample_frequency"] = float(v)
        elif k == "frame-length":
            cfg["framing"]["frame_length_ms"] = float(v)
        elif k == "frame-shift":
            cfg["framing"]["frame_shift_ms"] = float(v)
        elif k == "raw-energy":
            cfg["windowing"]["raw_energy"] = _b(v)
        elif k == "dither":
            cfg["windowing"]["dither"] = float(v)
        elif k == "low-freq":
            cfg["fbank"]["low_freq_cutoff"] = float(v)
        elif k == "high-freq":
            cfg["fbank"]["high_freq_cutoff"] = float(v)
        elif k == "num-mel-bins":
            cfg["fbank"]["num_bins"] = int(v)
        elif k == "use-log-fbank":
            cfg["fbank"]["use_log_fbank"] = _b(v)
        elif k == "use-power":
            cfg["fbank"]["use_power"] = _b(v)
        elif k == "snip-edges":
            cfg["snip_edges"] = _b(v)
        else:
            raise ValueError(k)
    return cfg


def cmvn_conf_to_kwargs(conf_json):
    conf = json.loads(conf_json) if isinstance(conf_json, str) else conf_json
    cfg = {"window": 600, "center": True, "norm_vars": False, "min_window": 100}
    for k, v in conf.items():
        if k == "cmn-window":
            cfg["window"] = int(v)
        elif k == "center":
            cfg["center"] = _b(v)
        elif k == "norm-vars":
            cfg["norm_vars"] = _b(v)
        elif k == "min-cmn-window":
            cfg["min_window"] = int(v)
        else:
            raise ValueError(k)
    return cfg


def vad_conf_to_kwargs(conf_json):
    conf = json.loads(conf_json) if isinstance(conf_json, str) else conf_json
    cfg = {"energy_mean_scale": 0.5, "energy_threshold": 5.0, "frames_context": 0,
           "proportion_threshold": 0.6, "return_indexes": False, "energy_coeff": 0}
    for k, v in conf.items():
        if k == "vad-energy-threshold":
            cfg["energy_threshold"] = float(v)
        elif k == "vad-energy-mean-scale":
            cfg["energy_mean_scale"] = float(v)
        elif k == "vad-frames-context":
            cfg["frames_context"] = int(v)
        elif k == "vad-proportion-threshold":
            cfg["proportion_threshold"] = float(v)
        else:
            raise ValueError(k)
    return cfg


STATS_CONFIGS = {
    "stats_mean": {"include_std": False},
    "stats_mean_std": {},
    "stats_mean_std_windowed": {"right_context": 4},
    "stats_mean_std_only_left_context": {"left_context": -4, "right_context": 0},
    "stats_mean_std_both_left_right_context": {"left_context": -4, "right_context": 4},
    "stats_mean_std_asymmetrical_context": {"left_context": -4, "right_context": 2},
    "stats_mean_std_subsampling": {"input_period": 4, "output_period": 4},
    "stats_mean_std_windowed_subsampling": {
        "left_context": -4, "right_context": 4, "input_period": 4, "output_period": 4},
}


def stats_default_cfg():
    # layers/stats/stats_pooling_test.py:33-46 (the keys StatsPooling actually accepts)
    return {"left_context": 0, "right_context": 16, "input_period": 1, "output_period": 1,
            "include_std": True, "padding": "SAME", "epsilon": 1e-10, "reduce_time_axis": False}


NARROW_LAYERS = [
    # layers/tdnn/tdnn_test.py:75-82
    ["tdnn1", 5, [-2, -1, 0, 1, 2], True, True],
    ["tdnn2", 8, [-2, 0, 2], True, True],
    ["tdnn3", 8, [-3, 0, 3], True, True],
    ["tdnn4", 8, [0], True, True],
    ["tdnn5", 8, [0], True, True],
    ["output", 1, [0], False, False],
]
