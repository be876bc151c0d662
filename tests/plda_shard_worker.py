"""
One rank of the two-rank sharded PLDA parity test (tests/test_gpu_round2.py).  Default: both ranks share cuda:0 (NCCL
refuses two ranks on one device, so gloo carries the CUDA tensors of the exchange).  With `nccl` as the second argument
every rank takes its own GPU and the exchange is the asynchronous NCCL all-gather (the third argument is the number of
test vectors: an odd count makes the shards ragged).  Every rank scores its shard of the ENROLLED columns with the real
kernels through kaldi_tflite_b200.parallel.plda_score_sharded and checks its (n_test x n_enroll / G) block against the
float64 oracle: |delta| <= 1e-3 * max(|s|, 1).
"""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main(out_dir, backend="gloo", n_test=1500):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    import kaldi_tflite_b200 as ktf
    from kaldi_tflite_b200 import parallel
    from oracle import ktf_oracle as O
    from test_gpu_tdnn_plda import synthetic_plda

    dim, n_enroll = 128, 1100                          # ragged against the 128 x 256 tiles and against the ranks
    mean, Tm, psi = synthetic_plda(dim)
    rng = np.random.default_rng(31)
    x = rng.standard_normal((n_test + n_enroll, dim))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True) * np.sqrt(dim)).astype(np.float32)
    xt, xe = x[:n_test], x[n_test:]
    t0, t1 = parallel.shard_range(n_test, rank, world)
    e0, e1 = parallel.shard_range(n_enroll, rank, world)
    counts = [parallel.shard_range(n_test, r, world)[1] - parallel.shard_range(n_test, r, world)[0] for r in range(world)]

    layer = ktf.layers.PLDA(dim, mean, Tm, psi, dtype=np.float32, return_transformed=False)
    n0 = ktf.launch_count()
    scores, u_all = parallel.plda_score_sharded(layer, torch.from_numpy(xt[t0:t1]).cuda(),
                                                torch.from_numpy(xe[e0:e1]).cuda(), test_counts=counts)
    torch.cuda.synchronize()
    launches = ktf.launch_count() - n0
    got = scores.cpu().numpy()

    uo = O.plda_transform(x, mean, Tm, psi, dtype=np.float64)
    want = O.plda_llr(uo, psi)[:n_test, n_test + e0:n_test + e1]
    ok = got.shape == want.shape
    rel = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0))) if ok else float("inf")
    gathered_ok = bool(np.allclose(u_all.cpu().numpy(), uo[:n_test], atol=1e-4))
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"ok": bool(ok and gathered_ok), "max_rel": rel, "launches": int(launches),
                   "shape": list(got.shape)}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3]), *(int(a) for a in sys.argv[3:4]))
