"""
World-size-2 `gloo` tests (CPU) of the multi-GPU host logic in kaldi_tflite_b200/parallel.py:
utterance sharding (no collective) and the PLDA exchange (one all-gather of the transformed test
vectors, enrolled columns stay sharded).  The score kernel itself is replaced by the float64 oracle
here -- the point is the partitioning / gather plumbing, which is device independent.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ktf_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_test, n_enroll, dim, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kaldi_tflite_b200 import parallel
    try:
        # --- utterance sharding: contiguous, balanced, covers every unit exactly once -----------------
        lo, hi = parallel.shard_range(1001)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([hi - lo]))
        assert sum(int(c) for c in counts) == 1001 and max(counts) - min(counts) <= 1

        # --- ragged all-gather of row blocks ----------------------------------------------------------
        rng = np.random.default_rng(0)
        x_test = rng.standard_normal((n_test, dim)).astype(np.float32)
        x_enroll = rng.standard_normal((n_enroll, dim)).astype(np.float32)
        tlo, thi = parallel.shard_range(n_test)
        elo, ehi = parallel.shard_range(n_enroll)
        gathered = parallel.gather_rows(torch.from_numpy(x_test[tlo:thi]))
        assert torch.equal(gathered, torch.from_numpy(x_test))

        # --- PLDA exchange: scores[all tests, local enrolled] with a stand-in scorer ----------------------
        psi = np.exp(np.linspace(2, -3, dim))
        mean, Tm = np.zeros(dim), np.eye(dim)

        class Stub:                                            # same two methods the PLDA layer offers
            def transformVector(self, x):
                return torch.from_numpy(O.plda_transform(x.numpy(), mean, Tm, psi, dtype=np.float64))

            def logLikelihoodRatio(self, u_test, u_enroll, out=None):
                u = np.concatenate([u_test.numpy(), u_enroll.numpy()])
                full = torch.from_numpy(O.plda_llr(u, psi)[:u_test.shape[0], u_test.shape[0]:])
                if out is None:
                    return full
                out.copy_(full)
                return out

        block, u_all = parallel.plda_score_sharded(Stub(), torch.from_numpy(x_test[tlo:thi]),
                                                   torch.from_numpy(x_enroll[elo:ehi]))
        counts = [parallel.shard_range(n_test, r, world)[1] - parallel.shard_range(n_test, r, world)[0]
                  for r in range(world)]
        block2, _ = parallel.plda_score_sharded(Stub(), torch.from_numpy(x_test[tlo:thi]),
                                                torch.from_numpy(x_enroll[elo:ehi]), test_counts=counts)
        assert torch.equal(block, block2)
        np.save(os.path.join(out_dir, f"block{rank}.npy"), block.numpy())
        assert u_all.shape == (n_test, dim)
    finally:
        dist.destroy_process_group()


def test_sharding_and_plda_exchange_world2(tmp_path):
    world, n_test, n_enroll, dim = 2, 37, 23, 16
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_test, n_enroll, dim, str(tmp_path)), nprocs=world, join=True)
    # reference: everything on one rank
    rng = np.random.default_rng(0)
    x_test = rng.standard_normal((n_test, dim)).astype(np.float32)
    x_enroll = rng.standard_normal((n_enroll, dim)).astype(np.float32)
    psi = np.exp(np.linspace(2, -3, dim))
    u = O.plda_transform(np.concatenate([x_test, x_enroll]), np.zeros(dim), np.eye(dim), psi, dtype=np.float64)
    want = O.plda_llr(u, psi)[:n_test, n_test:]
    got = np.concatenate([np.load(tmp_path / f"block{r}.npy") for r in range(world)], axis=1)
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=0, atol=1e-9)
