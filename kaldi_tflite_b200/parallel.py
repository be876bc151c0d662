"""
Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous.

* wav -> x-vector: utterances are independent, so they are sharded across ranks with NO
  data-path collective (`shard_range`); an optional final gather of the 128-d x-vectors is
  offered for callers that want one tensor (`gather_rows`).
* PLDA all-vs-all: ENROLLED vectors (score columns) are sharded by rank, the transformed TEST
  vectors are exchanged once with an all-gather over NCCL/NVLink, and every rank writes its own
  (n_test x n_enroll/G) block; the 10 GB score matrix is never gathered.
"""

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous, balanced [start, stop) of `n_items` units for this rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local, counts=None):
    """All-gather of row blocks that may differ in length; returns the concatenation on every rank."""
    rank, w = world()
    if w == 1:
        return local
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def plda_score_sharded(plda, x_test_local, x_enroll_local):
    """
    Row-sharded PLDA trial scoring.  Each rank holds a slice of the test x-vectors and a slice of
    the enrolled x-vectors (raw, (n, dim) float32).  Returns this rank's block
    scores[all tests, local enrolled] -- test rows ordered by rank -- plus the gathered transformed
    test vectors.  The only collective is one all-gather of the transformed test vectors.
    """
    u_test_local = plda.transformVector(x_test_local.contiguous())
    u_enroll_local = plda.transformVector(x_enroll_local.contiguous())
    u_test = gather_rows(u_test_local)
    return plda.logLikelihoodRatio(u_test, u_enroll_local), u_test
