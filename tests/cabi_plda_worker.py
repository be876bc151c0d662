"""
Sharded PLDA scoring driven through the C-ABI ALONE (ctypes + numpy: no torch, no torch.distributed) -- what a C / Go /
Java host would do with include/ktf_b200.h:

    ktf_ctx_create -> ktf_nccl_comm_init (id from rank 0 through a file) -> ktf_plda_create -> ktf_plda_transform of the
    local test / enrolled shards -> ktf_nccl_allgather_xvec of the transformed test vectors -> ktf_plda_score of
    (all tests x local enrolled) -> copy the block back and compare it with the float64 oracle.

usage: cabi_plda_worker.py <rank> <nranks> <device> <rendezvous dir>
"""

import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def main(rank, nranks, device, rdv):
    # the ctypes signatures only: _native.py is loaded as a stand-alone module (importing the package would pull in the
    # torch-based Python host, which is exactly what this worker does without)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ktf_native", os.path.join(ROOT, "kaldi_tflite_b200", "_native.py"))
    N = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(N)
    lib = N.lib()
    assert "torch" not in sys.modules
    P = ctypes.c_void_p

    def ok(rc):
        if rc != 0:
            raise RuntimeError(lib.ktf_last_error().decode())

    ctx = P()
    ok(lib.ktf_ctx_create(device, ctypes.byref(ctx)))
    stream = P(lib.ktf_ctx_stream(ctx))
    if nranks > 1:
        assert lib.ktf_nccl_available() == 1, "libnccl.so.2 not found"
        idfile = os.path.join(rdv, "nccl_id.bin")
        if rank == 0:
            buf = ctypes.create_string_buffer(N.KTF_NCCL_UNIQUE_ID_BYTES)
            ok(lib.ktf_nccl_unique_id(buf))
            with open(idfile + ".tmp", "wb") as f:
                f.write(buf.raw)
            os.replace(idfile + ".tmp", idfile)
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 120, "rendezvous timed out"
            time.sleep(0.05)
        with open(idfile, "rb") as f:
            idb = ctypes.create_string_buffer(f.read(), N.KTF_NCCL_UNIQUE_ID_BYTES)
        ok(lib.ktf_nccl_comm_init(ctx, nranks, rank, idb))

    dim, n_test, n_enroll = 128, 1200, 900
    prng = np.random.default_rng(1234)                   # same synthetic PLDA model as the other tests
    psi = np.exp(np.linspace(3, -4, dim))
    q, _ = np.linalg.qr(prng.standard_normal((dim, dim)))
    Tm = q * prng.uniform(0.5, 2.0, size=(1, dim))
    mean = prng.standard_normal(dim) * 0.05
    rng = np.random.default_rng(41)
    x = rng.standard_normal((n_test + n_enroll, dim))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True) * np.sqrt(dim)).astype(np.float32)
    xt, xe = x[:n_test], x[n_test:]
    per_t = (n_test + nranks - 1) // nranks              # equal blocks: the last shard is padded
    per_e = (n_enroll + nranks - 1) // nranks
    t0, t1 = rank * per_t, min((rank + 1) * per_t, n_test)
    e0, e1 = rank * per_e, min((rank + 1) * per_e, n_enroll)
    xt_loc = np.zeros((per_t, dim), np.float32)
    xt_loc[:t1 - t0] = xt[t0:t1]
    xe_loc = np.ascontiguousarray(xe[e0:e1])

    def dev_alloc(nbytes):
        p = P()
        ok(lib.ktf_ctx_malloc(ctx, nbytes, ctypes.byref(p)))
        return p

    def h2d(arr):
        arr = np.ascontiguousarray(arr)
        p = dev_alloc(arr.nbytes)
        ok(lib.ktf_ctx_memcpy_h2d(ctx, p, arr.ctypes.data_as(P), arr.nbytes))
        return p, arr                                   # keep the host array alive until the stream has consumed it

    plda = P()
    m64, T64, p64 = (np.ascontiguousarray(a, np.float64) for a in (mean, Tm, psi))
    ok(lib.ktf_plda_create(dim, m64.ctypes.data_as(P), T64.ctypes.data_as(P), p64.ctypes.data_as(P), 1, 0, 4,
                           ctypes.byref(plda)))
    d_xt, keep1 = h2d(xt_loc)
    d_xe, keep2 = h2d(xe_loc)
    d_ut = dev_alloc(per_t * dim * 4)
    d_ue = dev_alloc((e1 - e0) * dim * 4)
    d_uall = dev_alloc(nranks * per_t * dim * 4)
    n_all = nranks * per_t
    d_sc = dev_alloc(n_all * (e1 - e0) * 4)
    n0 = lib.ktf_launch_count()
    ok(lib.ktf_plda_transform(plda, d_xt, per_t, d_ut, stream))
    ok(lib.ktf_plda_transform(plda, d_xe, e1 - e0, d_ue, stream))
    ok(lib.ktf_nccl_allgather_xvec(ctx, d_ut, d_uall, per_t, dim, 4, None))
    ok(lib.ktf_plda_score(plda, d_uall, n_all, d_ue, e1 - e0, d_sc, e1 - e0, stream))
    launches = lib.ktf_launch_count() - n0
    got = np.empty((n_all, e1 - e0), np.float32)
    ok(lib.ktf_ctx_memcpy_d2h(ctx, got.ctypes.data_as(P), d_sc, got.nbytes))

    from oracle import ktf_oracle as O
    uo = O.plda_transform(x, mean, Tm, psi, dtype=np.float64)
    want_full = O.plda_llr(uo, psi)[:n_test, n_test + e0:n_test + e1]
    rows = np.concatenate([np.arange(r * per_t, r * per_t + min((r + 1) * per_t, n_test) - r * per_t)
                           for r in range(nranks)])       # the un-padded rows of every rank's block
    rel = float(np.max(np.abs(got[rows] - want_full) / np.maximum(np.abs(want_full), 1.0)))
    with open(os.path.join(rdv, f"cabi_rank{rank}.json"), "w") as f:
        json.dump({"max_rel": rel, "launches": int(launches), "rows": int(len(rows)), "nranks": nranks,
                   "torch_loaded": "torch" in sys.modules}, f)
    for p in (d_xt, d_xe, d_ut, d_ue, d_uall, d_sc):
        ok(lib.ktf_ctx_free(ctx, p))
    lib.ktf_plda_destroy(plda)
    if nranks > 1:
        ok(lib.ktf_nccl_comm_destroy(ctx))
    lib.ktf_ctx_destroy(ctx)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
