"""CPU, world_size 2, gloo: the expert-parallel dispatch/combine host logic (all-gather tokens, each rank computes
the (token, expert) pairs it owns, reduce the fp32 partials) gives the same result as the single-process MoE.
The per-rank compute is the CPU oracle with an expert-id offset — the same skip-non-owned-ids contract the CUDA
kernels implement (ktb200_moe_config.expert_id_offset)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

E, K, H, I = 8, 4, 512, 256
Q4_K, Q6_K, F32 = 12, 14, 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _weights():
    from ktransformers_b200.util.synth import synth_blocks
    return (synth_blocks(Q4_K, E * I * H, "cpu", 1).numpy(), synth_blocks(Q4_K, E * I * H, "cpu", 2).numpy(),
            synth_blocks(Q6_K, E * H * I, "cpu", 3).numpy())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ktransformers_b200.operators.expert_parallel import ExpertParallelCombine, shard_range
        from oracle.bindings import Oracle
        orc = Oracle()
        gate, up, down = _weights()
        lo, hi = shard_range(E, rank, world)
        gb, db = gate.size // E, down.size // E
        g_l, u_l, d_l = gate[lo * gb:hi * gb], up[lo * gb:hi * gb], down[lo * db:hi * db]

        def local_forward(xa, ia, wa, out):
            ids = ia.numpy() - lo                       # ids outside [0, E/N) are skipped by the MoE
            out.copy_(torch.from_numpy(orc.moe_forward(hi - lo, H, I, g_l, u_l, d_l, Q4_K, Q4_K, Q6_K, F32, ids, wa.numpy(), xa.numpy())))

        rng = np.random.default_rng(100 + rank)         # every rank has its own token
        x = torch.from_numpy((rng.standard_normal((1, H)) / 50).astype(np.float32))
        ids = torch.from_numpy(rng.permutation(E)[:K].astype(np.int64)[None, :])
        w = torch.from_numpy(rng.random((1, K)).astype(np.float32))
        ep = ExpertParallelCombine(H, K, 1, "cpu", in_dtype=torch.float32)
        got = ep.forward(x, ids, w, local_forward).numpy().copy()
        want = orc.moe_forward(E, H, I, gate, up, down, Q4_K, Q4_K, Q6_K, F32, ids.numpy(), w.numpy(), x.numpy())
        q.put((rank, float(np.abs(got - want).max() / np.abs(want).max()), ep.ids_all.numpy().tolist()))
    finally:
        dist.destroy_process_group()


def test_expert_parallel_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][2] == res[1][2] and len(res[0][2]) == world      # every rank saw the same gathered routing table
    for rank, err, _ in res:
        assert err < 1e-5, (rank, err)                             # fp32 re-association across the two shards only


def test_shard_range_partitions_experts():
    from ktransformers_b200.operators.expert_parallel import shard_range
    for n, w in ((256, 8), (384, 8), (64, 4), (256, 1)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(AssertionError):
        shard_range(10, 0, 4)
