"""Hardware test of the data-parallel machinery: the bucketed, backward-overlapped gradient
all-reduce (tpgan_b200.parallel.BucketReducer) driven through a real NCCL process group - a 1-rank group, so that one GPU
suffices and the reduced gradients must equal the un-reduced ones.  In deterministic mode the whole optimizer trajectory of
  (a) the plain single-GPU trainer,
  (b) the trainer with the bucket reducer (per-bucket export + NCCL all-reduce on the communication stream), eager, and
  (c) the same with the NCCL calls captured into the step's CUDA graph
is bit-identical: the per-bucket gradient export covers every parameter exactly once, buckets are reduced after their last
weight-gradient kernel and before the optimizer, and graph capture of the collectives changes nothing.
The N > 1 host logic (sharding, averaging) is covered on gloo in tests/test_parallel_cpu.py; N = 2/4/8 NCCL runs are the
driver's scaling bench (replica_checksum_spread in the bench line)."""
import os
import socket

import pytest
import torch

from test_model_gpu import _models, rel

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_bucket_reducer_on_one_rank_nccl_group():
    import torch.distributed as dist
    from oracle import step as ostep
    from tpgan_b200 import _lib
    from tpgan_b200.train_step import TPGANTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    prev = _lib.set_deterministic(True)
    try:
        B = 2
        b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}

        def run(**kw):
            G, D, _, _ = _models(False)
            tr = TPGANTrainer(G, D, B, **kw)
            ms = [tr.step(b, optimize=True) for _ in range(3)]
            torch.cuda.synchronize()
            return tr, ms[-1], tr.flat_g.data.clone(), tr.flat_d.data.clone(), tr.flat_g.grad.clone()

        _, m_a, pg_a, pd_a, gg_a = run()
        tr_b, m_b, pg_b, pd_b, gg_b = run(force_reducer=True, bucket_mb=8.0)
        assert len(tr_b.reducer.buckets) >= 8            # 551 MB of gradients in >= 8 MB buckets
        covered = sum(c for (_, _, _, c) in tr_b.reducer.buckets[:-1]) + (tr_b.flat_g.total - tr_b.reducer.buckets[-1][2])
        assert covered == tr_b.flat_g.total              # the buckets tile the flat gradient buffer exactly
        _, m_c, pg_c, pd_c, gg_c = run(force_reducer=True, bucket_mb=8.0, use_graphs=True, graph_collectives=True)
        assert _lib.kernel_status() == 0
        for name, x, y in (("G params eager", pg_b, pg_a), ("D params eager", pd_b, pd_a), ("G grads eager", gg_b, gg_a),
                           ("G params graph", pg_c, pg_a), ("D params graph", pd_c, pd_a), ("G grads graph", gg_c, gg_a)):
            assert torch.equal(x, y), (name, rel(x, y))
        for k in m_a:
            assert abs(m_a[k] - m_b[k]) <= 1e-5 * abs(m_a[k]) + 1e-7 and abs(m_a[k] - m_c[k]) <= 1e-5 * abs(m_a[k]) + 1e-7, k
    finally:
        _lib.set_deterministic(prev)
        dist.destroy_process_group()
