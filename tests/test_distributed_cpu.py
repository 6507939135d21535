"""N>1 path on the CPU: two gloo processes take the batch blocks of monteCarloDriver.f95:264-274, accumulate
moments and sum them with ONE all-reduce; the result must equal the single-process run of all batches
(per-batch results depend on (iseed, batch) only)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from i3rc_monte_carlo_model_b200 import fields
    from i3rc_monte_carlo_model_b200.driver import partition_batches, run_batches_host
    from oracle.binding import oracle_backend
    from tests.cases import make_integrator

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    I = make_integrator(oracle_backend(), fields.step_cloud(0.99), surfaceAlbedo=0.1, intensityMus=[1.0], intensityPhis=[0.0])
    nB, mine = partition_batches(5, world, rank)  # 5 -> 6 batches, 3 per rank
    st = run_batches_host(I, dict(solarMu=0.5, solarAzimuth=0.0), 500, mine, iseed=10)
    st.allreduce(dist)
    res = st.finish(1.0, nB)
    if rank == 0:
        q.put((nB, {k: (v[0].copy(), v[1].copy()) for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as mp

    from i3rc_monte_carlo_model_b200 import fields
    from i3rc_monte_carlo_model_b200.driver import partition_batches, run_batches_host
    from oracle.binding import oracle_backend
    from tests.cases import make_integrator

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    nB, dist_res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert nB == 6
    I = make_integrator(oracle_backend(), fields.step_cloud(0.99), surfaceAlbedo=0.1, intensityMus=[1.0], intensityPhis=[0.0])
    nB1, allb = partition_batches(6, 1, 0)
    single = run_batches_host(I, dict(solarMu=0.5, solarAzimuth=0.0), 500, allb, iseed=10).finish(1.0, nB1)
    for k in single:
        assert np.allclose(dist_res[k][0], single[k][0], rtol=1e-12, atol=1e-14), k
        assert np.allclose(dist_res[k][1], single[k][1], rtol=1e-9, atol=1e-12), k
