"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): one process per GPU, torch.distributed NCCL for the
plumbing; both exchange modes must reproduce the single-GPU result to 1e-12 and agree bitwise across ranks."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, exchange):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
    from bumpcosmology_b200.likelihood import ShardedHyperlikelihood
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cat = make_catalog("small", seed=17)
    like = ShardedHyperlikelihood(cat.as_args(), device=rank, exchange=exchange)
    res = []
    for th in np.vstack([THETA_DEFAULT, draw_prior_thetas(2, seed=3)]):
        r = like(th)
        res.append(np.concatenate([[r.loglike, r.log_mu_sel, r.log_mu2, r.neff_sel], r.dloglike, r.dlog_mu_sel]))
    np.save(os.path.join(out_dir, f"res_{exchange}_{rank}.npy"), np.array(res))
    np.save(os.path.join(out_dir, f"neff_{exchange}_{rank}.npy"), r.neff)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ("torch", "nccl", "p2p"))
def test_two_ranks_match_single_gpu(tmp_path, exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"res_{exchange}_{r}.npy") for r in range(world))
    assert r0.tobytes() == r1.tobytes()
    cat = make_catalog("small", seed=17)
    like = Hyperlikelihood(*cat.as_args())
    for k, th in enumerate(np.vstack([THETA_DEFAULT, draw_prior_thetas(2, seed=3)])):
        r = like(th)
        ref = np.concatenate([[r.loglike, r.log_mu_sel, r.log_mu2, r.neff_sel], r.dloglike, r.dlog_mu_sel])
        scale = np.maximum(np.abs(ref), max(1.0, float(np.max(np.abs(r.dloglike)))) * np.ones_like(ref))
        assert np.all(np.abs(r0[k] - ref) <= 1e-12 * scale), (k, np.max(np.abs(r0[k] - ref) / scale))
    neff = np.concatenate([np.load(tmp_path / f"neff_{exchange}_{r}.npy") for r in range(world)])
    assert np.allclose(neff, r.neff, rtol=1e-12)
    like.close()
