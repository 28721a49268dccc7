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


def _timeout_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import time

    import torch
    import torch.distributed as dist

    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    from bumpcosmology_b200.likelihood import ShardedHyperlikelihood
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cat = make_catalog("small", seed=17)
    like = ShardedHyperlikelihood(cat.as_args(), device=rank, exchange="p2p", timeout_s=1.0)
    assert like.exchange == "p2p"
    good = like(THETA_DEFAULT).loglike
    dist.barrier()
    if rank == 1:
        time.sleep(4.0)          # a host stall (GC, I/O, first-call graph instantiation ...) longer than the timeout
    outcome = []
    for _ in range(2):           # the failing evaluation and the one after it: the failure is sticky
        try:
            r = like(THETA_DEFAULT)
            outcome.append(("value", float(r.loglike)))
        except _lib.BumpError as e:
            outcome.append(("error", e.code))
    with open(os.path.join(out_dir, f"timeout_{rank}.txt"), "w") as f:
        f.write(repr((good, outcome)))
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_exchange_timeout_fails_on_every_rank(tmp_path):
    """One rank stalls for longer than the exchange timeout.  Round 1 returned NaN on the waiting rank and a VALID
    result on the late one (ranks then disagree about the chain state).  Now the evaluation fails on both, with
    BUMP_E_EXCHANGE, and keeps failing until the mailboxes are re-attached."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (kernels that wait for a peer must not share a GPU)")
    from bumpcosmology_b200 import _lib
    mp.spawn(_timeout_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        good, outcome = eval(open(tmp_path / f"timeout_{rank}.txt").read())
        assert np.isfinite(good)
        assert outcome == [("error", _lib.E_EXCHANGE), ("error", _lib.E_EXCHANGE)], (rank, outcome)
