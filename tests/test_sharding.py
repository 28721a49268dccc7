"""Host-side logic of the multi-rank path on CPU: sharding of the catalog, the gloo all-gather of the
per-rank partials (world_size 2) and the rank-ordered merge (bump_merge_partials, the same code as the device
finalize kernel).  The partials are built from the oracle's per-sample log-weights, so the merged
loglike / log_mu_sel / log_mu2 / neff_sel can be checked against the unsharded oracle."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bumpcosmology_b200 import _lib
from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
from bumpcosmology_b200.likelihood import merge_partials, shard_bounds, shard_catalog, unpack_header

# layout indices (csrc/bump_layout.cuh)
P_LLSUM, P_NOBS, P_NVALID_EVT, P_SEL_M, P_SEL_ACC0, P_NSEL, P_SCAL0 = 0, 1, 19, 20, 21, 41, 64
S_H, S_INV_H, S_M, S_FPL, S_CONST, S_LOG_NSAMP, S_LOG_NDRAW = 0, 1, 3, 16, 17, 34, 35


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 69, 5000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gwtc3_split_matches_survey():
    sizes = [shard_bounds(69, r, 8)[1] - shard_bounds(69, r, 8)[0] for r in range(8)]
    assert sizes == [9, 9, 9, 9, 9, 8, 8, 8]


def _oracle_partial(data, theta):
    """A partial in the library's layout from the oracle's log-weights (gradient features left at zero;
    theta-only constant folded in as S_CONST = 0)."""
    from oracle import bump_oracle as bo
    m1, q, dl, pd, m1s, qs, dls, pds, ndraw = data
    th = torch.tensor(np.asarray(theta, dtype=np.float64))
    with torch.no_grad():
        cosmo, log_dN = bo.build_model(th)
        p = np.zeros(_lib.PARTIAL_LEN)
        if m1.shape[0]:
            lw = bo.log_weights(cosmo, log_dN, m1, q, dl, pd)
            p[P_LLSUM] = float(torch.logsumexp(lw, dim=1).sum())
            p[P_NVALID_EVT] = float(torch.isfinite(lw).sum())
        p[P_NOBS] = m1.shape[0]
        if m1s.shape[0]:
            lws = bo.log_weights(cosmo, log_dN, m1s, qs, dls, pds)
            m = float(lws.max())
            p[P_SEL_M] = m
            p[P_SEL_ACC0] = float(torch.exp(lws - m).sum())
            p[P_SEL_ACC0 + 1] = float(torch.exp(2 * (lws - m)).sum())
        else:
            p[P_SEL_M] = -math.inf
        p[P_NSEL] = m1s.shape[0]
    sc = p[P_SCAL0:]
    sc[S_H], sc[S_INV_H], sc[S_M], sc[S_FPL] = theta[0], 1 / theta[0], theta[7], theta[9]
    sc[S_CONST] = 0.0
    sc[S_LOG_NSAMP] = math.log(max(m1.shape[1] if m1.ndim == 2 and m1.shape[0] else 1, 1))
    sc[S_LOG_NDRAW] = math.log(ndraw)
    return p


def _check_merged(hdr, cat):
    from oracle import bump_oracle as bo
    o = bo.evaluate(THETA_DEFAULT, cat.as_args(), grad=False)
    m = unpack_header(hdr, 14)
    assert abs(m["loglike"] - o["loglike"]) <= 1e-11 * abs(o["loglike"])
    assert abs(m["log_mu_sel"] - o["log_mu_sel"]) <= 1e-11 * abs(o["log_mu_sel"])
    assert abs(m["log_mu2"] - o["log_mu2"]) <= 1e-11 * abs(o["log_mu2"])
    assert abs(m["neff_sel"] - o["neff_sel"]) <= 1e-9 * o["neff_sel"]
    assert m["nobs"] == cat.nobs and m["nsel"] == cat.nsel


@pytest.mark.parametrize("world", (1, 2, 3, 8))
def test_merge_of_sharded_partials_equals_unsharded(world):
    cat = make_catalog("tiny")
    parts = [_oracle_partial(shard_catalog(cat.as_args(), r, world), THETA_DEFAULT) for r in range(world)]
    _check_merged(merge_partials(np.array(parts)), cat)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cat = make_catalog("tiny")
    part = torch.from_numpy(_oracle_partial(shard_catalog(cat.as_args(), rank, world), THETA_DEFAULT))
    gathered = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    hdr = merge_partials(torch.stack(gathered).numpy())
    np.save(os.path.join(out_dir, f"hdr{rank}.npy"), hdr)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_exchange_is_identical_on_every_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    h0, h1 = (np.load(tmp_path / f"hdr{r}.npy") for r in range(world))
    assert h0.tobytes() == h1.tobytes()          # bitwise identical merged result on every rank
    _check_merged(h0, make_catalog("tiny"))
