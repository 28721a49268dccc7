"""The device-pointer entries of the C ABI — what an XLA FFI handler (csrc/bump_xla_ffi.cc) forwards to — on ONE GPU:

* bump_eval_device on a caller stream with device pointers equals the host call bit for bit;
* the same call CAPTURED into a CUDA graph (how an XLA command buffer / torch.cuda.graph issues it) replays
  correctly for new theta values, allocates nothing and records no event of the library inside the capture;
* two emulated ranks (two shard contexts on one device): bump_eval_partial_device + bump_finalize_device with a
  device-side gather of the partials reproduce the unsharded evaluation — eagerly and as one captured graph;
* capture is refused (BUMP_E_INVALID, not a wrong answer) when the context shares its constant-bank slot.

No kernel here waits for another launch (the peer-memory exchange needs one GPU per rank: tests/test_gpu_multirank.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _flat(out, ntheta=14):
    return np.concatenate([out[:4], out[4:4 + ntheta], out[19:19 + ntheta], out[34:38]])


@pytest.fixture()
def small():
    from bumpcosmology_b200.catalogs import make_catalog
    return make_catalog("small", seed=31)


def _thetas():
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    return np.vstack([THETA_DEFAULT, draw_prior_thetas(3, seed=12)])


def test_eval_device_equals_host_call(small):
    import torch
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    like = Hyperlikelihood(*small.as_args())
    stream = torch.cuda.Stream()
    th_d = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64, device="cuda")
    out_d = torch.zeros(_lib.OUT_HEADER + like.nobs, dtype=torch.float64, device="cuda")
    for th in _thetas():
        ref = like.raw(th).copy()
        with torch.cuda.stream(stream):
            th_d[:14].copy_(torch.from_numpy(th))
            like.eval_device(th_d.data_ptr(), out_d.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        got = out_d.cpu().numpy()
        assert np.array_equal(got, ref, equal_nan=True)
        assert got[_lib.OUT_STATUS] == 0.0
    like.close()


def test_queued_evaluations_with_changing_theta(small):
    """48 evaluations queued back to back on one stream with no host synchronisation in between, theta changing every
    time: each kernel of evaluation k + 1 is on the GPU's queue while evaluation k still runs (and the dependent
    kernels of one evaluation are scheduled early), so a scalar or table taken from the neighbouring evaluation
    would give that evaluation's result.  Bitwise against the same theta evaluated alone."""
    import torch
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    like = Hyperlikelihood(*small.as_args())
    thetas = _thetas()
    alone = [like.raw(th).copy() for th in thetas]
    n = 48
    stream = torch.cuda.Stream()
    th_d = torch.zeros(n, _lib.NTHETA_MAX, dtype=torch.float64, device="cuda")
    out_d = torch.zeros(n, _lib.OUT_HEADER + like.nobs, dtype=torch.float64, device="cuda")
    order = [(5 * k + k // 4) % len(thetas) for k in range(n)]
    for k, j in enumerate(order):
        th_d[k, :14] = torch.from_numpy(thetas[j])
    torch.cuda.synchronize()
    for k in range(n):
        like.eval_device(th_d[k].data_ptr(), out_d[k].data_ptr(), stream.cuda_stream)
    stream.synchronize()
    got = out_d.cpu().numpy()
    for k, j in enumerate(order):
        assert np.array_equal(got[k], alone[j], equal_nan=True), (k, j)
    like.close()


def test_eval_device_under_stream_capture(small):
    import torch
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    like = Hyperlikelihood(*small.as_args())
    th_d = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64, device="cuda")
    out_d = torch.zeros(_lib.OUT_HEADER + like.nobs, dtype=torch.float64, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):   # global capture mode: any allocation or foreign event inside would abort the capture
        like.eval_device(th_d.data_ptr(), out_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    for th in _thetas():
        th_d[:14].copy_(torch.from_numpy(th))
        g.replay()
        torch.cuda.synchronize()
        got = out_d.cpu().numpy().copy()
        assert np.array_equal(got, like.raw(th), equal_nan=True)   # the library's own graph of the same kernels
    # the host call still works after the capture, and the captured graph after the host call
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out_d.cpu().numpy(), like.raw(_thetas()[-1]), equal_nan=True)
    del g
    like.close()


@pytest.mark.parametrize("captured", (False, True))
def test_two_emulated_ranks_through_the_device_entries(small, captured):
    import torch
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog
    world = 2
    full = Hyperlikelihood(*small.as_args())
    ranks = [Hyperlikelihood(*shard_catalog(small.as_args(), r, world)) for r in range(world)]
    th_d = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(world * _lib.PARTIAL_LEN, dtype=torch.float64, device="cuda")
    neff = [torch.zeros(max(r.nobs, 1), dtype=torch.float64, device="cuda") for r in ranks]
    hdr = torch.zeros(_lib.OUT_HEADER, dtype=torch.float64, device="cuda")

    def enqueue(s):
        for r, like in enumerate(ranks):   # each "rank" writes its partial straight into its slot of the gather buffer
            like.partial_device(th_d.data_ptr(), gathered.data_ptr() + 8 * r * _lib.PARTIAL_LEN, neff[r].data_ptr(), s)
        ranks[0].finalize_device(gathered.data_ptr(), world, hdr.data_ptr(), s)

    graph = None
    if captured:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            enqueue(torch.cuda.current_stream().cuda_stream)
    for th in _thetas():
        th_d[:14].copy_(torch.from_numpy(th))
        if captured:
            graph.replay()
        else:
            enqueue(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = full(th)
        got = hdr.cpu().numpy()
        gs = max(1.0, float(np.max(np.abs(ref.dloglike))))
        a = _flat(got)
        b = np.concatenate([[ref.loglike, ref.log_mu_sel, ref.log_mu2, ref.neff_sel], ref.dloglike, ref.dlog_mu_sel,
                            [ref.nvalid_evt, ref.nvalid_sel, ref.nobs, ref.nsel]])
        floor = np.ones_like(b)
        floor[4:18] = gs
        assert np.all(np.abs(a - b) <= 1e-12 * np.maximum(np.abs(b), floor)), float(np.max(np.abs(a - b)))
        ne = np.concatenate([neff[r].cpu().numpy()[:ranks[r].nobs] for r in range(world)])
        assert np.allclose(ne, ref.neff, rtol=1e-12)
    del graph
    for like in ranks + [full]:
        like.close()


def test_capture_is_refused_when_the_constant_slot_is_shared(small):
    """Five contexts on one device: the fifth shares a constant-bank slot with the first.  Capturing an evaluation of
    either must fail loudly (its replays could not be ordered against the other context's evaluations)."""
    import torch
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    likes = [Hyperlikelihood(*small.as_args()) for _ in range(5)]
    th_d = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64, device="cuda")
    th_d[:14].copy_(torch.from_numpy(_thetas()[0]))
    out_d = torch.zeros(_lib.OUT_HEADER + likes[0].nobs, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    shared = likes[4]   # the least-used slot was slot 0 again
    err = None
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=s):
            try:
                shared.eval_device(th_d.data_ptr(), out_d.data_ptr(), s.cuda_stream)
            except _lib.BumpError as e:   # raised inside the capture; the (empty) capture itself ends cleanly
                err = e
    except RuntimeError:
        pass
    assert err is not None and err.code == _lib.E_INVALID and "slot" in str(err)
    # eager evaluations of all five still agree
    ref = likes[0].raw(_thetas()[0]).copy()
    for like in likes[1:]:
        assert np.array_equal(like.raw(_thetas()[0]), ref, equal_nan=True)
    for like in likes:
        like.close()


def test_peer_memory_exchange_path_with_a_single_rank(small):
    """The fused exchange code path (partial -> mailbox -> flags -> merge from the mailbox) with one rank attached to
    its own mailbox: no kernel waits for another launch, so it may run on one GPU.  Must equal the plain evaluation
    bit for bit, from the first evaluation on, and report status 0."""
    import ctypes as C
    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    plain = Hyperlikelihood(*small.as_args())
    like = Hyperlikelihood(*small.as_args())
    raw = (C.c_char * 64)()
    _lib.check(like.lib.bump_p2p_export(like._ctx, raw))
    _lib.check(like.lib.bump_p2p_set_timeout(like._ctx, 2.0))
    _lib.check(like.lib.bump_p2p_attach(like._ctx, raw, 1, 0))
    for th in _thetas():
        a, b = like.raw(th).copy(), plain.raw(th).copy()
        assert a[_lib.OUT_STATUS] == 0.0
        assert np.array_equal(a, b, equal_nan=True), np.flatnonzero(a != b)[:10]
    _lib.check(like.lib.bump_p2p_detach(like._ctx))
    assert np.array_equal(like.raw(_thetas()[0]), plain.raw(_thetas()[0]), equal_nan=True)
    like.close()
    plain.close()


def test_clones_share_the_resident_catalog_and_evaluate_independently(small):
    """bump_ctx_clone: four chains on one upload.  Every clone returns what the original returns alone (bitwise), also
    when all of them evaluate different theta from concurrent host threads, and the catalog survives the original."""
    import threading
    from bumpcosmology_b200.likelihood import Hyperlikelihood
    first = Hyperlikelihood(*small.as_args())
    likes = [first] + [first.clone() for _ in range(3)]
    thetas = _thetas()
    alone = [first.raw(th).copy() for th in thetas]
    errors = []

    def work(i):
        for _ in range(100):
            if not np.array_equal(likes[i].raw(thetas[i]), alone[i], equal_nan=True):
                errors.append(i)
                return

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    first.close()                                            # the clones keep the shared columns alive
    assert np.array_equal(likes[2].raw(thetas[0]), alone[0], equal_nan=True)
    for like in likes[1:]:
        like.close()
