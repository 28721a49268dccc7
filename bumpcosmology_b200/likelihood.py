"""Host side of the hot path: the device-resident catalog and one-call evaluation of

    theta -> loglike, log_mu_sel, log_mu2, neff_sel, neff[nobs], d loglike/d theta, d log_mu_sel/d theta

i.e. the body of the reference's `pop_cosmo_model` (/root/reference/src/scripts/intensity_models.py:374-394,401)
plus its reverse pass, computed by the CUDA library behind include/bump.h.  No CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

THETA_NAMES = ("h", "Om", "w", "a", "b", "c", "mpisn", "mbhmax", "sigma", "fpl", "beta", "lam", "kappa", "zp")


@dataclass
class Evaluation:
    """Result of one evaluation; field names follow the reference's sites (intensity_models.py:383-401)."""
    loglike: float          # numpyro.factor('loglike', .)              :383
    log_mu_sel: float       # :389   ('selfactor' = -nobs * log_mu_sel, :390)
    log_mu2: float          # :392
    neff_sel: float         # :394
    neff: np.ndarray        # :401  (this rank's events)
    dloglike: np.ndarray    # d loglike / d theta      [14] (+ wa)
    dlog_mu_sel: np.ndarray  # d log_mu_sel / d theta  [14] (+ wa)
    nobs: int
    nsel: int
    nvalid_evt: int
    nvalid_sel: int

    @property
    def selfactor(self):
        return -self.nobs * self.log_mu_sel

    @property
    def logl(self):
        return self.loglike + self.selfactor

    @property
    def dlogl(self):
        return self.dloglike - self.nobs * self.dlog_mu_sel


def _c64(x):
    if hasattr(x, "to_numpy"):   # pandas Series, as run_cosmo_fit.py:47-49 passes them
        x = x.to_numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of n items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_catalog(data, rank, world):
    """This rank's share of the 9 model arguments: whole events (never split an event) and a contiguous
    range of injections.  Ndraw stays the TOTAL number of draws."""
    m1, q, dl, pd, m1s, qs, dls, pds, ndraw = data
    m1, q, dl, pd = (_c64(x) for x in (m1, q, dl, pd))
    m1s, qs, dls, pds = (_c64(x) for x in (m1s, qs, dls, pds))
    e0, e1 = shard_bounds(m1.shape[0], rank, world)
    s0, s1 = shard_bounds(m1s.shape[0], rank, world)
    return (m1[e0:e1], q[e0:e1], dl[e0:e1], pd[e0:e1], m1s[s0:s1], qs[s0:s1], dls[s0:s1], pds[s0:s1], ndraw)


def unpack_header(out, ntheta):
    h = out[:_lib.OUT_HEADER].tolist()   # one conversion to Python floats (this runs once per evaluation)
    return dict(loglike=h[_lib.OUT_LOGLIKE], log_mu_sel=h[_lib.OUT_LOG_MU_SEL], log_mu2=h[_lib.OUT_LOG_MU2],
                neff_sel=h[_lib.OUT_NEFF_SEL],
                dloglike=out[_lib.OUT_DLOGLIKE:_lib.OUT_DLOGLIKE + ntheta].copy(),
                dlog_mu_sel=out[_lib.OUT_DLOG_MU:_lib.OUT_DLOG_MU + ntheta].copy(),
                nobs=int(h[_lib.OUT_NOBS]), nsel=int(h[_lib.OUT_NSEL]),
                nvalid_evt=int(h[_lib.OUT_NVALID_EVT]), nvalid_sel=int(h[_lib.OUT_NVALID_SEL]))


class Hyperlikelihood:
    """The catalog resident in HBM on one GPU (one rank's shard) and its evaluator.

    Arguments are the 9 positional arguments of the reference's `pop_cosmo_model`
    (intensity_models.py:357): four [nobs, nsamp] arrays, four [nsel] arrays, Ndraw.
    """

    def __init__(self, m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, device=0,
                 wa=False, graph=True, sort=True, fixed_dvdzdt=None):
        self.lib = _lib.load()
        self.wa = bool(wa)
        self.ntheta = _lib.NTHETA_MAX if wa else _lib.NTHETA
        self.device = int(device)
        self._ctx = C.c_void_p()
        self.fixed = fixed_dvdzdt is not None
        flags = ((_lib.FLAG_WA if wa else 0) | (0 if graph else _lib.FLAG_NO_GRAPH)
                 | (0 if sort else _lib.FLAG_NO_SORT) | (_lib.FLAG_FIXED_COSMO if self.fixed else 0))
        _lib.check(self.lib.bump_ctx_create(C.byref(self._ctx), self.device, flags))
        if self.fixed:   # pop_model (intensity_models.py:313-355): arguments are (m1s, qs, zs, pdraw, ...) source frame
            tab = _c64(fixed_dvdzdt).ravel()
            _lib.check(self.lib.bump_set_fixed_dvdzdt(self._ctx, _lib.as_dp(tab), tab.shape[0]))
        ev = [_c64(x) for x in (m1s_det, qs, dls, pdraw)]
        if ev[0].ndim == 1:
            ev = [x.reshape(1, -1) if x.size else x.reshape(0, 0) for x in ev]
        if any(x.shape != ev[0].shape for x in ev):
            raise ValueError("event arrays must share one [nobs, nsamp] shape")
        sel = [_c64(x).ravel() for x in (m1s_det_sel, qs_sel, dls_sel, pdraw_sel)]
        if any(x.shape != sel[0].shape for x in sel):
            raise ValueError("injection arrays must share one [nsel] shape")
        self.nobs, self.nsamp = (int(ev[0].shape[0]), int(ev[0].shape[1])) if ev[0].size else (0, 0)
        self.nsel = int(sel[0].shape[0])
        self.Ndraw = float(Ndraw)
        _lib.check(self.lib.bump_upload_events(self._ctx, self.nobs, self.nsamp, *[_lib.as_dp(x) for x in ev]))
        _lib.check(self.lib.bump_upload_injections(self._ctx, self.nsel, *[_lib.as_dp(x) for x in sel],
                                                   self.Ndraw))
        self._out = np.empty(int(self.lib.bump_out_len(self._ctx)), dtype=np.float64)
        self._theta = np.zeros(_lib.NTHETA_MAX, dtype=np.float64)
        self._out_p, self._theta_p = _lib.as_dp(self._out), _lib.as_dp(self._theta)
        self.plan()   # builds the execution plan now: every later call (also inside a stream capture) allocates nothing

    def clone(self):
        """A second evaluator on the SAME resident catalog (bump_ctx_clone): no upload, no second copy in HBM; theta,
        tables, records and results are the clone's own, so the two evaluate concurrently (one per NUTS chain)."""
        new = object.__new__(Hyperlikelihood)
        for k in ("lib", "wa", "ntheta", "device", "fixed", "nobs", "nsamp", "nsel", "Ndraw"):
            setattr(new, k, getattr(self, k))
        new._ctx = C.c_void_p()
        _lib.check(self.lib.bump_ctx_clone(self._ctx, C.byref(new._ctx)))
        new._out = np.empty(int(self.lib.bump_out_len(new._ctx)), dtype=np.float64)
        new._theta = np.zeros(_lib.NTHETA_MAX, dtype=np.float64)
        new._out_p, new._theta_p = _lib.as_dp(new._out), _lib.as_dp(new._theta)
        new.plan()
        return new

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.bump_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _set_theta(self, theta):
        th = np.asarray(theta, dtype=np.float64).ravel()
        if th.shape[0] != self.ntheta:
            raise ValueError(f"theta must have {self.ntheta} entries {THETA_NAMES}{' + wa' if self.wa else ''}")
        self._theta[:self.ntheta] = th
        return self._theta

    # -- evaluation through HOST buffers (the reference-facing call: h2d of theta, d2h of the result inside)
    def __call__(self, theta):
        if np.shape(theta) != (self.ntheta,):
            self._set_theta(theta)           # raises with the list of parameter names (or accepts a column vector)
        else:
            self._theta[:self.ntheta] = theta
        code = self.lib.bump_eval(self._ctx, self._theta_p, self._out_p)
        if code:
            _lib.check(code)
        out = self._out
        return Evaluation(neff=out[_lib.OUT_HEADER:].copy(), **unpack_header(out, self.ntheta))

    def raw(self, theta):
        """Same evaluation, no wrapping: returns the library's output vector (a view that the next call overwrites):
        [0:40] = header (include/bump.h BUMP_OUT_*), [40:] = neff.  The sampler's inner loop uses this."""
        th = self._theta
        th[:self.ntheta] = theta
        _lib.check(self.lib.bump_eval(self._ctx, self._theta_p, self._out_p))
        return self._out

    # -- pieces for multi-rank drivers
    def partial(self, theta):
        """This rank's partial (host) and neff; merge with `merge_partials`."""
        th = self._set_theta(theta)
        part = np.empty(_lib.PARTIAL_LEN)
        neff = np.empty(self.nobs)
        _lib.check(self.lib.bump_eval_partial(self._ctx, _lib.as_dp(th), _lib.as_dp(part), _lib.as_dp(neff)))
        return part, neff

    def eval_device(self, theta_ptr, out_ptr, stream_ptr):
        _lib.check(self.lib.bump_eval_device(self._ctx, theta_ptr, out_ptr, stream_ptr))

    def partial_device(self, theta_ptr, partial_ptr, neff_ptr, stream_ptr):
        _lib.check(self.lib.bump_eval_partial_device(self._ctx, theta_ptr, partial_ptr, neff_ptr, stream_ptr))

    def finalize_device(self, partials_ptr, nranks, out_ptr, stream_ptr):
        _lib.check(self.lib.bump_finalize_device(self._ctx, partials_ptr, nranks, out_ptr, stream_ptr))

    def attach_nccl(self, unique_id, nranks, rank):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _lib.check(self.lib.bump_comm_attach(self._ctx, buf, nranks, rank))

    # -- introspection / timing
    def tables(self):
        """theta-dependent tables of the last evaluation (unit-level parity of the prologue)."""
        def get(which, n):
            buf = np.empty(n)
            _lib.check(self.lib.bump_debug_tables(self._ctx, which, _lib.as_dp(buf), n))
            return buf
        cos = get(0, 4 * 1024).reshape(4, 1024)
        tan = get(1, 9 * 1024).reshape(3, 3, 1024)
        g = get(2, 6 * 256).reshape(6, 256)
        return dict(zinterp=cos[0], dlinterp=cos[1], ddlinterp=cos[2], dvcinterp=cos[3],
                    d_dl=tan[0], d_ddl=tan[1], d_dvc=tan[2], log_dN_grid=g[0], d_log_dN_grid=g[1:],
                    scalars=get(3, 64))

    def time_evals(self, theta, iters, kernel=False):
        """(total ms for `iters` graph replays, ms in the streaming kernel alone or None); CUDA events on the
        context's stream."""
        th = self._set_theta(theta)
        tot, ker = C.c_float(), C.c_float()
        _lib.check(self.lib.bump_time_evals(self._ctx, _lib.as_dp(th), int(iters), C.byref(tot),
                                            C.byref(ker) if kernel else None))
        return tot.value, (ker.value if kernel else None)

    def timeline(self, theta):
        """Per-kernel timeline of one evaluation launched directly (no graph): {kernel: (start_us, end_us)} on the
        GPU's global timer, relative to the first kernel's first block."""
        th = self._set_theta(theta)
        names = ("prologue", "stream", "epilogue", "finalize", "prologue.rows", "prologue.cosmology", "prologue.last_block",
                 "stream.staged", "epilogue.blocks", "epilogue.last_block", "stream.warps_done")
        buf = np.empty(2 * len(names))
        _lib.check(self.lib.bump_debug_timeline(self._ctx, _lib.as_dp(th), _lib.as_dp(buf), buf.shape[0]))
        return {n: (float(buf[2 * i]), float(buf[2 * i + 1])) for i, n in enumerate(names) if buf[2 * i + 1] >= 0}

    def plan(self):
        info = (C.c_int64 * 8)()
        _lib.check(self.lib.bump_plan_info(self._ctx, info))
        keys = ("groups", "groups_per_warp", "records", "grid", "threads", "smem_bytes", "padded_samples", "sms")
        return dict(zip(keys, [int(v) for v in info]))

    @property
    def launches_per_eval(self):
        return int(self.lib.bump_launches_per_eval(self._ctx))


def merge_partials(partials):
    """Rank-ordered merge of host partials ([nranks, PARTIAL_LEN]) into the result header (same code as the
    device finalize kernel)."""
    lib = _lib.load()
    p = np.ascontiguousarray(np.asarray(partials, dtype=np.float64).reshape(-1, _lib.PARTIAL_LEN))
    out = np.empty(_lib.OUT_HEADER)
    _lib.check(lib.bump_merge_partials(_lib.as_dp(p), p.shape[0], _lib.as_dp(out)))
    return out


class ShardedHyperlikelihood:
    """One process per GPU (torch.distributed, backend nccl): events and injections sharded over the ranks,
    one all-gather of the 1 KiB per-rank partial per evaluation, identical merged result on every rank.

    exchange = 'torch' : torch.distributed.all_gather_into_tensor between the partial and finalize launches
               'nccl'  : ncclAllGather issued by the library inside its CUDA graph (bump_comm_attach)
               'p2p'   : no collective call at all — the epilogue kernel's last block writes the partial into every
                         peer's mailbox over NVLink (cudaIpc-mapped peer memory) and merges (bump_p2p_attach)
    """

    def __init__(self, data, device=None, wa=False, exchange="p2p", group=None, timeout_s=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.local = Hyperlikelihood(*shard_catalog(data, self.rank, self.world), device=self.device, wa=wa)
        self.ntheta = self.local.ntheta
        self.exchange = exchange
        dev = torch.device("cuda", self.device)
        self._theta_h = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64).pin_memory()
        self._theta = torch.zeros(_lib.NTHETA_MAX, dtype=torch.float64, device=dev)
        self._partial = torch.zeros(_lib.PARTIAL_LEN, dtype=torch.float64, device=dev)
        self._gathered = torch.zeros(self.world * _lib.PARTIAL_LEN, dtype=torch.float64, device=dev)
        self._out = torch.zeros(_lib.OUT_HEADER + self.local.nobs, dtype=torch.float64, device=dev)
        self._out_h = torch.zeros(_lib.OUT_HEADER + self.local.nobs, dtype=torch.float64).pin_memory()
        if exchange == "p2p":
            # peer mailboxes need CUDA IPC between the ranks' processes; if any rank cannot map its peers, every rank
            # falls back to the in-library NCCL all-gather
            lib, ctx = self.local.lib, self.local._ctx
            raw = (C.c_char * 64)()
            ok = lib.bump_p2p_export(ctx, raw) == 0
            mine = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone().to(dev)
            allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            hb = allh.cpu().numpy().tobytes()
            buf = (C.c_char * len(hb)).from_buffer_copy(hb)
            if timeout_s is not None:   # how long a rank waits for a peer's partial before the evaluation fails everywhere
                ok = ok and lib.bump_p2p_set_timeout(ctx, float(timeout_s)) == 0
            ok = ok and lib.bump_p2p_attach(ctx, buf, self.world, self.rank) == 0
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                lib.bump_p2p_detach(ctx)
                self.exchange = exchange = "nccl"
        if exchange == "nccl":
            idbuf = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                raw = (C.c_char * 128)()
                _lib.check(self.local.lib.bump_nccl_unique_id(raw))
                idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
            idbuf = idbuf.to(dev)
            dist.broadcast(idbuf, src=0, group=group)
            self.local.attach_nccl(idbuf.cpu().numpy().tobytes(), self.world, self.rank)

    def launch(self, stream=None):
        """Enqueue one evaluation (theta already in self._theta) on the current torch stream."""
        torch = self.torch
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if self.exchange in ("nccl", "p2p"):
            self.local.eval_device(self._theta.data_ptr(), self._out.data_ptr(), s)
        else:
            self.local.partial_device(self._theta.data_ptr(), self._partial.data_ptr(),
                                      self._out.data_ptr() + 8 * _lib.OUT_HEADER, s)
            self.dist.all_gather_into_tensor(self._gathered, self._partial, group=self.group)
            self.local.finalize_device(self._gathered.data_ptr(), self.world, self._out.data_ptr(), s)

    def raw(self, theta):
        return self.local.raw(theta) if self.exchange in ("nccl", "p2p") else None

    def close(self):
        self.local.close()

    def __call__(self, theta):
        """A failed peer-memory exchange (a rank did not arrive within the timeout) raises BumpError with code
        _lib.E_EXCHANGE on EVERY rank - never a result on some ranks and an error on others."""
        if self.exchange in ("nccl", "p2p"):   # the exchange lives inside the library's CUDA graph
            return self.local(theta)
        th = np.asarray(theta, dtype=np.float64).ravel()
        self._theta_h[:self.ntheta] = self.torch.from_numpy(th)
        self._theta.copy_(self._theta_h, non_blocking=True)
        self.launch()
        self._out_h.copy_(self._out, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        out = self._out_h.numpy()
        return Evaluation(neff=out[_lib.OUT_HEADER:].copy(), **unpack_header(out, self.ntheta))
