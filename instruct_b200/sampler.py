"""Host-side mirror of the reference's sampler interface, over the C-ABI.

``mcmc_updating(data, initial, chn, cvg)`` has the argument meaning of the reference's
``CHAIN mcmc_updating(SEQDATA data, INIT initial, int chn, CONVG *cvg)`` (mcmc.h:56,
mcmc.c:63-87): ``SeqData`` / ``Init`` / ``Convg`` / ``Chain`` carry the same-named fields of
SEQDATA (data_interface.h:10-56), INIT (initial.h:9-21), CONVG (check_converg.h:10-18) and
CHAIN (mcmc.h:29-53).  ``Sampler`` is the finer-grained handle the tests and bench.py use.
All computation happens in libinstruct_b200.so on the GPU; numpy here only carries buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import IgChainResult, IgConfig, InstructError, check


@dataclass
class SeqData:
    """The SEQDATA fields mcmc_updating() reads.  ``seqdata`` is the NEW packed genotype
    store: int16 [locinum][totalsize][ploid], negative = missing."""
    seqdata: np.ndarray
    allelenum: np.ndarray
    popnum: int
    ploid: int = 2
    mode: int = 2
    prior_flag: int = 0
    back_refl: int = 1
    type_freq: int = 1
    alpha_dpm: float = 10.0
    nstep_check_empty_cluster: int = 20
    print_iter: int = 0
    print_freq: int = 0
    autopoly: int = 1

    @property
    def locinum(self):
        return int(self.seqdata.shape[0])

    @property
    def totalsize(self):
        return int(self.seqdata.shape[1])

    @property
    def allelenum_max(self):
        return int(np.max(self.allelenum))

    @property
    def missindx(self):
        """[L][N] uint8: 1 if any copy is missing (get_missing, data_interface.c:812-835); for
        ploid 4, 1 if no allele was observed (get_missing_tetra, data_interface.c:722-741)."""
        if self.ploid == 4:
            return (self.seqdata < 0).all(axis=2).astype(np.uint8)
        return (self.seqdata < 0).any(axis=2).astype(np.uint8)


@dataclass
class Init:
    """INIT (initial.h:9-21)."""
    update: int
    burnin: int
    thinning: int
    chainnum: int = 1
    initd: np.ndarray | None = None          # float32 [chainnum][popnum]
    chn_name: list = field(default_factory=list)

    def name(self, chn):
        return self.chn_name[chn] if chn < len(self.chn_name) else f"Chain#{chn + 1}"


@dataclass
class Convg:
    """CONVG (check_converg.h:10-18): the first ``ckrep`` retained log-likelihoods per chain."""
    n_chain: int
    ckrep: int
    convg_ld: np.ndarray = None

    def __post_init__(self):
        if self.convg_ld is None:
            self.convg_ld = np.zeros(self.n_chain * max(self.ckrep, 1), dtype=np.float64)


@dataclass
class Chain:
    """CHAIN (mcmc.h:29-53): running first and second moments of the retained sweeps."""
    steps: int
    step: int
    chn_name: str
    flag_empty_cluster: int
    totallkh: float
    totallkh2: float
    indvlkh: np.ndarray
    qq: np.ndarray
    qq2: np.ndarray
    self_rates: np.ndarray
    self_rates2: np.ndarray
    gen: np.ndarray
    gen2: np.ndarray
    freq: np.ndarray | None = None
    freq2: np.ndarray | None = None


def _config(data: SeqData, update=1, burnin=1, thinning=1, ckrep=0, seed=1, device=0, shard_rank=0, shard_count=1,
            rng_rounds=0, totalsize=None, use_graph=0) -> IgConfig:
    cfg = IgConfig()
    cfg.ploid, cfg.popnum, cfg.locinum = data.ploid, data.popnum, data.locinum
    cfg.totalsize = data.totalsize if totalsize is None else totalsize
    cfg.mode, cfg.prior_flag, cfg.back_refl, cfg.type_freq = data.mode, data.prior_flag, data.back_refl, data.type_freq
    cfg.alpha_dpm = data.alpha_dpm
    cfg.nstep_check_empty_cluster = data.nstep_check_empty_cluster
    cfg.print_iter, cfg.print_freq, cfg.autopoly = data.print_iter, data.print_freq, data.autopoly
    cfg.update, cfg.burnin, cfg.thinning, cfg.ckrep = update, burnin, thinning, ckrep
    cfg.seed = seed
    cfg.device = device
    cfg.shard_rank, cfg.shard_count = shard_rank, shard_count
    cfg.shard_begin, cfg.shard_size = 0, 0
    cfg.rng_rounds = rng_rounds
    cfg.use_graph = use_graph
    return cfg


class _Result:
    """Caller-owned CHAIN buffers for ig_run_chain."""

    def __init__(self, N, K, ns, L=0, A=0, print_freq=0):
        self.indvlkh = np.zeros(N)
        self.qq, self.qq2 = np.zeros((N, K)), np.zeros((N, K))
        self.self_rates, self.self_rates2 = np.zeros(ns), np.zeros(ns)
        self.gen, self.gen2 = np.zeros(N), np.zeros(N)
        self.freq = np.zeros((K, L, A)) if print_freq else None
        self.freq2 = np.zeros((K, L, A)) if print_freq else None
        r = IgChainResult()
        for name in ("indvlkh", "qq", "qq2", "self_rates", "self_rates2", "gen", "gen2", "freq", "freq2"):
            arr = getattr(self, name)
            setattr(r, name, arr.ctypes.data if arr is not None else None)
        self.c = r

    def chain(self, name):
        r = self.c
        return Chain(steps=r.steps, step=r.step, chn_name=name, flag_empty_cluster=r.flag_empty_cluster,
                     totallkh=r.totallkh, totallkh2=r.totallkh2, indvlkh=self.indvlkh, qq=self.qq, qq2=self.qq2,
                     self_rates=self.self_rates, self_rates2=self.self_rates2, gen=self.gen, gen2=self.gen2,
                     freq=self.freq, freq2=self.freq2)


def mcmc_updating(data: SeqData, initial: Init, chn: int, cvg: Convg | None, seed: int = 1, device: int = 0) -> Chain:
    """Drop-in for the reference's mcmc_updating(): one chain, host buffers in, CHAIN out.
    Host-to-device copy of the genotype store and device-to-host copy of the moments are
    inside this call (it is what bench.py times as ``e2e``)."""
    lib = _lib.load()
    ckrep = cvg.ckrep if cvg is not None else 0
    cfg = _config(data, initial.update, initial.burnin, initial.thinning, ckrep, seed, device)
    x = np.ascontiguousarray(data.seqdata, dtype=np.int16)
    an = np.ascontiguousarray(data.allelenum, dtype=np.int32)
    ns = data.totalsize if (data.mode in (3, 5) and data.ploid == 2) else (0 if (data.mode <= 1 and data.ploid == 2) else data.popnum)
    res = _Result(data.totalsize, data.popnum, ns, data.locinum, data.allelenum_max, data.print_freq)
    initd = None
    if initial.initd is not None:
        initd = np.ascontiguousarray(initial.initd[chn], dtype=np.float32)
    cv = (C.c_double * max(ckrep, 1))()
    st = lib.ig_mcmc_updating(C.byref(cfg), x.ctypes.data, an.ctypes.data, chn,
                              initd.ctypes.data if initd is not None else None, C.byref(res.c), cv)
    check(st, (_lib.IG_OK, _lib.IG_EMPTY_CLUSTER))
    if cvg is not None and ckrep > 0:
        cvg.convg_ld[chn * ckrep:(chn + 1) * ckrep] = np.frombuffer(cv, dtype=np.float64)[:ckrep]
    return res.chain(initial.name(chn))


class Sampler:
    """A prepared context: genotype store resident in HBM, one chain at a time."""

    def __init__(self, data: SeqData, update=1, burnin=1, thinning=1, ckrep=0, seed=1, device=0,
                 shard_rank=0, shard_count=1, rng_rounds=0, totalsize=None, x_device_ptr=None,
                 allelenum_device_ptr=None, use_graph=0):
        self.lib = _lib.load()
        self.data = data
        self.cfg = _config(data, update, burnin, thinning, ckrep, seed, device, shard_rank, shard_count, rng_rounds,
                           totalsize, use_graph)
        self.N = self.cfg.totalsize
        self.K, self.L = data.popnum, data.locinum
        self.h = C.c_void_p()
        check(self.lib.ig_create(C.byref(self.cfg), C.byref(self.h)))
        cap = -(-self.N // shard_count)
        self.i0 = shard_rank * cap
        self.Nloc = min(cap, self.N - self.i0)
        self.ns = self.N if (data.mode in (3, 5) and data.ploid == 2) else (0 if (data.mode <= 1 and data.ploid == 2) else self.K)
        self.ploid = data.ploid
        if x_device_ptr is not None:
            check(self.lib.ig_load_genotypes_device(self.h, x_device_ptr, allelenum_device_ptr))
            self.A = None
        else:
            x = np.ascontiguousarray(data.seqdata, dtype=np.int16)
            if x.shape[1] != self.Nloc:
                raise InstructError(f"shard holds {self.Nloc} individuals, seqdata has {x.shape[1]}")
            an = np.ascontiguousarray(data.allelenum, dtype=np.int32)
            check(self.lib.ig_load_genotypes(self.h, x.ctypes.data, an.ctypes.data))
        geo = self.geometry()
        self.A = geo["A"]

    def close(self):
        if self.h:
            self.lib.ig_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-GPU -------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_char * 128)()
        check(_lib.load().ig_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid: bytes):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        check(self.lib.ig_comm_init(self.h, buf))

    # -- chain control ---------------------------------------------------------------
    def chain_init(self, chain_id=0, initd=None):
        p = None
        if initd is not None:
            self._initd = np.ascontiguousarray(initd, dtype=np.float32)
            p = self._initd.ctypes.data
        check(self.lib.ig_chain_init(self.h, chain_id, p))

    def sweep(self, n=1):
        check(self.lib.ig_sweep(self.h, n))

    def sync(self):
        check(self.lib.ig_sync(self.h))

    def time_sweeps(self, n):
        """n sweeps bracketed by CUDA events on the library's stream; returns milliseconds."""
        ms = C.c_double()
        check(self.lib.ig_time_sweeps(self.h, n, C.byref(ms)))
        return ms.value

    def run_phase(self, mask):
        check(self.lib.ig_run_phase(self.h, mask))

    def run_chain(self, chain_id=0, initd=None, name=None):
        res = _Result(self.N, self.K, self.ns, self.L, self.A or 0, self.cfg.print_freq)
        p = None
        if initd is not None:
            self._initd = np.ascontiguousarray(initd, dtype=np.float32)
            p = self._initd.ctypes.data
        cv = np.zeros(max(self.cfg.ckrep, 1))
        st = self.lib.ig_run_chain(self.h, chain_id, p, C.byref(res.c), cv.ctypes.data)
        check(st, (_lib.IG_OK, _lib.IG_EMPTY_CLUSTER))
        ch = res.chain(name or f"Chain#{chain_id + 1}")
        return ch, cv[: self.cfg.ckrep]

    def trace(self, update, burnin, thinning=1):
        """Sweeps an initialised chain and keeps the retained draws (same schedule as mcmc_updating: sweep `step` is kept
        when step >= burnin and (step + 1 - burnin) % thinning == 0) of the log-likelihood, the rates S and Q -- the
        input of converge.chain_diagnostics.  One device -> host read per retained sweep: a diagnostic, not the hot path."""
        ll, S, Q = [], [], []
        for step in range(update):
            self.sweep(1)
            if step >= burnin and (step + 1 - burnin) % thinning == 0:
                ll.append(float(self.get(_lib.STATE_TOTALLKH)[0])); S.append(self.get(_lib.STATE_S)); Q.append(self.get(_lib.STATE_Q))
        return np.array(ll), np.array(S), np.array(Q)

    # -- state hooks -------------------------------------------------------------------
    def _shape(self, sid):
        L, N, K, Nl, A = self.L, self.N, self.K, self.Nloc, self.A
        pl = self.ploid
        if sid in (_lib.STATE_TABLES, _lib.STATE_TABLES_PROP, _lib.STATE_EXFREQ):
            return (K, L, self.gmax()), np.float32
        return {
            _lib.STATE_X: ((L, Nl, pl), np.int16), _lib.STATE_Z: ((L, Nl, pl), np.int8),
            _lib.STATE_GENO: ((L, Nl, 4), np.int8), _lib.STATE_SPROP: ((K,), np.float64),
            _lib.STATE_DSTAT: ((K,), np.float64), _lib.STATE_GMAX: ((1,), np.int32),
            _lib.STATE_Q: ((N, K), np.float64), _lib.STATE_P: ((K, L, A), np.float64),
            _lib.STATE_ALPHA: ((1,), np.float64), _lib.STATE_S: ((self.ns,), np.float64),
            _lib.STATE_G: ((N,), np.int32), _lib.STATE_INDVLKH: ((N,), np.float64),
            _lib.STATE_TOTALLKH: ((1,), np.float64), _lib.STATE_TALLY: ((K, L, A), np.int32),
            _lib.STATE_CNT: ((Nl, K), np.int32), _lib.STATE_GPROP: ((N,), np.int32),
            _lib.STATE_STATE: ((K,), np.int32), _lib.STATE_MASK: ((L, Nl), np.uint8),
            _lib.STATE_ITER: ((1,), np.int64), _lib.STATE_LLPARTS: ((Nl, 4), np.float64),
            _lib.STATE_GEOMETRY: ((8,), np.int32), _lib.STATE_FPROP: ((self.ns,), np.float64),
            _lib.STATE_FK: ((N, 2, K), np.float64),
            _lib.STATE_DPWEIGHTS: ((N, N + 1), np.float64), _lib.STATE_DPCLUSTERS: ((1,), np.int64),
            _lib.STATE_P2: ((K, L, A), np.float64), _lib.STATE_TALLY2: ((K, L, A), np.int32),
        }[sid]

    def get(self, sid):
        shape, dt = self._shape(sid)
        a = np.zeros(shape, dtype=dt)
        check(self.lib.ig_get_state(self.h, sid, a.ctypes.data, a.nbytes))
        return a

    def set(self, sid, value):
        shape, dt = self._shape(sid)
        a = np.ascontiguousarray(np.asarray(value, dtype=dt).reshape(shape))
        check(self.lib.ig_set_state(self.h, sid, a.ctypes.data, a.nbytes))

    def gmax(self):
        a = np.zeros(1, dtype=np.int32)
        check(self.lib.ig_get_state(self.h, _lib.STATE_GMAX, a.ctypes.data, a.nbytes))
        return int(a[0])

    def refresh_tables(self):
        """ploid 4: recompute exfreq and both selfing tables from the current P, S and S'."""
        a = np.zeros(1, dtype=np.int32)
        check(self.lib.ig_set_state(self.h, _lib.STATE_TABLES, a.ctypes.data, a.nbytes))

    def geometry(self):
        a = np.zeros(8, dtype=np.int32)
        check(self.lib.ig_get_state(self.h, _lib.STATE_GEOMETRY, a.ctypes.data, a.nbytes))
        return dict(TL=int(a[0]), nchunks=int(a[1]), nblk=int(a[2]), subs_per_blk=int(a[3]), R=int(a[4]),
                    smem=int(a[5]), KP=int(a[6]), A=int(a[7]))

    def loglik(self, gen):
        g = np.ascontiguousarray(gen, dtype=np.int32)
        out = np.zeros(self.Nloc)
        check(self.lib.ig_loglik(self.h, g.ctypes.data, out.ctypes.data))
        return out

    def proposal_loglik(self, S):
        s = np.ascontiguousarray(S, dtype=np.float64)
        out = np.zeros(1)
        check(self.lib.ig_proposal_loglik(self.h, s.ctypes.data, out.ctypes.data))
        return float(out[0])

    def alpha_logratio(self, ralpha):
        out = np.zeros(1)
        check(self.lib.ig_alpha_logratio(self.h, float(ralpha), out.ctypes.data))
        return float(out[0])

    # -- profiling -----------------------------------------------------------------------
    def profile(self, enable=True):
        check(self.lib.ig_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        n = C.c_int32()
        ms = C.c_double()
        k = C.c_int64()
        check(self.lib.ig_profile_read(self.h, C.byref(n), C.byref(ms), C.byref(k)))
        return n.value, ms.value, k.value

    def algorithmic_bytes(self):
        b, c = C.c_double(), C.c_double()
        check(self.lib.ig_algorithmic_bytes(self.h, C.byref(b), C.byref(c)))
        return b.value, c.value
