"""ctypes bindings for the CHECKERS: the in-repo CPU restatement (liboracle.so) and the
include-the-.c harness around the unmodified reference (oracle/_ref/libinstruct_ref.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from instruct_b200/ (the product).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libinstruct_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "InStruct")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)


def build(quiet=True):
    """Compile liboracle.so (always) and, when /root/reference is present, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_int] * 8 + [C.c_double, C.c_void_p, C.c_void_p]
        for name, rt in [("orc_z", C.c_void_p), ("orc_qq", c_dp), ("orc_qqnum", c_dp), ("orc_freq", c_dp),
                         ("orc_self", c_dp), ("orc_state", c_ip), ("orc_gen", c_ip), ("orc_indvlkh", c_dp),
                         ("orc_alpha", c_dp), ("orc_totallkh", c_dp)]:
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_amax.argtypes = [C.c_void_p]
        L.orc_setseeds.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long]
        L.orc_getseeds.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.orc_ran1.restype = C.c_double
        L.orc_ran1.argtypes = [C.c_void_p]
        L.orc_rgamma.restype = C.c_double
        L.orc_rgamma.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_rbeta.restype = C.c_double
        L.orc_rbeta.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_rnormal.restype = C.c_double
        L.orc_rnormal.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_rgeom.argtypes = [C.c_void_p, C.c_double]
        L.orc_disc_unif.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.orc_missing_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tally.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tally_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_count_z.argtypes = [C.c_void_p, c_dp]
        L.orc_genofreq.restype = C.c_double
        L.orc_genofreq.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_log_ld_indv.restype = C.c_double
        L.orc_log_ld_indv.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_proposal.restype = C.c_double
        L.orc_proposal.argtypes = [C.c_void_p, c_dp]
        L.orc_dgeom.restype = C.c_double
        L.orc_dgeom.argtypes = [C.c_double, C.c_int]
        L.orc_dt_stat.argtypes = [C.c_double]
        L.orc_alpha_logratio.restype = C.c_double
        L.orc_alpha_logratio.argtypes = [C.c_void_p, C.c_double]
        L.orc_alpha_ratio_product.restype = C.c_double
        L.orc_alpha_ratio_product.argtypes = [C.c_void_p, C.c_double]
        L.orc_check_empty_cluster.argtypes = [C.c_void_p]
        L.orc_z_conditional.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]
        L.orc_zz.restype = c_ip
        L.orc_zz.argtypes = [C.c_void_p]
        L.orc_update_Z.argtypes = [C.c_void_p, C.c_int]
        L.orc_log_ld_indv_K.restype = C.c_double
        L.orc_log_ld_indv_K.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_log_ld_F.restype = C.c_double
        L.orc_log_ld_F.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int]
        L.orc_log_ld_F_total.restype = C.c_double
        L.orc_log_ld_F_total.argtypes = [C.c_void_p, c_dp]
        for name in ["orc_update_P", "orc_update_S_POP", "orc_update_S_IND", "orc_update_F_POP", "orc_update_F_IND", "orc_update_G", "orc_update_alpha",
                     "orc_cal_lkh", "orc_init_DP", "orc_update_DP"]:
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_update_ZQ.argtypes = [C.c_void_p, C.c_int]
        L.orc_dp_nclusters.argtypes = [C.c_void_p]
        L.orc_dp_from_values.argtypes = [C.c_void_p]
        L.orc_dp_from_values.restype = None
        L.orc_sweeps.argtypes = [C.c_void_p, C.c_int]
        L.orc_chain_new.restype = C.c_void_p
        L.orc_chain_new.argtypes = [C.c_void_p, C.c_int]
        L.orc_chain_free.argtypes = [C.c_void_p]
        L.orc_run_chain.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, c_fp, C.c_void_p]
        L.orc_gelman_rubin_ref.restype = C.c_double
        L.orc_gelman_rubin_ref.argtypes = [c_dp, C.c_int, C.c_int]
        L.orc_gelman_rubin.restype = C.c_double
        L.orc_gelman_rubin.argtypes = [c_dp, C.c_int, C.c_int]
        _oracle_lib = L
    return _oracle_lib


class _ChainStruct(C.Structure):
    _fields_ = [("steps", C.c_long), ("step", C.c_long), ("flag_empty_cluster", C.c_int),
                ("totallkh", C.c_double), ("totallkh2", C.c_double),
                ("indvlkh", c_dp), ("qq", c_dp), ("qq2", c_dp), ("self_rates", c_dp), ("self_rates2", c_dp),
                ("gen", c_dp), ("gen2", c_dp), ("convg", c_dp)]


class Oracle:
    """The CPU restatement on one data set.  ``x`` is int16 [L][N][ploid] (new packed layout)."""

    def __init__(self, x, allelenum, K, mode=2, prior_flag=0, back_refl=1, type_freq=1, alpha_dpm=10.0):
        self.lib = oracle_lib()
        self.x = np.ascontiguousarray(x, dtype=np.int16)
        self.allelenum = np.ascontiguousarray(allelenum, dtype=np.int32)
        self.L, self.N, self.ploid = self.x.shape
        self.K, self.mode = K, mode
        self.h = self.lib.orc_new(self.N, self.L, K, self.ploid, mode, prior_flag, back_refl, type_freq,
                                  float(alpha_dpm), self.x.ctypes.data, self.allelenum.ctypes.data)
        self.Amax = self.lib.orc_amax(self.h)
        ns = self.N if mode in (3, 5) else K
        as_arr = np.ctypeslib.as_array
        self.z = as_arr(C.cast(self.lib.orc_z(self.h), C.POINTER(C.c_int8)), (self.L, self.N, self.ploid))
        self.zz = as_arr(self.lib.orc_zz(self.h), (self.N,))
        self.qq = as_arr(self.lib.orc_qq(self.h), (self.N, K))
        self.qqnum = as_arr(self.lib.orc_qqnum(self.h), (self.N, K))
        self.freq = as_arr(self.lib.orc_freq(self.h), (K, self.L, self.Amax))
        self.self_rates = as_arr(self.lib.orc_self(self.h), (ns,))
        self.state = as_arr(self.lib.orc_state(self.h), (K,))
        self.gen = as_arr(self.lib.orc_gen(self.h), (self.N,))
        self.indvlkh = as_arr(self.lib.orc_indvlkh(self.h), (self.N,))
        self._alpha = as_arr(self.lib.orc_alpha(self.h), (1,))
        self._tot = as_arr(self.lib.orc_totallkh(self.h), (1,))

    def __del__(self):
        try:
            self.lib.orc_free(self.h)
        except Exception:
            pass

    alpha = property(lambda s: float(s._alpha[0]), lambda s, v: s._alpha.__setitem__(0, v))
    totallkh = property(lambda s: float(s._tot[0]))

    def setseeds(self, a, b, c):
        self.lib.orc_setseeds(self.h, a, b, c)

    def getseeds(self):
        o = (C.c_long * 3)()
        self.lib.orc_getseeds(self.h, o)
        return tuple(o)

    def ran1(self):
        return self.lib.orc_ran1(self.h)

    def missing_mask(self):
        m = np.zeros((self.L, self.N), dtype=np.uint8)
        self.lib.orc_missing_mask(self.h, m.ctypes.data)
        return m

    def tally(self, i0=None, i1=None):
        n = np.zeros((self.K, self.L, self.Amax), dtype=np.int32)
        if i0 is None:
            self.lib.orc_tally(self.h, n.ctypes.data)
        else:
            self.lib.orc_tally_range(self.h, i0, i1, n.ctypes.data)
        return n

    def count_z(self):
        c = np.zeros((self.N, self.K))
        self.lib.orc_count_z(self.h, _dp(c))
        return c

    def log_ld_indv(self, gen, i):
        return self.lib.orc_log_ld_indv(self.h, int(gen), int(i))

    def log_ld_all(self, gens=None):
        g = self.gen if gens is None else gens
        return np.array([self.log_ld_indv(g[i], i) for i in range(self.N)])

    def proposal(self, S):
        S = np.ascontiguousarray(S, dtype=np.float64)
        return self.lib.orc_proposal(self.h, _dp(S))

    def alpha_logratio(self, ralpha):
        return self.lib.orc_alpha_logratio(self.h, float(ralpha))

    def alpha_ratio_product(self, ralpha):
        return self.lib.orc_alpha_ratio_product(self.h, float(ralpha))

    def check_empty_cluster(self):
        return self.lib.orc_check_empty_cluster(self.h)

    def z_conditional(self, i, l, c):
        p = np.zeros(self.K)
        self.lib.orc_z_conditional(self.h, i, l, c, _dp(p))
        return p

    def update_P(self):
        self.lib.orc_update_P(self.h)

    def update_S_POP(self):
        self.lib.orc_update_S_POP(self.h)

    def update_S_IND(self):
        self.lib.orc_update_S_IND(self.h)

    def update_Z(self, init_flag=0):
        self.lib.orc_update_Z(self.h, init_flag)

    def log_ld_indv_K(self, i, k):
        return self.lib.orc_log_ld_indv_K(self.h, int(i), int(k))

    def update_F_POP(self):
        self.lib.orc_update_F_POP(self.h)

    def update_F_IND(self):
        self.lib.orc_update_F_IND(self.h)

    def log_ld_F(self, inbreed, by_pop, i):
        """log_ld_F_pop (by_pop=1, inbreed[K]) / log_ld_F_indv (by_pop=0, inbreed[0]), mcmc.c:1776,1812"""
        v = np.ascontiguousarray(np.atleast_1d(inbreed), dtype=np.float64)
        return self.lib.orc_log_ld_F(self.h, _dp(v), int(by_pop), int(i))

    def log_ld_F_total(self, inbreed):
        v = np.ascontiguousarray(inbreed, dtype=np.float64)
        return self.lib.orc_log_ld_F_total(self.h, _dp(v))

    def update_G(self):
        self.lib.orc_update_G(self.h)

    def update_ZQ(self, init_flag=0):
        self.lib.orc_update_ZQ(self.h, init_flag)

    def update_alpha(self):
        self.lib.orc_update_alpha(self.h)

    def cal_lkh(self):
        self.lib.orc_cal_lkh(self.h)

    def init_DP(self):
        self.lib.orc_init_DP(self.h)

    def update_DP(self):
        self.lib.orc_update_DP(self.h)

    def dp_nclusters(self):
        return self.lib.orc_dp_nclusters(self.h)

    def dp_from_values(self):
        """test hook: rebuild the Dirichlet-process clusters from self_rates (equal values share a cluster)"""
        self.lib.orc_dp_from_values(self.h)

    def dgeom(self, s, g):
        return self.lib.orc_dgeom(float(s), int(g))

    def sweeps(self, n):
        self.lib.orc_sweeps(self.h, n)

    def run_chain(self, update, burnin, thinning, ckrep=0, nstep_check_empty=20, initd=None):
        ns = self.N if self.mode in (3, 5) else self.K
        ch = self.lib.orc_chain_new(self.h, ckrep)
        initd = np.ascontiguousarray(initd if initd is not None else np.full(self.K, 0.5), dtype=np.float32)
        flag = self.lib.orc_run_chain(self.h, update, burnin, thinning, ckrep, nstep_check_empty,
                                      initd.ctypes.data_as(c_fp), ch)
        s = _ChainStruct.from_address(ch)
        cp = lambda p, n: np.ctypeslib.as_array(p, (n,)).copy()
        res = dict(flag_empty_cluster=flag, steps=s.steps, step=s.step, totallkh=s.totallkh, totallkh2=s.totallkh2,
                   indvlkh=cp(s.indvlkh, self.N), qq=cp(s.qq, self.N * self.K).reshape(self.N, self.K),
                   qq2=cp(s.qq2, self.N * self.K).reshape(self.N, self.K), self_rates=cp(s.self_rates, ns),
                   self_rates2=cp(s.self_rates2, ns), gen=cp(s.gen, self.N), gen2=cp(s.gen2, self.N),
                   convg=cp(s.convg, max(ckrep, 1))[:ckrep])
        self.lib.orc_chain_free(ch)
        return res


def gelman_rubin_ref(vec, numchains, totrep):
    v = np.ascontiguousarray(vec, dtype=np.float64)
    return oracle_lib().orc_gelman_rubin_ref(_dp(v), numchains, totrep)


def gelman_rubin(vec, numchains, n):
    v = np.ascontiguousarray(vec, dtype=np.float64)
    return oracle_lib().orc_gelman_rubin(_dp(v), numchains, n)


# ----------------------------------------------------------------------------------------
# the unmodified reference, through the include-the-.c harness
# ----------------------------------------------------------------------------------------
_ref_lib = None


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        if not have_ref():
            raise FileNotFoundError(REF_SO + " (built only where /root/reference exists: make -C oracle ref)")
        L = C.CDLL(REF_SO)
        L.refh_new.restype = C.c_void_p
        L.refh_new.argtypes = [C.c_int] * 8 + [C.c_double, c_ip, c_ip]
        L.refh_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refh_amax.argtypes = [C.c_void_p]
        L.refh_get_missindx.argtypes = [C.c_void_p, c_ip]
        for n_, t in [("refh_z", c_ip), ("refh_qq", c_dp), ("refh_qqnum", c_dp), ("refh_freq", c_dp),
                      ("refh_gen", c_ip), ("refh_self", c_dp), ("refh_state", c_ip), ("refh_alpha", c_dp)]:
            getattr(L, n_).argtypes = [C.c_void_p, t, C.c_int]
        L.refh_lkh.argtypes = [C.c_void_p, c_dp, c_dp]
        L.refh_setseeds.argtypes = [C.c_int] * 3
        L.refh_ran1.restype = C.c_double
        L.refh_update_P.argtypes = [C.c_void_p, c_ip]
        L.refh_update_ZQ.argtypes = [C.c_void_p, C.c_int]
        L.refh_zz.argtypes = [C.c_void_p, c_ip, C.c_int]
        L.refh_update_Z.argtypes = [C.c_void_p, C.c_int]
        L.refh_log_ld_indv_K.restype = C.c_double
        L.refh_log_ld_indv_K.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_log_ld_F.restype = C.c_double
        L.refh_log_ld_F.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int]
        L.refh_log_ld_F_total.restype = C.c_double
        L.refh_log_ld_F_total.argtypes = [C.c_void_p, c_dp]
        for n_ in ["refh_update_G", "refh_update_S_POP", "refh_update_S_IND", "refh_update_F_POP", "refh_update_F_IND", "refh_update_alpha", "refh_cal_lkh",
                   "refh_init_DP", "refh_update_DP", "refh_check_empty_cluster", "refh_dp_nclusters"]:
            getattr(L, n_).argtypes = [C.c_void_p]
        L.refh_log_ld_indv.restype = C.c_double
        L.refh_log_ld_indv.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_proposal.restype = C.c_double
        L.refh_proposal.argtypes = [C.c_void_p, c_dp]
        L.refh_dgeom.restype = C.c_double
        L.refh_dgeom.argtypes = [C.c_double, C.c_int]
        L.refh_genofreq.restype = C.c_double
        L.refh_genofreq.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.refh_dt_stat.argtypes = [C.c_double]
        L.refh_rgeom.argtypes = [C.c_double]
        L.refh_disc_unif.argtypes = [c_dp, C.c_int]
        L.refh_rgamma.restype = C.c_double
        L.refh_rgamma.argtypes = [C.c_double, C.c_double]
        L.refh_rbeta.restype = C.c_double
        L.refh_rbeta.argtypes = [C.c_double, C.c_double]
        L.refh_rnormal.restype = C.c_double
        L.refh_rnormal.argtypes = [C.c_double, C.c_double]
        L.refh_sweeps.argtypes = [C.c_void_p, C.c_int]
        L.refh_mcmc_updating.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, c_fp] + [c_dp] * 9
        L.refd_read_data.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p] + [C.c_int] * 5 + \
                                    [c_ip] * 3 + [c_ip] * 4
        _ref_lib = L
    return _ref_lib


class Reference:
    """Injected-state access to the unmodified reference's static hot functions.
    ``x`` is int16 [L][N][ploid] in the new layout; it is transposed to the reference's
    individual-major int tensors here.  NOTE the reference's RNG is process-global."""

    def __init__(self, x, allelenum, K, mode=2, prior_flag=0, back_refl=1, type_freq=1, alpha_dpm=10.0):
        self.lib = ref_lib()
        self.L, self.N, self.ploid = x.shape
        self.K, self.mode = K, mode
        xi = np.ascontiguousarray(np.transpose(np.asarray(x, dtype=np.int32), (1, 0, 2)))   # [N][L][ploid]
        an = np.ascontiguousarray(allelenum, dtype=np.int32)
        self.h = self.lib.refh_new(self.N, self.L, K, self.ploid, mode, prior_flag, back_refl, type_freq,
                                   float(alpha_dpm), _ip(xi), _ip(an))
        self.Amax = self.lib.refh_amax(self.h)

    # -- state exchange in the NEW layouts ---------------------------------------------
    def set_z(self, z):            # z int8 [L][N][ploid]
        zi = np.ascontiguousarray(np.transpose(np.asarray(z, dtype=np.int32), (1, 0, 2)))
        self.lib.refh_z(self.h, _ip(zi), 1)

    def get_z(self):
        zi = np.zeros((self.N, self.L, self.ploid), dtype=np.int32)
        self.lib.refh_z(self.h, _ip(zi), 0)
        return np.ascontiguousarray(np.transpose(zi, (1, 0, 2)).astype(np.int8))

    def _xd(self, fn, shape, val=None):
        if val is None:
            a = np.zeros(shape)
            fn(self.h, _dp(a), 0)
            return a
        a = np.ascontiguousarray(val, dtype=np.float64).reshape(shape)
        fn(self.h, _dp(a), 1)

    def set_qq(self, q): self._xd(self.lib.refh_qq, (self.N, self.K), q)
    def get_qq(self): return self._xd(self.lib.refh_qq, (self.N, self.K))
    def set_qqnum(self, q): self._xd(self.lib.refh_qqnum, (self.N, self.K), q)
    def get_qqnum(self): return self._xd(self.lib.refh_qqnum, (self.N, self.K))
    def set_freq(self, f): self._xd(self.lib.refh_freq, (self.K, self.L, self.Amax), f)
    def get_freq(self): return self._xd(self.lib.refh_freq, (self.K, self.L, self.Amax))

    def set_self(self, s):
        self._xd(self.lib.refh_self, (self.N if self.mode in (3, 5) else self.K,), s)

    def get_self(self):
        return self._xd(self.lib.refh_self, (self.N if self.mode in (3, 5) else self.K,))

    def set_gen(self, g):
        a = np.ascontiguousarray(g, dtype=np.int32)
        self.lib.refh_gen(self.h, _ip(a), 1)

    def get_gen(self):
        a = np.zeros(self.N, dtype=np.int32)
        self.lib.refh_gen(self.h, _ip(a), 0)
        return a

    def set_alpha(self, v):
        a = np.array([v], dtype=np.float64)
        self.lib.refh_alpha(self.h, _dp(a), 1)

    def get_alpha(self):
        a = np.zeros(1)
        self.lib.refh_alpha(self.h, _dp(a), 0)
        return float(a[0])

    def get_lkh(self):
        a = np.zeros(self.N)
        t = np.zeros(1)
        self.lib.refh_lkh(self.h, _dp(a), _dp(t))
        return a, float(t[0])

    def missindx(self):            # returned as [L][N]
        a = np.zeros((self.N, self.L), dtype=np.int32)
        self.lib.refh_get_missindx(self.h, _ip(a))
        return np.ascontiguousarray(a.T)

    def setseeds(self, a, b, c): self.lib.refh_setseeds(a, b, c)

    def update_P(self, want_tally=True):
        if not want_tally:
            self.lib.refh_update_P(self.h, None)
            return None
        t = np.zeros((self.K, self.L, self.Amax), dtype=np.int32)
        self.lib.refh_update_P(self.h, _ip(t))
        return t

    def update_ZQ(self, init_flag=0): self.lib.refh_update_ZQ(self.h, init_flag)
    def update_G(self): self.lib.refh_update_G(self.h)
    def update_S_POP(self): self.lib.refh_update_S_POP(self.h)
    def update_S_IND(self): self.lib.refh_update_S_IND(self.h)
    def update_Z(self, init_flag=0): self.lib.refh_update_Z(self.h, init_flag)
    def log_ld_indv_K(self, i, k): return self.lib.refh_log_ld_indv_K(self.h, int(i), int(k))

    def set_zz(self, zz):
        a = np.ascontiguousarray(zz, dtype=np.int32)
        self.lib.refh_zz(self.h, _ip(a), 1)

    def get_zz(self):
        a = np.zeros(self.N, dtype=np.int32)
        self.lib.refh_zz(self.h, _ip(a), 0)
        return a

    def update_F_POP(self): self.lib.refh_update_F_POP(self.h)
    def update_F_IND(self): self.lib.refh_update_F_IND(self.h)

    def log_ld_F(self, inbreed, by_pop, i):
        v = np.ascontiguousarray(np.atleast_1d(inbreed), dtype=np.float64)
        return self.lib.refh_log_ld_F(self.h, _dp(v), int(by_pop), int(i))

    def log_ld_F_total(self, inbreed):
        v = np.ascontiguousarray(inbreed, dtype=np.float64)
        return self.lib.refh_log_ld_F_total(self.h, _dp(v))
    def update_alpha(self): self.lib.refh_update_alpha(self.h)
    def cal_lkh(self): self.lib.refh_cal_lkh(self.h)
    def init_DP(self): self.lib.refh_init_DP(self.h)
    def update_DP(self): self.lib.refh_update_DP(self.h)
    def dp_nclusters(self): return self.lib.refh_dp_nclusters(self.h)
    def log_ld_indv(self, gen, i): return self.lib.refh_log_ld_indv(self.h, int(gen), int(i))
    def check_empty_cluster(self): return self.lib.refh_check_empty_cluster(self.h)

    def proposal(self, S):
        S = np.ascontiguousarray(S, dtype=np.float64)
        return self.lib.refh_proposal(self.h, _dp(S))

    def sweeps(self, n): self.lib.refh_sweeps(self.h, n)

    def mcmc_updating(self, update, burnin, thinning, ckrep=1, nstep_check_empty=20, initd=None):
        """One chain through the reference's own driver (mcmc.c:63)."""
        N, K = self.N, self.K
        ns = N if self.mode in (3, 5) else K
        self.lib.refh_set_flags(self.h, nstep_check_empty, 0, 0)
        tot, ind = np.zeros(2), np.zeros(N)
        qq, qq2 = np.zeros((N, K)), np.zeros((N, K))
        s, s2, g, g2 = np.zeros(ns), np.zeros(ns), np.zeros(N), np.zeros(N)
        cv = np.zeros(max(ckrep, 1))
        initd = np.ascontiguousarray(initd if initd is not None else np.full(K, 0.5), dtype=np.float32)
        flag = self.lib.refh_mcmc_updating(self.h, update, burnin, thinning, ckrep, initd.ctypes.data_as(c_fp),
                                           _dp(tot), _dp(ind), _dp(qq), _dp(qq2), _dp(s), _dp(s2), _dp(g), _dp(g2), _dp(cv))
        return dict(flag_empty_cluster=flag, totallkh=tot[0], totallkh2=tot[1], indvlkh=ind, qq=qq, qq2=qq2,
                    self_rates=s, self_rates2=s2, gen=g, gen2=g2, convg=cv[:ckrep])


def ref_read_data(path, ploid, N, K, L, missing="-9", label=1, popdata=1, n_extra_col=0, markername_flag=0,
                  datafmt=0):
    """Parse a text file with the reference's own reader (data_interface.c:36).  Returns
    (x int16 [L][N][ploid], allelenum, missindx [L][N], alleleid or None)."""
    lib = ref_lib()
    oN, oL, oA = C.c_int(), C.c_int(), C.c_int()
    x = np.zeros((N, L, ploid), dtype=np.int32)
    an = np.zeros(L, dtype=np.int32)
    mi = np.zeros((N, L), dtype=np.int32)
    aid = np.zeros((N, L), dtype=np.int32)
    lib.refd_read_data(path.encode(), ploid, N, K, L, missing.encode(), label, popdata, n_extra_col,
                       markername_flag, datafmt, C.byref(oN), C.byref(oL), C.byref(oA), _ip(x), _ip(an), _ip(mi),
                       _ip(aid))
    n, l = oN.value, oL.value
    # the reader writes rows with stride l (its own locinum), so re-view the flat buffers
    xv = x.reshape(-1)[: n * l * ploid].reshape(n, l, ploid)
    miv = mi.reshape(-1)[: n * l].reshape(n, l)
    aidv = aid.reshape(-1)[: n * l].reshape(n, l)
    return (np.ascontiguousarray(np.transpose(xv, (1, 0, 2)).astype(np.int16)), an[:l].copy(),
            np.ascontiguousarray(miv.T), np.ascontiguousarray(aidv.T) if ploid == 4 else None)
