"""ctypes wrappers of the autotetraploid CHECKERS -- TEST INFRASTRUCTURE ONLY.

  TetraOracle   oracle/tetra_oracle.c, the in-repo CPU restatement (liboracle.so)
  RefTetra      the unmodified reference's poly_geno.c through oracle/ref_harness_poly.c
                (oracle/_ref/libinstruct_ref.so)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C

import numpy as np

from oracle.pyoracle import _ChainStruct as OrcChain, oracle_lib, ref_lib

c_dp = C.POINTER(C.c_double)


def _tet_lib():
    lib = oracle_lib()
    if getattr(lib, "_tet_ready", False):
        return lib
    lib.tet_new.restype = C.c_void_p
    lib.tet_new.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
    lib.tet_new2.restype = C.c_void_p
    lib.tet_new2.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
    lib.tet_freq2.restype = c_dp
    lib.tet_freq2.argtypes = [C.c_void_p]
    lib.tet_tally_allo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tet_geno_conditional_allo.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for f, rt in [("tet_z", C.c_void_p), ("tet_geno", C.c_void_p), ("tet_qq", c_dp), ("tet_qqnum", c_dp), ("tet_freq", c_dp),
                  ("tet_self", c_dp), ("tet_state", C.POINTER(C.c_int)), ("tet_alpha", c_dp), ("tet_indvlkh", c_dp),
                  ("tet_totallkh", c_dp), ("tet_exfreq", C.POINTER(C.c_float)), ("tet_genofreq", C.POINTER(C.c_float))]:
        getattr(lib, f).restype = rt
        getattr(lib, f).argtypes = [C.c_void_p]
    for f in ["tet_free", "tet_calc_exfreq", "tet_initial_geno", "tet_update_P", "tet_update_S", "tet_update_geno"]:
        getattr(lib, f).restype = None
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.tet_setseeds.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long]
    lib.tet_gmax.argtypes = [C.c_void_p]
    lib.tet_amax.argtypes = [C.c_void_p]
    lib.tet_genolist.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tet_geno_index.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tet_tally.argtypes = [C.c_void_p, C.c_void_p]
    lib.tet_count_z.argtypes = [C.c_void_p, C.c_void_p]
    lib.tet_calc_genofreq.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    lib.tet_site_loglik.restype = C.c_double
    lib.tet_site_loglik.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tet_cal_lkd.restype = C.c_double
    lib.tet_cal_lkd.argtypes = [C.c_void_p]
    lib.tet_cal_lkd_props.restype = C.c_double
    lib.tet_cal_lkd_props.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tet_geno_conditional.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tet_z_conditional.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.tet_update_ZQ.argtypes = [C.c_void_p, C.c_int]
    lib.tet_sweeps.argtypes = [C.c_void_p, C.c_int]
    lib.tet_chain_new.restype = C.POINTER(OrcChain)
    lib.tet_chain_new.argtypes = [C.c_void_p, C.c_int]
    lib.tet_run_chain.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(OrcChain)]
    lib._tet_ready = True
    return lib


class TetraOracle:
    def __init__(self, x, nd, allelenum, K, back_refl=1, autopoly=1):
        self.lib = _tet_lib()
        self.autopoly = autopoly
        self.x = np.ascontiguousarray(x, dtype=np.int16)
        self.nd = np.ascontiguousarray(nd, dtype=np.uint8)
        self.allelenum = np.ascontiguousarray(allelenum, dtype=np.int32)
        self.L, self.N, _ = self.x.shape
        self.K = K
        self.h = self.lib.tet_new2(self.N, self.L, K, back_refl, autopoly, self.x.ctypes.data, self.nd.ctypes.data,
                                   self.allelenum.ctypes.data)
        self.Amax = self.lib.tet_amax(self.h)
        self.Gmax = self.lib.tet_gmax(self.h)
        as_arr = np.ctypeslib.as_array
        self.z = as_arr(C.cast(self.lib.tet_z(self.h), C.POINTER(C.c_int8)), (self.L, self.N, 4))
        self.geno = as_arr(C.cast(self.lib.tet_geno(self.h), C.POINTER(C.c_int8)), (self.L, self.N, 4))
        self.qq = as_arr(self.lib.tet_qq(self.h), (self.N, K))
        self.qqnum = as_arr(self.lib.tet_qqnum(self.h), (self.N, K))
        self.freq = as_arr(self.lib.tet_freq(self.h), (K, self.L, self.Amax))
        self.freq2 = as_arr(self.lib.tet_freq2(self.h), (K, self.L, self.Amax))     # allotetraploid: second subgenome
        self.self_rates = as_arr(self.lib.tet_self(self.h), (K,))
        self.state = as_arr(self.lib.tet_state(self.h), (K,))
        self.indvlkh = as_arr(self.lib.tet_indvlkh(self.h), (self.N,))
        self._alpha = as_arr(self.lib.tet_alpha(self.h), (1,))
        self._tot = as_arr(self.lib.tet_totallkh(self.h), (1,))
        self.exfreq = as_arr(self.lib.tet_exfreq(self.h), (K, self.L, self.Gmax))
        self.genofreq = as_arr(self.lib.tet_genofreq(self.h), (K, self.L, self.Gmax))

    def __del__(self):
        try:
            self.lib.tet_free(self.h)
        except Exception:
            pass

    alpha = property(lambda s: float(s._alpha[0]), lambda s, v: s._alpha.__setitem__(0, v))
    totallkh = property(lambda s: float(s._tot[0]))

    def setseeds(self, a, b, c):
        self.lib.tet_setseeds(self.h, a, b, c)

    def genolist(self, l):
        buf = np.zeros(self.Gmax, dtype=np.int32)
        n = self.lib.tet_genolist(self.h, l, buf.ctypes.data)
        return buf[:n].copy()

    def geno_index(self, l, g4):
        g = np.ascontiguousarray(g4, dtype=np.int8)
        return self.lib.tet_geno_index(self.h, l, g.ctypes.data)

    def tally(self):
        n = np.zeros((self.K, self.L, self.Amax), dtype=np.int32)
        self.lib.tet_tally(self.h, n.ctypes.data)
        return n

    def tally_allo(self):
        n1 = np.zeros((self.K, self.L, self.Amax), dtype=np.int32)
        n2 = np.zeros((self.K, self.L, self.Amax), dtype=np.int32)
        self.lib.tet_tally_allo(self.h, n1.ctypes.data, n2.ctypes.data)
        return n1, n2

    def geno_conditional_allo(self, i, l):
        p = np.zeros(12)
        n = self.lib.tet_geno_conditional_allo(self.h, i, l, p.ctypes.data)
        return p[:n]

    def count_z(self):
        c = np.zeros((self.N, self.K))
        self.lib.tet_count_z(self.h, c.ctypes.data)
        return c

    def calc_exfreq(self):
        self.lib.tet_calc_exfreq(self.h)

    def calc_genofreq(self, k, s):
        out = np.zeros((self.L, self.Gmax), dtype=np.float32)
        self.lib.tet_calc_genofreq(self.h, k, float(s), out.ctypes.data)
        return out

    def tables(self):
        """exfreq from the current freq, genofreq of every population at the current rates"""
        self.calc_exfreq()
        for k in range(self.K):
            self.genofreq[k] = self.calc_genofreq(k, self.self_rates[k])

    def cal_lkd(self):
        return self.lib.tet_cal_lkd(self.h)

    def cal_lkd_props(self, k, tab):
        t = np.ascontiguousarray(tab, dtype=np.float32)
        return self.lib.tet_cal_lkd_props(self.h, k, t.ctypes.data)

    def geno_conditional(self, i, l):
        p = np.zeros(3)
        self.lib.tet_geno_conditional(self.h, i, l, p.ctypes.data)
        return p

    def z_conditional(self, i, l, c):
        p = np.zeros(self.K)
        self.lib.tet_z_conditional(self.h, i, l, c, p.ctypes.data)
        return p

    def initial_geno(self):
        self.lib.tet_initial_geno(self.h)

    def update_P(self):
        self.lib.tet_update_P(self.h)

    def update_S(self):
        self.lib.tet_update_S(self.h)

    def update_ZQ(self, init_flag=0):
        self.lib.tet_update_ZQ(self.h, init_flag)

    def update_geno(self):
        self.lib.tet_update_geno(self.h)

    def sweeps(self, n):
        self.lib.tet_sweeps(self.h, n)

    def run_chain(self, update, burnin, thinning, ckrep=1, nstep_check_empty=1 << 30, initd=None):
        initd = np.ascontiguousarray(initd if initd is not None else np.full(self.K, 0.5), dtype=np.float32)
        ch = self.lib.tet_chain_new(self.h, ckrep)
        flag = self.lib.tet_run_chain(self.h, update, burnin, thinning, ckrep, nstep_check_empty, initd.ctypes.data, ch)
        c = ch.contents
        as_arr = np.ctypeslib.as_array
        out = dict(flag=flag, totallkh=c.totallkh, totallkh2=c.totallkh2,
                   indvlkh=as_arr(c.indvlkh, (self.N,)).copy(), qq=as_arr(c.qq, (self.N, self.K)).copy(),
                   qq2=as_arr(c.qq2, (self.N, self.K)).copy(), self_rates=as_arr(c.self_rates, (self.K,)).copy(),
                   self_rates2=as_arr(c.self_rates2, (self.K,)).copy(), convg=as_arr(c.convg, (max(ckrep, 1),)).copy())
        return out


class RefTetra:
    """The reference's own poly_geno.c on the same data (individual-major, as the reference holds it)."""

    def __init__(self, x, nd, allelenum, K, back_refl=1, autopoly=1):
        self.lib = ref_lib()
        self.lib.refp_new2.restype = C.c_void_p
        self.lib.refp_new2.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        self.lib.refp_gmax.argtypes = [C.c_void_p]
        self.lib.refp_genolist.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.lib.refp_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self.lib.refp_mcmc_updating.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int] + [C.c_void_p] * 8
        L, N, _ = x.shape
        self.L, self.N, self.K = L, N, K
        xi = np.ascontiguousarray(np.transpose(x, (1, 0, 2)).astype(np.int32))       # [N][L][4]
        miss = np.transpose(nd, (1, 0)) == 0
        xi[miss, 0] = -9                                                             # data_interface.c:648-649
        self._x = xi
        self._nd = np.ascontiguousarray(np.transpose(nd, (1, 0)).astype(np.int32))
        self._an = np.ascontiguousarray(allelenum, dtype=np.int32)
        self.h = self.lib.refp_new2(N, L, K, back_refl, autopoly, self._x.ctypes.data, self._nd.ctypes.data, self._an.ctypes.data)
        self.Gmax = self.lib.refp_gmax(self.h)
        self.Amax = int(self._an.max())

    def setseeds(self, a, b, c):
        self.lib.refh_setseeds(a, b, c)

    def genolist(self, l):
        buf = np.zeros(self.Gmax, dtype=np.int32)
        n = self.lib.refp_genolist(self.h, l, buf.ctypes.data)
        return buf[:n].copy()

    def tables(self, freq, S, freq2=None):
        """freq2: the second subgenome's frequencies (allotetraploid harness)"""
        f = np.ascontiguousarray(freq if freq2 is None else np.stack([freq, freq2]), dtype=np.float64)
        s = np.ascontiguousarray(S, dtype=np.float64)
        ex = np.zeros((self.K, self.L, self.Gmax), dtype=np.float32)
        gf = np.zeros((self.K, self.L, self.Gmax), dtype=np.float32)
        self.lib.refp_tables(self.h, f.ctypes.data, s.ctypes.data, ex.ctypes.data, gf.ctypes.data, self.Gmax)
        return ex, gf

    def run_chain(self, update, burnin, thinning, ckrep=1, initd=None):
        initd = np.ascontiguousarray(initd if initd is not None else np.full(self.K, 0.5), dtype=np.float32)
        tot = np.zeros(2); lk = np.zeros(self.N); qq = np.zeros((self.N, self.K)); qq2 = np.zeros((self.N, self.K))
        s = np.zeros(self.K); s2 = np.zeros(self.K); cv = np.zeros(max(ckrep, 1))
        flag = self.lib.refp_mcmc_updating(self.h, update, burnin, thinning, ckrep, initd.ctypes.data, tot.ctypes.data,
                                           lk.ctypes.data, qq.ctypes.data, qq2.ctypes.data, s.ctypes.data, s2.ctypes.data,
                                           cv.ctypes.data)
        return dict(flag=flag, totallkh=tot[0], totallkh2=tot[1], indvlkh=lk, qq=qq, qq2=qq2, self_rates=s, self_rates2=s2, convg=cv)
