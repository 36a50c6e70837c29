"""The C host side (instruct_b200/host): the packer against the reference's own text reader,
and the result-file writer against the reference's own chain_stat(), byte for byte.
CPU-only (the reference lives in oracle/_ref, built from /root/reference by oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from instruct_b200.synth import make_dataset
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "instruct_b200", "host")
HOST_SO = os.path.join(HOST, "libinbreed_host.so")

pytestmark = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built")


class GsOptions(C.Structure):
    _fields_ = [("ploid", C.c_int), ("totalsize", C.c_int), ("locinum", C.c_int), ("missing", C.c_char_p),
                ("label", C.c_int), ("popdata", C.c_int), ("n_extra_col", C.c_int), ("markername_flag", C.c_int),
                ("datafmt", C.c_int), ("quiet", C.c_int)]


class GsStore(C.Structure):
    _fields_ = [("ploid", C.c_int), ("totalsize", C.c_int), ("locinum", C.c_int), ("locinum_file", C.c_int),
                ("allelenum_max", C.c_int), ("x", C.POINTER(C.c_int16)), ("allelenum", C.POINTER(C.c_int32)),
                ("alleletype", C.c_void_p), ("locus_of", C.POINTER(C.c_int)), ("marker_names", C.c_void_p),
                ("indvname", C.POINTER(C.c_char_p)), ("popindx", C.POINTER(C.c_int)), ("poptype", C.POINTER(C.c_char_p)),
                ("pop_count", C.c_int), ("extra_col", C.c_void_p), ("n_extra_col", C.c_int), ("missvec", C.POINTER(C.c_int))]


class WrRun(C.Structure):
    _fields_ = [("datafilename", C.c_char_p), ("initialfilename", C.c_char_p), ("missingdata", C.c_char_p)] + \
               [(n, C.c_int) for n in ("chainnum", "thinning", "ploid", "autopoly", "totalsize", "locinum", "popnum", "mode",
                                       "inf_K", "prior_flag", "back_refl", "print_freq", "GR_flag", "ckrep", "distr_fmt",
                                       "label", "popdata", "markername_flag")] + \
               [("update", C.c_long), ("burnin", C.c_long), ("siglevel", C.c_double), ("alpha_dpm", C.c_double)]


class WrData(C.Structure):
    _fields_ = [("indvname", C.POINTER(C.c_char_p)), ("poptype", C.POINTER(C.c_char_p)), ("marker_names", C.c_void_p),
                ("alleletype", C.c_void_p), ("popindx", C.POINTER(C.c_int)), ("missvec", C.POINTER(C.c_int)),
                ("allelenum", C.POINTER(C.c_int)), ("pop_count", C.c_int), ("allelenum_max", C.c_int)]


class WrChain(C.Structure):
    _fields_ = [("chn_name", C.c_char_p), ("name_len", C.c_int), ("totallkh", C.c_double), ("totallkh2", C.c_double)] + \
               [(n, C.POINTER(C.c_double)) for n in ("indvlkh", "qq", "qq2", "self_rates", "self_rates2", "gen", "gen2",
                                                      "freq", "freq2")]


@pytest.fixture(scope="module")
def host():
    out = subprocess.run(["make", "-C", HOST, "libinbreed_host.so"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lib = C.CDLL(HOST_SO)
    lib.gs_read.argtypes = [C.c_char_p, C.POINTER(GsOptions), C.POINTER(GsStore), C.c_char_p, C.c_int]
    lib.gs_free.argtypes = [C.POINTER(GsStore)]
    lib.gs_save.argtypes = [C.c_char_p, C.POINTER(GsStore), C.c_char_p, C.c_int]
    lib.gs_load.argtypes = [C.c_char_p, C.POINTER(GsStore), C.c_char_p, C.c_int]
    lib.wr_chain.argtypes = [C.c_char_p, C.POINTER(WrRun), C.POINTER(WrData), C.POINTER(WrChain), C.POINTER(C.c_double)]
    lib.wr_gelman_rubin.restype = C.c_double
    lib.wr_gelman_rubin.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    return lib


def _write_text(path, x, pops, label, popdata, extra, fmt, markers, alleles):
    L, N, _ = x.shape
    with open(path, "w") as fh:
        if markers:
            fh.write(" ".join(f"m{l}" for l in range(L)) + "\n")
        for i in range(N):
            meta = ([f"id{i}"] if label else []) + ([f"P{pops[i]}"] if popdata else []) + [f"e{i}_{k}" for k in range(extra)]
            tok = lambda l, c: "-9" if x[l, i, c] < 0 else alleles[l][x[l, i, c]]
            if fmt == 0:
                for c in range(2):
                    fh.write("\t".join(meta + [tok(l, c) for l in range(L)]) + "\n")
            else:
                fh.write(" ".join(meta + [tok(l, c) for l in range(L) for c in range(2)]) + "\n")


@pytest.mark.parametrize("label,popdata,extra,fmt,markers", [(1, 1, 0, 0, 0), (0, 0, 0, 0, 0), (1, 0, 2, 1, 0), (1, 1, 1, 0, 1)])
def test_packer_matches_reference_reader(host, tmp_path, label, popdata, extra, fmt, markers):
    rng = np.random.default_rng(label * 8 + popdata * 4 + extra * 2 + fmt)
    d = make_dataset(N=23, L=14, K=3, A=5, miss=0.08, seed=77)
    x = d.x.copy()
    x[3] = np.where(x[3] >= 0, 0, x[3])                   # a monomorphic locus: both readers must drop it
    x[9, :, :] = -9                                       # an all-missing locus
    alleles = [[str(int(v)) for v in rng.permutation(np.arange(100, 100 + 40))[:8]] for _ in range(x.shape[0])]
    p = str(tmp_path / "geno.txt")
    _write_text(p, x, d.pop, label, popdata, extra, fmt, markers, alleles)
    opt = GsOptions(2, 23, 14, b"-9", label, popdata, extra, markers, fmt, 1)
    st = GsStore()
    err = C.create_string_buffer(512)
    assert host.gs_read(p.encode(), C.byref(opt), C.byref(st), err, 512) == 0, err.value
    N, L = st.totalsize, st.locinum
    mine = np.ctypeslib.as_array(st.x, (L, N, 2)).copy()
    an = np.ctypeslib.as_array(st.allelenum, (L,)).copy()
    rx, ran, rmiss, _ = pyoracle.ref_read_data(p, 2, 23, 3, 14, label=label, popdata=popdata, n_extra_col=extra,
                                               markername_flag=markers, datafmt=fmt)
    assert (L, N) == (rx.shape[0], rx.shape[1]) == (12, 23)
    assert np.array_equal(an, ran)
    assert np.array_equal(mine, rx)                       # same recoding, same missing code, same layout
    mv = np.ctypeslib.as_array(st.missvec, (N,))
    assert np.array_equal(mv, rmiss.sum(axis=0))
    if popdata:
        assert st.pop_count == 3
        assert [st.poptype[i].decode() for i in range(3)] == ["P0", "P1", "P2"]
    if label:
        assert st.indvname[5].decode() == "id5"
    host.gs_free(C.byref(st))


@pytest.mark.parametrize("mode,label,popdata,distr", [(2, 1, 1, 1), (2, 0, 0, 0), (3, 0, 1, 1), (4, 1, 1, 1), (4, 0, 1, 0), (5, 1, 1, 1), (5, 0, 0, 1),
                                                       (0, 1, 1, 1), (0, 0, 0, 1)])
def test_result_file_bytes_match_reference_writer(host, tmp_path, mode, label, popdata, distr):
    _writer_case(host, tmp_path, mode, label, popdata, distr)


@pytest.mark.parametrize("case", range(12))
def test_result_file_bytes_fuzz(host, tmp_path, case):
    """Random sizes, models, column switches and value magnitudes (tiny variances, rates at the ends of (0,1), large
    likelihoods): the result file stays byte-identical to the reference writer's."""
    rng = np.random.default_rng(7000 + case)
    _writer_case(host, tmp_path, int(rng.choice([0, 2, 4, 5])), int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2)),
                 seed=7100 + case, N=int(rng.integers(3, 40)), K=int(rng.integers(2, 7)), scale=float(rng.choice([1.0, 1e-3, 40.0])))


def _writer_case(host, tmp_path, mode, label, popdata, distr, seed=3, N=17, K=3, scale=1.0):
    d = make_dataset(N=N, L=9, K=K, A=4, miss=0.1, seed=5)
    rng = np.random.default_rng(seed)
    ns = N if mode in (3, 5) else K
    tot, tot2 = -1234.5678 * scale, (1234.5678 * scale) ** 2 + 33.3 * scale
    indv = rng.normal(-70, 5, N) * scale
    qq = rng.dirichlet(np.ones(K), N); qq2 = qq ** 2 + rng.uniform(0, 0.01, (N, K))
    if mode == 0:                                   # CHAIN.z / steps: shares of 10 retained samples
        qq = rng.multinomial(10, np.ones(K) / K, N) / 10.0
    s = rng.uniform(0.05, 0.95, ns) if scale == 1.0 else rng.choice([1e-4, 0.5, 1 - 1e-4], ns); s2 = s ** 2 + rng.uniform(0, 0.01, ns) * min(scale, 1.0)
    g = rng.uniform(1, 6, N); g2 = g ** 2 + rng.uniform(0, 1, N)
    pops = (np.arange(N) % 2).astype(np.int32)
    missvec = (d.x < 0).any(axis=2).sum(axis=0).astype(np.int32)
    ref = pyoracle.Reference(d.x, d.allelenum, K, mode=mode)
    L = pyoracle.ref_lib()
    L.refh_chain_stat.restype = C.c_double
    L.refh_chain_stat.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, pyoracle.c_ip, C.c_int, pyoracle.c_ip,
                                  C.c_char_p, C.c_double, C.c_double] + [pyoracle.c_dp] * 7
    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(pyoracle.c_dp)
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (indv, qq, qq2, s, s2, g, g2)]
    f_ref = str(tmp_path / "ref.out")
    dic_ref = L.refh_chain_stat(ref.h, f_ref.encode(), label, popdata, distr, pops.ctypes.data_as(pyoracle.c_ip), 2,
                                missvec.ctypes.data_as(pyoracle.c_ip), b"Chain#1", tot, tot2,
                                *[k.ctypes.data_as(pyoracle.c_dp) for k in keep])
    # ---- ours
    names = (C.c_char_p * N)(*[f"ind{i}".encode() for i in range(N)])
    ptypes = (C.c_char_p * 2)(b"pop0", b"pop1")
    run = WrRun()
    run.ploid, run.totalsize, run.locinum, run.popnum, run.mode = 2, N, d.L, K, mode
    run.label, run.popdata, run.distr_fmt, run.print_freq = label, popdata, distr, 0
    wd = WrData(names, ptypes, None, None, pops.ctypes.data_as(C.POINTER(C.c_int)), missvec.ctypes.data_as(C.POINTER(C.c_int)),
                None, 2, 4)
    ch = WrChain(b"Chain#1", 8, tot, tot2, *[k.ctypes.data_as(C.POINTER(C.c_double)) for k in keep], None, None)
    f_my = str(tmp_path / "my.out")
    dic = C.c_double()
    assert host.wr_chain(f_my.encode(), C.byref(run), C.byref(wd), C.byref(ch), C.byref(dic)) == 0
    a, b = open(f_ref, "rb").read(), open(f_my, "rb").read()
    assert abs(dic.value - dic_ref) < 1e-9
    if mode == 3:
        # the reference's mode-3 table is shifted by one individual and reads past the arrays
        # (SURVEY.md App. B #2); everything outside that table must still match byte for byte
        def strip(t):
            i0 = t.index(b"The Posterior distribution of Selfing Rates:")
            i1 = t.index(b"The Posterior distribution of Generations:")
            return t[:i0] + t[i1:]
        assert strip(a) == strip(b)
        rows = b[b.index(b"The Posterior distribution of Selfing Rates:"):b.index(b"The Posterior distribution of Generations:")]
        assert rows.count(b"Indv ") == N and b"Indv 0\t\t" in rows
    else:
        assert a == b
    assert b"Chain#1\x00:" in b                            # the NUL the reference writes (App. B #6)


def test_gelman_rubin_c_matches_python(host):
    from instruct_b200.converge import gelman_rubin
    rng = np.random.default_rng(1)
    tr = np.ascontiguousarray(rng.normal(size=(5, 9)) + np.arange(5)[:, None] * 0.3)
    got = host.wr_gelman_rubin(tr.ctypes.data_as(C.POINTER(C.c_double)), 5, 9)
    assert abs(got - gelman_rubin(tr)) < 1e-12


def test_packer_matches_reference_reader_tetraploid(host, tmp_path):
    """-p 4: the store holds the ascending distinct-allele set per genotype (transform_data2,
    data_interface.c:571-669), every locus is kept, missing = no allele observed."""
    from instruct_b200.synth import make_tetra_dataset, write_reference_text_tetra
    d = make_tetra_dataset(N=19, L=11, K=2, A=5, miss=0.1, seed=9)
    copies = d.dosage.copy()
    copies[4, :, :] = np.where(copies[4] >= 0, 1, copies[4])   # a monomorphic locus: kept for ploid 4
    copies[2, 5, 1] = -9                                         # a partly missing genotype: the observed alleles still count
    p = str(tmp_path / "geno4.txt")
    write_reference_text_tetra(p, copies, pop=d.pop)
    opt = GsOptions(4, 19, 11, b"-9", 1, 1, 0, 0, 1, 1)
    st = GsStore()
    err = C.create_string_buffer(512)
    assert host.gs_read(p.encode(), C.byref(opt), C.byref(st), err, 512) == 0, err.value
    N, L = st.totalsize, st.locinum
    mine = np.ctypeslib.as_array(st.x, (L, N, 4)).copy()
    an = np.ctypeslib.as_array(st.allelenum, (L,)).copy()
    rx, ran, rmiss, rid = pyoracle.ref_read_data(p, 4, 19, 2, 11, label=1, popdata=1, datafmt=1)
    assert (L, N) == (rx.shape[0], rx.shape[1]) == (11, 19)
    assert np.array_equal(an, ran)
    want = rx.copy()
    want[(rid == 0)] = -1                                  # the reference writes -9 into slot 0 of a missing genotype
    assert np.array_equal(mine, want)
    assert np.array_equal((mine >= 0).sum(axis=2), rid)    # alleleid
    mv = np.ctypeslib.as_array(st.missvec, (N,))
    assert np.array_equal(mv, rmiss.sum(axis=0))
    host.gs_free(C.byref(st))


def _strings(pp, n):
    a = C.cast(pp, C.POINTER(C.c_char_p))
    return [a[i] for i in range(n)]


@pytest.mark.parametrize("label,popdata,extra,fmt,markers", [(1, 1, 2, 0, 1), (0, 0, 0, 1, 0)])
def test_packed_store_round_trip(host, tmp_path, label, popdata, extra, fmt, markers):
    """gs_save / gs_load: the packed store carries the int16 matrix exactly as
    ig_load_genotypes() takes it and every table the result writer reads; a damaged file is refused."""
    rng = np.random.default_rng(5)
    d = make_dataset(N=23, L=14, K=3, A=5, miss=0.08, seed=78)
    x = d.x.copy()
    x[3] = np.where(x[3] >= 0, 0, x[3])
    alleles = [[str(int(v)) for v in rng.permutation(np.arange(100, 140))[:8]] for _ in range(x.shape[0])]
    p = str(tmp_path / "geno.txt")
    _write_text(p, x, d.pop, label, popdata, extra, fmt, markers, alleles)
    opt = GsOptions(2, 23, 14, b"-9", label, popdata, extra, markers, fmt, 1)
    a, b = GsStore(), GsStore()
    err = C.create_string_buffer(512)
    assert host.gs_read(p.encode(), C.byref(opt), C.byref(a), err, 512) == 0, err.value
    q = str(tmp_path / "geno.igs")
    assert host.gs_save(q.encode(), C.byref(a), err, 512) == 0, err.value
    assert host.gs_load(q.encode(), C.byref(b), err, 512) == 0, err.value
    for f in ("ploid", "totalsize", "locinum", "locinum_file", "allelenum_max", "pop_count", "n_extra_col"):
        assert getattr(a, f) == getattr(b, f), f
    N, L = a.totalsize, a.locinum
    assert np.array_equal(np.ctypeslib.as_array(a.x, (L, N, 2)), np.ctypeslib.as_array(b.x, (L, N, 2)))
    for f, n in (("allelenum", L), ("locus_of", L), ("missvec", N)):
        assert np.array_equal(np.ctypeslib.as_array(getattr(a, f), (n,)), np.ctypeslib.as_array(getattr(b, f), (n,))), f
    an = np.ctypeslib.as_array(a.allelenum, (L,))
    ta, tb = C.cast(a.alleletype, C.POINTER(C.c_void_p)), C.cast(b.alleletype, C.POINTER(C.c_void_p))
    for l in range(L):
        assert _strings(ta[l], int(an[l])) == _strings(tb[l], int(an[l]))
    assert bool(a.marker_names) == bool(b.marker_names) == bool(markers)
    if markers:
        assert _strings(a.marker_names, a.locinum_file) == _strings(b.marker_names, a.locinum_file)
    assert bool(a.indvname) == bool(b.indvname)
    if label:
        assert [a.indvname[i] for i in range(N)] == [b.indvname[i] for i in range(N)]
    if popdata:
        assert np.array_equal(np.ctypeslib.as_array(a.popindx, (N,)), np.ctypeslib.as_array(b.popindx, (N,)))
        assert [a.poptype[i] for i in range(a.pop_count)] == [b.poptype[i] for i in range(a.pop_count)]
    if extra:
        ea, eb = C.cast(a.extra_col, C.POINTER(C.c_void_p)), C.cast(b.extra_col, C.POINTER(C.c_void_p))
        for i in range(N):
            assert _strings(ea[i], extra) == _strings(eb[i], extra)
    host.gs_free(C.byref(b))
    # a truncated file and a foreign file are refused with a message, not read
    raw = open(q, "rb").read()
    open(q, "wb").write(raw[: len(raw) // 2])
    assert host.gs_load(q.encode(), C.byref(b), err, 512) != 0 and b"packed store" in err.value
    open(q, "wb").write(b"NOTASTORE" + raw[9:])
    assert host.gs_load(q.encode(), C.byref(b), err, 512) != 0 and b"magic" in err.value
    host.gs_free(C.byref(a))


def test_cli_pack_only_needs_no_gpu(tmp_path):
    """`inbreed --pack-only --save-store f`: text -> packed conversion stops before any CUDA call."""
    out = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    d = make_dataset(N=30, L=12, K=2, A=4, miss=0.05, seed=79)
    alleles = [[str(100 + a) for a in range(8)] for _ in range(d.x.shape[0])]
    p = str(tmp_path / "g.txt")
    _write_text(p, d.x, d.pop, 1, 1, 0, 0, 0, alleles)
    q = str(tmp_path / "g.igs")
    r = subprocess.run([os.path.join(HOST, "inbreed"), "-d", p, "-o", str(tmp_path / "o.txt"), "-K", "2", "-L", "12", "-N", "30",
                        "-lb", "1", "-a", "1", "--quiet-data", "--pack-only", "--save-store", q], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(q, "rb").read()
    assert raw[:8] == b"IGSTORE1" and raw[-4:] == b"ENDE"
    hdr = np.frombuffer(raw[8:8 + 20], dtype=np.int32)
    assert tuple(hdr[:3]) == (2, 30, 12)


@pytest.mark.parametrize("case", range(24))
def test_packer_fuzz_against_reference_reader(host, tmp_path, case):
    """Random small files (shape, allele count, missingness up to whole individuals and whole loci, monomorphic loci,
    both line formats, every combination of label / population / extra / marker columns): the packer and the reference's
    own reader must agree on every cell, allele count and missing count."""
    rng = np.random.default_rng(1000 + case)
    N, L, A = int(rng.integers(2, 30)), int(rng.integers(1, 25)), int(rng.integers(2, 9))
    label, popdata, extra, fmt, markers = (int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 3)),
                                           int(rng.integers(0, 2)), int(rng.integers(0, 2)))
    x = rng.integers(0, A, size=(L, N, 2)).astype(np.int16)
    miss = rng.random((L, N, 2)) < rng.choice([0.0, 0.05, 0.4])
    x[miss] = -9
    for l in range(L):
        r = rng.random()
        if r < 0.15:
            x[l] = np.where(x[l] >= 0, x[l].max(), x[l])          # monomorphic (possibly with missing cells)
        elif r < 0.22:
            x[l] = -9                                             # nothing observed at this locus
    if N > 3 and rng.random() < 0.3:
        x[:, int(rng.integers(0, N)), :] = -9                     # an individual with no data at all
    pops = rng.integers(0, 3, size=N)
    alleles = [[str(int(v)) for v in rng.permutation(np.arange(100, 160))[:A]] for _ in range(L)]
    p = str(tmp_path / "geno.txt")
    _write_text(p, x, pops, label, popdata, extra, fmt, markers, alleles)
    opt = GsOptions(2, N, L, b"-9", label, popdata, extra, markers, fmt, 1)
    st = GsStore()
    err = C.create_string_buffer(512)
    rc = host.gs_read(p.encode(), C.byref(opt), C.byref(st), err, 512)
    try:
        rx, ran, rmiss, _ = pyoracle.ref_read_data(p, 2, N, 3, L, label=label, popdata=popdata, n_extra_col=extra,
                                                   markername_flag=markers, datafmt=fmt)
    except Exception:
        rx = None
    if rc != 0:
        # the only files the packer may refuse are those the reference cannot use either (no polymorphic locus left)
        assert rx is None or rx.shape[0] == 0, err.value
        return
    assert rx is not None
    Ls, Ns = st.locinum, st.totalsize
    assert (Ls, Ns) == (rx.shape[0], rx.shape[1])
    if Ls:
        assert np.array_equal(np.ctypeslib.as_array(st.allelenum, (Ls,)), ran)
        assert np.array_equal(np.ctypeslib.as_array(st.x, (Ls, Ns, 2)), rx)
        assert np.array_equal(np.ctypeslib.as_array(st.missvec, (Ns,)), rmiss.sum(axis=0))
    host.gs_free(C.byref(st))


@pytest.mark.parametrize("case", range(12))
def test_packer_fuzz_tetraploid(host, tmp_path, case):
    """-p 4, random files: copies in any order, partly and wholly missing genotypes, monomorphic loci."""
    from instruct_b200.synth import write_reference_text_tetra
    rng = np.random.default_rng(2000 + case)
    N, L, A = int(rng.integers(2, 25)), int(rng.integers(1, 16)), int(rng.integers(2, 7))
    copies = rng.integers(0, A, size=(L, N, 4)).astype(np.int16)
    copies[rng.random((L, N, 4)) < rng.choice([0.0, 0.1, 0.5])] = -9
    for l in range(L):
        if rng.random() < 0.2:
            copies[l] = np.where(copies[l] >= 0, 0, copies[l])
    pops = rng.integers(0, 2, size=N)
    p = str(tmp_path / "geno4.txt")
    write_reference_text_tetra(p, copies, pop=pops)
    opt = GsOptions(4, N, L, b"-9", 1, 1, 0, 0, 1, 1)
    st = GsStore()
    err = C.create_string_buffer(512)
    assert host.gs_read(p.encode(), C.byref(opt), C.byref(st), err, 512) == 0, err.value
    Ls, Ns = st.locinum, st.totalsize
    rx, ran, rmiss, rid = pyoracle.ref_read_data(p, 4, N, 2, L, label=1, popdata=1, datafmt=1)
    assert (Ls, Ns) == (rx.shape[0], rx.shape[1])
    mine = np.ctypeslib.as_array(st.x, (Ls, Ns, 4)).copy()
    assert np.array_equal(np.ctypeslib.as_array(st.allelenum, (Ls,)), ran)
    want = rx.copy()
    want[(rid == 0)] = -1
    assert np.array_equal(mine, want)
    assert np.array_equal((mine >= 0).sum(axis=2), rid)
    assert np.array_equal(np.ctypeslib.as_array(st.missvec, (Ns,)), rmiss.sum(axis=0))
    host.gs_free(C.byref(st))


def test_reference_program_bound_to_the_library_fails_loudly_without_a_gpu(tmp_path):
    """oracle/_ref/InStruct_b200 is the reference PROGRAM (its own flag parser, text reader and writers, compiled from
    the sources in place) with mcmc_updating() bound to libinstruct_b200.so by instruct_b200/host/reference_binding/
    mcmc_gpu.c.  Here, without a GPU: it links, reads the file with the reference's reader,
    reaches the library, and the library's refusal ("no CPU path") comes back through the reference's own nrerror()."""
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "InStruct_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/InStruct_b200 not built (needs /root/reference)")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: see tests/test_gpu_cli.py")
    d = make_dataset(N=30, L=8, K=2, A=4, miss=0.03, seed=3)
    from instruct_b200.synth import write_reference_text
    p = str(tmp_path / "g.txt")
    write_reference_text(p, d.x, pop=d.pop)
    r = subprocess.run([exe, "-d", p, "-o", str(tmp_path / "o.txt"), "-K", "2", "-L", "8", "-N", "30", "-u", "50", "-b", "10", "-t", "2", "-c", "1",
                        "-v", "2", "-g", "1", "-r", "5", "-pi", "0", "-s", "13", "4", "1972"], capture_output=True, text=True, timeout=120,
                       cwd=str(tmp_path))
    assert r.returncode != 0
    assert "Chain#1 Starts:" in r.stdout and "no CUDA device: instruct_b200 has no CPU path" in (r.stdout + r.stderr)
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libinstruct_b200.so" in ldd and "not found" not in ldd


def test_rate_convergence_aligns_clusters_and_matches_plain_gelman_rubin(host, tmp_path):
    """wr_rate_convergence (per-parameter Gelman-Rubin of the CLI): chains whose clusters are labelled differently are
    matched through their posterior mean Q before the statistic is taken; the statistic itself is wr_gelman_rubin."""
    rng = np.random.default_rng(3)
    m, n, K, N = 3, 60, 3, 40
    base = rng.normal([0.2, 0.5, 0.8], 0.02, size=(m, n, K))           # well-mixed traces around three rates
    home = rng.integers(0, K, N)
    q0 = np.full((N, K), 0.05); q0[np.arange(N), home] = 0.9
    perms = [np.arange(K), np.array([2, 0, 1]), np.array([1, 2, 0])]       # chain c calls cluster a of chain 0 "perms[c][a]"
    tr = np.zeros((m, n, K)); qq = []
    for c in range(m):
        tr[c][:, perms[c]] = base[c]
        q = np.zeros((N, K)); q[:, perms[c]] = q0
        qq.append(np.ascontiguousarray(q))
    host.wr_rate_convergence.restype = C.c_int
    host.wr_rate_convergence.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_double)), C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    ptrs = (C.POINTER(C.c_double) * m)(*[q.ctypes.data_as(C.POINTER(C.c_double)) for q in qq])
    out = str(tmp_path / "gr.txt")
    trc = np.ascontiguousarray(tr)
    flag = host.wr_rate_convergence(out.encode(), trc.ctypes.data_as(C.POINTER(C.c_double)), ptrs, m, n, K, N, b"selfing rate")
    lines = open(out).read().strip().split("\n")
    assert len(lines) == K
    grs = []
    for a in range(K):
        one = np.ascontiguousarray(base[:, :, a])
        want = host.wr_gelman_rubin(one.ctypes.data_as(C.POINTER(C.c_double)), m, n)
        got = float(lines[a].rstrip(".").split()[-1])
        assert abs(got - want) < 1e-6 and f"cluster {a + 1}" in lines[a]
        grs.append(want)
    assert flag == int(max(grs) > 1.1) and max(grs) < 1.3
    # without the alignment the statistic explodes (means 0.2 / 0.5 / 0.8 mixed up): the matching is what makes it usable
    mixed = np.ascontiguousarray(tr[:, :, 0])
    assert host.wr_gelman_rubin(mixed.ctypes.data_as(C.POINTER(C.c_double)), m, n) > 5.0
