"""End-to-end drop-in check of the host program: `inbreed` (GPU) and the compiled reference
`InStruct` (CPU) are run on the same file with the same flags; the result files must have
the same banner bytes and the same table structure, and agree numerically within MCMC noise."""
import os
import re
import subprocess

import numpy as np
import pytest

from instruct_b200.synth import make_dataset, write_reference_text

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INBREED = os.path.join(ROOT, "instruct_b200", "host", "inbreed")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "InStruct")


def _floats(line):
    return [float(v) for v in re.findall(r"-?\d+\.\d+", line)]


@pytest.mark.skipif(not os.path.exists(REFBIN), reason="reference binary not built")
def test_cli_matches_reference_program(tmp_path):
    assert os.path.exists(INBREED), "build the host program: make -C instruct_b200/host"
    d = make_dataset(N=120, L=12, K=2, A=6, miss=0.03, seed=2024, pure=True)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "3000", "-b", "1000", "-t", "5", "-c", "2",
             "-v", "2", "-f", "0", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972"]
    outs = {}
    for name, exe, extra in (("ref", REFBIN, []), ("gpu", INBREED, ["--quiet-data"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=600,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
        outs[name] = open(out, "rb").read()
        if name == "gpu":      # per-parameter convergence (stdout only: the result file keeps the reference's format)
            gr = [float(v) for v in re.findall(r"Gelman-Rubin statistics of the selfing rate of cluster \d+ is (-?\d+\.\d+)", p.stdout)]
            assert len(gr) == 2 and all(0.8 < v < 2.5 for v in gr), p.stdout[-1500:]
    # ---- banner: identical bytes apart from the echoed command line
    def banner(t):
        t = t[: t.index(b"Chain#1")]
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        return re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t)
    assert banner(outs["ref"]) == banner(outs["gpu"])
    # ---- same sequence of section headers and the same number of table rows
    def skeleton(t):
        return [re.sub(rb"-?\d+\.\d+", b"#", ln) for ln in t[t.index(b"Chain#1"):].split(b"\n")
                if not ln.startswith(b"The Gelman-Rubin")]
    sr, sg = skeleton(outs["ref"]), skeleton(outs["gpu"])
    assert len(sr) == len(sg)
    assert sum(a == b for a, b in zip(sr, sg)) > 0.97 * len(sr)      # a few rows differ in a printed digit count only
    # ---- numbers: selfing rates and log-likelihood of both chains agree with the reference's
    def selfing(t):
        rows = [ln for ln in t.decode(errors="ignore").split("\n") if ln.startswith("Cluster ")]
        return np.array([_floats(r)[0] for r in rows]).reshape(2, 2)
    def loglik(t):
        return np.array([_floats(ln)[0] for ln in t.decode(errors="ignore").split("\n") if "Posterior Mean" in ln])
    assert np.abs(selfing(outs["ref"]).mean(0) - selfing(outs["gpu"]).mean(0)).max() < 0.08
    assert np.abs(loglik(outs["ref"]).mean() - loglik(outs["gpu"]).mean()) < 15.0


@pytest.mark.skipif(not os.path.exists(REFBIN), reason="reference binary not built")
@pytest.mark.parametrize("ap", [1, 0])
def test_cli_tetraploid_matches_reference_program(tmp_path, ap):
    """`-p 4 -ap 1` / `-ap 0`: the one-line-per-individual tetraploid format, the autotetraploid and
    allotetraploid drivers and the ploid-4 result tables, against the compiled reference on the same file."""
    from instruct_b200.synth import make_tetra_dataset, write_reference_text_tetra
    d = make_tetra_dataset(N=80, L=16, K=2, A=4, miss=0.03, seed=31)
    data = str(tmp_path / "geno4.txt")
    write_reference_text_tetra(data, d.dosage, pop=d.pop)
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "4", "-ap", str(ap), "-u", "1200", "-b", "400", "-t", "5", "-c", "2",
             "-v", "2", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972"]
    outs = {}
    for name, exe, extra in (("ref", REFBIN, []), ("gpu", INBREED, ["--quiet-data"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=900,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
        outs[name] = open(out, "rb").read()

    def banner(t):
        t = t[: t.index(b"Chain#1")]
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        return re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t)
    assert banner(outs["ref"]) == banner(outs["gpu"])

    def skeleton(t):
        return [re.sub(rb"-?\d+\.\d+", b"#", ln) for ln in t[t.index(b"Chain#1"):].split(b"\n")
                if not ln.startswith(b"The Gelman-Rubin")]
    sr, sg = skeleton(outs["ref"]), skeleton(outs["gpu"])
    assert len(sr) == len(sg)
    assert sum(a == b for a, b in zip(sr, sg)) > 0.97 * len(sr)

    def selfing(t):
        rows = [ln for ln in t.decode(errors="ignore").split("\n") if ln.startswith("Cluster ")]
        return np.array([_floats(r)[0] for r in rows]).reshape(2, 2)

    def loglik(t):
        return np.array([_floats(ln)[0] for ln in t.decode(errors="ignore").split("\n") if "Posterior Mean" in ln])
    # each chain draws its own alpha once (poly_geno.c:386), so chains differ more than MCMC noise alone
    assert np.abs(np.sort(selfing(outs["ref"]).mean(0)) - np.sort(selfing(outs["gpu"]).mean(0))).max() < 0.15
    # (the allotetraploid posterior moves ~6 % in log-likelihood across the alphas of the 12 reference chains of
    # tests/golden/posterior_allo.npz; the alpha-paired comparison is tests/test_gpu_posterior.py)
    tol = 0.03 if ap == 1 else 0.10
    assert np.abs(loglik(outs["ref"]).mean() - loglik(outs["gpu"]).mean()) < tol * abs(loglik(outs["ref"]).mean())


@pytest.mark.skipif(not os.path.exists(REFBIN), reason="reference binary not built")
def test_cli_k_inference(tmp_path):
    """`-ik 1 -kv 1 3` (inf_K_val, InStruct.c:536-601): every K of the range is run and tabulated, the
    K with the smallest DIC is reported.  Same file, same flags, both programs."""
    d = make_dataset(N=150, L=20, K=2, A=6, miss=0.02, seed=77, pure=True)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "1500", "-b", "500", "-t", "5", "-c", "2",
             "-v", "1", "-g", "1", "-r", "10", "-pi", "0", "-ik", "1", "-kv", "1", "3", "-s", "13", "4", "1972"]
    best, dics = {}, {}
    for name, exe, extra in (("ref", REFBIN, []), ("gpu", INBREED, ["--quiet-data"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=900,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        t = open(out, "rb").read().decode(errors="ignore")
        for k in (1, 2, 3):
            assert f"The current K is {k}" in t
        assert "The range of value for K is (1 - 3)!" in t
        best[name] = int(re.search(r"The optimal K is (\d+)", t).group(1))
        dics[name] = [float(v) for v in re.findall(r"Deviance information criterion of this model is (-?\d+\.\d+)", t)]
        assert len(dics[name]) == 6                       # 3 values of K x 2 chains
    # K = 1 has no admixture freedom: its DIC is the same model on both sides up to MCMC noise
    assert abs(np.mean(dics["ref"][:2]) - np.mean(dics["gpu"][:2])) < 0.01 * abs(np.mean(dics["ref"][:2]))
    # two well-separated pure clusters: K = 1 must lose on both sides
    assert best["ref"] >= 2 and best["gpu"] >= 2, (best, dics)


@pytest.mark.skipif(not os.path.exists(REFBIN), reason="reference binary not built")
@pytest.mark.parametrize("mode", [4, 5])
def test_cli_inbreeding_modes_match_reference_program(tmp_path, mode):
    """`-v 4` / `-v 5`: inbreeding coefficients per population / per individual (uniform prior)."""
    d = make_dataset(N=120, L=12, K=2, A=6, miss=0.03, seed=2025, pure=True)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "3000", "-b", "1000", "-t", "5", "-c", "2",
             "-v", str(mode), "-f", "0", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972"]
    outs = {}
    for name, exe, extra in (("ref", REFBIN, []), ("gpu", INBREED, ["--quiet-data"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=600,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
        outs[name] = open(out, "rb").read()

    def banner(t):
        t = t[: t.index(b"Chain#1")]
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        return re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t)
    assert banner(outs["ref"]) == banner(outs["gpu"])

    def skeleton(t):
        return [re.sub(rb"-?\d+\.\d+", b"#", ln) for ln in t[t.index(b"Chain#1"):].split(b"\n")
                if not ln.startswith(b"The Gelman-Rubin")]
    sr, sg = skeleton(outs["ref"]), skeleton(outs["gpu"])
    assert len(sr) == len(sg)
    assert sum(a == b for a, b in zip(sr, sg)) > 0.97 * len(sr)
    assert b"The Posterior distribution of Inbreeding Coefficients:" in outs["gpu"]

    def coeff(t):
        key = "Cluster " if mode == 4 else "Indv "
        sec = t.decode(errors="ignore")
        sec = sec[sec.index("Inbreeding Coefficients"):]
        rows = [ln for ln in sec.split("\n") if ln.startswith(key)]
        n = 2 if mode == 4 else d.N
        return np.array([_floats(r)[0] for r in rows[:n]])

    def loglik(t):
        return np.array([_floats(ln)[0] for ln in t.decode(errors="ignore").split("\n") if "Posterior Mean" in ln])
    a, b = coeff(outs["ref"]), coeff(outs["gpu"])
    if mode == 4:
        assert np.abs(a - b).max() < 0.08
    else:
        assert abs(a.mean() - b.mean()) < 0.05 and np.corrcoef(a, b)[0, 1] > 0.6
    assert np.abs(loglik(outs["ref"]).mean() - loglik(outs["gpu"]).mean()) < 15.0


@pytest.mark.skipif(not os.path.exists(REFBIN), reason="reference binary not built")
def test_cli_no_admixture_matches_reference_program(tmp_path):
    """`-v 0`: the no-admixture model; the result file lists the classification probabilities."""
    d = make_dataset(N=120, L=12, K=2, A=6, miss=0.03, seed=2026, pure=True)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "1500", "-b", "500", "-t", "5", "-c", "2",
             "-v", "0", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972"]
    outs = {}
    for name, exe, extra in (("ref", REFBIN, []), ("gpu", INBREED, ["--quiet-data"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=600,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
        outs[name] = open(out, "rb").read()

    def banner(t):
        t = t[: t.index(b"Chain#1")]
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        return re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t)
    assert banner(outs["ref"]) == banner(outs["gpu"])

    def skeleton(t):
        return [re.sub(rb"-?\d+\.\d+", b"#", ln) for ln in t[t.index(b"Chain#1"):].split(b"\n")
                if not ln.startswith(b"The Gelman-Rubin")]
    sr, sg = skeleton(outs["ref"]), skeleton(outs["gpu"])
    assert len(sr) == len(sg)
    assert sum(a == b for a, b in zip(sr, sg)) > 0.97 * len(sr)
    assert b"Inferred Classification of individuals:" in outs["gpu"]

    def probs(t):
        sec = t.decode(errors="ignore")
        sec = sec[sec.index("Inferred Classification"):]
        rows = [ln for ln in sec.split("\n") if re.match(r"^\d+\t", ln)][: d.N]
        return np.array([_floats(r)[-2:] for r in rows])

    def loglik(t):
        return np.array([_floats(ln)[0] for ln in t.decode(errors="ignore").split("\n") if "Posterior Mean" in ln])
    a, b = probs(outs["ref"]), probs(outs["gpu"])
    assert a.shape == b.shape == (d.N, 2)
    # cluster labels may be swapped between the programs: compare the partition
    agree = max(np.mean((a[:, 0] > 0.5) == (b[:, 0] > 0.5)), np.mean((a[:, 0] > 0.5) == (b[:, 1] > 0.5)))
    assert agree > 0.95
    assert np.abs(loglik(outs["ref"]).mean() - loglik(outs["gpu"]).mean()) < 15.0


def test_cli_multi_gpu_result_files_are_identical(tmp_path):
    """`--gpus 2`: chains spread over the GPUs, or every chain sharded by individuals over them --
    either way the chains are the same pure functions of (seed, chain), so the result file is
    byte-identical to the one-GPU run."""
    import instruct_b200
    if instruct_b200.load().ig_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    d = make_dataset(N=301, L=40, K=3, A=5, miss=0.04, seed=77)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    flags = ["-K", "3", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "400", "-b", "100", "-t", "5", "-c", "4",
             "-v", "2", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972", "--quiet-data"]
    outs = {}
    for name, extra in (("one", []), ("chains", ["--gpus", "2", "--shard", "chains"]),
                        ("individuals", ["--gpus", "2", "--shard", "individuals"])):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([INBREED, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=600,
                           cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        t = open(out, "rb").read()
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        outs[name] = re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t)
    assert outs["one"] == outs["chains"]
    assert outs["one"] == outs["individuals"]


def test_cli_packed_store_gives_the_same_result_file(tmp_path):
    """--save-store / --load-store: a run that reads the packed store instead of the
    text file writes the same result file, byte for byte apart from the echoed command line."""
    d = make_dataset(N=90, L=15, K=2, A=5, miss=0.04, seed=77)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    store = str(tmp_path / "geno.igs")
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "400", "-b", "100", "-t", "5", "-c", "2",
             "-v", "2", "-g", "1", "-r", "10", "-pi", "0", "-s", "13", "4", "1972", "--quiet-data"]
    outs = []
    for extra in (["--save-store", store], ["--load-store", store]):
        out = str(tmp_path / f"o{len(outs)}.txt")
        p = subprocess.run([INBREED, "-d", data, "-o", out] + flags + extra, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        t = open(out, "rb").read()
        t = re.sub(rb"Command line arguments:\n.*\n", b"", t)
        outs.append(re.sub(rb"Output File:   .*\n", b"Output File:   X\n", t))
    assert os.path.getsize(store) > 2 * d.N * d.L
    assert outs[0] == outs[1]


REFBIN_B200 = os.path.join(ROOT, "oracle", "_ref", "InStruct_b200")


BIND_CASES = {
    # name: (flags beyond the common ones, tetraploid?)
    "mode2": (["-p", "2", "-v", "2", "-f", "0"], False),
    "mode0": (["-p", "2", "-v", "0"], False),
    "mode3": (["-p", "2", "-v", "3", "-f", "0", "-lb", "0"], False),          # -lb 0: the reference's writer crashes on labels in mode 3
    "mode4": (["-p", "2", "-v", "4"], False),
    "mode2_pf": (["-p", "2", "-v", "2", "-f", "0", "-pf", "1"], False),      # print_freq: CHAIN.freq through the binding
    "tetra": (["-p", "4", "-ap", "1", "-v", "2"], True),
}


@pytest.mark.skipif(not (os.path.exists(REFBIN) and os.path.exists(REFBIN_B200)), reason="reference binaries not built")
@pytest.mark.parametrize("case", sorted(BIND_CASES))
def test_reference_program_bound_to_the_library_matches_the_reference(tmp_path, case):
    """The drop-in for real: the reference PROGRAM with its mcmc_updating() bound to libinstruct_b200.so (INTEGRATION.md
    section 2, instruct_b200/host/reference_binding/mcmc_gpu.c) against the stock reference program on the same file:
    same banner bytes, same table skeleton, posterior log-likelihood (and selfing rates where the mode has them) within
    MCMC noise.  Modes 0 (z reconstruction), 2, 3 (N selfing rates), 4 (inbreed slots), print_freq and ploid 4."""
    extra, tetra = BIND_CASES[case]
    if tetra:
        from instruct_b200.synth import make_tetra_dataset, write_reference_text_tetra
        d = make_tetra_dataset(N=80, L=16, K=2, A=4, miss=0.03, seed=31)
        data = str(tmp_path / "geno4.txt")
        write_reference_text_tetra(data, d.dosage, pop=d.pop)
        upd = ["-u", "1200", "-b", "400"]
    else:
        d = make_dataset(N=120, L=12, K=2, A=6, miss=0.03, seed=2024, pure=True)
        data = str(tmp_path / "geno.txt")
        write_reference_text(data, d.x, pop=d.pop, labels="-lb" not in extra)
        upd = ["-u", "3000", "-b", "1000"]
    flags = ["-K", "2", "-L", str(d.L), "-N", str(d.N)] + upd + ["-t", "5", "-c", "2", "-g", "1", "-r", "10", "-pi", "0",
                                                                  "-s", "13", "4", "1972"] + extra
    outs = {}
    for name, exe in (("ref", REFBIN), ("gpu", REFBIN_B200)):
        out = str(tmp_path / f"{name}.out")
        p = subprocess.run([exe, "-d", data, "-o", out] + flags, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
        outs[name] = open(out, "rb").read()
    head = lambda t: re.sub(rb"Output File:   .*\n", b"", re.sub(rb"Command line arguments:\n.*\n", b"", t[: t.index(b"Chain#1")]))
    assert head(outs["ref"]) == head(outs["gpu"])

    def skeleton(t):
        return [re.sub(rb"-?\d+\.\d+", b"#", ln) for ln in t[t.index(b"Chain#1"):].split(b"\n")
                if not ln.startswith(b"The Gelman-Rubin")]
    sr, sg = skeleton(outs["ref"]), skeleton(outs["gpu"])
    assert len(sr) == len(sg)
    assert sum(a == b for a, b in zip(sr, sg)) > 0.95 * len(sr)

    def loglik(t):
        return np.array([_floats(ln)[0] for ln in t.decode(errors="ignore").split("\n") if "Posterior Mean" in ln])
    lr, lg = loglik(outs["ref"]), loglik(outs["gpu"])
    assert lr.size == 2 and lg.size == 2
    assert abs(lr.mean() - lg.mean()) < 0.03 * abs(lr.mean())
    if case in ("mode2", "mode2_pf", "mode4", "tetra"):
        def rates(t):
            rows = [ln for ln in t.decode(errors="ignore").split("\n") if ln.startswith("Cluster ")]
            return np.array([_floats(r)[0] for r in rows]).reshape(2, 2)
        tol = 0.15 if tetra else 0.08
        assert np.abs(np.sort(rates(outs["ref"]).mean(0)) - np.sort(rates(outs["gpu"]).mean(0))).max() < tol


@pytest.mark.skipif(not os.path.exists(REFBIN_B200), reason="reference binding not built")
def test_reference_binding_without_gelman_rubin(tmp_path):
    """`-g 0`: the reference never allocates CONVG (InStruct.c:181), so the binding must not read it.  The stock
    reference itself segfaults here; the bound program has to finish."""
    d = make_dataset(N=60, L=8, K=2, A=4, miss=0.0, seed=5, pure=True)
    data = str(tmp_path / "geno.txt")
    write_reference_text(data, d.x, pop=d.pop)
    out = str(tmp_path / "o.txt")
    p = subprocess.run([REFBIN_B200, "-d", data, "-o", out, "-K", "2", "-L", str(d.L), "-N", str(d.N), "-p", "2", "-u", "400", "-b", "200",
                        "-t", "5", "-c", "1", "-v", "2", "-g", "0", "-pi", "0"], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "THE JOB IS SUCCESSFULLY FINISHED" in p.stdout
