"""Build-time guards on the hot kernels, checked without a GPU from `-Xptxas -v` and `cuobjdump -sass` of the in-tree
library (B200_PROFILING.md: look at both before spending GPU time).  They pin the properties DESIGN.md section 4 rests on:
the sweep kernels compile for sm_100a without local-memory spills inside their register budgets (2 / 3 resident CTAs per
SM), stage P through one TMA bulk copy (SASS UBLKCP), read it with 128-bit shared loads, tally with shared-memory
reductions, take logarithms on the XU pipe and never touch local memory or convert float -> int in the inner loop."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "instruct_b200", "csrc")
LIB = os.path.join(ROOT, "instruct_b200", "libinstruct_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="CUDA toolchain not installed")

ZQ = "_ZN2ig15zq_sweep_kernelILi8ELi7ELb0ELi3ELi0EEEvNS_6ZQArgsE"           # K <= 8, Philox-7, modes 1-3: the config-4 instance
ZS = "_ZN2ig15tetra_zs_kernelILi8ELi7EEEvNS_6ZsArgsE"
GENO = "_ZN2ig17tetra_geno_kernelILi8ELi7ELb1EEEvNS_8GenoArgsE"
ALLO = "_ZN2ig22tetra_geno_allo_kernelILi8ELi7ELb1EEEvNS_12GenoAlloArgsE"


@pytest.fixture(scope="module")
def built():
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    logs = {}
    for name in ("zq_sweep", "tetra"):
        txt = open(os.path.join(CSRC, "build", f"{name}.ptxas.log")).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s*: Used (\d+) registers", txt):
            logs[m.group(1)] = dict(stack=int(m.group(2)), spill=int(m.group(3)) + int(m.group(4)), regs=int(m.group(5)))
    return logs


def _sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True)
    assert out.returncode == 0 and "Function :" in out.stdout, out.stderr[-1000:]
    return out.stdout


def test_register_budgets_and_no_spills(built):
    assert built[ZQ]["regs"] <= 128 and built[ZQ]["spill"] == 0 and built[ZQ]["stack"] == 0        # 2 CTAs of 256 threads per SM
    assert built[ZS]["regs"] <= 80 and built[ZS]["spill"] == 0                                      # 3 CTAs per SM
    assert built[GENO]["regs"] <= 80 and built[GENO]["spill"] <= 16                                 # 3 CTAs per SM
    assert built[ALLO]["regs"] <= 128 and built[ALLO]["spill"] == 0 and built[ALLO]["stack"] == 0   # no dynamically indexed arrays


def test_every_instance_targets_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_zq_sweep_sass_shape():
    s = _sass(ZQ)
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", s, flags=re.M)
    cnt = lambda pat: sum(1 for o in ops if re.match(pat, o))
    assert cnt(r"UBLKCP") == 1                                   # the P chunk: ONE TMA bulk copy on an mbarrier
    assert cnt(r"LDS\.128") >= 64 and cnt(r"ATOMS") >= 32        # 128-bit reads of the P rows, shared-memory tally / counters
    assert cnt(r"MUFU\.LG2") >= 24                               # log-likelihood pieces on the XU pipe
    assert cnt(r"LDL|STL") == 0                                  # no local memory
    assert cnt(r"F2I") == 0 and cnt(r"IMAD\.HI") == 0            # the slow conversions / high multiplies stay out
    assert cnt(r"LDG\.E.*128") >= 3 and cnt(r"STG\.E.*128") >= 1  # genotype store and Z as 128-bit vectors


@pytest.mark.parametrize("fun", [ZS, GENO, ALLO])
def test_tetra_passes_sass_shape(fun):
    s = _sass(fun)
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", s, flags=re.M)
    cnt = lambda pat: sum(1 for o in ops if re.match(pat, o))
    assert cnt(r"UBLKCP") >= 1
    assert cnt(r"ATOMS") >= 4
    assert cnt(r"LDG\.E.*128") >= 2 and cnt(r"STG\.E.*128") >= 1
    if fun != GENO:
        assert cnt(r"LDL|STL") == 0
