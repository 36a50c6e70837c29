"""instruct_b200 -- B200-native per-sweep MCMC hot path of InStruct (modes 2/3 of
slowkoni/InStruct behind the reference's own mcmc_updating() boundary).

Everything numerical runs in instruct_b200/libinstruct_b200.so (hand-written sm_100a CUDA
behind the C-ABI of include/instruct_b200.h).  This package is the host-side mirror of the
reference interface for tests and benchmarks; it has no CPU or PyTorch fallback.
"""
from ._lib import InstructError, LIB_PATH, load            # noqa: F401
from .sampler import Chain, Convg, Init, Sampler, SeqData, mcmc_updating   # noqa: F401

__all__ = ["InstructError", "LIB_PATH", "load", "Chain", "Convg", "Init", "Sampler", "SeqData", "mcmc_updating"]
