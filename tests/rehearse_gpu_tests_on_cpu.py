#!/usr/bin/env python
"""Development aid, NOT part of either test run: executes `-m gpu` test files without a GPU by putting the CPU oracle's
build of the C ABI under the product's host layer (tests/cpu_device.py) and redirecting torch's "cuda" placements to host
memory.  It rehearses the TEST CODE and the product's Python host code after a change — API drift, program validity,
byte equality with the CPU restatement — before GPU time is spent; it says nothing about the kernels.  Expected
failures under the rehearsal: tests that assert `.is_cuda`, tests that launch the C++ binaries (they need a device) and
argument-validation cases the CPU build does not replicate.

    python tests/rehearse_gpu_tests_on_cpu.py tests/test_gpu_stark.py tests/test_gpu_deep.py ...

Used at the end of round 2 (GPU budget spent) for the compiler rewrites of DESIGN.md §8.1: 153 of the 162 selected GPU
tests passed here, the 9 others for the reasons above."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import cpu_device  # noqa: E402

cpu_device.install()
import torch  # noqa: E402

CPU = torch.device("cpu")


def _host(d):
    if isinstance(d, int) or (isinstance(d, (str, torch.device)) and "cuda" in str(d)):
        return CPU
    return d


def _factory(orig):
    def f(*a, **k):
        if "device" in k:
            k["device"] = _host(k["device"])
        k.pop("pin_memory", None)
        return orig(*a, **k)
    return f


for _name in ("empty", "zeros", "ones", "tensor", "arange", "full", "randint", "empty_like", "zeros_like", "frombuffer", "as_tensor"):
    setattr(torch, _name, _factory(getattr(torch, _name)))
_to = torch.Tensor.to


def _to_host(self, *a, **k):
    a = tuple(_host(x) if isinstance(x, (str, torch.device)) else x for x in a)
    if "device" in k:
        k["device"] = _host(k["device"])
    k.pop("non_blocking", None)
    return _to(self, *a, **k)


torch.Tensor.to = _to_host
torch.Tensor.cuda = lambda self, *a, **k: self.clone()
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.device_count = lambda: 1
torch.cuda.set_device = lambda *a, **k: None


class _Stream(cpu_device._Stream):
    def __init__(self, *a, **k):
        pass


torch.cuda.Stream = _Stream
torch.cuda.current_stream = lambda *a, **k: _Stream()

if __name__ == "__main__":
    import pytest
    sys.exit(pytest.main(sys.argv[1:] + ["-q", "-m", "gpu", "-p", "no:cacheprovider", "--tb=line"]))
