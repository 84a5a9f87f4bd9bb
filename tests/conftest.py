import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")
    config.addinivalue_line("markers", "simt: kernel sources executed on the host by the SIMT emulator (tests/simt)")
    config.addinivalue_line("markers", "simt_skip: too slow under the SIMT emulator, GPU only")


def record_parity(test, **counts):
    """parity bookkeeping: how many elements differed / were exposed in a tolerance-based comparison.  Printed (pytest
    -s / -rA show it) and, on the GPU box, appended to gpurun_out/parity_counts.jsonl so that the numbers measured on
    the B200 can be quoted (profiles/r02_parity_counts.md)."""
    line = json.dumps(dict(test=test, **counts))
    print("PARITY " + line)
    try:
        import torch
        if torch.cuda.is_available():
            out = os.path.join(REPO, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "parity_counts.jsonl"), "a") as f:
                f.write(line + "\n")
    except Exception:
        pass


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def to_dev(api, t):
    """tensor -> the device the kernels run on (cuda:N on the GPU box, host memory under the emulator)"""
    return t.to(api._device())


# Every kernel parity test runs twice: [cuda] on the B200 through libcdnet_b200.so (-m gpu, the parity tests
# proper) and [simt] on the host, where tests/simt executes the very same kernel sources under a SIMT emulator
# (-m "not gpu"; catches logic regressions where there is no GPU -- it is test infrastructure, not a CPU path
# of the product).
KERNEL_BACKENDS = [pytest.param("cuda", marks=pytest.mark.gpu), pytest.param("simt", marks=pytest.mark.simt)]


@pytest.fixture(params=KERNEL_BACKENDS)
def kernel_api(request):
    if request.param == "cuda":
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        from cdnet_b200 import api
        yield api
        return
    if request.node.get_closest_marker("simt_skip"):
        pytest.skip("GPU only (too slow under the emulator)")
    from simt import emulated_api
    with emulated_api() as api:
        yield api


@pytest.fixture
def cuda_api():
    """GPU-only tests (pinned staging, streams, NCCL): mark them @pytest.mark.gpu"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cdnet_b200 import api
    return api
