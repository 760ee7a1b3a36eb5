import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---- parity report: every GPU parity comparison records achieved vs allowed error; written at session end --------------
_PARITY = []


def parity_record(test: str, tensor: str, err: float, scale: float, allowed: float, note: str = "") -> None:
    """err / allowed are absolute; scale = max|reference|.  tests/test_gpu_*.py call this next to their asserts."""
    _PARITY.append({"test": test, "tensor": tensor, "max_abs_err": float(err), "ref_scale": float(scale),
                    "rel_err": float(err / scale) if scale > 0 else 0.0, "allowed_abs": float(allowed),
                    "allowed_rel": float(allowed / scale) if scale > 0 else 0.0, "note": note})


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump({"exitstatus": int(exitstatus), "entries": _PARITY}, f, indent=1)
