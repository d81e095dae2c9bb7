import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "polymer-stats_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat_energy.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cli_table():
    with open(os.path.join(GOLDEN, "cli_table.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pm():
    """The product's Python host package; the GPU tests call the CUDA path through its C ABI."""
    import polymc
    polymc.load()
    return polymc


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.lib()
    return oracle


def both_cases(pm, O, **kw):
    """The same case for the CUDA library and for the oracle."""
    okw = {k: v for k, v in kw.items() if k not in ("force_init", "accum_mode")}
    return pm.make_case(**kw), O.make_case(**okw)
