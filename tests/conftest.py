import os
import sys

import pytest

# a few tests A/B developer knobs (environment variables the library only honours in developer mode)
os.environ.setdefault("TFFT_DEVELOPER", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensor-fft_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build every native artefact once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
