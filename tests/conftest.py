import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no libb200sv.so (built artefacts are git-ignored): build it once if nvcc is here.  Only when it is
    # MISSING -- a library that travelled with a snapshot is used as it is, whatever the file times say.  If the build fails
    # the tests that need the library fail loudly.
    try:
        from moveit2_b200 import build as _build
        if not os.path.exists(_build.LIB):
            _build.build(force=True)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("conftest: libb200sv.so not built: %s\n" % e)


@pytest.fixture(scope="session")
def resources():
    return os.path.join(ROOT, "moveit2_b200", "resources")


@pytest.fixture(scope="session")
def golden():
    return os.path.join(ROOT, "tests", "golden")


def load_text(*parts):
    with open(os.path.join(*parts)) as f:
        return f.read()
