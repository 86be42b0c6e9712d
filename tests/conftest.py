import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle) if it is not there; on the GPU box the prebuilt
    .so files travel with the snapshot and /root/reference does not exist."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwtref.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwtrefbw.so")) \
            and os.path.exists(os.path.join(ROOT, "wiggletools_b200", "libwiggletools_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refbw"])
    if os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "integration", "_build", "wiggletools_b200")) \
            and os.path.exists(os.path.join(ROOT, "wiggletools_b200", "libwiggletools_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "integration")])


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Oracle
    try:
        return Oracle("reference")
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libwtref.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="session")
def engine():
    from wiggletools_b200.api import Engine
    return Engine(0)
