import os
import sys

import pytest

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_REPO, os.path.join(_REPO, "oracle"), os.path.join(_REPO, "tests")):
	if p not in sys.path:
		sys.path.insert(0, p)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
	import pyoracle
	pyoracle.build()
	return pyoracle
