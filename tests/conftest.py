import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
  """libpclseg.so is compiled in-tree (nvcc cross-compiles without a GPU); tests never build the oracle into it."""
  from pclsegmentation_b200 import build
  return build.build()


@pytest.fixture(scope="session")
def golden_dir():
  return os.path.join(ROOT, "tests", "golden")
