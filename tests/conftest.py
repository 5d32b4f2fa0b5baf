import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: long-running exhaustive check')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    """The oracle is test infrastructure: (re)build it next to its source if missing or stale."""
    so = os.path.join(ROOT, 'oracle', 'liborb_oracle.so')
    src = os.path.join(ROOT, 'oracle', 'orb_oracle.cpp')
    have_ref = os.path.exists('/root/reference/src/ORBextractor.cc')   # oracle/_ref is (re)built only where the mount exists
    if have_ref or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    yield
