"""The driver's entry points: `__graft_entry__.smoke()` (two tiny collect + update iterations on cuda:0 through the public
API, checked against the oracle at 1e-4) must pass on the GPU box with the shipped library."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def test_smoke_entry_point():
    import __graft_entry__ as entry

    entry.smoke()
