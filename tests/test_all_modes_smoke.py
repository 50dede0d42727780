"""Every kernel family for every supported segment count (2..10): all policy-storage modes, moments in
registers and shared memory, both dynamics variants, screening, trajectories, a ragged last block and
the ARS epilogue must launch and give finite results (tools/sanitize_cases.py)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_every_kernel_family_launches_for_every_n(S):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sanitize_cases
    sanitize_cases.main()
