"""The reference's robot demo imports the module in the UI process and first CALLS gicp() in a
forked worker, with lists of tuples in and the result pickled back through a Queue
(robot-visualization.py:151-166,195-200).  CUDA must therefore initialise lazily, inside the call.
The scenario runs in a fresh interpreter (tests/fork_demo.py): pytest's own process has usually
initialised CUDA already, and a forked child of such a process cannot use it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_first_call_in_forked_worker():
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "fork_demo.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "FORK-OK" in out.stdout
