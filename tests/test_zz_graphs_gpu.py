"""GPU: panic3d_b200.graphs.GraphedCallable - capture once per signature, bit-identical replay, fresh noise per replay,
autograd and failing captures stay eager.  The bodies live in tests/graphs_gpu_cases.py and run in their own interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('case', ['capture_replay_equals_eager_and_noise_is_fresh', 'autograd_calls_and_failing_captures_stay_eager'])
def test_graphs(case):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'graphs_gpu_cases.py'), case], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'CASE_OK ' + case in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
