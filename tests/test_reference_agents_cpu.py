"""CPU, this container only (needs /root/reference): the reference's own agents -- agents/mcts.py, agents/sarsa.py,
agents/dqn.py, unmodified -- import ``shipping`` from compat/ and train on the drop-in facade.  The facade's launches
are computed by the oracle here (tests/facade_on_oracle.py), so what is exercised is the Python contract of SURVEY 8b:
constructor, attributes assigned from outside, state dict, exception types and messages, sample_action / step
bookkeeping.  The same facade on the real kernels is covered by tests/test_gpu_facade.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "agents")), reason="the reference checkout is not on this box")
def test_unmodified_reference_agents_train_on_the_facade():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, os.path.join(ROOT, "tests"), REF])
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_agents.py")], capture_output=True,
                         text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert r["mcts"]["episodes"] >= 1 and r["sarsa"]["episodes"] >= 1 and r["dqn"]["episodes"] >= 1
    assert r["sarsa"]["q_entries"] > 10 and r["launches"] > 500
