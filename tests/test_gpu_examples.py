"""GPU: the consumers of the path that BASELINE configs 4 and 5 name run end to end on the device."""
import importlib.util
import math
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("envs,steps,graph", [(1, 300, False), (512, 60, False), (64, 200, True)])
def test_dqn_loop_on_device(envs, steps, graph):
    out = _load("dqn_device").train(envs=envs, steps=steps, max_steps=50, graph=graph)
    assert out["cuda_graph"] == graph
    assert out["agent_steps_per_s"] > 0 and out["final_epsilon"] < 1.0
    assert all(math.isfinite(v) for v in out["loss_first_last"])
    if envs > 1:
        assert out["episodes"]["total_episodes"] > 0


@pytest.mark.parametrize("graph", [False, True])
def test_sarsa_loop_on_device(graph):
    """tabular SARSA over N envs, keys from shipenv_discretize, Q-table in HBM."""
    out = _load("sarsa_device").train(envs=2048, steps=80, max_steps=40, graph=graph)
    assert out["cuda_graph"] == graph and out["agent_steps_per_s"] > 0
    assert out["states_visited"] > 100 and math.isfinite(out["mean_abs_td_last"]) and out["mean_abs_td_last"] > 0
    assert out["episodes"]["total_episodes"] > 0 and out["epsilon"] < 1.0


def test_mcts_agent_on_device_beats_the_random_policy():
    """BASELINE config 4 as an agent: choose_action for every ship each step, against sample_action."""
    mod = _load("mcts_device")
    planned = mod.play(envs=512, steps=40, simulations=20, rollout_steps=30)
    rand = mod.play(envs=512, steps=40, planner=False)
    assert planned["decisions_per_s"] > 0
    assert planned["mean_reward_per_step"] > rand["mean_reward_per_step"]
