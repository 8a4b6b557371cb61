"""Subprocess body of tests/test_reference_agents_cpu.py: the reference's agents, unmodified, on the facade.

    PYTHONPATH=<repo>/compat:<repo>:<repo>/tests:/root/reference python tests/run_reference_agents.py

SHIPENV_AGENTS_ON=cuda runs them on the real kernels (tests/test_gpu_facade.py, where a reference checkout exists:
$SHIPPINGENV_REF or /root/reference); otherwise the facade's launches are computed by the oracle.
"""
import json
import os
import random
import sys
import types

for name in ("matplotlib", "matplotlib.pyplot"):          # agents/sarsa.py imports pyplot; not installed here
    sys.modules.setdefault(name, types.ModuleType(name))

REF = os.environ.get("SHIPPINGENV_REF", "/root/reference")
ON_CUDA = os.environ.get("SHIPENV_AGENTS_ON") == "cuda"    # tests/test_gpu_facade.py: the real kernels
contexts = {}
if not ON_CUDA:
    import facade_on_oracle
    contexts = facade_on_oracle.install()

import shipping                                            # compat/shipping: the drop-in
assert "compat" in shipping.__file__, shipping.__file__
from shipping import Environment
from agents.mcts import MCTSAgent                          # /root/reference/agents, unmodified
from agents.sarsa import SARSAAgent
from agents.dqn import DQNAgent
import agents
assert agents.__file__.startswith(REF.rstrip("/") + "/"), agents.__file__
from utils.constants import DEFAULT_PORTS

os.chdir(REF)                                # MCTSAgent re-opens "mapa_mundi_binario.jpg" by relative path
random.seed(7)
out = {}


def make_env():
    env = Environment("mapa_mundi_binario.jpg")
    for p in DEFAULT_PORTS:
        env.add_port(p)
    return env


env = make_env()
res = MCTSAgent(env, num_simulations=6, max_rollout_steps=15).train(episodes=2, render=False, verbose=False, max_steps=12,
                                                                   max_total_steps=20)
s = res.get_summary()
out["mcts"] = {"episodes": s["total_episodes"], "mean_reward": s["mean_reward"]}

# SARSAAgent.train() evaluates the greedy policy after the first episode with `while not done`, which never ends once
# the greedy action is a port selection (no fuel is used; true of the reference environment as well), so its loop is
# driven from here, method by method (agents/sarsa.py:271-296)
env = make_env()
agent = SARSAAgent(env)
steps, episodes, total = 0, 0, 0.0
while steps < 400:
    state = env.reset()
    action = agent.choose_action(state)
    done = False
    while not done and steps < 400:
        next_state, reward, done, _ = agent._try_action(action)
        next_action = agent.choose_action(next_state)
        agent.update(state, action, reward, next_state, next_action)
        state, action, total, steps = next_state, next_action, total + reward, steps + 1
    agent.decay_epsilon()
    episodes += 1
out["sarsa"] = {"episodes": episodes, "q_entries": len(agent.q_table), "mean_reward": total / episodes}

env = make_env()
agent = DQNAgent(env)
res = agent.train(episodes=3, render=False, verbose=False, max_steps=40, max_total_steps=100)
s = res.get_summary()
out["dqn"] = {"episodes": s["total_episodes"], "mean_reward": s["mean_reward"]}
if ON_CUDA:
    from shippingenv_b200 import launch_count
    out["launches"] = launch_count()
else:
    out["launches"] = sum(c[0].launches for c in contexts.values())
print("RESULT " + json.dumps(out))
