"""``from shipping import environment`` (agents/mcts.py:9, agents/sarsa.py:11): the names those modules use."""
from shippingenv_b200.environment import Environment  # noqa: F401
from shippingenv_b200.types import ActionType, Color, Entity, Initial, Penalty, Reward, ShipMove  # noqa: F401
