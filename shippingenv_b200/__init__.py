"""shippingenv_b200 -- the ShippingEnv environment step as sm_100a CUDA kernels behind a C ABI.

``BatchedEnvironment``: N ships in HBM (torch tensors), stepped by ``libshipenv.so``.
``Environment``: the reference's single-env Python API (shipping/environment.py) on the same kernels.
There is no CPU fallback; importing works anywhere, constructing an environment needs a B200.
"""
from ._native import ShipEnvError, launch_count  # noqa: F401
from .batched import (BatchedEnvironment, StepResult, unpack_bits, DEFAULT_PORTS,  # noqa: F401
                      DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO)
from .environment import Environment  # noqa: F401
from .types import ActionType, Color, Entity, Initial, Penalty, Reward, ShipMove  # noqa: F401
