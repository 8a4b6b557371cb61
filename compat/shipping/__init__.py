"""Drop-in ``shipping`` package: put ``<repo>/compat`` before the reference checkout on PYTHONPATH and
the reference's agents (agents/*.py, main.py) import the B200 environment instead of shipping/environment.py.
Mirrors shipping/__init__.py:1-2."""
from shippingenv_b200.environment import Environment  # noqa: F401
from shippingenv_b200.types import ShipMove  # noqa: F401
