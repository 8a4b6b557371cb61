from shippingenv_b200.types import ActionType, Color, Entity, ShipMove  # noqa: F401
