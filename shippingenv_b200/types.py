"""Enumerations and constants of the environment, with the reference's names and values
(shipping/type.py:1-46, shipping/environment.py:8-26) so that agent code written against
``shipping`` keeps working."""


class ActionType:            # shipping/type.py:1-5
    MOVE_SHIP = 1
    SELECT_PORT = 2
    TAKE_FUEL = 3
    TAKE_CARGO = 4


class ShipMove:              # shipping/type.py:8-16, values are (dx, dy) on np_game[x, y]
    NORTH = (0, -1)
    SOUTH = (0, 1)
    EAST = (-1, 0)
    WEST = (1, 0)


# flat move index (utils/preprocessing.py:126, shipping/environment.py:166): N, E, S, W
MOVES = (ShipMove.NORTH, ShipMove.EAST, ShipMove.SOUTH, ShipMove.WEST)
MOVE_INDEX = {m: k for k, m in enumerate(MOVES)}


class Entity:                # shipping/type.py:19-30
    GROUND = 0
    WATER = 1
    PORT = 2
    DESTINY_PORT = 3
    BOAT = 5
    STORM = 6
    TRAVEL = 7


class Color:                 # shipping/type.py:33-46 (BGR)
    WATER = [255, 0, 0]
    GROUND = [37, 73, 141]
    TRAVEL = [0, 0, 255]
    PORT = [154, 152, 150]
    DESTINY_PORT = [0, 255, 0]
    STORM = [128, 128, 128]
    BOAT = [65, 138, 222]
    WHITE = [255, 255, 255]
    BLACK = [0, 0, 0]


class Initial:               # shipping/environment.py:8-11
    CARGO = 0
    FUEL = 200
    MAX_CARGO_CAPACITY = 50


class Reward:                # shipping/environment.py:13-18
    CARGO_DELIVER = 2
    REACH_DESTINATION = 10
    CLOSER_TO_DESTINATION = 2
    TAKE_FUEL = 0.05
    TAKE_CARGO = 0.05


class Penalty:               # shipping/environment.py:20-26
    RUN_OUT_OF_FUEL = -10
    MOVE_ON_GROUND = -5
    MOVE_ON_WATER = -1
    CARGO_LOSS = -3
    FARTHER_FROM_DESTINATION = -2
    USE_FUEL = -0.0001
