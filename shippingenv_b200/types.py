"""Enumerations and constants of the environment, with the reference's names and values
(shipping/type.py:1-46, shipping/environment.py:8-26), so that agent code written against
``shipping`` keeps working.  Values only -- there is no logic here."""


class ActionType:            # shipping/type.py:1-5
    MOVE_SHIP, SELECT_PORT, TAKE_FUEL, TAKE_CARGO = 1, 2, 3, 4


class ShipMove:              # shipping/type.py:8-16: (dx, dy) on np_game[x, y]
    NORTH, SOUTH, EAST, WEST = (0, -1), (0, 1), (-1, 0), (1, 0)


# flat move index (utils/preprocessing.py:126, shipping/environment.py:166): N, E, S, W
MOVES = (ShipMove.NORTH, ShipMove.EAST, ShipMove.SOUTH, ShipMove.WEST)
MOVE_INDEX = {m: k for k, m in enumerate(MOVES)}


class Entity:                # shipping/type.py:19-30 (cell codes of np_game; 4 is unused upstream)
    GROUND, WATER, PORT, DESTINY_PORT = 0, 1, 2, 3
    BOAT, STORM, TRAVEL = 5, 6, 7


class Color:                 # shipping/type.py:33-46, BGR triples used by the renderer only
    WATER, GROUND, TRAVEL = [255, 0, 0], [37, 73, 141], [0, 0, 255]
    PORT, DESTINY_PORT, STORM = [154, 152, 150], [0, 255, 0], [128, 128, 128]
    BOAT, WHITE, BLACK = [65, 138, 222], [255, 255, 255], [0, 0, 0]


class Initial:               # shipping/environment.py:8-11
    CARGO, FUEL, MAX_CARGO_CAPACITY = 0, 200, 50


class Reward:                # shipping/environment.py:13-18
    CARGO_DELIVER, REACH_DESTINATION, CLOSER_TO_DESTINATION = 2, 10, 2
    TAKE_FUEL = TAKE_CARGO = 0.05


class Penalty:               # shipping/environment.py:20-26
    RUN_OUT_OF_FUEL, MOVE_ON_GROUND, MOVE_ON_WATER = -10, -5, -1
    CARGO_LOSS, FARTHER_FROM_DESTINATION, USE_FUEL = -3, -2, -0.0001
