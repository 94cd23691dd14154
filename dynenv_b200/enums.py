"""Public enums of the drop-in API: same names and values as DynEnv/cutils.py:10-64,144-165."""
from enum import IntEnum


class _ArgEnum(IntEnum):
    def __str__(self):
        return self.name.lower()

    def __repr__(self):
        return str(self)

    @classmethod
    def argparse(cls, s):
        try:
            return cls[s.upper()]
        except KeyError:
            return s


class DynEnvType(_ArgEnum):
    ROBO_CUP = 0
    DRIVE = 1


class NoiseType(_ArgEnum):
    RANDOM = 0
    REALISTIC = 1


class ObservationType(_ArgEnum):
    FULL = 0
    PARTIAL = 1
    IMAGE = 2


class SightingType(IntEnum):
    NoSighting = 0
    Partial = 1
    Distant = 2
    Normal = 3
    Misclassified = 4


class InteractionType(IntEnum):
    NoInter = 0
    Nearby = 1
    Occlude = 2


class LanePosition(IntEnum):
    AtGoal = 0
    InRightLane = 1
    InOpposingLane = 2
    OverRoad = 3
    OffRoad = 4
