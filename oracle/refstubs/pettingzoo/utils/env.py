from typing import TypeVar

ActionType = TypeVar("ActionType")
AgentID = str
ObsType = TypeVar("ObsType")


class AECEnv:
    pass


class ParallelEnv:
    pass
