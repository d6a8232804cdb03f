class BaseWrapper:
    def __init__(self, env):
        self.env = env


class OrderEnforcingWrapper(BaseWrapper):
    pass


class AssertOutOfBoundsWrapper(BaseWrapper):
    pass
