class Error(Exception):
    pass


class CustomSpaceError(Error):
    pass


class AlreadyPendingCallError(Error):
    def __init__(self, message, name=None):
        super().__init__(message)
        self.name = name


class NoAsyncCallError(Error):
    def __init__(self, message, name=None):
        super().__init__(message)
        self.name = name


class ClosedEnvironmentError(Error):
    pass


class ResetNeeded(Error):
    pass
