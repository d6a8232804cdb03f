import warnings


def warn(msg, *args, **kwargs):
    warnings.warn(str(msg) % args if args else str(msg))


def error(msg, *args):
    print("ERROR:", msg % args if args else msg)


def info(msg, *args):
    pass


def debug(msg, *args):
    pass
