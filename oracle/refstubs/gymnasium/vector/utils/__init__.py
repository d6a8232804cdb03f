import contextlib
import os

import cloudpickle
import pickle


class CloudpickleWrapper:
    def __init__(self, fn):
        self.fn = fn

    def __getstate__(self):
        return cloudpickle.dumps(self.fn)

    def __setstate__(self, ob):
        self.fn = pickle.loads(ob)

    def __call__(self):
        return self.fn()


@contextlib.contextmanager
def clear_mpi_env_vars():
    removed = {}
    for k, v in list(os.environ.items()):
        for prefix in ["OMPI_", "PMI_"]:
            if k.startswith(prefix):
                removed[k] = v
                del os.environ[k]
    try:
        yield
    finally:
        os.environ.update(removed)
