"""Logger with the reference's call surface (openrl/utils/logger.py:31: info(), log_info(dict, step));
prints to stdout, optional file.  wandb/tensorboard sinks are out of scope."""
import logging
import os
import sys


class Logger:
    def __init__(self, cfg=None, project_name="openrl_b200", scenario_name="default", exp_name="default",
                 log_path=None, log_to_terminal=True, quiet=False):
        self.quiet = quiet
        self._log = logging.getLogger(f"openrl_b200.{id(self)}")
        self._log.setLevel(logging.INFO)
        self._log.propagate = False
        if log_to_terminal and not quiet:
            self._log.addHandler(logging.StreamHandler(sys.stdout))
        if log_path:
            os.makedirs(log_path, exist_ok=True)
            self._log.addHandler(logging.FileHandler(os.path.join(log_path, "log.txt")))
        self.history = []

    def info(self, msg):
        self._log.info(msg)

    def log_info(self, infos, step):
        self.history.append((step, dict(infos)))
        if not self.quiet:
            self._log.info("step %d: " % step + ", ".join(f"{k}: {float(v):.6g}" for k, v in infos.items()))

    def close(self):
        for h in list(self._log.handlers):
            h.close()
            self._log.removeHandler(h)
