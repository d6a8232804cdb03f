"""Callback hook points of the training loop (reference: openrl/utils/callbacks/callbacks.py:14-172;
call sites ppo_agent.py:128-132, onpolicy_driver.py:155,174-178,196).  The same duck type is
accepted: on_training_start(locals, globals), on_rollout_start(), update_locals(locals),
on_step() -> bool, on_rollout_end(), on_training_end()."""


class BaseCallback:
    def __init__(self, verbose=0):
        self.agent = None
        self.n_calls = 0
        self.num_time_steps = 0
        self.verbose = verbose
        self.locals = {}
        self.globals = {}
        self.parent = None

    def init_callback(self, agent):
        self.agent = agent
        self._init_callback()

    def _init_callback(self):
        pass

    def set_parent(self, parent):
        self.parent = parent

    def on_training_start(self, locals_, globals_):
        self.locals, self.globals = locals_, globals_
        self.num_time_steps = self.agent.num_time_steps
        self._on_training_start()

    def _on_training_start(self):
        pass

    def on_rollout_start(self):
        self._on_rollout_start()

    def _on_rollout_start(self):
        pass

    def _on_step(self):
        return True

    def on_step(self):
        self.n_calls += 1
        self.num_time_steps = self.agent.num_time_steps
        return self._on_step()

    def on_training_end(self):
        self._on_training_end()

    def _on_training_end(self):
        pass

    def on_rollout_end(self):
        self._on_rollout_end()

    def _on_rollout_end(self):
        pass

    def update_locals(self, locals_):
        self.locals.update(locals_)
        self.update_child_locals(locals_)

    def update_child_locals(self, locals_):
        pass

    # True when the callback never looks at per-step locals, so the whole rollout may run as one
    # device launch.
    needs_per_step = True


class CallbackList(BaseCallback):
    def __init__(self, callbacks):
        super().__init__()
        self.callbacks = list(callbacks)

    @property
    def needs_per_step(self):
        return any(getattr(c, "needs_per_step", True) for c in self.callbacks)

    def _init_callback(self):
        for c in self.callbacks:
            c.init_callback(self.agent)

    def _on_training_start(self):
        for c in self.callbacks:
            c.on_training_start(self.locals, self.globals)

    def _on_rollout_start(self):
        for c in self.callbacks:
            c.on_rollout_start()

    def _on_step(self):
        cont = True
        for c in self.callbacks:
            cont = c.on_step() and cont
        return cont

    def _on_rollout_end(self):
        for c in self.callbacks:
            c.on_rollout_end()

    def _on_training_end(self):
        for c in self.callbacks:
            c.on_training_end()

    def update_child_locals(self, locals_):
        for c in self.callbacks:
            c.update_locals(locals_)


class StopTrainingOnMaxSteps(BaseCallback):
    """Minimal stand-in used by tests: stops after max_calls on_step() calls."""

    def __init__(self, max_calls):
        super().__init__()
        self.max_calls = max_calls

    def _on_step(self):
        return self.n_calls < self.max_calls
