class MultiAgentEnv:
    """No-op base: the reference only calls super().__init__() and super().reset(seed=...)."""

    def __init__(self, *args, **kwargs):
        pass

    def reset(self, *, seed=None, options=None):
        return None

    def close(self):
        pass
