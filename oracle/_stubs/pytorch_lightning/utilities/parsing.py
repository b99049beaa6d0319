class AttributeDict(dict):
    """dict with attribute access (enough to unpickle Lightning 0.9 checkpoints)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, val):
        self[key] = val
