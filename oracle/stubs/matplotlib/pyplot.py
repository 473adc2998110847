def __getattr__(name):
    raise NotImplementedError("matplotlib stub")
