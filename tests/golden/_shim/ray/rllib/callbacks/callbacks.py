class RLlibCallback:
    pass
