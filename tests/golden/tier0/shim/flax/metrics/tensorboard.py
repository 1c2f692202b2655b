class SummaryWriter:
    def __init__(self, *a, **k): pass
