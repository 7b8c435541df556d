"""Small helpers with the reference's names (xw_mcts/utils.py:1-22)."""


class AverageMeter(object):
    """Running average of a scalar (used for the training losses, NNet.py:36-37,60-62)."""

    def __init__(self):
        self.val = self.avg = self.sum = self.count = 0

    def __repr__(self):
        return f"{self.avg:.2e}"

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class dotdict(dict):
    """dict with attribute access, the `args` container of main_bpp.py:54-87."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value
