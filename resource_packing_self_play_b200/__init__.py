"""B200-native self-play hot path for the resource-packing (bin-packing) game.

Drop-in for the reference's Game / MCTS / NeuralNet / Coach API (Wang-Xiaoyang/resource_packing_self_play,
xw_mcts/), with the search, environment and leaf evaluation running in hand-written sm_100a CUDA kernels behind the
C ABI of include/bpp_b200.h.  There is no CPU fallback.
"""
from .utils import AverageMeter, dotdict  # noqa: F401

__all__ = ["AverageMeter", "dotdict"]
__version__ = "0.1.0"
