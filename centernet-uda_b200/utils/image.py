"""``utils.image.entropy_map`` (reference utils/image.py:121-124): weighted self-information map
``-p*log2(p+1e-30)/log2(C)`` of the channel softmax, differentiable (the ADVENT generator loss flows
through it, uda/adversarial_entropy_minimization.py:91-110); fwd and bwd are sm_100a kernels
(csrc/softmax_stat.cu).  The remaining helpers of the reference module (gaussian rasteriser, FDA) are
re-exported from a reference checkout when one is on sys.path."""
from cnhead import functional as _F
from cnhead._dropin import reexport as _reexport


def entropy_map(hm):
    return _F.entropy_map(hm)


_reexport(__name__, __file__, globals())
