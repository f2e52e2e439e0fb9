from .tchutils import tch, npy, print_nnparams
from .nnbase import MLPBase
from .mlp import MLP, Expon
from .rnet import RNet, Poly, Lin, Quad, Cubic, NonPar, LayerFcn
from .losses import NegLogPost, NegLogPrior
from .nnwrap import NNWrap, nn_p, nnwrapper
from .nnfit import nnfit
