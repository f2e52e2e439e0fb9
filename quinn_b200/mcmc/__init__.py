from .mcmc import MCMCBase, DeviceLogPost
from .admcmc import AMCMC
from .hmc import HMC
from .mala import MALA
