"""quinn_b200 -- B200-native implementation of QUiNN's posterior-sampling hot path.

Same Python surface as the reference for that path (quinn.solvers.NN_MCMC / NN_Ens / NN_VI,
quinn.mcmc.AMCMC / HMC / MALA, quinn.nns.MLP / RNet / NNWrap / NegLogPost, predict_* calls); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/quinn_b200.h.
"""
__version__ = '0.1.0'
