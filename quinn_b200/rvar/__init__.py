from .rvs import RV, Gaussian_1d, GMM2_1d
