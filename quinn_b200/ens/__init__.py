from .learner import Learner
