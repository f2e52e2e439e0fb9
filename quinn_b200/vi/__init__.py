from .bnet import BNet
