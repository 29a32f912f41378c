from .mnf_feed_forward import MNFFeedForward as MNFFeedForward
from .mnf_lenet import MNFLeNet as MNFLeNet
from .stock import BatchNormalization, Flatten, MaxPool2D, ReLU, ReLUMaxPool2D, Sequential, Softmax  # noqa: F401
from .mnf_vgg import MNFVGG as MNFVGG
