from .mnf_conv import MNFConv2D as MNFConv2D
from .mnf_dense import MNFDense as MNFDense
