from tf_mnf_b200.layers import MNFConv2D as MNFConv2D
from tf_mnf_b200.layers import MNFDense as MNFDense
