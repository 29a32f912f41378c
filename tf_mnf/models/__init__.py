from tf_mnf_b200.models import MNFFeedForward as MNFFeedForward
from tf_mnf_b200.models import MNFLeNet as MNFLeNet
