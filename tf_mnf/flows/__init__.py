from tf_mnf_b200.flows import IAF as IAF
from tf_mnf_b200.flows import MADE as MADE
from tf_mnf_b200.flows import MAF as MAF
from tf_mnf_b200.flows import RNVP as RNVP
from tf_mnf_b200.flows import MaskedDense as MaskedDense
from tf_mnf_b200.flows import NormalizingFlow as NormalizingFlow
from tf_mnf_b200.flows import NormalizingFlowModel as NormalizingFlowModel
from tf_mnf_b200.flows import PlanarFlow as PlanarFlow
