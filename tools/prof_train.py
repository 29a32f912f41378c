"""Host-side cProfile of Trainer.step for one of the training configs (diagnostic)."""
import cProfile, pstats, sys, os, io
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M
from tf_mnf_b200.models import MNFLeNet, MNFVGG, MNFFeedForward
from tf_mnf_b200.train import Trainer
cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
M.set_seed(0)
ctx = M.default_context(0)
ctx.set_math("tf32x3")
rng = np.random.RandomState(0)
if cfg == "c1":
    model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0, flow_type="rnvp")
    x = rng.uniform(size=(128, 28, 28, 1)).astype(np.float32); y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=128)]
    tr = Trainer(model, loss="xent", kl_weight=1 / 60000)
elif cfg == "c2":
    model = MNFFeedForward(layer_sizes=(100, 50, 10), std_init=1e-2)
    x = rng.standard_normal((1024, 136)).astype(np.float32); y = rng.standard_normal((1024, 10)).astype(np.float32)
    tr = Trainer(model, loss="mse", kl_weight=1 / 128)
else:
    model = MNFVGG()
    x = rng.uniform(size=(256, 32, 32, 3)).astype(np.float32); y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=256)]
    tr = Trainer(model, loss="xent", kl_weight=1 / 50000)
xd, yd = ctx.to_device(x), ctx.to_device(y)
for _ in range(3):
    tr.step(xd, yd)
ctx.sync()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    tr.step(xd, yd)
ctx.sync()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
