"""Throughput of every BASELINE.json config on one B200 (device-resident synthetic inputs, CUDA events).
    python scripts/bench_configs.py > profiles/configs_r1.json
C1 MNIST ODE block fwd B=128 | C2 premetanode10 inference B=512 | C3 solver / model ensembling over 4 RK2 (+RK4)
C4 FGSM-random adversarial training step with solver smoothing B=256 | C5 PGD-7 robust-accuracy evaluation B=512."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import numpy as np, torch, torch.nn.functional as F
import metasolver_b200 as msb
from metasolver_b200.sopa.src.solvers.utils import create_solver, sample_solver_by_noising_params
from metasolver_b200.sopa.src.models.odenet_cifar10.layers import premetanode10
from metasolver_b200.sopa.src.models.odenet_cifar10.utils import Identity
from metasolver_b200.sopa.src.models.odenet_mnist.layers import MetaODEBlock as MnistBlock
from metasolver_b200.MegaAdversarial.src.attacks import FGSMRandom, PGD, FGSM2Ensemble, ensemble_logits

MEAN, STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
dev = torch.device("cuda")
torch.manual_seed(0); np.random.seed(0)


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def solver(m="rk2", p="u", n=8, u=0.5, v=-1):
    s = create_solver(m, p, n, -1, u, v, torch.float32, dev); s.freeze_params(); return s


out = {}
# ---- C1
blk = MnistBlock().to(dev)
x = torch.rand(128, 64, 6, 6, device=dev).contiguous(memory_format=torch.channels_last)
s05 = solver()
opts = Namespace(solver_mode="standalone")
with torch.no_grad():
    ms = timed(lambda: blk(x, [s05], opts), 20)
    gs = msb.GraphedStep(lambda xx: blk(xx, [s05], opts), (x,))
    ms_g = timed(lambda: gs(x), 20)
msb.set_option("mnist_fused", 0)
with torch.no_grad():
    ms_unfused = timed(lambda: blk(x, [s05], opts), 20)
msb.set_option("mnist_fused", 1)
out["C1_mnist_odeblock_fwd_B128"] = dict(ms=ms, ms_cuda_graph=ms_g, images_per_s=128 / ms_g * 1e3, ms_multi_launch_simt_path=ms_unfused,
                                         note="one persistent tcgen05 launch for the whole solve (mnist_fused.cu); reference: 46 ms on 8 CPU cores (SURVEY 8a9)")
# ---- C2..C5 model
model = premetanode10((Identity,) * 3, (lambda t: t,) * 3, (F.gelu,) * 3, in_planes=64, is_odenet=True).to(dev)
model = model.to(memory_format=torch.channels_last).eval()
B = 512
img = torch.rand(B, 3, 32, 32, device=dev)
xin = ((img - torch.tensor(MEAN, device=dev).view(1, 3, 1, 1)) / torch.tensor(STD, device=dev).view(1, 3, 1, 1)).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (B,), device=dev)
kw = {"solvers": [s05], "solver_options": opts}
with torch.no_grad():
    ms = timed(lambda: model(xin, **kw))
    g2 = msb.GraphedStep(lambda xx: model(xx, **kw), (xin,))
    ms_g = timed(lambda: g2(xin))
out["C2_premetanode10_inference_B512"] = dict(ms=ms, ms_cuda_graph=ms_g, images_per_s=B / ms_g * 1e3)
# ---- C3
us = [0.3, 0.5, 2 / 3., 1.0]
rk2s = [solver(u=u) for u in us]
ens = Namespace(solver_mode="ensemble", ensemble_prob=1.0, ensemble_weights=None)
xb = xin[:256]
with torch.no_grad():
    ms_st = timed(lambda: model(xb, rk2s, ens))
    ms_seq = timed(lambda: [model(xb, [s], opts) for s in rk2s])
    kwa = [{"solvers": [s], "solver_options": opts} for s in rk2s]
    ms_me = timed(lambda: ensemble_logits([model] * 4, xb, kwa))
    rk4s = [solver("rk4", "u2", 8, 1 / 3.), solver("rk4", "uv", 8, 1 / 3., 2 / 3.)]
    ms_rk4 = timed(lambda: model(xb, rk4s, ens))
out["C3_ensembles_B256"] = dict(solver_ensemble_4xRK2_stacked_ms=ms_st, four_standalone_passes_ms=ms_seq,
                                model_ensemble_4xRK2_stacked_ms=ms_me, solver_ensemble_2xRK4_stacked_ms=ms_rk4,
                                images_per_s_solver_ensemble=256 / ms_st * 1e3, images_per_s_model_ensemble=256 / ms_me * 1e3)
# ---- C4: FGSM-random training step, solver smoothing (u ~ N(0.5, 0.0125) redrawn per batch), SGD momentum
model.train()
opt = msb.FusedSGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4)     # flat buffers, one update kernel
atk = FGSMRandom(model, alpha=10 / 255., epsilon=8 / 255., mu=MEAN, std=STD)
base = create_solver("rk2", "u", 8, -1, 0.5, -1, torch.float32, dev); base.freeze_params()
xt, yt = xin[:256], y[:256]


def train_step():
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):      # the sampler prints its draw (as the reference does)
        s = sample_solver_by_noising_params(base, std=0.0125, bernoulli_p=1.0, noise_type="normal")
    kws = {"solvers": [s], "solver_options": opts}
    opt.zero_grad()
    xa, _ = atk(xt, yt, kws)
    loss = F.cross_entropy(model(xa, **kws), yt)
    loss.backward()
    opt.step(grad_scale=opt.all_reduce())


ms = timed(train_step, 5, 2)
out["C4_fgsm_random_train_step_B256"] = dict(ms=ms, images_per_s=256 / ms * 1e3,
                                             note="attack pass (fwd+bwd) + training pass (fwd+bwd) + fused SGD; fused attack steps; u redrawn per batch")
# ---- C5: PGD-7 evaluation
model.eval()
pgd = PGD(model, eps=8 / 255., lr=2 / 255., n_iter=7, mean=MEAN, std=STD)


def pgd_eval():
    xa, _ = pgd(xin, y, kw)
    with torch.no_grad():
        return (model(xa, **kw).argmax(1) == y).sum()


ms = timed(pgd_eval, 3, 1)
try:
    def pgd_eval_xy(xx, yy):
        xa, _ = pgd(xx, yy, kw)
        with torch.no_grad():
            return (model(xa, **kw).argmax(1) == yy).sum()
    g5 = msb.GraphedStep(pgd_eval_xy, (xin, y), warmup=1)
    ms_g = timed(lambda: g5(xin, y), 3, 1)
except Exception as exc:
    ms_g = None
    print("C5 graph capture failed:", exc, file=sys.stderr)
out["C5_pgd7_eval_B512"] = dict(ms=ms, ms_cuda_graph=ms_g, images_per_s=B / (ms_g or ms) * 1e3,
                                note="7 x (fwd + input-gradient bwd) + 1 fwd per batch")
out["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(out, indent=1))
