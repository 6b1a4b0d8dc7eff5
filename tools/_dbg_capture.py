import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cellbender_b200 import run, synth, ops
from cellbender_b200.posterior import Posterior
dev = "cuda:0"
ds = synth.make_config_dataset("c1", seed=0, device=dev)
model = run.build_model(ds, run.default_args(), device=dev)
model.eval()
model.encoder["other"].offset = {"logit_p": 0.0, "epsilon": 0.0}
post = Posterior(ds, model)
rows = torch.arange(512, device=dev)
G = model.n_genes
lin = model.decoder.out_linear
S = 20
logits_t = torch.empty((G, ops.row_pitch(S * 512)), device=dev)
state = {}
def s1():
    _, enc = post._encode(rows); state["enc"] = enc
def s2():
    state["smp"] = post.sample_rate_coefficients(state["enc"], S)
def s3():
    z = state["smp"][0]; state["h"] = model.decoder.hidden(z.reshape(512 * S, -1))
def s4():
    ops.gemm_tf32(lin.weight, state["h"].contiguous(), G, 512 * S, lin.weight.shape[1], out=logits_t[:, :512 * S], split_k=1)
    state["lse"] = ops.col_logsumexp(logits_t, G, 512 * S, row_add=lin.bias)
for name, fn in (("encode", s1), ("sample", s2), ("decoder_hidden", s3), ("gemm+lse", s4)):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "capture OK", flush=True)
    except Exception as e:
        print(name, "capture FAILED:", str(e).split("\n")[0], flush=True)
        break
