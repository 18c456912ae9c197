import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200 import models
from baton_b200.ops import nn as bnn
from baton_b200.parallel.arena import ParamArena
BF16 = torch.bfloat16
dev = torch.device("cuda:0")
x = torch.randn(32, 32, 32, 3, device=dev).to(BF16)
y = torch.randint(0, 10, (32,), device=dev)
def rnd(m):
    with torch.no_grad():
        for mod in m.modules():
            if hasattr(mod, "running_mean") and hasattr(mod, "weight"):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
def run(mode):
    torch.manual_seed(3)
    m = models.resnet18(10); rnd(m)
    arena = ParamArena(m, dev); m.build_workspace(dev); m.train()
    outs = []
    if mode == "explicit":
        orig = m._block_fwd
        def wrap(blk, h, tape):
            o = orig(blk, h, tape); outs.append(o.float().clone()); return o
        m._block_fwd = wrap
        st = m.explicit_step(x, y)
    else:
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                blk.register_forward_hook(lambda mod, i, o: outs.append(o.detach().float().clone()))
        logits = m(x); loss, st = bnn.cross_entropy(logits, y); loss.backward(); bnn.WGRAD.join()
    torch.cuda.synchronize()
    return st.clone(), arena.grad.clone(), outs
for overlap in (True, False):
    bnn.BRANCH.enabled = overlap
    a = run("autograd"); b = run("autograd"); c = run("explicit")
    print("overlap", overlap, "loss a/a/e", a[0].tolist(), b[0].tolist(), c[0].tolist())
    for i in range(len(a[2])):
        print("   block", i, "a-b", float((a[2][i]-b[2][i]).abs().max()), "a-c", float((a[2][i]-c[2][i]).abs().max()), "scale", float(a[2][i].abs().max()))
    cos = torch.nn.functional.cosine_similarity
    print("    grad cos a,b", float(cos(a[1], b[1], dim=0)), "a,c", float(cos(a[1], c[1], dim=0)))
