"""Layer-by-layer comparison of the tensor-core tier against the fp32 tier (GPU vs GPU) and the
oracle (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from speech_recognition_b200 import Engine, synth, arch as A, TTA_8

arch = int(sys.argv[1]) if len(sys.argv) > 1 else 195
nclips = int(sys.argv[2]) if len(sys.argv) > 2 else 3
last = int(sys.argv[3]) if len(sys.argv) > 3 else 11
eng = Engine(device=0, max_rows=64, precision="fp32")
w = synth.synthetic_weights(arch)
eng.load_model(0, arch, w)
x = torch.from_numpy(synth.make_clips(nclips, seed=21)).cuda()
views = ((0, 1.0), (-1500, 1.2))
Ts = A.layer_lengths(arch)[1:]
Cs = [A.ARCHS[arch]["conv1"]] + [c for c, _ in A.ARCHS[arch]["blocks"]]
for layer in range(0, last + 1):
    eng.set_precision("fp32")
    ref = eng.debug_activation(x, layer, (Ts[layer], Cs[layer]), views=views)
    eng.set_precision("tc")
    got = eng.debug_activation(x, layer, (Ts[layer], Cs[layer]), views=views)
    torch.cuda.synchronize()
    d = (got - ref).abs()
    print(f"layer {layer:2d} T={Ts[layer]:3d} C={Cs[layer]:3d}  max|ref|={ref.abs().max():.3f} "
          f"max err={d.max():.4e} mean err={d.mean():.3e} frac>0.05={(d > 0.05).float().mean():.4f}", flush=True)
    if d.max() > 0.5:
        bad = (d > 0.05).nonzero()
        print("   first bad idx:", bad[:5].tolist(), " rows bad:", sorted(set((bad[:, 1]).tolist()))[:20],
              " chans bad:", sorted(set((bad[:, 2]).tolist()))[:20])
        print("   got", got[tuple(bad[0].tolist())].item(), "ref", ref[tuple(bad[0].tolist())].item())
eng.set_precision("fp32"); p32, a32 = eng.forward(x, views=TTA_8)
eng.set_precision("tc"); ptc, atc = eng.forward(x, views=TTA_8)
print("probs max abs diff", (p32 - ptc).abs().max().item(), "labels", a32.tolist(), atc.tolist())

# label agreement at scale (GPU fp32 tier == oracle to 1e-6, so it stands in for the oracle here)
eng2 = Engine(device=0, max_rows=2048, precision="fp32")
eng2.load_model(0, arch, w)
N = 8192
xb = torch.from_numpy(synth.make_clips(N, seed=77)).cuda()
for views, name in ((((0, 1.0),), "1 view"), (TTA_8, "8 views")):
    eng2.set_precision("fp32"); p32, a32 = eng2.forward(xb, views=views)
    eng2.set_precision("tc"); ptc, atc = eng2.forward(xb, views=views)
    e = (p32 - ptc).abs()
    top2 = p32.topk(2, dim=1).values
    print(f"{name}: N={N} label agreement {(a32 == atc).float().mean().item():.5f}  prob err max {e.max().item():.4f} "
          f"q99 {e.flatten().quantile(0.99).item():.5f} mean {e.mean().item():.2e}; disagreeing clips' margin max "
          f"{((top2[:,0]-top2[:,1])[a32 != atc]).max().item() if (a32 != atc).any() else 0:.4f}")
