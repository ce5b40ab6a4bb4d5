"""Quick per-stage timing on one GPU (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from speech_recognition_b200 import Engine, synth, TTA_8

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
MR = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
eng = Engine(device=0, max_rows=MR, precision=prec)
bank, offs = synth.make_noise_bank(seconds=60)
eng.set_noise_bank(torch.from_numpy(bank).cuda(), offs)
eng.frontend_config(480, 160, 40, 40)
eng.load_model(0, 195, synth.synthetic_weights(195))
pool = torch.from_numpy(synth.make_clips(256, seed=1)).cuda()
x = pool.repeat((B + 255) // 256, 1)[:B].contiguous()
p = synth.make_params(B, offs, seed=2)
pt = {k: torch.from_numpy(v).cuda() for k, v in p.items()}


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


aug = torch.empty_like(x)
t = timeit(lambda: eng.augment(x, pt["time_shift"], pt["bg_index"], pt["bg_offset"], pt["bg_volume"], pt["fg_volume"], out_t=aug))
print(f"augment   B={B}: {t:.3f} ms  {B/t*1e3:.0f} clips/s  {B*192020/t/1e6:.1f} GB/s")
for kind in ("spec", "logmel", "mfcc"):
    t = timeit(lambda: eng.features(aug, kind))
    print(f"features[{kind}] {prec}: {t:.3f} ms  {B/t*1e3:.0f} clips/s  {B*50.7e6/t/1e9:.1f} TFLOP/s-equiv")
for views, name in ((TTA_8, "8 views"),):
    t = timeit(lambda: eng.forward(aug, views=views), n=3, warm=1)
    nv = len(views)
    print(f"forward[{name}] {prec} max_rows={MR}: {t:.3f} ms  {B/t*1e3:.0f} clips/s  {B*nv*112.48e6/t/1e9:.1f} TFLOP/s")
