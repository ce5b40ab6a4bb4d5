"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun), NCCL.  Every rank holds the whole job
(N clips, N not divisible by the world size), ShardedPredictor runs the rank's shard through the CUDA path and
assembles the job with ONE padded all-gather; rank 0 also runs the whole job alone and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_b200 import Engine, synth, TTA_8  # noqa: E402
from speech_recognition_b200 import sharded  # noqa: E402
from speech_recognition_b200.classes import class_map_32_to_12  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
    eng = Engine(device=local, max_rows=2048, precision="tc")
    try:
        eng.load_model(0, 195, synth.trained_weights(195))
        eng.load_model(1, 106, synth.trained_weights(106))
        clips = synth.make_word_clips(n, 12, seed=31337).numpy()          # the same job on every rank
        probs, amax = sharded.ShardedPredictor(eng, views=TTA_8).predict(clips)
        assert probs.shape == (n, 12) and amax.shape == (n,)
        # config 4 on shards: only uint8 probabilities / labels / keep flags are gathered
        s, e = sharded.shard_range(n, world, rank)
        p32, _ = eng.forward(torch.from_numpy(clips[s:e]).cuda(), views=TTA_8, slot=1)
        u8, label, keep = sharded.sharded_pseudo_labels(eng, p32, n, 0.6)
        ok = True
        if rank == 0:
            ref_p, ref_a = eng.predict_host(clips, views=TTA_8)
            ok = np.array_equal(probs, ref_p) and np.array_equal(amax, ref_a)
            rp32, _ = eng.forward(torch.from_numpy(clips).cuda(), views=TTA_8, slot=1)
            _, ru8 = eng.convert_classes(rp32, class_map_32_to_12("heng"), 12)
            rl, rk = eng.select(ru8, 0.6)
            ok = ok and torch.equal(u8, ru8) and torch.equal(label, rl) and torch.equal(keep, rk)
            print(f"sharded NCCL world={world} n={n}: shards "
                  f"{[sharded.shard_range(n, world, r) for r in range(world)]} match single-rank result: {ok}")
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1
    finally:
        eng.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
