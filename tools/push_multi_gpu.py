#!/usr/bin/env python
"""The push path across GPUs (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/push_multi_gpu.py

Every rank triangulates its own frame block and its kernel stores the packed outputs straight into rank 0's
memory (sharding.PeerGather); rank 0 compares what arrived with an NCCL gather of the same results, bit for
bit, over several steps with rotating buffers, and prints one JSON line (also gpurun_out/push_multi_gpu.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pose2sim_b200 import ops, sharding, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = ops.get_engine(local)
    eng.set_output_mode(os.environ.get("P2S_OUTPUT_MODE", "vector"))
    F = 20_000 + 1000 * rank                                     # ragged blocks
    wl = synth.make_triangulation_workload(8, F, 1, 26, seed=202 + rank, lik_thr=None, frame0=rank * 100_000)
    x, y, lik = (torch.from_numpy(wl[k]).to(dev) for k in ("x", "y", "lik"))
    U = x.shape[0]
    units = [26 * (20_000 + 1000 * r) for r in range(world)]
    ref = eng.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, 2)
    pack = torch.empty(sharding.PACK_BYTES * U, dtype=torch.uint8, device=dev)
    v = sharding.packed_views(pack, U)
    for k in v:
        v[k].copy_(ref[k])
    bufs, _ = sharding.gather_packed(pack, units, dst=0)
    pg = sharding.PeerGather(eng, units, dst=0, n_buffers=2)
    side = torch.cuda.Stream(device=dev)
    bad = 0
    for step in range(6):
        eng.triangulate_planes_push(x, y, lik, wl["P"], 0.3, 15.0, 2, **pg.push_args(step))
        if rank == 0:
            with torch.cuda.stream(side):
                pg.collect(step, stream=side.cuda_stream)
            side.synchronize()
            got = pg.views(step % 2)
            for r in range(world):
                want = sharding.packed_views(bufs[r], units[r])
                for k in ("Q", "err", "mask", "nexcl"):
                    a, b = got[r][k].cpu().numpy(), want[k].cpu().numpy()
                    same = np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b)
                    bad += int(not same)
    torch.cuda.synchronize()
    err = torch.tensor([eng.peer_error()], device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    dist.barrier()
    pg.close()
    if rank == 0:
        line = {"tool": "push_multi_gpu", "world": world, "steps": 6, "units_per_rank": units, "mismatching_planes": bad,
                "peer_error_bits": int(err.item()), "output_mode": os.environ.get("P2S_OUTPUT_MODE", "vector"), "ok": bad == 0 and int(err.item()) == 0}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(line, open(os.path.join(ROOT, "gpurun_out", "push_multi_gpu.json"), "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
