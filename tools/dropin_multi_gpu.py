#!/usr/bin/env python
"""Run the drop-ins `triangulate_all` and `associate_all` under torchrun (one process per GPU, NCCL) on the golden
trials and compare the files they write with what the reference wrote.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dropin_multi_gpu.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from dropin_util import assert_trc_equal, associated_people, golden_trcs, rebuild_trial, written_trcs
    import pose2sim_b200
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    for tag in ("e2e_tri_single", "e2e_tri_multi", "e2e_tri_undistort"):
        g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"), allow_pickle=False)
        base = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(base, src=0)
        if rank == 0:
            rebuild_trial(g, base[0], "trial_demo")
        dist.barrier()
        proj = os.path.join(base[0], "trial_demo")
        from pose2sim_b200 import synth_project
        import json
        extra = json.loads(str(g["extra"])) if "extra" in g.files else {}
        cfg = synth_project.base_config(proj, multi_person=bool(g["multi_person"]), **extra)
        os.chdir(proj)
        pose2sim_b200.triangulate_all(cfg)
        dist.barrier()
        if rank == 0:
            got, ref = written_trcs(proj), golden_trcs(g)
            assert sorted(got) == sorted(ref), (sorted(got), sorted(ref))
            worst = max(assert_trc_equal(got[n], ref[n], tol=1e-6) for n in ref)
            print(f"{tag}: {world} ranks, TRC equal to the reference's, max |d| = {worst:.2e} m", flush=True)
    for tag in ("e2e_assoc_single", "e2e_assoc_multi"):
        g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"), allow_pickle=False)
        base = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(base, src=0)
        cfg = None
        if rank == 0:
            _, cfg = rebuild_trial(g, base[0], "trial_assoc")
        box = [cfg]
        dist.broadcast_object_list(box, src=0)
        cfg = box[0]
        proj = os.path.join(base[0], "trial_assoc")
        dist.barrier()
        os.chdir(proj)
        pose2sim_b200.associate_all(cfg)
        dist.barrier()
        if rank == 0:
            if tag.endswith("multi"):
                from dropin_util import assert_multi_person_json_equal
                assert_multi_person_json_equal(proj, g)
            else:
                chosen, exists = associated_people(proj, [str(c) for c in g["cams"]], g["kp"].shape[0], g["chosen"].shape[2])
                assert np.array_equal(exists, g["exists"]) and np.array_equal(np.isnan(chosen), np.isnan(g["chosen"]))
                assert np.array_equal(np.nan_to_num(chosen).astype(np.float32), np.nan_to_num(g["chosen"]))
            print(f"{tag}: {world} ranks, pose-associated/ equal to the reference's", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
