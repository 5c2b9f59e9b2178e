"""Worker of tests/test_gpu_fused.py::test_two_rank_tile_gather (run under torch.distributed.run, one rank per GPU):
each rank renders its ray shard and the tiles are all-gathered through libpnr's NCCL entry points; every rank must
end with exactly the single-GPU frame (fp32 maps bit for bit, label tiles = argmax of the single-GPU maps)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import panopticnerf_b200 as PN                                     # noqa: E402
from panopticnerf_b200 import parallel, synthetic as S             # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    cfg = PN.make_cfg("cfg1", num_classes=5, num_instances=6, N_importance=16)
    net = S.init_network_weights(PN.make_network(cfg), seed=2).to(dev)
    ren = PN.make_renderer(cfg, net)
    batch = {k: v.to(dev) for k, v in S.make_batch(cfg).items()}
    R = batch["rays"].shape[0] - 5                                  # ragged: not a multiple of the world size
    batch["rays"] = batch["rays"][:R].contiguous()
    whole = ren.render(batch)
    local = ren.render(parallel.shard_batch(batch, rank, world))
    tg = parallel.TileGather(dev)
    keys = ("rgb_map", "depth_map", "acc_map", "semantic_map", "instance_map")
    got = tg.gather_maps(local, R, keys)
    for k in keys:
        assert torch.equal(got[k], whole[k]), f"rank {rank}: {k} differs from the single-GPU frame"
    ref = parallel.all_gather_maps(local, R, keys)                  # the torch.distributed path gives the same
    for k in keys:
        assert torch.equal(got[k], ref[k]), k
    lab = tg.gather_labels(local, R)
    assert torch.equal(lab["sem_label"].long(), whole["semantic_map"].argmax(-1))
    assert torch.equal(lab["inst_label"].long(), whole["instance_map"].argmax(-1))
    assert torch.equal(lab["depth"], whole["depth_map"])
    assert torch.equal(lab["rgb8"].float(), torch.round(whole["rgb_map"].clamp(0, 1) * 255))
    tg.close()
    dist.barrier()
    if rank == 0:
        print(f"COMM2 OK world={world} bytes_per_rank={lab['bytes_per_rank']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
