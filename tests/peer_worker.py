"""Worker of tests/test_peer_exchange.py (launched by torch.distributed.run, one process per GPU): a sharded render whose per-round exchange
runs over peer memory (cgrt_peer_*), checked on rank 0 against the same render on one GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from cgraytracing_b200 import Context, RenderConfig, preset
    from cgraytracing_b200.distributed import GpuEngine, ShardedRenderer, make_native_comm

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    accum, how = int(sys.argv[1]), sys.argv[2]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, ROUNDS, PER = 160, 120, 4, 50001
    cfg = RenderConfig(width=W, height=H, into_rule=1, update_mode=1)
    scene = preset("c2_bunny_chess")
    comm = make_native_comm(local, rank, world)
    with Context(local) as g:
        g.set_config(cfg, accum_mode=accum)
        scene.build_into(g); g.commit()
        R = ShardedRenderer(GpuEngine(g, local, comm, world, rank=rank, peer=(how == "peer")), rank, world)
        img = R.render(H, ROUNDS, PER)
        hp = g.download_hitpoints()
    # every rank holds the same replica
    t = torch.tensor([float(hp["n"].sum()), float(hp["flux"].sum()), float(img.sum())], device=f"cuda:{local}", dtype=torch.float64)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), (lo, hi)
    if rank == 0:
        with Context(local) as s:
            s.set_config(cfg, accum_mode=accum)
            scene.build_into(s); s.commit(); s.eye_pass(); s.build_grid()
            for r in range(ROUNDS):
                s.photon_pass(r * PER, PER); s.round_update()
            ref = s.download_hitpoints()
            rimg = s.gather_image(float(ROUNDS * PER))
        assert np.array_equal(hp["pos"], ref["pos"]) and np.array_equal(hp["key"], ref["key"])
        assert np.array_equal(hp["n"], ref["n"]) and hp["n"].sum() > 0          # accepted photons per hitpoint: equal
        assert np.array_equal(hp["r2"], ref["r2"])                              # hence the same radii, bit for bit
        tol = 1e-9 if accum == 0 else 1e-4
        assert np.allclose(hp["flux"], ref["flux"], rtol=tol, atol=1e-12)
        assert np.allclose(img, rimg, rtol=tol, atol=1e-12)
        print("PEER_OK", how, accum, int(hp["n"].sum()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
