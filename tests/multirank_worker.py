"""Worker of tests/test_gpu_multirank.py: one process per GPU under torchrun (NCCL).  Checks, on real ranks and real
collectives (libfesr's own communicator, fesr_allgatherv_pred / fesr_allreduce_grads):
  * sharded scheduler.predict + reconstruct_from_partition == the single-rank result, bit for bit, for device-resident
    and host inputs; the ranks' field slices tile the single-rank field;
  * MeshPredictor.step (the resident path bench.py times): slice and full field bit-identical to one rank;
  * FlatAdam: parameters broadcast from rank 0 at construction, bit-identical on every rank after k steps on
    different per-rank data.
Exit code 0 = all checks passed on this rank."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fesr_b200 import comm
    from fesr_b200.dataset.GraphDataset import AnsysDataset
    from fesr_b200.models import scheduler_gnn as sg
    from fesr_b200.models.model import KernelNN
    from fesr_b200.models.training import FlatAdam, train_step
    from fesr_b200.pipeline import MeshPredictor, node_slice

    z = np.load(os.path.join(ROOT, "tests", "golden", "shipped_w43_weights.npz"))
    sd = {k[len("neuralop::"):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("neuralop::")}
    os.chdir(tempfile.mkdtemp(prefix=f"fesr_rank{rank}_"))
    os.makedirs("logs/models/collection_t", exist_ok=True)
    torch.save(sd, "logs/models/collection_t/partition_0.pth")
    ds = AnsysDataset(mesh_n=12, num_meshes=1, sub_size=16)
    sched = sg.GNNPartitionScheduler("t", 1, ds, KernelNN(43, 43, 5, in_width=4, out_width=4), train=False)
    base = ds.get_one_full_sample(0)
    c = ds._mesh(0)
    N = c["mesh"].num_nodes

    for prec in ("fp32", "f16"):
        sched.models[0].precision = prec
        # single-rank reference inside this process
        real = sg._dist
        sg._dist = lambda: (None, 0, 1)
        p1, r1, mi1, w1 = sched.predict(base)
        out1 = ds.reconstruct_from_partition(p1, r1, 0, mi1, w1)
        full_p, full_w, full_field = torch.cat(list(p1)).clone(), torch.stack([t[0] for t in w1]).clone(), out1.field.clone()
        sg._dist = real
        xh, yh = c["x"].cpu().pin_memory(), c["y"].cpu().pin_memory()
        for sample in (base, base.with_host_inputs(xh, yh)):
            p, r, mi, w = sched.predict(sample)
            assert comm.ready() and comm.world() == world and comm.rank() == rank
            out = ds.reconstruct_from_partition(p, r, 0, mi, w)
            lo, hi = out.node_range
            assert (lo, hi) == node_slice(N, rank, world)
            assert torch.equal(out.field_local, full_field[lo:hi]), "field slice differs from the single-rank field"
            assert torch.equal(out.field, full_field)
            assert float((out.ref_field - torch.from_numpy(c["mesh"].y)).abs().max()) < 1e-6
            assert torch.equal(torch.cat(list(p)), full_p), "gathered predictions differ from the single-rank ones"
            assert torch.equal(torch.stack([t[0] for t in w]), full_w)

        # resident step (bench.py's `value` path)
        model = sched.models[0]
        pos, cells = torch.from_numpy(c["mesh"].pos).to(dev), torch.from_numpy(c["mesh"].cells).to(dev)
        mp1 = MeshPredictor(model, pos, cells, ds.levels, rank=0, world=1)
        x_all = torch.from_numpy(c["mesh"].x).to(dev)
        y_all = torch.from_numpy(c["mesh"].y).to(dev)
        f1, w_1, _ = mp1.step(x_all[mp1.shard.global_ids], y_all[mp1.shard.global_ids])
        mpn = MeshPredictor(model, pos, cells, ds.levels, rank=rank, world=world)
        g = mpn.shard.global_ids
        fs, ws, _ = mpn.step(x_all[g], y_all[g])
        lo, hi = node_slice(N, rank, world)
        assert torch.equal(fs, f1[lo:hi])
        ff, _, _ = mpn.step(x_all[g], y_all[g], full_field=True)
        assert torch.equal(ff, f1)
        assert torch.equal(ws, w_1[mpn.bounds[rank]:mpn.bounds[rank + 1]])

    # FlatAdam: different initial weights per rank -> broadcast; different data per rank -> identical after k steps
    torch.manual_seed(100 + rank)
    m = KernelNN(16, 16, 3, in_width=4, out_width=4).to(dev).train()
    opt = FlatAdam(m, lr=0.01)
    b = c["batch"]
    gen = torch.Generator(device=dev).manual_seed(rank)
    for _ in range(3):
        x = c["x"] + 0.01 * torch.randn(c["x"].shape, device=dev, generator=gen)
        train_step(m, opt, x, b.csr, b.edge_attr, c["y"])
    flats = [torch.empty_like(opt.flat) for _ in range(world)]
    dist.all_gather(flats, opt.flat)
    for f in flats[1:]:
        assert torch.equal(f, flats[0]), "ranks diverged"
    assert bool(torch.isfinite(opt.flat).all())
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()
    print(f"rank {rank}: multirank checks OK")


if __name__ == "__main__":
    main()
