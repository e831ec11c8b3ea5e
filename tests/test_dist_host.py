"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — node partition, halo
plan (both exported by the C-ABI library, no device needed) and the communication schedule of
the distributed CG (r-halo pushed before the second allreduce, p-halo recomputed locally) — run
with numpy standing in for the kernels, against the oracle's single-process solve."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from magnetite_b200 import dist as mdist, meshgen
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = meshgen.jitter(meshgen.plate(nx, ny))
    meta = meshgen.EXAMPLE_MATERIAL
    om = O.Mesh(mesh)
    full = O.assemble_sparse(om, O.element_stiffness(om, meta))
    (rp, col, val), rhs, fmap = O.partition(om, full, dense=False)
    n = len(rhs)
    # reduced-row boundaries from the node partition (rows of a node range are contiguous)
    rowmap = np.concatenate([[0], np.cumsum(((mesh.known[:, None] >> np.array([2, 3])) & 1).ravel())])
    bounds = [mdist.partition_nodes(mesh.n_nodes, world, r)[0] for r in range(world)] + [mesh.n_nodes]
    row_lo = np.array([rowmap[2 * b] for b in bounds], np.uint32)
    lo, hi = int(row_lo[rank]), int(row_lo[rank + 1])
    my_cols = col[rp[lo]:rp[hi]]
    ext = torch.tensor([min(lo, int(my_cols.min())), max(hi, int(my_cols.max()) + 1)])
    exts = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(exts, ext)
    ext_lo = np.array([int(e[0]) for e in exts], np.uint32); ext_hi = np.array([int(e[1]) for e in exts], np.uint32)
    plan = mdist.halo_plan(world, rank, row_lo, ext_lo, ext_hi)
    recv_plan = [(a, b, r) for r in range(world) if r != rank for (a, b, d) in mdist.halo_plan(world, r, row_lo, ext_lo, ext_hi) if d == rank]

    def push(vec):          # what the peer stores of kernel B do, expressed as send/recv
        reqs = [dist.isend(torch.from_numpy(vec[a:b].copy()), dst=d) for a, b, d in plan]
        for a, b, src in recv_plan:
            t = torch.empty(b - a, dtype=torch.float64)
            dist.recv(t, src=src)
            vec[a:b] = t.numpy()
        for q in reqs:
            q.wait()

    def allreduce(*vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.tolist()

    diag = np.array([val[rp[i]:rp[i + 1]][col[rp[i]:rp[i + 1]] == i][0] for i in range(lo, hi)])
    r_ext = np.zeros(n); dinv = np.zeros(n); p = np.zeros(n); x = np.zeros(hi - lo)
    r_ext[lo:hi] = rhs[lo:hi]; dinv[lo:hi] = 1.0 / diag
    push(r_ext); push(dinv)
    e0, e1 = int(ext_lo[rank]), int(ext_hi[rank])
    rz, bb = allreduce(float(r_ext[lo:hi] @ (dinv[lo:hi] * r_ext[lo:hi])), float(r_ext[lo:hi] @ r_ext[lo:hi]))
    p[e0:e1] = r_ext[e0:e1] * dinv[e0:e1]
    it, rr = 0, bb
    while rr > 1e-24 * bb and it < 5000:
        q = np.array([val[rp[i]:rp[i + 1]] @ p[col[rp[i]:rp[i + 1]]] for i in range(lo, hi)])
        (pq,) = allreduce(float(p[lo:hi] @ q))
        alpha = rz / pq
        x += alpha * p[lo:hi]
        r_ext[lo:hi] -= alpha * q
        push(r_ext)                                   # halo of r rides on kernel B
        rz_new, rr = allreduce(float(r_ext[lo:hi] @ (dinv[lo:hi] * r_ext[lo:hi])), float(r_ext[lo:hi] @ r_ext[lo:hi]))
        p[e0:e1] = r_ext[e0:e1] * dinv[e0:e1] + (rz_new / rz) * p[e0:e1]   # owned rows AND halo rows
        rz = rz_new
        it += 1
    np.save(Path(out_dir) / f"x_{rank}.npy", np.concatenate([[lo, hi, it], x]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_partitioned_cg_matches_oracle(tmp_path, built):
    nx, ny, world = 14, 9, 2
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, str(tmp_path)), nprocs=world, join=True)
    from magnetite_b200 import meshgen
    from oracle import oracle as O
    mesh = meshgen.jitter(meshgen.plate(nx, ny))
    ref = O.run(O.Mesh(mesh), meshgen.EXAMPLE_MATERIAL, O.cg_options(), dense=False)
    u = np.stack([ref["ux"], ref["uy"]], 1).ravel()
    free = np.array([(mesh.known[d // 2] >> (d % 2)) & 1 == 0 for d in range(2 * mesh.n_nodes)])
    x_ref = u[free]
    parts = [np.load(tmp_path / f"x_{r}.npy") for r in range(world)]
    assert int(parts[0][0]) == 0 and int(parts[0][1]) == int(parts[1][0]) and int(parts[1][1]) == len(x_ref)
    x = np.concatenate([p[3:] for p in parts])
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-9
    assert parts[0][2] == parts[1][2] > 10


def test_halo_plan_properties(built):
    from magnetite_b200 import dist as mdist
    row_lo = np.array([0, 100, 250, 300, 420], np.uint32)
    ext_lo = np.array([0, 80, 90, 260, 299], np.uint32)      # rank 2 reaches back into rank 0
    ext_hi = np.array([130, 270, 310, 421, 420], np.uint32)
    R = 4
    plans = [mdist.halo_plan(R, r, row_lo, ext_lo, ext_hi) for r in range(R)]
    for me, plan in enumerate(plans):
        for lo, hi, dst in plan:
            assert dst != me and row_lo[me] <= lo < hi <= row_lo[me + 1]
            assert (ext_lo[dst] <= lo and hi <= row_lo[dst]) or (row_lo[dst + 1] <= lo and hi <= ext_hi[dst])
    # every halo index of every rank is supplied exactly once
    for dst in range(R):
        need = set(range(ext_lo[dst], row_lo[dst])) | set(range(row_lo[dst + 1], min(ext_hi[dst], row_lo[R])))
        got = [i for plan in plans for lo, hi, d in plan if d == dst for i in range(lo, hi)]
        assert sorted(got) == sorted(need)
    assert (90, 100, 2) in plans[0] and (100, 250, 2) in plans[1]
