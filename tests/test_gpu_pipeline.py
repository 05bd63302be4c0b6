"""slab.GroupPipeline with its peer mailboxes (csrc/hp_peer.cu) between two processes that share ONE GPU: CUDA IPC mapping
of the mailboxes, rows stored into the neighbour's staging area, stream-level waits for the sequence numbers, the
application-finished signal of rank 0, solver contexts and per-group threads/streams.  The collectives run over gloo here
(NCCL refuses two ranks on one device), point-to-point messages are staged through the host by a test-only subclass.
Runs on the B200 box: pytest -m gpu."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, b, ret):
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(300, exit=True, file=sys.stderr)     # a hang ends with the stacks of all threads
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")   # no context-wide synchronisation at the first launch of a kernel
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        dev = torch.device("cuda:0")
        import helmholtz_preconditioner_b200 as hp
        from helmholtz_preconditioner_b200.slab import SlabSolver, GroupPipeline, slab_bounds
        from helmholtz_preconditioner_b200.gmres import DeviceVectors, gmres_batch

        class HostStaged(SlabSolver):
            def _exchange(self, ops):                        # gloo has no point-to-point for device tensors
                if not ops:
                    return
                torch.cuda.current_stream().synchronize()
                cpu, back = [], []
                for q in ops:
                    t = q.tensor.cpu() if q.op is dist.isend else torch.empty(q.tensor.shape, dtype=q.tensor.dtype)
                    cpu.append(dist.P2POp(q.op, t, q.peer, q.group))
                    if q.op is dist.irecv:
                        back.append((q.tensor, t))
                for w in dist.batch_isend_irecv(cpu):
                    w.wait()
                for dst, t in back:
                    dst.copy_(t)

        class HostReduced(DeviceVectors):                    # dot products summed over gloo on host copies
            def _all_reduce(self, t):
                c = t.cpu()
                dist.all_reduce(torch.view_as_real(c), group=self.group)
                t.copy_(c)

        omega = 2 * np.pi * (n / 10) + 2j
        c_mat, f_mat = hp.init_layered_f1(omega, n)
        R = slab_bounds(n, b, world)
        s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat, device=dev)
        s.setup_preconditioner(0, 0, max(b + 1, R[rank] + 1), min(n, R[rank + 1]))
        S = HostStaged(s, n, b, rank, world, None, device=dev)
        sizes = [8, 8, 3]
        fs = [np.roll(f_mat, 17 * i, axis=1).astype(np.complex128) for i in range(sum(sizes))]
        loc = [torch.from_numpy(np.ascontiguousarray(f[S.j0:S.j1].ravel())).to(dev) for f in fs]
        groups, i = [], 0
        for g in sizes:
            groups.append(loc[i:i + g]); i += g
        kw = dict(rtol=1e-3, restart=20, maxiter=6)
        print(f"[rank {rank}] strips factored", file=sys.stderr, flush=True)
        pipe = GroupPipeline(S, len(sizes), backend="gloo", solver_cls=HostStaged)
        assert len(pipe.mails) == len(sizes)
        print(f"[rank {rank}] pipeline built", file=sys.stderr, flush=True)
        res = pipe.gmres(groups, lambda nloc, pg: HostReduced(nloc, dev, group=pg), diag="paper", nglobal=n * n, **kw)
        torch.cuda.synchronize()
        print(f"[rank {rank}] groups solved", file=sys.stderr, flush=True)
        st = pipe.sweep_status()
        pipe.close()
        s.close()
        # the same systems on one solver that holds the whole grid, in lock step
        s1 = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat, device=dev).setup_preconditioner()
        full = [torch.from_numpy(f.ravel()).to(dev) for f in fs]
        ref = gmres_batch(lambda x, o: s1.matvec(x, o), lambda reqs: s1.precond_apply_batch(reqs, diag="paper"), full,
                          vec=DeviceVectors(n * n, dev), **kw)
        errs, iters = [], True
        for (u, info, hist), (u0, info0, hist0) in zip([x for grp in res for x in grp], ref):
            u0l = u0.reshape(n, n)[S.j0:S.j1].reshape(-1)
            errs.append((torch.linalg.norm(u - u0l) / torch.linalg.norm(u0l)).item())
            iters = iters and info == info0 and len(hist) == len(hist0) and np.allclose(hist, hist0, rtol=1e-8)
        s1.close()
        ret[rank] = dict(err=max(errs), iters=iters, status=st)
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()
    except Exception as e:                                   # noqa: BLE001 - reported to the parent
        import traceback
        ret[rank] = dict(error=traceback.format_exc())
        raise e


def test_group_pipeline_two_ranks_share_a_gpu():
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    n, b, world = 512, 12, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, b, ret), nprocs=world, join=True)
    for r in range(world):
        assert "error" not in ret[r], ret[r].get("error")
        assert ret[r]["status"] == 0 and ret[r]["iters"] and ret[r]["err"] < 1e-9, dict(ret[r])
