"""World-size-2 gloo test of the data-parallel gradient path (CPU): the bucketed all-reduce must equal the mean
of the per-rank gradients, with several buckets in flight -- in both schedules (reduced after backward with the
gradients left as views into the buckets: the default; hook-driven and overlapped with backward)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for overlap in (False, True):
            _check_schedule(rank, world, overlap)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _check_schedule(rank, world, overlap):
    if True:
        from adaptersis_b200.dp import BucketedGradAllReduce
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 16), torch.nn.Linear(16, 4))
        unused = torch.nn.Linear(3, 3)              # never takes part in the step (like cls_token / pos_embed here)
        red = BucketedGradAllReduce(list(net.parameters()) + list(unused.parameters()), bucket_bytes=600, overlap=overlap)
        assert len(red.buckets) >= 3 and red.overlap == overlap
        for step in range(2):
            g = torch.Generator().manual_seed(100 + step)
            data = torch.randn(world, 5, 8, generator=g)
            net.zero_grad(set_to_none=True)
            net(data[rank]).square().sum().backward()
            red.finish()
            assert all(p.grad is None for p in unused.parameters())
            got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
            ref_net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 16), torch.nn.Linear(16, 4))
            ref_net.load_state_dict(net.state_dict())
            tot = 0
            for r in range(world):
                tot = tot + ref_net(data[r]).square().sum()
            (tot / world).backward()
            want = torch.cat([p.grad.reshape(-1) for p in ref_net.parameters()])
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (rank, step, float((got - want).abs().max()))
        # gradient accumulation: first micro-batch under no_sync(), second one reduces the accumulated gradients
        g = torch.Generator().manual_seed(300)
        data = torch.randn(2, world, 5, 8, generator=g)
        net.zero_grad(set_to_none=True)
        with red.no_sync():
            net(data[0, rank]).square().sum().backward()
        net(data[1, rank]).square().sum().backward()
        red.finish()
        got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        ref_net.load_state_dict(net.state_dict())
        ref_net.zero_grad(set_to_none=True)
        tot = sum(ref_net(data[mb, r]).square().sum() for mb in range(2) for r in range(world))
        (tot / world).backward()
        want = torch.cat([p.grad.reshape(-1) for p in ref_net.parameters()])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (rank, "accumulate", float((got - want).abs().max()))
        # a second backward before finish() without no_sync() must fail loudly, not drop gradients silently
        net.zero_grad(set_to_none=True)
        net(data[0, rank]).square().sum().backward()
        try:
            net(data[1, rank]).square().sum().backward()
            raise AssertionError("second backward before finish() did not raise")
        except RuntimeError as e:
            assert "no_sync" in str(e)
        red.finish()
        if not overlap:     # the default schedule leaves every gradient as a view into its reduced bucket (no unpack pass)
            assert all(p.grad._base is not None and p.grad._base.dim() == 1 for p in net.parameters())
        red.remove()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
