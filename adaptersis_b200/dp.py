"""Data-parallel plumbing: one process per GPU, parameters replicated, batch sharded, gradients
all-reduced in buckets over NCCL (NVLink 5 / NVSwitch) on a side stream while backward is still
running -- the only collective the path has (reference: four
DistributedDataParallel wrappers, train.py:84,96,112,116).

Buckets are formed in reverse parameter-registration order (the order backward produces
gradients).  Two schedules:

  * ``overlap=False`` (default): the buckets are packed and reduced back to back when backward has finished
    (`finish()`), and every ``p.grad`` is left as a VIEW into its reduced flat bucket (no unpack pass).  The step's hot
    kernels are persistent one-CTA-per-SM grids (GEMM, attention): an NCCL kernel that runs beside them takes SMs
    away, the displaced CTAs start only when the collective ends and the static tile schedule makes the whole kernel
    wait for them (measured at 8 GPUs: GEMMs 53.0 -> 59.6 ms, LayerNorm backward 6.4 -> 7.8 ms, step +10..12 ms),
    whereas the 1.31 GB of gradients take a few ms over NVLink 5 / NVSwitch when reduced alone.
  * ``overlap=True`` (``ASIS_DP_OVERLAP=1``): a bucket is reduced on a side stream as soon as all its gradients have
    been accumulated (post-accumulate-grad hooks), then unpacked -- the classic DDP schedule.

Works with any torch.distributed backend (gloo on CPU for tests)."""
import contextlib
import os

import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, params, bucket_bytes=None, process_group=None, average=True, overlap=None):
        self.params = [p for p in params if p.requires_grad]
        self.overlap = (os.environ.get("ASIS_DP_OVERLAP", "0") == "1") if overlap is None else bool(overlap)
        if bucket_bytes is None:
            bucket_bytes = (64 << 20) if self.overlap else (256 << 20)
        self.bucket_bytes = bucket_bytes
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nb = p.numel() * 4
            if cur and cur_bytes + nb > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [0] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._works = []
        self._comm_stream = None
        self._defer = False
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._fired = set()
        self._works = []

    def _stream(self, device):
        if device.type != "cuda":
            return None
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device)
        return self._comm_stream

    def _on_grad(self, p):
        if self._defer:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if id(p) in self._fired:
            raise RuntimeError(
                "BucketedGradAllReduce: a gradient was accumulated twice before finish() -- a second backward() "
                "(gradient accumulation) or a parameter used by two graphs.  Run the micro-batches that should not "
                "be reduced yet under `with reducer.no_sync():`; the last backward (outside it) + finish() reduces "
                "the accumulated gradients.")
        self._fired.add(id(p))
        if self._pending[i] == 0 and self.overlap:
            self._launch(i)

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward() calls inside this context only accumulate into p.grad (no bucket is
        launched); the first backward outside it reduces the accumulated gradients (DDP.no_sync semantics)."""
        old = self._defer
        self._defer = True
        try:
            yield
        finally:
            self._defer = old

    def _launch(self, i, partial=False):
        if self.world == 1:
            return
        bucket = self.buckets[i]
        if partial:
            # parameters that took no part in this step (e.g. cls_token / pos_embed / final norm, which
            # only the forward-only taps pass touches): reduce the rest of the bucket.  The set is the
            # same on every rank because the graph is.
            bucket = [p for p in bucket if p.grad is not None]
            if not bucket:
                return
        dev = bucket[0].device
        n = sum(p.numel() for p in bucket)
        if partial:
            flat = torch.empty(n, dtype=torch.float32, device=dev)
        else:
            if self._flat[i] is None:
                self._flat[i] = torch.empty(n, dtype=torch.float32, device=dev)
            flat = self._flat[i]
        side = self._stream(dev)
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(dev))
            ctx = torch.cuda.stream(side)
        else:
            ctx = contextlib.nullcontext()
        # NCCL averages inside the collective; other backends (gloo in the CPU tests) sum and divide after
        avg_in_op = self.average and dev.type == "cuda" and dist.get_backend(self.group) == "nccl"
        with ctx:
            torch._foreach_copy_(list(flat.split([p.numel() for p in bucket])), [p.grad.reshape(-1) for p in bucket])
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg_in_op else dist.ReduceOp.SUM, group=self.group,
                                   async_op=True)
        self._works.append((bucket, flat, work, avg_in_op))

    def _finish_deferred(self):
        """All buckets packed, then reduced back to back (the collectives queue on NCCL's stream while the current
        stream has nothing else to run), gradients left as views into the reduced buffers."""
        jobs = []
        for i, bucket in enumerate(self.buckets):
            live = [p for p in bucket if p.grad is not None]     # same set on every rank: the graph is the same
            if not live:
                continue
            sizes = [p.numel() for p in live]
            if len(live) == len(bucket):
                if self._flat[i] is None:
                    self._flat[i] = torch.empty(sum(sizes), dtype=torch.float32, device=live[0].device)
                flat = self._flat[i]
            else:
                flat = torch.empty(sum(sizes), dtype=torch.float32, device=live[0].device)
            chunks = list(flat.split(sizes))
            torch._foreach_copy_(chunks, [p.grad.reshape(-1) for p in live])
            jobs.append((live, flat, chunks))
        works = []
        for live, flat, _ in jobs:
            avg_in_op = self.average and flat.device.type == "cuda" and dist.get_backend(self.group) == "nccl"
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg_in_op else dist.ReduceOp.SUM,
                                          group=self.group, async_op=True), avg_in_op))
        for (live, flat, chunks), (work, averaged) in zip(jobs, works):
            work.wait()
            if self.average and not averaged:
                flat.div_(self.world)
            for p, c in zip(live, chunks):
                p.grad = c.view_as(p)
        self.reset()

    def finish(self):
        """Wait for all buckets, write the reduced (averaged) gradients back."""
        if self.world == 1:
            self.reset()
            return
        if not self.overlap:
            return self._finish_deferred()
        for i, left in enumerate(self._pending):
            if left != 0:
                self._launch(i, partial=True)
        for bucket, flat, work, averaged in self._works:
            work.wait()
            if self.average and not averaged:
                flat.div_(self.world)
            # one multi-tensor copy per bucket (a per-parameter loop costs ~420 launches per step)
            torch._foreach_copy_([p.grad.view(-1) if p.grad.is_contiguous() else p.grad for p in bucket],
                                 [c.view_as(p.grad) if not p.grad.is_contiguous() else c
                                  for p, c in zip(bucket, flat.split([p.numel() for p in bucket]))])
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
