"""One training iteration of the reference's train() loop (train.py:268-441) as a callable:
H2D copy -> adapter encoder -> FeatureDecoder -> bilinear resize -> Softmax -> dice loss ->
backward -> (bucketed gradient all-reduce) -> SGD step -> loss.item().

Differences from the script as shipped, all deliberate and documented in DESIGN.md:
  * the graph is left connected (no torch.no_grad around the backbone blocks / the final concat),
    so the backbone runs forward AND backward as BASELINE.json asks;
  * deform_inputs and the interpolated position embedding are cached instead of recomputed."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .decoders import FeatureDecoder
from .dp import BucketedGradAllReduce
from .encoder import AdapterEncoder


class FusedSGD(torch.optim.SGD):
    """torch.optim.SGD's constructor, param_groups, state (``momentum_buffer``), state_dict and scheduler interface
    (train.py:178-189, :191: CosineAnnealingLR steps it); the update itself is asis_sgd_step -- one kernel over 48
    tensors at a time, pointer tables in the kernel parameters, so the launches sit in the step's CUDA graph."""

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedSGD.step: closures are not supported")
        from . import kernels as K
        for group in self.param_groups:
            if group["dampening"] != 0 or group["nesterov"] or group.get("maximize", False):
                raise NotImplementedError("FusedSGD: dampening / Nesterov / maximize are not implemented (the reference uses none)")
            ps, gs, ms = [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if st.get("momentum_buffer") is None:       # zero buffer: the first update then equals PyTorch's (buf = g')
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ps.append(p)
                gs.append(p.grad)
                ms.append(st["momentum_buffer"])
            K.sgd_step(ps, gs, ms, group["lr"], group["momentum"], group["weight_decay"])
        return None


def dice_loss_after_softmax(prob, target, n_classes):
    """DC(n).forward on the output of nn.Softmax(1) (train.py:424-428; segloss/dice.py:5-36): the
    reference applies softmax a second time inside the loss; kept as is."""
    p = torch.softmax(prob, 1)
    onehot = F.one_hot(target.long(), n_classes).permute(0, 3, 1, 2).to(p.dtype)
    inter = (p * onehot).sum((2, 3))
    dice = (2 * inter) / (p.sum((2, 3)) + onehot.sum((2, 3)) + 10e-20)
    return 1.0 - dice.mean()


class TrainStep(nn.Module):
    def __init__(self, arch="vit_large", num_classes=2, adapter_heads=8, inplanes=64, lr=0.01, momentum=0.99,
                 weight_decay=3e-5, train_backbone=True, dec_features=None, device="cuda", precision="bf16",
                 bucket_bytes=None):
        super().__init__()
        self.precision = precision
        self.encoder = AdapterEncoder(arch=arch, adapter_heads=adapter_heads, inplanes=inplanes,
                                      frozen_backbone=not train_backbone, injector_init=0.0)
        C = self.encoder.model.embed_dim
        feats = dec_features or [C, 512, 256, 128, 64]
        self.seg_decoder = FeatureDecoder(embed_dim=C, num_classes=num_classes, features=feats)
        self.num_classes = num_classes
        self.to(device)
        self.encoder.model.eval()
        if not train_backbone:
            for p in self.encoder.model.parameters():
                p.requires_grad_(False)
        params = [p for p in self.parameters() if p.requires_grad]
        # reference: SGD(momentum=0.99, weight_decay=3e-5), train.py:178-189
        on_cuda = torch.device(device).type == "cuda"
        if on_cuda:
            self.optimizer = FusedSGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        else:        # host-side logic tests only
            self.optimizer = torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay, foreach=True)
        # bucket size: 64 MB when the buckets are reduced while backward still runs (granularity of the overlap), 256 MB
        # when they are reduced back to back afterwards (the default schedule, dp.py: fewer, larger collectives)
        self.reducer = BucketedGradAllReduce(params, bucket_bytes=bucket_bytes)
        self.device = torch.device(device)
        self._graph = None            # (CUDAGraph, static input, static target, static loss, our kernel launches per replay)
        self._replayed = False

    def forward_loss(self, inp, target, internals=None):
        """Loss of one batch.  ``internals`` (optional dict) receives the tensors the parity tests compare with
        the oracle: encoder state ("x", "c", "feat"), full-resolution "logits", and the "loss" itself."""
        with Fn.precision(self.precision):
            res = self.encoder(inp)
            feat = res["feat"]
            out = self.seg_decoder(feat)          # [B, n_cls, 672, 672] f32 (channels-last memory)
            H, W = target.shape[1], target.shape[2]
            out = F.interpolate(out, size=(H, W), mode="bilinear")
            if internals is not None:
                internals.update(x=res["x"], c=res["c"], feat=feat, logits=out)
            out = torch.softmax(out, 1)
            loss = dice_loss_after_softmax(out, target, self.num_classes)
            if internals is not None:
                internals["loss"] = loss
            return loss

    # ---- whole-step CUDA graph -------------------------------------------------------------------------------------
    def capture(self, inp, target, warmup=2):
        """Capture ONE whole training iteration (forward, backward, gradient all-reduce, SGD update: ~1800 kernel
        launches, ~1400 of them ours through the C ABI) into a CUDA graph that `step_device` / `step` then replay.
        The step has static shapes and no host reads (deform_inputs, the interpolated position embedding and the
        spatial-shape checks are cached after the first call), so replaying it removes the launch gaps of the
        eager step (kernel time 142 ms vs 148 ms wall in round 1).  `inp` / `target`: a device batch of the shape
        every later call will have.  The learning rate is baked into the graph: re-capture after changing it."""
        from . import _lib
        if self._graph is not None:
            raise RuntimeError("TrainStep.capture: already captured")
        try:        # warm-up runs on a side stream while the AccumulateGrad nodes were created on the default one: expected here
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):               # warm-up on a side stream, as torch.cuda.graph asks
            for _ in range(max(1, warmup)):
                self._eager_step(inp, target)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g_inp, g_tgt = inp.clone(), target.clone()
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            loss = self._eager_step(g_inp, g_tgt, zero=False)
        launches = _lib.launch_count() - n0
        self._graph = (graph, g_inp, g_tgt, loss, launches)
        return launches

    @property
    def graph_launches(self):
        """libasis_b200 kernel launches inside one replay of the captured step (0 when not captured)."""
        return 0 if self._graph is None else self._graph[4]

    def _eager_step(self, inp, target, zero=True):
        if self._replayed:
            # a replay updates the parameters without bumping their version counters: cached bf16 copies are stale
            Fn.invalidate_weight_cache()
            self._replayed = False
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        with Fn.precision(self.precision):
            loss = self.forward_loss(inp, target)
            loss.backward()
        self.reducer.finish()
        self.optimizer.step()
        return loss

    def step_device(self, inp, target, eager=False):
        """inputs already on the device; returns the loss tensor (no host sync).  Replays the captured graph when
        there is one (`capture`), unless `eager`."""
        if self._graph is None or eager:
            return self._eager_step(inp, target)
        graph, g_inp, g_tgt, loss, _ = self._graph
        if inp.data_ptr() != g_inp.data_ptr():
            g_inp.copy_(inp, non_blocking=True)
            g_tgt.copy_(target, non_blocking=True)
        graph.replay()
        self._replayed = True
        return loss

    def step_frames(self, frames_host, masks_host):
        """The same call fed as the dataset produces its samples (tools/dataset.py:111-118): uint8 HWC frames
        [B, H, W, 3] and uint8 masks [B, H, W] in pinned host memory.  The / 255, the HWC -> CHW transposition and the
        .long() run on the device (asis_frames_to_batch) straight into the step's input buffers: a quarter of the
        host-to-device bytes of the float batch, bit-identical tensors."""
        from . import kernels as K
        fr = frames_host.to(self.device, non_blocking=True)
        mk = masks_host.to(self.device, non_blocking=True)
        if self._graph is not None:
            graph, g_inp, g_tgt, loss, _ = self._graph
            K.frames_to_batch(fr, mk, img_out=g_inp, target_out=g_tgt)
            graph.replay()
            self._replayed = True
            return float(loss.item())
        inp, target = K.frames_to_batch(fr, mk)
        return float(self.step_device(inp, target).item())

    def step(self, inp_host, target_host):
        """The call a user makes (train.py:270-271, :432-440): pinned host batch in, python float out."""
        if self._graph is not None:           # straight into the graph's static input buffers
            graph, g_inp, g_tgt, loss, _ = self._graph
            g_inp.copy_(inp_host, non_blocking=True)
            g_tgt.copy_(target_host, non_blocking=True)
            graph.replay()
            self._replayed = True
            return float(loss.item())
        inp = inp_host.to(self.device, non_blocking=True)
        target = target_host.to(self.device, non_blocking=True)
        loss = self.step_device(inp, target)
        return float(loss.item())
