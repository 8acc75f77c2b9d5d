"""A whole training step captured once into a CUDA graph and replayed on static slots.

The reference drives every block from Python, one op at a time (src/as_rigid_as_possible/main.py:217-230 calling
models.py:142-146): ~3000 kernel launches per step, which on a B200 are bound by host launch overhead (eager step
26 ms vs 15 ms replayed, bench.py ``ms_per_step_eager``).  ``CapturedTrainStep`` turns the reference loop

    outputs = model(Di, DiA, mask, inputs); loss = criterion(...); optimizer.zero_grad(); loss.backward(); optimizer.step()

into ``install(batch)`` + ``replay()``:

    step = CapturedTrainStep(model, loss_fn, optimizer,
                             tensors={"inputs": x, "targets": y, "mask": m}, operators={"Di": D, "DiA": DA})
    for batch in loader:                                  # any iterable of host / device batches
        step.install(tensors=batch.tensors, operators=batch.operators)
        loss = step.replay()                              # static device scalar; float(loss) synchronises

* Static slots: the tensors / operators handed to the constructor ARE the buffers the captured kernels read.  New batches
  are copied into them (``install``; host sources should be pinned) -- operator slots take any operator of the same
  batch shape whose block / nnz count fits the slot's capacity (``Bsr4Operator.load_from``), transposes included, or
  are written in place by ``MeshOperatorCache.assemble(out=slot)`` / ``build_dirac_operators``.
* Gradients: parameters are reset with ``grad = None`` inside the step, so autograd hands every gradient over without an
  accumulation kernel (with pre-existing ``.grad`` views it launches one add per parameter: 124 per step for the
  15-block models).  With more than one rank the gradients are packed into one flat buffer (two multi-tensor copies),
  summed with ONE NCCL all-reduce inside the captured step and unpacked.
* ``capture=False`` (or a failed capture) runs the same step eagerly; ``eager_step()`` is always available and is what
  the graph was captured from, so replayed and eager steps produce the same bits (tests/test_gpu_graph.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .operators import Bsr4Operator, CsrOperator

__all__ = ["CapturedTrainStep", "load_operator"]


def load_operator(slot, new, clamp=False):
    """Overwrite the device arrays of operator ``slot`` (and of its transpose) with ``new``'s, keeping addresses.
    ``clamp``: see Bsr4Operator.load_from (operators built without a read-back report their capacity as count)."""
    if not isinstance(slot, (Bsr4Operator, CsrOperator)):
        raise TypeError("operator slots must be CsrOperator / Bsr4Operator, got %r" % type(slot))
    slot.load_from(new, clamp)
    if slot._T is not None:
        slot._T.load_from(new.T, clamp)
    return slot


class CapturedTrainStep:
    """forward + loss + backward + (N > 1) gradient all-reduce + optimizer step as one replayable CUDA graph.

    model      the nn.Module to train (on the current CUDA device)
    loss_fn    callable(model, tensors, operators) -> scalar loss tensor, e.g.
               ``lambda m, t, o: arap_loss(m(o["Di"], o["DiA"], t["mask"], t["inputs"]), t["targets"], t["mask"], B)``
    optimizer  a torch optimizer whose step is capturable (Adam(..., fused=True, capturable=True))
    tensors    {name: static device tensor}; operators: {name: CsrOperator | Bsr4Operator} (transposes are built here so
               the backward structures are part of the slots)
    warmup     eager steps before the capture (initialises optimizer state, autotunes nothing: kernels are hand-picked)
    """

    def __init__(self, model, loss_fn, optimizer, tensors, operators=None, warmup=3, capture=True, world_size=None):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.tensors = dict(tensors)
        self.operators = dict(operators or {})
        for op in self.operators.values():
            op.T  # noqa: B018  (build the backward structure now, outside any timed / captured region)
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        if world_size is None:
            world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.world = world_size
        self.flat = self.views = None
        if self.world > 1:
            p0 = self.params[0]
            self.flat = torch.zeros(sum(p.numel() for p in self.params), device=p0.device, dtype=p0.dtype)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        self.graph, self.loss, self.mode = None, None, "eager"
        for _ in range(max(int(warmup), 0)):
            self.eager_step()
        if capture:
            self._capture()

    # ------------------------------------------------------------------------------------------------- the step
    @property
    def grad_bytes(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    def eager_step(self):
        """One training step on the current slots, launched op by op; returns the loss tensor."""
        for p in self.params:
            p.grad = None                       # autograd then hands gradients over instead of accumulating (no add kernels)
        loss = self.loss_fn(self.model, self.tensors, self.operators)
        loss.backward()
        if self.world > 1:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            torch._foreach_copy_(self.views, grads)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / self.world)
            for p, g, v in zip(self.params, grads, self.views):
                if p.grad is None:
                    p.grad = g
            torch._foreach_copy_(grads, self.views)
        self.optimizer.step()
        return loss

    def _capture(self):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):       # capture needs a non-default stream history for the allocator pools
                for _ in range(2):
                    self.eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = self.eager_step()
            self.graph, self.loss, self.mode = graph, loss, "cuda_graph_replay"
        except Exception as exc:                # capture is an optimisation, not a requirement
            self.graph, self.loss = None, None
            self.mode = "eager (graph capture failed: %s)" % str(exc).splitlines()[0][:120]
            torch.cuda.synchronize()

    def replay(self):
        """Run one step on whatever the slots hold now; returns the (static, device) loss tensor."""
        if self.graph is not None:
            self.graph.replay()
            return self.loss
        return self.eager_step()

    __call__ = replay

    # ------------------------------------------------------------------------------------------------- slots
    def install(self, tensors=None, operators=None, wait_event=None, clamp_operators=False):
        """Copy a new batch into the static slots on the current stream (stream-ordered before the next ``replay``).
        ``wait_event``: a CUDA event the copies must wait for (e.g. the staging stream's upload + conversion);
        ``clamp_operators``: the new operators were built without a read-back (see ``load_operator``)."""
        if wait_event is not None:
            torch.cuda.current_stream().wait_event(wait_event)
        for k, src in (tensors or {}).items():
            slot = self.tensors[k]
            if slot.shape != src.shape:
                raise ValueError("tensor slot %r has shape %s, got %s" % (k, tuple(slot.shape), tuple(src.shape)))
            slot.copy_(src, non_blocking=True)
        for k, new in (operators or {}).items():
            load_operator(self.operators[k], new, clamp_operators)
        return self
