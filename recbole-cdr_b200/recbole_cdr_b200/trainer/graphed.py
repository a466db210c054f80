"""CUDA-graph capture of one whole training step (forward + backward [+ optimizer]) of a drop-in model.

The composed models (EMCDR map phase, CoNet, DTCDR, BiTGCF) are chains of 20-60 small libxdr kernels per step; launched
one by one from Python the step is pure host overhead (measured: DTCDR 1.24 ms/step eager for ~0.15 ms of GPU work).
Static batch shapes make the chain capturable: ids/labels live in static device buffers, the captured graph is replayed
per batch with no Python between kernels.  This is the "CUDA streams and graphs instead of a tracing compiler" path: the
kernels are the same hand-written ones, only the launch mechanism changes.

Embedding-table gradients follow ``ops`` table-grad mode ``'inplace'`` during capture: ``loss.backward()`` scatter-adds
into the persistent ``weight.grad`` of every table (allocated once, before capture), so no dense ``[N, D]`` tensor is
created or zero-filled inside the graph.  Small dense parameters get fresh ``.grad`` tensors from the graph's pool.
"""
from typing import Callable, Optional

import torch

from .. import ops
from ..data.interaction import Interaction

_TABLE_NUMEL = 1 << 18  # parameters at least this large are treated as embedding tables (persistent .grad)


class GraphedTrainStep(object):
    def __init__(self, model, example: Interaction, optimizer: Optional[torch.optim.Optimizer] = None,
                 loss_fn: Optional[Callable] = None, warmup: int = 3):
        self.model, self.optimizer = model, optimizer
        self.loss_fn = loss_fn or model.calculate_loss
        self.static = Interaction({k: example[k].clone() for k in example.columns})
        self._prev_mode = ops.get_table_grad_mode()
        ops.set_table_grad_mode('inplace')
        self._table_grads = {}
        # The warm-up steps below are REAL steps on the example batch (they allocate what the capture needs: cuBLAS-free here,
        # but the graph pool, the small parameters' .grad tensors and the optimizer's state).  Everything they change is put
        # back afterwards -- in place, so that the captured graph keeps pointing at the same memory: table gradients (zeroed
        # if created here, restored if the caller had some), and, when an optimizer steps inside the graph, the parameters
        # and the optimizer state (a fresh optimizer ends up with allocated, zeroed state: what its first real step expects).
        saved_grads, saved_params, saved_state = {}, {}, None
        for p in model.parameters():
            if p.numel() >= _TABLE_NUMEL:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                else:
                    saved_grads[p] = p.grad.clone()
                self._table_grads[p] = p.grad   # the captured kernels scatter-add into exactly this memory
        if optimizer is not None:
            saved_params = {p: p.detach().clone() for p in model.parameters()}
            saved_state = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                           for p, st in optimizer.state.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        dev = next(model.parameters()).device
        if dev.type == 'cuda':   # streams that ops.cross_pair forks its independent launches onto inside the capture
            ops.side_streams(dev, max(1, ops.CROSS_STREAMS - 1))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero_small()
                self._step()
            torch.cuda.synchronize()
            self._zero_small()
            with torch.no_grad():
                for p, gbuf in self._table_grads.items():
                    if p in saved_grads:
                        gbuf.copy_(saved_grads[p])
                    else:
                        gbuf.zero_()
                for p, v in saved_params.items():
                    p.copy_(v)
                if optimizer is not None:
                    for p, st in optimizer.state.items():
                        before = saved_state.get(p, {})
                        for k, v in st.items():
                            if torch.is_tensor(v):
                                v.copy_(before[k]) if k in before else v.zero_()
                            elif k in before:
                                st[k] = before[k]
            del saved_grads, saved_params, saved_state
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.loss = self._step()
        torch.cuda.current_stream().wait_stream(side)
        ops.set_table_grad_mode(self._prev_mode)

    def zero_table_grads(self):
        for gbuf in self._table_grads.values():
            gbuf.zero_()

    def _zero_small(self):
        for p in self.model.parameters():
            if p.numel() < _TABLE_NUMEL:
                p.grad = None

    def _step(self):
        losses = self.loss_fn(self.static)
        loss = sum(losses) if isinstance(losses, tuple) else losses
        if loss.dim() > 0:   # (a 0-d loss needs no reduction kernel and no expand in its backward)
            loss = loss.sum()
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def __call__(self, interaction: Interaction) -> torch.Tensor:
        """Copy the batch into the static buffers (shapes must match the example) and replay the captured step.
        Table gradients ACCUMULATE in the persistent ``weight.grad`` buffers across calls (``zero_table_grads()`` clears
        them); if the caller dropped or replaced ``weight.grad`` it is re-attached here."""
        for p, gbuf in self._table_grads.items():
            if p.grad is not gbuf:
                p.grad = gbuf
        for k in self.static.columns:
            src = interaction[k]
            if src.shape != self.static[k].shape:
                raise ValueError(f'field {k}: shape {tuple(src.shape)} differs from the captured {tuple(self.static[k].shape)}')
            self.static[k].copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
