"""Row-sparse optimizers for embedding tables.

The reference steps a dense ``torch.optim`` optimizer over every row of every table each batch (recbole
``Trainer._build_optimizer`` / ``_train_epoch`` [recbole-1.0.1]).  ``RowSparseOptimizer`` updates only the rows the batch
touched, through ``ops.sparse_optim_rows`` (one kernel per table per step), and leaves the gradient tables zero again, so no
dense ``zero_grad`` is needed either.  Semantics per ``kind``:

* ``'sgd'``      -- ``torch.optim.SGD`` (no momentum, no weight decay): identical to the dense optimizer;
* ``'adagrad'``  -- ``torch.optim.Adagrad`` (lr_decay 0, weight_decay 0): identical to the dense optimizer, because a zero
  gradient row leaves the dense optimizer's row unchanged as well;
* ``'lazy_adam'``-- ``torch.optim.SparseAdam``: moments and weights of touched rows only (dense Adam would keep moving
  untouched rows along their decaying first moment -- a different algorithm, stated here rather than hidden).

Use: build the model with ``ops.set_table_grad_mode('inplace')`` so ``loss.backward()`` scatter-adds into ``table.grad``;
after ``backward()`` call ``opt.step([(table_param, ids), ...])`` with the batch's ids per table.
"""
from typing import Iterable, Tuple

import torch

from .. import _lib, ops

KINDS = {'sgd': _lib.OPT_SGD, 'adagrad': _lib.OPT_ADAGRAD, 'lazy_adam': _lib.OPT_LAZY_ADAM, 'sparse_adam': _lib.OPT_LAZY_ADAM}


class RowSparseOptimizer:
    def __init__(self, kind: str = 'adagrad', lr: float = 1e-3, eps: float = None, betas=(0.9, 0.999)):
        if kind not in KINDS:
            raise ValueError(f'row-sparse optimizer kind must be one of {sorted(KINDS)}, got {kind!r}')
        self.kind_name, self.kind = kind, KINDS[kind]
        self.lr = float(lr)
        # torch defaults: Adagrad eps 1e-10, (Sparse)Adam eps 1e-8
        self.eps = float(eps) if eps is not None else (1e-10 if self.kind == _lib.OPT_ADAGRAD else 1e-8)
        self.betas = (float(betas[0]), float(betas[1]))
        self.t = 0            # optimizer steps taken (Adam bias correction)
        self._state = {}      # id(table) -> dict(stamp, s1, s2, calls)

    def _table_state(self, table: torch.Tensor):
        st = self._state.get(id(table))
        if st is None:
            st = {'stamp': torch.zeros(table.shape[0], dtype=torch.int32, device=table.device), 'calls': 0,
                  's1': torch.zeros_like(table) if self.kind != _lib.OPT_SGD else None,
                  's2': torch.zeros_like(table) if self.kind == _lib.OPT_LAZY_ADAM else None}
            self._state[id(table)] = st
        return st

    def step(self, touched: Iterable[Tuple[torch.Tensor, torch.Tensor]]):
        """``touched``: (table parameter with a dense ``.grad``, ids of the batch for that table) pairs.  A table may
        appear several times (e.g. positive and negative item ids): its rows are still updated once, by the first call
        that visits them, because every call of one step shares the step id."""
        self.t += 1
        groups = {}
        for table, ids in touched:
            groups.setdefault(id(table), (table, []))[1].append(ids.reshape(-1))
        for table, id_list in groups.values():
            if table.grad is None:
                continue
            st = self._table_state(table)
            st['calls'] += 1
            ids = id_list[0] if len(id_list) == 1 else torch.cat(id_list)
            ops.sparse_optim_rows(self.kind, table.data, table.grad, ids, st['stamp'], st['calls'], self.lr,
                                  state1=st['s1'], state2=st['s2'], adam_t=self.t, eps=self.eps, beta1=self.betas[0],
                                  beta2=self.betas[1])

    def state_of(self, table: torch.Tensor):
        return self._table_state(table)
