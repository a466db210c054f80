"""CrossDomainTrainer on the xdr hot path -- mirror of reference trainer/trainer.py:19-76 plus the per-batch inner
loop of ``recbole.trainer.Trainer._train_epoch`` [recbole-1.0.1] that it inherits.

Scope: the phase loop and the ``zero_grad -> calculate_loss -> backward -> step`` inner
loop.  Evaluation, early stopping, checkpointing and logging stay with the host framework (out of scope); ``fit``
accepts the reference's signature and ignores what it does not implement.

Two inner loops:

* ``_train_epoch``: the reference's loop, batch by batch through ``model.calculate_loss`` (any model, any
  ``torch.optim`` learner).  The only change: the running loss stays on the device and is read once per epoch instead
  of ``loss.item()`` every step (a device->host sync per batch in recbole).
* ``_train_epoch_fused`` (``config['xdr_fused_steps'] = K > 0``, learner ``sgd``, models exposing ``fused_step_spec``):
  K batches are packed into one pinned ``[K, rows, B]`` id block, copied H2D on a copy stream, and run as ONE
  persistent launch (``xdr_train_steps``) that does forward, backward and the SGD update (scatter-add of ``-lr * grad``
  into the tables); per-step losses come back in one D2H copy.  Chunks are double-buffered so copies overlap compute.
"""
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib, ops
from ..utils import train_mode2state
from .row_optim import RowSparseOptimizer


class FusedStepRunner:
    """K-step persistent launches fed from pinned host id blocks (the engine under ``_train_epoch_fused``).

    ``spec`` = dict(user_tab, item_tab, pairwise, loss_kind, reg_weight, gamma): see ``fused_step_spec`` of the models.
    ``run(host_block)``: ``host_block`` is a pinned int64 ``[K, R, B]`` tensor (R = 3: user, item+, item-; R = 2 with a
    separate ``host_label`` ``[K, B]`` float block for pointwise losses).  Returns the ``[K]`` per-step losses as a
    pinned host tensor (valid after ``synchronize()``).
    """

    def __init__(self, spec: Dict, lr: Optional[float] = None, grad_tables=None, n_buffers: int = 2, launch=None,
                 device=None):
        self.spec = spec
        self.launch = launch    # optional callable(ids [K,R,B], label, out8): e.g. the row-sharded multi-GPU launch
        if launch is None:
            self.ut, self.it = spec['user_tab'], spec['item_tab']
            self.dev = self.ut.device
            if lr is not None:      # fused SGD: the scatter-add target is the weight table itself
                self.dst_u, self.dst_i = self.ut.data, self.it.data
                self.scale = -float(lr) * float(spec.get('loss_weight', 1.0))
            else:                   # gradient accumulation into dense .grad-style tables
                gu, gi = grad_tables if grad_tables is not None else (torch.zeros_like(self.ut), torch.zeros_like(self.it))
                self.dst_u, self.dst_i, self.scale = gu, gi, 1.0
        else:
            self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.one_call = self.dev.type == 'cuda'   # pairwise chunks from pinned blocks go through xdr_train_steps_host
        self.steps_kw = {}      # extra keyword arguments of ops.train_steps (e.g. the touch maps of lazily zeroed gradient tables)
        self.n_buffers = n_buffers
        self._bufs = []  # per in-flight chunk: (dev ids, dev label, out8, host loss, ready event, done event)
        self._turn = 0
        self.launches = 0

    def _buffers(self, shape, with_label):
        key = (tuple(shape), with_label)
        if not self._bufs or self._bufs[0]['key'] != key:
            # The block shape changed (e.g. the short last chunk of an epoch).  Launches that read the old device buffers may
            # still be queued on the main stream: keep the old buffers alive until their 'done' events have completed, and
            # make the copy stream wait for everything the main stream has queued before it writes into freshly allocated
            # memory (the caching allocator may hand out blocks whose last use is still in flight on the main stream).
            main = torch.cuda.current_stream(self.dev)
            self._retired = [b for b in getattr(self, '_retired', []) if b['used'] and not b['done'].query()]
            self._retired += [b for b in self._bufs if b['used'] and not b['done'].query()]
            self.copy_stream.wait_stream(main)
            self._bufs = []
            K = shape[0]
            for _ in range(self.n_buffers):
                ids = torch.empty(shape, dtype=torch.int64, device=self.dev)
                lab = torch.empty(shape, dtype=torch.float32, device=self.dev) if with_label else None
                for t in (ids, lab):
                    if t is not None:
                        t.record_stream(self.copy_stream)   # written on the copy stream, read on the main stream
                ready, done = torch.cuda.Event(), torch.cuda.Event()
                ready.record(self.copy_stream)   # (recorded once so that the CUDA events exist: the one-call path below hands
                done.record(main)                #  their raw handles to the library)
                self._bufs.append({
                    'key': key, 'ids': ids,
                    # the kernel walks ids and labels with ONE step stride: give the label rows the stride of the id rows
                    'label': (lab[:, 0] if with_label else None),
                    'out8': torch.empty((K, 8), dtype=torch.float32, device=self.dev),
                    'ready': ready, 'done': done, 'used': False, 'rec': None, 'src': None})
        b = self._bufs[self._turn % self.n_buffers]
        self._turn += 1
        return b

    def run(self, host_block: torch.Tensor, host_label: Optional[torch.Tensor] = None) -> torch.Tensor:
        sp = self.spec
        K, R, B = host_block.shape
        if R != (3 if sp['pairwise'] else 2):
            raise ValueError(f'id block must be [K, {3 if sp["pairwise"] else 2}, B]')
        buf = self._buffers(host_block.shape, host_label is not None)
        main = torch.cuda.current_stream(self.dev)
        if (self.one_call and self.launch is None and sp['pairwise'] and host_label is None and not self.steps_kw and
                host_block.dtype == torch.int64 and host_block.is_contiguous() and host_block.is_pinned()):
            # the whole chunk -- H2D ids on the copy stream, events, the persistent launch, D2H loss records -- in ONE library
            # call (xdr_train_steps_host): a dozen torch calls per chunk made a K = 20 pass host-bound
            rec = torch.empty((K, 8), dtype=torch.float32, pin_memory=True)
            ws = ops._steps_workspace(self.dev, K)
            ops.call('xdr_train_steps_host', ops.ptr(self.ut.data), ops.ptr(self.it.data), self.ut.shape[0], self.it.shape[0],
                     self.ut.shape[1], host_block.data_ptr(), ops.ptr(buf['ids']), B, K, float(sp.get('gamma', 1e-10)),
                     float(sp['reg_weight']), None, float(self.scale), ops.ptr(self.dst_u), ops.ptr(self.dst_i),
                     ops.ptr(buf['out8']), rec.data_ptr(), ops.ptr(ws), ws.numel(), ops._oob(self.dev),
                     self.copy_stream.cuda_stream, buf['done'].cuda_event if buf['used'] else None, buf['ready'].cuda_event,
                     buf['done'].cuda_event, main.cuda_stream)
            self.launches += 1
            buf['used'] = True
            # both pinned blocks stay referenced until this device buffer's next turn: their copies are raw CUDA calls that torch's
            # caching host allocator does not know about (a block the caller drops right after run() must not be recycled mid-DMA)
            buf['rec'], buf['src'] = rec, host_block
            return rec[:, 0]
        with torch.cuda.stream(self.copy_stream):
            if buf['used']:
                self.copy_stream.wait_event(buf['done'])  # the launch that last read this buffer has finished
            buf['ids'].copy_(host_block, non_blocking=True)
            if host_label is not None:
                buf['label'].copy_(host_label, non_blocking=True)
            buf['ready'].record(self.copy_stream)
        main.wait_event(buf['ready'])
        ids = buf['ids']
        if self.launch is not None:
            self.launch(ids, buf['label'], buf['out8'])
        else:
            ops.train_steps(self.ut.data, self.it.data, ids[:, 0], ids[:, 1], ids[:, 2] if sp['pairwise'] else None,
                            buf['label'], loss_kind=sp.get('loss_kind', _lib.LOSS_MSE), reg_weight=sp['reg_weight'],
                            gamma=sp.get('gamma', 1e-10), user_dst=self.dst_u, item_dst=self.dst_i, scale=self.scale,
                            out8=buf['out8'], **self.steps_kw)
        self.launches += 1
        # one pinned tensor PER RUN (torch's caching host allocator makes this cheap): the caller owns it, so a later chunk
        # that reuses this device buffer cannot overwrite losses the caller has not read yet
        # (the whole [K, 8] record in one contiguous device->host copy: a strided column would cost a gather kernel first)
        rec = torch.empty((K, 8), dtype=torch.float32, pin_memory=True)
        rec.copy_(buf['out8'], non_blocking=True)
        buf['done'].record(main)
        buf['used'] = True
        return rec[:, 0]

    def synchronize(self):
        torch.cuda.current_stream(self.dev).synchronize()


class CrossDomainTrainer(object):
    r"""Trainer for cross-domain models with the four training modes SOURCE, TARGET, BOTH, OVERLAP set by
    ``train_epochs`` (parsed by the host config into ``train_modes`` / ``epoch_num``, reference trainer.py:26-31)."""

    def __init__(self, config, model):
        self.config = config
        self.model = model
        self.device = config['device']
        self.learner = (config['learner'] if 'learner' in config else 'adam').lower()
        self.learning_rate = config['learning_rate'] if 'learning_rate' in config else 1e-3
        self.weight_decay = config['weight_decay'] if 'weight_decay' in config else 0.0
        self.clip_grad_norm = config['clip_grad_norm'] if 'clip_grad_norm' in config else None
        self.train_modes = config['train_modes']
        self.train_epochs = config['epoch_num']
        self.split_valid_flag = config['source_split'] if 'source_split' in config else False
        self.fused_steps = int(config['xdr_fused_steps']) if 'xdr_fused_steps' in config else 0
        self.valid_metric_bigger = config['valid_metric_bigger'] if 'valid_metric_bigger' in config else True
        # row-sparse optimizer for the embedding tables (section 8 F1): 'sgd' | 'adagrad' | 'lazy_adam'; the dense learner keeps
        # the remaining (MLP) parameters.  Needs a model with ``touched_rows(interaction)``.
        self.row_optimizer_kind = config['xdr_row_optimizer'] if 'xdr_row_optimizer' in config else None
        self.row_optimizer = None
        if self.row_optimizer_kind:
            if not hasattr(model, 'touched_rows'):
                raise ValueError(f'{type(model).__name__} does not expose touched_rows(): no row-sparse optimizer for it')
            self.row_optimizer = RowSparseOptimizer(self.row_optimizer_kind, lr=self.learning_rate)
        self.optimizer = self._build_optimizer()
        self.train_loss_dict = dict()
        self.best_valid_score, self.best_valid_result = -np.inf, None
        self.epochs = 0
        # per-phase validation / early stopping / checkpoint of recbole's Trainer.fit [recbole-1.0.1], which the reference runs
        # once per phase (trainer.py:59-73).  The evaluator itself (metrics over full_sort_predict) is outside this package's
        # scope: ``valid_fn(model, valid_data) -> (score, result)`` is supplied by the caller
        # (``set_valid_fn``); without one, passing ``valid_data`` warns instead of being silently dropped.
        self.eval_step = int(config['eval_step']) if 'eval_step' in config else 1
        self.stopping_step = int(config['stopping_step']) if 'stopping_step' in config else 10
        self.checkpoint_dir = config['checkpoint_dir'] if 'checkpoint_dir' in config else None
        self.saved_model_file = None
        self.valid_fn = None
        self.cur_step = 0
        self.start_epoch = 0

    def set_valid_fn(self, valid_fn):
        """``valid_fn(model, valid_data) -> (score: float, result)``: the evaluation used by ``fit`` every ``eval_step`` epochs."""
        self.valid_fn = valid_fn

    # ---- checkpoints in the reference's format (recbole Trainer._save_checkpoint / resume_checkpoint) ----------------
    def _save_checkpoint(self, epoch):
        if not self.checkpoint_dir:
            return None
        import os
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        if self.saved_model_file is None:
            self.saved_model_file = os.path.join(self.checkpoint_dir, f'{type(self.model).__name__}-xdr.pth')
        other = self.model.other_parameter() if hasattr(self.model, 'other_parameter') else None
        state = {'config': dict(self.config) if hasattr(self.config, 'keys') else None, 'epoch': epoch, 'cur_step': self.cur_step,
                 'best_valid_score': self.best_valid_score, 'state_dict': self.model.state_dict(), 'other_parameter': other,
                 'optimizer': self.optimizer.state_dict() if self.optimizer is not None else None}
        torch.save(state, self.saved_model_file)
        return self.saved_model_file

    def resume_checkpoint(self, resume_file):
        ckpt = torch.load(resume_file, map_location=self.device, weights_only=False)
        self.start_epoch = ckpt['epoch'] + 1
        self.cur_step = ckpt['cur_step']
        self.best_valid_score = ckpt['best_valid_score']
        self.model.load_state_dict(ckpt['state_dict'])
        if ckpt.get('other_parameter') is not None and hasattr(self.model, 'load_other_parameter'):
            self.model.load_other_parameter(ckpt['other_parameter'])
        if self.optimizer is not None and ckpt.get('optimizer') is not None:
            self.optimizer.load_state_dict(ckpt['optimizer'])

    def _build_optimizer(self):
        """recbole Trainer._build_optimizer [recbole-1.0.1]: dense torch optimizers keyed by ``learner``."""
        params = self.model.parameters()
        if self.row_optimizer is not None:   # the tables belong to the row-sparse optimizer
            params = [p for n, p in self.model.named_parameters() if not n.endswith('_embedding.weight')]
            if not params:
                return None
        lr, wd = self.learning_rate, self.weight_decay
        table = {'adam': torch.optim.Adam, 'sgd': torch.optim.SGD, 'adagrad': torch.optim.Adagrad,
                 'rmsprop': torch.optim.RMSprop}
        if self.learner == 'sparse_adam':
            return torch.optim.SparseAdam(params, lr=lr)
        return table.get(self.learner, torch.optim.Adam)(params, lr=lr, weight_decay=wd)

    def _reinit(self, phase):
        """Reset per-phase state (reference trainer.py:30-41)."""
        self.start_epoch = 0
        self.cur_step = 0
        self.best_valid_score = -np.inf if self.valid_metric_bigger else np.inf
        self.best_valid_result = None
        self.train_loss_dict = dict()
        self.epochs = int(self.train_epochs[phase])

    @staticmethod
    def _check_nan(loss):
        """End-of-epoch checks at the one place the epoch reads from the device anyway: the loss is a number, and no kernel of
        the epoch met an out-of-range id (the reference's nn.Embedding would have raised at the op)."""
        if torch.isnan(loss).any():
            raise ValueError('Training loss is nan')
        if loss.is_cuda:
            ops.check_ids_now(loss.device)

    # ---- inner loops -----------------------------------------------------------------------------------------
    def _train_epoch(self, train_data, epoch_idx, loss_func=None, show_progress=False):
        """recbole Trainer._train_epoch: zero_grad -> calculate_loss -> (sum tuple) -> backward -> [clip] -> step."""
        self.model.train()
        loss_func = loss_func or self.model.calculate_loss
        total = None
        if self.row_optimizer is not None:
            return self._train_epoch_row_sparse(train_data, loss_func)
        for interaction in train_data:
            interaction = interaction.to(self.device)
            self.optimizer.zero_grad()
            losses = loss_func(interaction)
            loss = sum(losses) if isinstance(losses, tuple) else losses
            if loss.dim() > 0:
                loss = loss.sum()
            total = loss.detach() if total is None else total + loss.detach()
            loss.backward()
            if self.clip_grad_norm:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), **self.clip_grad_norm)
            self.optimizer.step()
        if total is None:
            return 0.0
        self._check_nan(total)
        return float(total.item())  # one device->host read per epoch

    def _train_epoch_row_sparse(self, train_data, loss_func):
        """The reference loop with the table part of ``zero_grad`` / ``step`` replaced by one row-sparse kernel per table:
        ``backward()`` scatter-adds into ``table.grad`` ('inplace' table-gradient mode), the optimizer visits the batch's
        rows only and leaves ``table.grad`` zero again.  Dense parameters keep the configured torch learner."""
        if self.clip_grad_norm:
            raise ValueError('clip_grad_norm needs the dense gradient norm: not available with xdr_row_optimizer')
        prev_mode = ops.get_table_grad_mode()
        ops.set_table_grad_mode('inplace')
        total = None
        try:
            for interaction in train_data:
                interaction = interaction.to(self.device)
                if self.optimizer is not None:
                    self.optimizer.zero_grad()
                losses = loss_func(interaction)
                loss = sum(losses) if isinstance(losses, tuple) else losses
                loss = loss.sum()
                total = loss.detach() if total is None else total + loss.detach()
                loss.backward()
                if self.optimizer is not None:
                    self.optimizer.step()
                self.row_optimizer.step(self.model.touched_rows(interaction))
        finally:
            ops.set_table_grad_mode(prev_mode)
        if total is None:
            return 0.0
        self._check_nan(total)
        return float(total.item())

    def _fused_spec(self):
        spec_fn = getattr(self.model, 'fused_step_spec', None)
        if self.fused_steps <= 0 or spec_fn is None or self.learner != 'sgd' or self.weight_decay:
            return None
        return spec_fn()

    def _train_epoch_fused(self, train_data, epoch_idx, spec):
        """K batches per persistent launch with the SGD update fused.  ``spec`` is one term (EMCDR phases) or a list of
        weighted terms on shared tables (CMF BOTH: source term then target term, one launch each per chunk)."""
        self.model.train()
        specs = spec if isinstance(spec, (list, tuple)) else [spec]
        runners = [FusedStepRunner(sp, lr=self.learning_rate) for sp in specs]
        K = self.fused_steps
        pending = []   # (weight, pinned per-step losses)
        total = 0.0
        chunks = [[] for _ in specs]
        labels = [[] for _ in specs]
        widths = [None] * len(specs)

        def flush():
            for n, (sp, runner) in enumerate(zip(specs, runners)):
                if not chunks[n]:
                    continue
                block = torch.stack(chunks[n]).pin_memory()
                lab = torch.stack(labels[n]).pin_memory() if labels[n] else None
                pending.append((float(sp.get('loss_weight', 1.0)), runner.run(block, lab)))
                chunks[n], labels[n] = [], []

        for interaction in train_data:
            ids = [torch.stack([interaction[f].reshape(-1).cpu() for f in sp['fields']]) for sp in specs]
            ok = all(ops.train_steps_supported(i.shape[1], sp['user_tab'].shape[1], sp['pairwise'], self.device) and
                     (w is None or i.shape[1] == w) for i, sp, w in zip(ids, specs, widths))
            if not ok:
                # a ragged last batch (or a shape the persistent kernel does not take) goes through the per-step path
                flush()
                runners[0].synchronize()
                total += self._train_epoch([interaction], epoch_idx)
                continue
            for n, (i, sp) in enumerate(zip(ids, specs)):
                widths[n] = i.shape[1]
                chunks[n].append(i)
                if sp.get('label_field'):
                    labels[n].append(interaction[sp['label_field']].reshape(-1).float().cpu())
            if len(chunks[0]) == K:
                flush()
        flush()
        runners[0].synchronize()
        for w, l in pending:
            total += w * float(l.sum().item())
        if total != total:
            raise ValueError('Training loss is nan')
        return total

    def train_epoch_device(self, domain_data, batch_size, spec=None, steps_per_launch=None, generator=None):
        """One epoch with NOTHING on the host per step: ``domain_data`` (``data.DeviceDomainData``) permutes the positives
        and draws the negatives on the GPU, every block of K steps is one persistent launch with the SGD update fused.
        Returns the summed per-step loss (one device->host read per epoch)."""
        spec = spec or self.model.fused_step_spec()
        if spec is None or isinstance(spec, (list, tuple)):
            raise ValueError('train_epoch_device needs a single-term fused step spec (e.g. EMCDR SOURCE / TARGET phase)')
        K = steps_per_launch or max(self.fused_steps, 1)
        total = torch.zeros((), dtype=torch.float32, device=spec['user_tab'].device)
        if self.row_optimizer is not None:
            return self._train_epoch_device_row_sparse(domain_data, batch_size, spec, K, generator, total)
        self._require_plain_sgd('train_epoch_device')
        scale = -float(self.learning_rate) * float(spec.get('loss_weight', 1.0))
        for ids, label in domain_data.epoch_blocks(batch_size, K, pairwise=spec['pairwise'], generator=generator, drop_last=False):
            out8, _, _ = ops.train_steps(spec['user_tab'].data, spec['item_tab'].data, ids[:, 0], ids[:, 1],
                                         ids[:, 2] if spec['pairwise'] else None, label,
                                         loss_kind=spec.get('loss_kind', _lib.LOSS_MSE), reg_weight=spec['reg_weight'],
                                         gamma=spec.get('gamma', 1e-10), user_dst=spec['user_tab'].data,
                                         item_dst=spec['item_tab'].data, scale=scale)
            total = total + out8[:, 0].sum()
        self._check_nan(total)
        domain_data.sampler.check_status()
        return float(total.item())

    def _require_plain_sgd(self, what):
        """The fused launches apply ``-lr * grad`` inside the scatter: that IS plain SGD and nothing else (the
        trainer's default learner is adam)."""
        if self.learner != 'sgd' or self.weight_decay:
            raise ValueError(f"{what} applies a fused plain-SGD update; it needs learner='sgd' and weight_decay=0 "
                             f"(got learner={self.learner!r}, weight_decay={self.weight_decay!r}), or an xdr_row_optimizer")

    def train_epoch_device_both(self, source_data, target_data, batch_size, specs=None, steps_per_launch=None,
                                generator=None):
        """BOTH-mode epoch on the device for models whose loss is a weighted sum of one term per domain on shared tables
        (CMF, cmf.py:81-99): the epoch has as many steps as the TARGET domain has batches and the source domain's batches
        restart when they run out (reference data/dataloader.py:129-137,156-159).  Per block of K steps: one persistent
        launch per term, SGD update fused (``-lr * loss_weight``).  Returns the summed weighted per-step loss."""
        specs = specs or self.model.fused_step_spec()
        if not isinstance(specs, (list, tuple)) or len(specs) != 2:
            raise ValueError('train_epoch_device_both needs a [source term, target term] fused step spec (e.g. CMF)')
        K = steps_per_launch or max(self.fused_steps, 1)
        dev = specs[0]['user_tab'].device
        total = torch.zeros((), dtype=torch.float32, device=dev)
        self._require_plain_sgd('train_epoch_device_both')

        def source_blocks():
            while True:  # the source loader silently restarts (dataloader.py:156-159)
                got = False
                for blk in source_data.epoch_blocks(batch_size, K, pairwise=specs[0]['pairwise'], generator=generator):
                    got = True
                    yield blk
                if not got:
                    raise ValueError('the source domain has fewer interactions than one batch')

        def run(sp, ids, lab):
            w = float(sp.get('loss_weight', 1.0))
            out8, _, _ = ops.train_steps(sp['user_tab'].data, sp['item_tab'].data, ids[:, 0], ids[:, 1],
                                         ids[:, 2] if sp['pairwise'] else None, lab,
                                         loss_kind=sp.get('loss_kind', _lib.LOSS_MSE), reg_weight=sp['reg_weight'],
                                         gamma=sp.get('gamma', 1e-10), user_dst=sp['user_tab'].data,
                                         item_dst=sp['item_tab'].data, scale=-float(self.learning_rate) * w)
            return w * out8[:, 0].sum()

        src = source_blocks()
        pending = None   # source steps left over from the previous block
        for t_ids, t_lab in target_data.epoch_blocks(batch_size, K, pairwise=specs[1]['pairwise'], generator=generator):
            need = t_ids.shape[0]
            while need > 0:      # as many source steps as target steps in this block (views: the step stride is kept)
                if pending is None:
                    pending = next(src)
                s_ids, s_lab = pending
                take = min(need, s_ids.shape[0])
                total = total + run(specs[0], s_ids[:take], None if s_lab is None else s_lab[:take])
                pending = (s_ids[take:], None if s_lab is None else s_lab[take:]) if take < s_ids.shape[0] else None
                need -= take
            total = total + run(specs[1], t_ids, t_lab)
        self._check_nan(total)
        source_data.sampler.check_status()
        target_data.sampler.check_status()
        return float(total.item())

    def _train_epoch_device_row_sparse(self, domain_data, batch_size, spec, K, generator, total):
        """Device pipeline with a row-sparse Adagrad / lazy-Adam / SGD step after EVERY batch (sequential semantics, unlike
        the asynchronous fused-SGD launch): per step one persistent-kernel launch that scatter-adds the batch's gradient
        rows into the gradient tables, then one optimizer kernel per table over the batch's ids, which also re-zeroes the
        rows it consumed.  No host work per step; the losses are read once per epoch."""
        ut, it = spec['user_tab'], spec['item_tab']
        for t in (ut, it):
            if t.grad is None:
                t.grad = torch.zeros_like(t)
        weight = float(spec.get('loss_weight', 1.0))
        for ids, label in domain_data.epoch_blocks(batch_size, K, pairwise=spec['pairwise'], generator=generator):
            for k in range(ids.shape[0]):
                out8, _, _ = ops.train_steps(ut.data, it.data, ids[k:k + 1, 0], ids[k:k + 1, 1],
                                             ids[k:k + 1, 2] if spec['pairwise'] else None,
                                             None if label is None else label[k:k + 1],
                                             loss_kind=spec.get('loss_kind', _lib.LOSS_MSE), reg_weight=spec['reg_weight'],
                                             gamma=spec.get('gamma', 1e-10), user_dst=ut.grad, item_dst=it.grad, scale=weight)
                self.row_optimizer.step([(ut, ids[k, 0]), (it, ids[k, 1:])])
                total = total + out8[0, 0]
        self._check_nan(total)
        domain_data.sampler.check_status()
        return float(total.item())

    # ---- phase loop ------------------------------------------------------------------------------------------
    def _fit_phase(self, train_data, valid_data, verbose, saved, show_progress, callback_fn):
        """One phase of recbole's Trainer.fit: train; every ``eval_step`` epochs validate, keep the best score, save the
        checkpoint on improvement, stop after ``stopping_step`` validations without one."""
        spec = self._fused_spec()
        if valid_data is not None and self.valid_fn is None:
            import warnings
            warnings.warn('CrossDomainTrainer.fit: valid_data was given but no evaluation function is set (set_valid_fn); this '
                          'phase trains all its epochs without validation, early stopping or best-model selection', stacklevel=3)
        for epoch_idx in range(self.start_epoch, self.epochs):
            if spec is not None:
                loss = self._train_epoch_fused(train_data, epoch_idx, spec)
            else:
                loss = self._train_epoch(train_data, epoch_idx, show_progress=show_progress)
            self.train_loss_dict[epoch_idx] = loss
            if callback_fn:
                callback_fn(epoch_idx, loss)
            if valid_data is None or self.valid_fn is None or self.eval_step <= 0:
                if saved and (epoch_idx + 1 == self.epochs):
                    self._save_checkpoint(epoch_idx)
                continue
            if (epoch_idx + 1) % self.eval_step == 0:
                self.model.eval()
                with torch.no_grad():
                    score, result = self.valid_fn(self.model, valid_data)
                better = score > self.best_valid_score if self.valid_metric_bigger else score < self.best_valid_score
                if better:
                    self.best_valid_score, self.best_valid_result, self.cur_step = score, result, 0
                    if saved:
                        self._save_checkpoint(epoch_idx)
                else:
                    self.cur_step += 1
                    if self.cur_step > self.stopping_step:
                        break
        return self.best_valid_score, self.best_valid_result

    def fit(self, train_data, valid_data=None, verbose=True, saved=True, show_progress=False, callback_fn=None):
        r"""For each ``train_epochs`` entry: reset, switch the dataloader state, switch the model phase, train
        (reference trainer.py:59-73); ends with ``model.set_phase('OVERLAP')`` (trainer.py:75)."""
        for phase in range(len(self.train_modes)):
            self._reinit(phase)
            scheme = self.train_modes[phase]
            train_data.set_mode(train_mode2state[scheme])
            self.model.set_phase(scheme)
            if self.split_valid_flag and valid_data is not None:
                source_valid_data, target_valid_data = valid_data
                vd = source_valid_data if scheme == 'SOURCE' else target_valid_data
            else:
                vd = valid_data
            self._fit_phase(train_data, vd, verbose, saved, show_progress, callback_fn)
        self.model.set_phase('OVERLAP')
        return self.best_valid_score, self.best_valid_result


class DCDCSRTrainer(CrossDomainTrainer):
    r"""Trainer of DCDCSR (reference trainer/trainer.py:79-131): the CrossDomainTrainer phase loop, except that the BOTH
    phase (benchmark + mapping fit) trains without validation data."""

    def fit(self, train_data, valid_data=None, verbose=True, saved=True, show_progress=False, callback_fn=None):
        for phase in range(len(self.train_modes)):
            self._reinit(phase)
            scheme = self.train_modes[phase]
            train_data.set_mode(train_mode2state[scheme])
            self.model.set_phase(scheme)
            if scheme == 'BOTH':
                vd = None
            elif self.split_valid_flag and valid_data is not None:
                source_valid_data, target_valid_data = valid_data
                vd = source_valid_data if scheme == 'SOURCE' else target_valid_data
            else:
                vd = valid_data
            self._fit_phase(train_data, vd, verbose, saved, show_progress, callback_fn)
        self.model.set_phase('OVERLAP')
        return self.best_valid_score, self.best_valid_result
