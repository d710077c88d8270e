"""Optimiser plumbing of the module / distiller mirrors.

``ArenaAdam`` is what ``self.optimizer_G`` / ``self.optimizer_D`` are in the mirrors: the Adam update itself is the
``catb_adam`` kernel over a flat arena (one launch per network), this object carries what the reference touches from
Python -- ``param_groups[i]['lr']`` (``update_learning_rate``, models/base_model.py:146-156) and ``state_dict()`` /
``load_state_dict()`` in the layout of ``torch.optim.Adam`` (``save_networks`` writes ``<epoch>_optim-<i>.pth`` with
``optimizer.state_dict()``, distillers/base_inception_distiller.py:393-396; ``load_networks`` restores it through
``utils/util.py:143-148`` for ``--restore_O_path``), so the files are interchangeable with the reference's.

``carry_engine_state`` moves the complete training state from one compiled step engine to another of the same
architectures (a different batch shape: the reference DataLoader has no ``drop_last``, data/__init__.py:82, so every
epoch ends with a partial batch): parameters, Adam moments, step counters, learning rates, running statistics, image
pools.  Engines are cached per shape by the mirrors, so an epoch boundary costs two arena copies, no re-tuning and no
graph capture.
"""
from collections import OrderedDict

import torch


def _locate(p, arena):
    """(offset, numel) of parameter tensor `p` inside `arena.p` when it aliases the arena, else None."""
    if arena is None or arena.p is None or not isinstance(p, torch.Tensor):
        return None
    base, esz = arena.p.data_ptr(), arena.p.element_size()
    off = p.data_ptr() - base
    if p.dtype != arena.p.dtype or off < 0 or off % esz or off // esz + p.numel() > arena.p.numel():
        return None
    return off // esz, p.numel()


class ArenaAdam:
    """torch.optim.Adam look-alike over engine arenas.

    bind(groups): one entry per param group, each a list of members ``(params, arena, step_counter)`` in the order in
    which the reference hands the parameters to Adam (e.g. optimizer_G of the Inception distiller = group 0: the
    student's parameters, group 1: the four adaptor convs, base_inception_distiller.py:205-214); ``params`` are the mirror
    module's parameters (which alias ``arena.p``), ``arena`` / ``step_counter`` may be None for parameters that no
    kernel updates (the adaptors under the 'ka' loss: they never receive a gradient in the reference either)."""

    def __init__(self, lr, betas, n_groups=1):
        self.defaults = {'lr': lr, 'betas': tuple(betas), 'eps': 1e-8, 'weight_decay': 0, 'amsgrad': False}
        self.param_groups = [dict(self.defaults, params=[]) for _ in range(n_groups)]
        self._groups = None
        self._pending = None

    def bind(self, groups):
        assert len(groups) == len(self.param_groups)
        self._groups = [[(list(params), arena, counter) for (params, arena, counter) in members] for members in groups]
        idx = 0
        for pg, members in zip(self.param_groups, self._groups):
            pg['params'] = list(range(idx, idx + sum(len(m[0]) for m in members)))
            idx += len(pg['params'])
        if self._pending is not None:
            sd, self._pending = self._pending, None
            self.load_state_dict(sd)

    def _walk(self):
        idx = 0
        for members in (self._groups or []):
            for (params, arena, counter) in members:
                for p in params:
                    yield idx, p, arena, counter, _locate(p, arena)
                    idx += 1

    def state_dict(self):
        state = OrderedDict()
        for idx, p, arena, counter, loc in self._walk():
            if loc is None or counter is None:
                continue
            step = int(counter.item())
            if step == 0:          # torch.optim.Adam has no state before the first step
                continue
            off, n = loc
            state[idx] = {'step': torch.tensor(float(step)),
                          'exp_avg': arena.m[off:off + n].view(p.shape).detach().cpu().clone(),
                          'exp_avg_sq': arena.v[off:off + n].view(p.shape).detach().cpu().clone()}
        return {'state': state, 'param_groups': [dict(pg) for pg in self.param_groups]}

    def load_state_dict(self, sd):
        if self._groups is None:       # the engine is compiled with the first batch: applied by bind()
            self._pending = sd
            return
        if len(sd.get('param_groups', [])) != len(self.param_groups):
            raise ValueError('loaded state dict has a different number of parameter groups')
        for mine, theirs in zip(self.param_groups, sd['param_groups']):
            if len(mine['params']) != len(theirs['params']):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k, v in theirs.items():
                if k != 'params':
                    mine[k] = v
        ids = [i for pg in sd['param_groups'] for i in pg['params']]       # saved id of the k-th parameter
        state = sd.get('state', {})
        steps = {}
        for k, (idx, p, arena, counter, loc) in enumerate(self._walk()):
            st = state.get(ids[k])
            if st is None or loc is None:
                continue
            off, n = loc
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError('optimizer state of parameter %d has shape %s, expected %s' % (idx, tuple(st['exp_avg'].shape), tuple(p.shape)))
            arena.m[off:off + n].copy_(st['exp_avg'].reshape(-1))
            arena.v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            if counter is not None:
                steps[id(counter)] = (counter, max(steps.get(id(counter), (None, 0))[1], int(float(st['step']))))
        for counter, step in steps.values():
            counter.fill_(step)

    def zero_grad(self, set_to_none=False):
        pass   # the step zeroes its gradient arenas itself


def carry_engine_state(old, new):
    """Copy everything that persists between steps from `old` to `new` (same architectures and hyper-parameters, any batch
    shape).  Both engines expose ``persistent_state()`` -> {name: tensor}; shapes that differ (image pools at another
    resolution) are skipped."""
    src, dst = old.persistent_state(), new.persistent_state()
    for name, t in dst.items():
        s = src.get(name)
        if s is not None and s.shape == t.shape:
            t.copy_(s)
    for name in ('pool_A', 'pool_B'):          # CycleGAN history buffers: host-side fill count
        po, pn = getattr(old, name, None), getattr(new, name, None)
        if po is not None and pn is not None and getattr(po, 'pool_size', 0) > 0 and po.images.shape == pn.images.shape:
            pn.num = po.num
    new.after_state_load()


def engine_state_from_nets(nets, extra):
    """Helper for persistent_state(): arenas (p, m, v), buffer arenas and adaptor arenas of `nets` = {name: net}."""
    out = OrderedDict()
    for name, net in nets.items():
        if net is None:
            continue
        for which in ('p', 'm', 'v'):
            t = getattr(net.arena, which, None)
            if t is not None:
                out['%s.%s' % (name, which)] = t
        bufs = getattr(net, 'bufs', None)
        if bufs is not None and bufs.p is not None:
            out['%s.bufs' % name] = bufs.p
    out.update(extra)
    return out


class EngineOwner:
    """Mixin of the distiller / model mirrors: step engines are compiled per batch shape, cached (the two most recent
    shapes: the full batch and the partial batch that ends an epoch) and handed the complete training state whenever the
    shape changes.  Subclasses provide ``_make_engine(B, H, W)`` (compile only) and ``_bind_engine(engine)`` (point the
    module mirrors and the optimizer objects at the engine's arenas)."""

    def _ensure_engine(self, B, H, W):
        key = (B, H, W)
        old = getattr(self, 'engine', None)
        if old is not None and (old.B, old.H, old.W) == key:
            return
        cache = self.__dict__.setdefault('_engine_cache', OrderedDict())
        eng = cache.get(key)
        if eng is None:
            eng = self._make_engine(B, H, W)
        # binding copies the modules' current weights (a checkpoint, or the previous engine's arena they alias) into the
        # engine's arenas and re-points the modules there
        self._bind_engine(eng)
        if old is not None:
            carry_engine_state(old, eng)       # Adam moments, step counters, learning rates, running statistics, pools
        else:
            eng.set_lr(*[o.param_groups[0]['lr'] for o in self.optimizers])     # restored / decayed learning rates
        cache[key] = eng
        cache.move_to_end(key)
        while len(cache) > 2:
            cache.popitem(last=False)
        self.engine = eng
