"""Fused clip + AdamW over the engine's flat arenas (one pass over HBM per update).

Stands in for ``timm.optim.create_optimizer_v2(model, 'adamw', lr, eps, betas, layer_decay)`` ->
``torch.optim.AdamW`` and ``timm.utils.dispatch_clip_grad(..., mode='norm')``
(/root/reference/src/pixparse/task/task_cruller_pretrain.py:196-203, 259-278). ``weight_decay`` is accepted but, as
in the reference (which never forwards it), defaults to 0. It is a ``torch.optim.Optimizer`` so that
schedulers, ``param_groups`` consumers (``get_current_lr``) and ``state_dict()`` keep working.
"""
import numpy as np
import torch

from . import ops

_SEG_DTYPE = np.dtype([("end", "<i8"), ("lr_scale", "<f4"), ("weight_decay", "<f4")])


def layer_decay_groups(model, layer_decay, weight_decay=0.0, layers_per_group=12):
    """timm ``param_groups_layer_decay`` for a model without ``group_matcher``: the fallback layer map only chunks
    parameters whose name does NOT start with ``pretrained_cfg['classifier']``; a plain nn.Module such as Cruller has
    no pretrained_cfg, every parameter lands in the single head group and lr_scale is 1.0 for all of them. Groups
    are still split into 1-D ("no_decay") and other ("decay") tensors, both with weight_decay 0 here."""
    head_prefix = getattr(model, 'pretrained_cfg', {}).get('classifier', None)
    names = [n for n, _ in model.named_parameters()]

    def in_head(n):
        if not head_prefix:
            return True
        if isinstance(head_prefix, (tuple, list)):
            return any(n.startswith(h) for h in head_prefix)
        return n.startswith(head_prefix)

    trunk = [n for n in names if not in_head(n)]
    chunks = [trunk[i:i + layers_per_group] for i in range(0, len(trunk), layers_per_group)]
    layer_of = {n: i for i, c in enumerate(chunks) for n in c}
    n_layers = len(chunks) + 1
    scales = [layer_decay ** (n_layers - 1 - i) for i in range(n_layers)]
    groups = {}
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        lid = layer_of.get(n, n_layers - 1)
        kind = "no_decay" if p.ndim == 1 else "decay"
        g = groups.setdefault((lid, kind), {"params": [], "lr_scale": scales[lid],
                                            "weight_decay": 0.0 if kind == "no_decay" else weight_decay})
        g["params"].append(p)
    return list(groups.values())


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, engine, lr=5e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, layer_decay=None):
        arena = engine.ensure_bound()
        if layer_decay is not None:
            params = layer_decay_groups(model, layer_decay, weight_decay)
            weight_decay = 0.0
        else:
            params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.engine = engine
        self.arena = arena
        self.exp_avg = torch.zeros_like(arena.p32)
        self.exp_avg_sq = torch.zeros_like(arena.p32)
        self._step = 0      # update ATTEMPTS (host side); the number really applied lives in norm_stats[3]
        # [sum of squares, global L2 norm, clip coefficient, applied updates] -- written by the grad-norm kernel, read by
        # the AdamW kernel; nothing here ever comes back to the host inside a step
        self.norm_stats = torch.zeros(4, device=arena.device, dtype=torch.float32)
        self._segs_key = None
        self._segs = None
        self._nseg = 0
        for key, p in zip(arena.keys, arena.params):
            o, n, shape = arena.index[key]
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.exp_avg[o:o + n].view(shape),
                             "exp_avg_sq": self.exp_avg_sq[o:o + n].view(shape)}

    def _segments(self):
        """Per-tensor {end, lr_scale, weight_decay} table in arena order; consecutive equal entries are merged."""
        g0 = self.param_groups[0]
        base_lr = g0['lr'] / g0.get('lr_scale', 1.0) if g0.get('lr_scale', 1.0) != 0 else g0['lr']
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            for p in g['params']:
                group_of[id(p)] = gi
        key = (base_lr == 0.0,) + tuple((g['lr'] / base_lr if base_lr != 0 else 1.0, g['weight_decay'])
                                        for g in self.param_groups)
        if key != self._segs_key:
            rows = []
            ar = self.arena
            for k, p in zip(ar.keys, ar.params):
                o, n, _ = ar.index[k]
                g = self.param_groups[group_of[id(p)]]
                scale = g['lr'] / base_lr if base_lr != 0 else 1.0
                end = o + (n + 63) // 64 * 64
                if rows and rows[-1][1] == scale and rows[-1][2] == g['weight_decay']:
                    rows[-1][0] = end
                else:
                    rows.append([end, scale, g['weight_decay']])
            rows[-1][0] = ar.total
            arr = np.zeros(len(rows), dtype=_SEG_DTYPE)
            for i, (end, s, wd) in enumerate(rows):
                arr[i] = (end, s, wd)
            self._segs = torch.from_numpy(arr.view(np.uint8).copy()).to(ar.device)
            self._nseg = len(rows)
            self._segs_key = key
        return base_lr

    @torch.no_grad()
    def step(self, closure=None, clip_grad_norm=None, grad_scale=1.0):
        """One update over the whole arena. clip_grad_norm: max global L2 norm (timm clip mode 'norm') or None.

        The global gradient norm is always computed: a non-finite norm skips the parameter / moment update exactly as
        timm's NativeScaler (GradScaler.step) does in the reference (task_cruller_pretrain.py:259-268) -- the skipped
        step does not advance AdamW's bias-correction step either -- and the decision stays on the device (no host
        sync). Gradients are zeroed in the same pass, also on a skipped step (the reference calls
        optimizer.zero_grad() right after, :295)."""
        assert closure is None
        if not self.arena.intact():
            raise RuntimeError("parameters were re-allocated after the optimizer was created; rebuild the optimizer")
        base_lr = self._segments()
        g0 = self.param_groups[0]
        self._step += 1
        stats = ops.grad_norm(self.arena.g32, max_norm=float(clip_grad_norm or 0.0), pre_scale=float(grad_scale),
                              out=self.norm_stats)
        ops.adamw_step(self.arena.p32, self.arena.g32, self.exp_avg, self.exp_avg_sq, self.arena.p16, self._segs,
                       self._nseg, lr=base_lr, beta1=g0['betas'][0], beta2=g0['betas'][1], eps=g0['eps'],
                       norm_stats=stats, grad_scale=grad_scale, zero_grad=True)
        self.engine.mark_params_updated_by_kernel(True)

    def applied_steps(self):
        """Number of updates really applied (device -> host read): attempts minus the steps skipped for non-finite
        gradients. This is what torch.optim.AdamW keeps in state['step']."""
        return int(self.norm_stats[3].item())

    def state_dict(self):
        n = float(self.applied_steps())
        for st in self.state.values():
            st["step"] = torch.tensor(n)
        return super().state_dict()

    def zero_grad(self, set_to_none=False):
        # gradients live in the arena and were zeroed by step(); keep param.grad attached to the arena views
        self.arena.attach_grads()

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.AdamW-style state_dict; moments are copied INTO the arena views."""
        sd_state = state_dict["state"]
        ids = [i for g in state_dict["param_groups"] for i in g["params"]]
        params = [p for g in self.param_groups for p in g["params"]]
        with torch.no_grad():
            for i, p in zip(ids, params):
                if i in sd_state:
                    self.state[p]["exp_avg"].copy_(sd_state[i]["exp_avg"])
                    self.state[p]["exp_avg_sq"].copy_(sd_state[i]["exp_avg_sq"])
                    self._step = max(self._step, int(sd_state[i]["step"]))
            self.norm_stats[3] = float(self._step)
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
        self._segs_key = None
