"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of the timm training helpers the reference train step calls
(/root/reference/src/pixparse/task/task_cruller_pretrain.py:196-223, 206, 259-278):

    timm.optim.create_optimizer_v2, timm.scheduler.create_scheduler_v2,
    timm.utils.NativeScaler, timm.utils.dispatch_clip_grad

timm is an unpinned third-party dependency (pyproject.toml:32-37) that is not installed here; the
algorithms below restate timm 0.9.x. Parity unpinned: the reference ships no
tests or golden vectors for these helpers.
"""
import math

import torch


# --------------------------------------------------------------------------------------------------
# optimizer factory  (timm/optim/optim_factory.py)
# --------------------------------------------------------------------------------------------------
def _layer_map(model, layers_per_group=12):
    """timm's fallback layer map for models without a ``group_matcher``.

    ``head_prefix = model.pretrained_cfg.get('classifier')``; a missing prefix makes ``_in_head`` return
    True for every name, so a plain nn.Module such as Cruller (no pretrained_cfg) ends up with ALL
    parameters in the single "head" group -> one layer, lr_scale 1.0. A model that does carry a
    classifier prefix gets its remaining parameters chunked 12 names at a time.
    """
    head_prefix = getattr(model, 'pretrained_cfg', {}).get('classifier', None)

    def in_head(n):
        if not head_prefix:
            return True
        if isinstance(head_prefix, (tuple, list)):
            return any(n.startswith(h) for h in head_prefix)
        return n.startswith(head_prefix)

    names_trunk, names_head = [], []
    for n, _ in model.named_parameters():
        (names_head if in_head(n) else names_trunk).append(n)
    groups = [names_trunk[i:i + layers_per_group] for i in range(0, len(names_trunk), layers_per_group)]
    layer_map = {n: i for i, g in enumerate(groups) for n in g}
    layer_map.update({n: len(groups) for n in names_head})
    return layer_map


def param_groups_layer_decay(model, weight_decay=0.05, no_weight_decay_list=(), layer_decay=.75):
    no_weight_decay_list = set(no_weight_decay_list)
    layer_map = _layer_map(model)
    num_layers = max(layer_map.values()) + 1
    layer_max = num_layers - 1
    layer_scales = [layer_decay ** (layer_max - i) for i in range(num_layers)]
    groups = {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if param.ndim == 1 or name in no_weight_decay_list:
            g_decay, this_decay = "no_decay", 0.
        else:
            g_decay, this_decay = "decay", weight_decay
        layer_id = layer_map.get(name, layer_max)
        key = "layer_%d_%s" % (layer_id, g_decay)
        if key not in groups:
            groups[key] = {"lr_scale": layer_scales[layer_id], "weight_decay": this_decay, "params": [],
                           "param_names": []}
        groups[key]["params"].append(param)
        groups[key]["param_names"].append(name)
    return list(groups.values())


def create_optimizer_v2(model_or_params, opt='adamw', lr=None, weight_decay=0., momentum=0.9,
                        filter_bias_and_bn=True, layer_decay=None, **kwargs):
    """Only the branches pixparse can reach: opt='adamw', weight_decay never forwarded (=0)."""
    assert opt.lower() == 'adamw', "pixparse's OptimizationCfg default (framework/config.py:8)"
    if isinstance(model_or_params, torch.nn.Module):
        if layer_decay is not None:
            parameters = param_groups_layer_decay(model_or_params, weight_decay=weight_decay, layer_decay=layer_decay)
            for g in parameters:
                g.pop("param_names")
            weight_decay = 0.
        else:
            # weight_decay == 0 -> no decay/no-decay split
            parameters = model_or_params.parameters()
    else:
        parameters = model_or_params
    opt_args = dict(weight_decay=weight_decay, **kwargs)
    if lr is not None:
        opt_args['lr'] = lr
    opt_args.pop('momentum', None)
    return torch.optim.AdamW(parameters, **opt_args)


# --------------------------------------------------------------------------------------------------
# scheduler  (timm/scheduler/{scheduler_factory,cosine_lr,scheduler}.py)
# --------------------------------------------------------------------------------------------------
class CosineLRScheduler:
    def __init__(self, optimizer, t_initial, lr_min=0., warmup_t=0, warmup_lr_init=0., warmup_prefix=False,
                 cycle_limit=1, t_in_epochs=True):
        self.optimizer = optimizer
        self.t_initial = t_initial
        self.lr_min = lr_min
        self.warmup_t = warmup_t
        self.warmup_lr_init = warmup_lr_init
        self.warmup_prefix = warmup_prefix
        self.cycle_limit = cycle_limit
        self.t_in_epochs = t_in_epochs
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self.base_values = [g['initial_lr'] for g in optimizer.param_groups]
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self._update_groups(self.warmup_lr_init)
        else:
            self.warmup_steps = [1 for _ in self.base_values]
            self._update_groups(self.base_values)

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        if self.warmup_prefix:
            t = t - self.warmup_t
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        if i < self.cycle_limit:
            return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t_curr / self.t_initial))
                    for v in self.base_values]
        return [self.lr_min for _ in self.base_values]

    def _update_groups(self, values):
        if not isinstance(values, (list, tuple)):
            values = [values] * len(self.optimizer.param_groups)
        for g, v in zip(self.optimizer.param_groups, values):
            g['lr'] = v * g['lr_scale'] if 'lr_scale' in g else v

    def step(self, epoch, metric=None):
        if self.t_in_epochs:
            self._update_groups(self._get_lr(epoch))

    def step_update(self, num_updates, metric=None):
        if not self.t_in_epochs:
            self._update_groups(self._get_lr(num_updates))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != 'optimizer'}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def create_scheduler_v2(optimizer, sched='cosine', num_epochs=300, warmup_lr=1e-5, warmup_epochs=0, min_lr=0.,
                        warmup_prefix=False, step_on_epochs=True, updates_per_epoch=0, **kwargs):
    assert sched == 'cosine', "pixparse's OptimizationCfg default (framework/config.py:9)"
    t_initial, warmup_t = num_epochs, warmup_epochs
    if not step_on_epochs:
        assert updates_per_epoch, 'updates_per_epoch must be set to number of dataloader batches'
        t_initial *= updates_per_epoch
        warmup_t *= updates_per_epoch
    s = CosineLRScheduler(optimizer, t_initial=t_initial, lr_min=min_lr, warmup_t=warmup_t,
                          warmup_lr_init=warmup_lr, warmup_prefix=warmup_prefix, cycle_limit=1,
                          t_in_epochs=step_on_epochs)
    return s, num_epochs


# --------------------------------------------------------------------------------------------------
# clipping / AMP scaler  (timm/utils/{clip_grad,cuda}.py)
# --------------------------------------------------------------------------------------------------
def dispatch_clip_grad(parameters, value, mode='norm', norm_type=2.0):
    if mode == 'norm':
        torch.nn.utils.clip_grad_norm_(parameters, value, norm_type=norm_type)
    elif mode == 'value':
        torch.nn.utils.clip_grad_value_(parameters, value)
    else:
        raise AssertionError(f"Unknown clip mode ({mode}).")


class NativeScaler:
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scaler = torch.amp.GradScaler("cuda", enabled=torch.cuda.is_available())

    def __call__(self, loss, optimizer, clip_grad=None, clip_mode='norm', parameters=None, create_graph=False,
                 need_update=True):
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if need_update:
            if clip_grad is not None:
                assert parameters is not None
                self._scaler.unscale_(optimizer)
                dispatch_clip_grad(parameters, clip_grad, mode=clip_mode)
            self._scaler.step(optimizer)
            self._scaler.update()

    def state_dict(self):
        return self._scaler.state_dict()

    def load_state_dict(self, state_dict):
        self._scaler.load_state_dict(state_dict)
