"""Host-side mirror of the parts of ``pixparse.framework`` the hot path's callers touch.

    DeviceEnv / world_info_from_env     /root/reference/src/pixparse/framework/device.py:21-45, 56-166
    OptimizationCfg / TaskTrainCfg /... /root/reference/src/pixparse/framework/config.py:5-39
    TaskTrain / TaskEval                /root/reference/src/pixparse/framework/task.py:9-90
    train_one_interval / evaluate       /root/reference/src/pixparse/framework/train.py:5-14, eval.py:4-24

Same names, fields and call sequence; logging (Monitor) is duck-typed and optional.
"""
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def world_info_from_env():
    local_rank = 0
    for v in ("LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "OMPI_COMM_WORLD_LOCAL_RANK"):
        if v in os.environ:
            local_rank = int(os.environ[v])
            break
    global_rank = 0
    for v in ("RANK", "PMI_RANK", "SLURM_PROCID", "OMPI_COMM_WORLD_RANK"):
        if v in os.environ:
            global_rank = int(os.environ[v])
            break
    world_size = 1
    for v in ("WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", "OMPI_COMM_WORLD_SIZE"):
        if v in os.environ:
            world_size = int(os.environ[v])
            break
    return local_rank, global_rank, world_size


class DeviceEnv:
    """One process per GPU; NCCL over NVLink for the gradient exchange (device.py:104-151).

    ``backend`` may be 'gloo' with ``device_type='cpu'`` for host-side tests of the multi-rank logic only --
    the compute path itself refuses to run anywhere but on a B200."""

    def __init__(self, device_type: Optional[str] = None, device_index: Optional[int] = None, backend: str = "nccl",
                 dist_url: str = "env://"):
        device_type = device_type or "cuda"
        if device_type == "cuda":
            assert torch.cuda.device_count(), "pixparse_b200 needs a CUDA device (B200)"
        local_rank, global_rank, world_size = world_info_from_env()
        if world_size > 1:
            assert device_index is None
            if not dist.is_initialized():
                if "SLURM_PROCID" in os.environ:
                    dist.init_process_group(backend=backend, init_method=dist_url, world_size=world_size,
                                            rank=global_rank)
                else:
                    dist.init_process_group(backend=backend, init_method=dist_url)
            self.world_size = dist.get_world_size()
            self.global_rank = dist.get_rank()
            self.local_rank = int(local_rank)
            if device_type == "cuda":
                self.device = torch.device("cuda:%d" % self.local_rank)
                torch.cuda.set_device(self.local_rank)
            else:
                self.device = torch.device("cpu")
        else:
            if device_type == "cuda":
                self.device = torch.device("cuda" if device_index is None else f"cuda:{device_index}")
            else:
                self.device = torch.device("cpu")
            self.local_rank, self.world_size, self.global_rank = 0, 1, 0

    def is_global_primary(self):
        return self.global_rank == 0

    def is_local_primary(self):
        return self.local_rank == 0

    def is_primary(self, local=False):
        return self.is_local_primary() if local else self.is_global_primary()

    def broadcast_object(self, obj, src=0):
        objects = [obj] if self.global_rank == src else [None]
        dist.broadcast_object_list(objects, src=src)
        return objects[0]

    def all_gather_object(self, obj, dst=0):
        objects = [None for _ in range(self.world_size)]
        dist.all_gather_object(objects, obj)
        return objects


@dataclass
class OptimizationCfg:
    optimizer: str = 'adamw'
    scheduler: str = 'cosine'
    learning_rate: float = 5e-4
    warmup_learning_rate: float = 0.
    weight_decay: float = .02        # never forwarded to the optimizer by the reference tasks
    eps: float = 1e-6
    clip_grad_value: Optional[float] = None
    clip_grad_mode: Optional[str] = None
    grad_accum_steps: int = 1
    momentum: Optional[float] = None
    betas: Optional[Tuple[float, float]] = None
    layer_decay: Optional[float] = None


@dataclass
class TaskTrainCfg:
    num_intervals: int = 100
    num_warmup_intervals: int = 5
    eval_frequency: int = 1000
    opt: OptimizationCfg = field(default_factory=OptimizationCfg)
    dtype: Optional[str] = None
    amp: bool = True
    model_name: str = ""


@dataclass
class TaskEvalCfg:
    dtype: Optional[str] = None
    amp: bool = True
    model_name: str = ""
    model_state_dict: dict = field(default_factory=dict)


class Task:
    def __init__(self, device_env, monitor=None):
        self.device_env = device_env
        self.monitor = monitor


class TaskEval(Task):
    def __init__(self, cfg: TaskEvalCfg, device_env, monitor=None):
        super().__init__(device_env=device_env, monitor=monitor)

    def collate_fn(self, batch):
        pass

    def setup(self, *args, **kwargs):
        pass

    def prepare_for_evaluation(self, *args, **kwargs):
        pass

    def step(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        pass

    def end(self):
        pass


class TaskTrain(Task):
    def __init__(self, cfg: TaskTrainCfg, device_env, monitor=None):
        super().__init__(device_env=device_env, monitor=monitor)
        self.num_intervals = cfg.num_intervals
        self.num_warmup_intervals = cfg.num_warmup_intervals
        self.eval_frequency = cfg.eval_frequency
        self.num_steps_per_interval = None
        self.start_interval = 0
        self.step = 0
        self.batch_idx = 0
        self.interval_idx = 0
        self.interval_batch_idx = 0
        self.optimizer = None
        self.scheduler = None
        self.scaler = None
        self.autocast = None

    def collate_fn(self, batch):
        pass

    def train_setup(self, *args, **kwargs):
        pass

    def train_interval_start(self):
        pass

    def train_interval_end(self):
        pass

    def train_step(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        pass

    def eval_step(self, sample: Dict[str, Any]) -> Dict[str, Any]:
        pass

    def get_current_lr(self):
        lrl = [param_group['lr'] for param_group in self.optimizer.param_groups]
        return sum(lrl) / len(lrl)


def train_one_interval(task: TaskTrain, loader):
    task.train_interval_start()
    for i, sample in enumerate(loader.loader):
        task.train_step(sample)
    task.train_interval_end()


def evaluate(task: TaskEval, loaders):
    metrics = dict()
    authorized_loaders = task.prepare_for_evaluation(loaders)
    for key, loader in authorized_loaders.items():
        metrics[key] = dict()
        for index_batch, sample in enumerate(loader.loader):
            metrics[key][index_batch] = task.step(sample)
        if hasattr(task, 'average_metrics'):
            averaged_metrics = task.average_metrics(metrics[key])
            metrics[key] = {}
            metrics[key]["average"] = averaged_metrics
    return metrics
