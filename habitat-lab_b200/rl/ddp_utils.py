"""Process-group bootstrap and resume-state helpers with the reference's names and env-var contract
(habitat-baselines/habitat_baselines/rl/ddppo/ddp_utils.py:54-352): LOCAL_RANK / RANK / WORLD_SIZE or the
SLURM_* variables, MAIN_ADDR / MAIN_PORT (+ MAIN_PORT_RANGE offset by SLURM_JOB_ID), an explicit TCPStore that
doubles as DD-PPO's preemption counter, `rank0_only`, resume-state save/load.  Host plumbing only."""
from __future__ import annotations

import functools
import os
import socket
from typing import Any, Callable, Optional, Tuple

import torch
import torch.distributed as distrib

DEFAULT_PORT = 8738
DEFAULT_PORT_RANGE = 127
DEFAULT_MAIN_ADDR = "127.0.0.1"
SLURM_JOBID = os.environ.get("SLURM_JOB_ID", None)
RESUME_STATE_BASE_NAME = ".habitat-resume-state"


def is_slurm_job() -> bool:
    return SLURM_JOBID is not None


def is_slurm_batch_job() -> bool:
    return is_slurm_job() and os.environ.get("SLURM_JOB_NAME", None) not in (None, "bash", "zsh", "fish", "tcsh", "sh", "interactive")


def resume_state_filename(config, filename_key: str = "") -> str:
    fname = RESUME_STATE_BASE_NAME
    if is_slurm_job() and SLURM_JOBID is not None:
        fname += "-{}".format(SLURM_JOBID)
    if filename_key:
        fname += "-" + filename_key
    return os.path.join(getattr(config.habitat_baselines, "checkpoint_folder", "."), fname + ".pth")


def get_distrib_size() -> Tuple[int, int, int]:
    """(local_rank, world_rank, world_size): torch.distributed.launch / torchrun vars first, then SLURM's."""
    if os.environ.get("LOCAL_RANK", None) is not None:
        return int(os.environ["LOCAL_RANK"]), int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if os.environ.get("SLURM_JOBID", None) is not None:
        return int(os.environ["SLURM_LOCALID"]), int(os.environ["SLURM_PROCID"]), int(os.environ["SLURM_NTASKS"])
    return 0, 0, 1


def get_main_addr() -> str:
    return os.environ.get("MAIN_ADDR", os.environ.get("MASTER_ADDR", DEFAULT_MAIN_ADDR))


def find_free_port() -> int:
    with socket.socket() as s:
        s.bind(("", 0))
        return s.getsockname()[1]


def init_distrib_slurm(backend: str = "nccl") -> Tuple[int, torch.distributed.TCPStore]:
    """TCPStore on rank 0 + init_process_group(store=...); returns (local_rank, tcp_store) like the reference."""
    assert torch.distributed.is_available(), "torch.distributed must be available"
    local_rank, world_rank, world_size = get_distrib_size()
    main_port = int(os.environ.get("MAIN_PORT", os.environ.get("MASTER_PORT", DEFAULT_PORT)))
    if SLURM_JOBID is not None:
        main_port += int(SLURM_JOBID) % int(os.environ.get("MAIN_PORT_RANGE", DEFAULT_PORT_RANGE))
    if backend.lower() == "nccl" and torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    tcp_store = distrib.TCPStore(get_main_addr(), main_port, world_size, world_rank == 0)
    distrib.init_process_group(backend.lower(), store=tcp_store, rank=world_rank, world_size=world_size)
    return local_rank, tcp_store


def rank0_only(fn: Optional[Callable] = None):
    """Decorator / predicate: run only on world rank 0 (ddp_utils.py:312-352)."""
    if fn is None:
        return (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0

    @functools.wraps(fn)
    def _wrapper(*args, **kwargs):
        if rank0_only():
            return fn(*args, **kwargs)
        return None

    return _wrapper


@rank0_only
def save_resume_state(state: Any, filename_or_config, filename_key: str = ""):
    filename = filename_or_config if isinstance(filename_or_config, str) else resume_state_filename(filename_or_config, filename_key)
    torch.save(state, filename)


def load_resume_state(filename_or_config, filename_key: str = "") -> Optional[Any]:
    filename = filename_or_config if isinstance(filename_or_config, str) else resume_state_filename(filename_or_config, filename_key)
    if not os.path.exists(filename):
        return None
    return torch.load(filename, map_location="cpu", weights_only=False)
