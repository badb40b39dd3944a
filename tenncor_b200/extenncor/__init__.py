"""Trainers that drive the evaluation path from python (the reference's `extenncor/`): a replay-buffer DQN environment with
checkpoint / resume, config C5 end to end."""
from . import dqn_trainer, trainer_cache  # noqa: F401
from .dqn_trainer import DQNEnv, get_dqnerror, get_dqnupdate  # noqa: F401
from .trainer_cache import EnvManager, SessionCache  # noqa: F401
