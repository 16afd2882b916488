"""carle_b200 — B200-native implementation of the CARLE environment step.

Drop-in for the hot path of riveSunder/carle: ``carle_b200.CARLE`` mirrors
``carle.env.CARLE`` and ``carle_b200.mcl`` mirrors the grid-reduction reward wrappers
of ``carle.mcl``.  All compute runs in hand-written sm_100a CUDA kernels behind the C
ABI of ``include/carle_b200.h`` (``carle_b200/lib/libcarle_b200.so``); there is no CPU
or torch-op fallback."""
from .env import CARLE, PackedAction, RandomAction                                # noqa: F401
from .mcl import (Motivator, ParsimonyBonus, CornerBonus, SpeedDetector,  # noqa: F401
                  PufferDetector, MorphoBonus)
from .agents import RandomAgent, DeviceRandomAgent                  # noqa: F401
from .rollout import RolloutPlan, host_rollout, train_loop         # noqa: F401
from .sharding import ShardedCARLE, ShardedSpeedDetector, shard_range   # noqa: F401

__all__ = ["CARLE", "PackedAction", "DeviceRandomAgent", "Motivator", "ParsimonyBonus", "CornerBonus", "SpeedDetector",
           "PufferDetector", "MorphoBonus", "RandomAgent", "RolloutPlan", "host_rollout", "train_loop", "ShardedCARLE",
           "ShardedSpeedDetector", "shard_range"]
