"""B200-native KNARPE attention + closed-loop rollout hot path of TrafficBots V1.5."""
from . import config, params, synth  # noqa: F401
