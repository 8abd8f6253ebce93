"""Minimal stand-in for hydra (only hydra.utils.instantiate/get_class), golden generation only."""
from . import utils  # noqa: F401
