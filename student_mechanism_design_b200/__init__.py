"""B200-native batched Scotland Yard environment (drop-in for the reference's
src/environment hot path).  See DESIGN.md / INTEGRATION.md at the repo root."""
from ._cabi import LIB_PATH, SyError, load_library  # noqa: F401
from .graphs import GraphSpec, generate_connected_graph, generate_graph_pool, pack_csr  # noqa: F401
from .env import (  # noqa: F401
    DEFAULT_REWARD_WEIGHTS,
    REWARD_WEIGHT_NAMES,
    BatchedScotlandYardEnv,
    dense_action_mask,
    numpy_reward_tables,
)

__all__ = [
    "BatchedScotlandYardEnv", "GraphSpec", "generate_connected_graph", "generate_graph_pool", "pack_csr",
    "dense_action_mask", "numpy_reward_tables", "DEFAULT_REWARD_WEIGHTS", "REWARD_WEIGHT_NAMES", "SyError",
    "load_library", "LIB_PATH",
]
