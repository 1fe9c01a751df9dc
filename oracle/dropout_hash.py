"""TEST INFRASTRUCTURE: numpy restatement of the library's counter-based dropout hash
(coper_b200/csrc/common.cuh: hash32 / keep_threshold) so that golden vectors generated from the reference
can use exactly the keep-masks the CUDA kernels draw for a given (seed, salt)."""
import numpy as np

SALT_FEATURE_MAP = 1 << 40
SALT_OUTPUT = 2 << 40
SALT_CTX = 3 << 40
M64 = (1 << 64) - 1


def hash32(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & M64) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint64)


def keep_threshold(keep: float) -> int:
    t = float(np.float64(np.float32(keep)) * 4294967296.0)
    if t >= 4294967295.0:
        return 0xFFFFFFFF
    return max(0, int(t))


def keep_mask(n: int, keep: float, seed_dev: int, salt: int) -> np.ndarray:
    """Boolean keep-mask of n elements; keep >= 1 -> all True."""
    if np.float32(keep) >= 1.0:
        return np.ones(n, bool)
    return hash32(seed_dev + salt, np.arange(n, dtype=np.uint64)) < np.uint64(keep_threshold(keep))


def ctx_salt(net_id: int, layer: int) -> int:
    return SALT_CTX + ((net_id * 64 + layer) << 32)


def model_seed0(seed: int) -> int:
    """Initial value of ConvE.seed_dev for ConvE(seed=seed); step k (0-based) draws with seed0 + k + 1."""
    return seed * 1000003 + 12345
