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


# ---- on-device label sampling (coper_b200/csrc/sampled.cu: prp / sample_labels_kernel), restated ------------------
SALT_SAMPLE = 4 << 40


SAMPLE_SMALL = 1024


def prp(x: np.ndarray, n: int, key: int) -> np.ndarray:
    """Keyed pseudorandom permutation of [0, n): 6-round balanced Feistel over the next even power of two, cycle-walked
    back into [0, n).  x: uint array with values < n."""
    bits = 2
    while bits < 32 and (1 << bits) < n:
        bits += 2
    half = bits >> 1
    mask = np.uint64((1 << half) - 1)
    x = x.astype(np.uint64).copy()
    todo = np.ones(x.shape, bool)
    while todo.any():
        l, r = x[todo] >> np.uint64(half), x[todo] & mask
        for rnd in range(6):
            t = l ^ (hash32(key, (np.uint64(rnd) << np.uint64(32)) | r) & mask)
            l, r = r, t
        x[todo] = (l << np.uint64(half)) | r
        todo = x >= np.uint64(n)
    return x


def random_prefix(n: int, count: int, key: int) -> np.ndarray:
    """First ``count`` elements of the keyed random permutation of [0, n) the kernel uses: the order of the n hashes
    (ties by index) for n <= SAMPLE_SMALL, the Feistel permutation otherwise."""
    if n <= SAMPLE_SMALL:
        h = hash32(key, np.arange(n, dtype=np.uint64))
        return np.lexsort((np.arange(n), h))[:count].astype(np.int64)
    return prp(np.arange(count, dtype=np.uint64), n, key).astype(np.int64)


def sample_labels(rowptr, col, N: int, L: int, n_pos_needed: int, seed_dev: int, salt: int = SALT_SAMPLE):
    """data.py:228-277 as the CUDA kernel draws it for (seed_dev, salt).  Returns lookup int32 [B, L], labels f32."""
    B = len(rowptr) - 1
    lookup = np.zeros((B, L), np.int32)
    labels = np.zeros((B, L), np.float32)
    seed = (seed_dev + salt) & M64
    for b in range(B):
        pos = np.asarray(col[rowptr[b]:rowptr[b + 1]])
        P = len(pos)
        n_pos = P
        if P > n_pos_needed:
            n_pos = L - min(L - n_pos_needed, N)
        n_pos = min(n_pos, P, L)
        k = []
        for j in (2 * b, 2 * b + 1):
            hi = int(hash32(seed, np.array([j], np.uint64))[0])
            lo = int(hash32(seed ^ 0x5bd1e995, np.array([j], np.uint64))[0])
            k.append((hi << 32) | lo)
        if n_pos:
            lookup[b, :n_pos] = pos[random_prefix(P, n_pos, k[0])]
            labels[b, :n_pos] = 1.0
        neg = random_prefix(N, L - n_pos, k[1])
        lookup[b, n_pos:] = neg
        labels[b, n_pos:] = np.isin(neg, pos).astype(np.float32)
    return lookup, labels
