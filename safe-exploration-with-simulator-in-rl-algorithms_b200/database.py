"""Trajectory store.

File format = the reference's: one `.npz` with `policies[K, n-1, 2n+2]` and
`trajectories[K, H, 2n+2]` (ars/database.py:16-37), so stores written by either side load on the
other.  The in-memory side is its own design: records arrive in bulk from the rollout kernel's
time-major trajectory output (`extend_from_rollout`), are kept as per-record views of those
blocks instead of being copied one Python list at a time, and `as_arrays()` hands the estimator
objective contiguous batches.  The reference's attribute surface (`.policies`, `.trajectories`,
`.size`, `load`, `add_trajectory`, `save`) is preserved for ARSAgent / Estimator.
"""
import numpy as np

_KEYS = ("policies", "trajectories")


class Database:
    def __init__(self):
        self._records = []  # (trajectory[H, 2n+2], policy[n-1, 2n+2]) pairs, possibly views of a block

    # ---- reference surface ----
    @property
    def size(self):
        return len(self._records)

    @property
    def trajectories(self):
        return [t for t, _ in self._records]

    @property
    def policies(self):
        return [p for _, p in self._records]

    def add_trajectory(self, trajectory, policy):
        self._records.append((trajectory, policy))

    def load(self, path):
        with np.load(path) as store:
            missing = [k for k in _KEYS if k not in store.files]
            if missing:
                raise AssertionError("%s has no array %s (expected %s)" % (path, missing, list(_KEYS)))
            pol, traj = store["policies"], store["trajectories"]
        if pol.shape[0] != traj.shape[0]:
            raise AssertionError("%d policies for %d trajectories in %s" % (pol.shape[0], traj.shape[0], path))
        self._records.extend(zip(traj, pol))

    def save(self, path):
        traj, pol = self.as_arrays()
        np.savez(path, policies=pol, trajectories=traj)

    # ---- bulk side ----
    def extend_from_rollout(self, trajectory_hbo, policies):
        """trajectory_hbo[H, B, 2n+2] (the kernel's time-major layout, host array) and policies[B, ...]:
        adds B records that are views of one transposed block."""
        block = np.ascontiguousarray(np.transpose(np.asarray(trajectory_hbo), (1, 0, 2)))
        policies = np.asarray(policies)
        if policies.shape[0] != block.shape[0]:
            raise ValueError("%d policies for %d trajectories" % (policies.shape[0], block.shape[0]))
        self._records.extend(zip(block, policies))

    def as_arrays(self):
        """-> (trajectories[K, H, 2n+2], policies[K, n-1, 2n+2]) as float64 arrays."""
        if not self._records:
            return np.zeros((0, 0, 0)), np.zeros((0, 0, 0))
        traj = np.stack([np.asarray(t, dtype=np.float64) for t, _ in self._records])
        pol = np.stack([np.asarray(p, dtype=np.float64) for _, p in self._records])
        return traj, pol


def pick_sub_database(data_path, size, sub_data_path):
    """Writes a store of `size` records drawn with replacement from the store at `data_path`
    (ars/database.py:40-48; consumes np.random like the reference)."""
    full = Database()
    full.load(data_path)
    picked = Database()
    recs = full._records
    for i in np.random.randint(0, full.size, size):
        picked.add_trajectory(*recs[int(i)])
    picked.save(sub_data_path)
