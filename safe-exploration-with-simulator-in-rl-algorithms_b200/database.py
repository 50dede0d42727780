"""Host-side trajectory store with the reference's npz format (ars/database.py:16-37):
keys `policies` [(n-1) x (2n+2)] and `trajectories` [H x (2n+2)].  Pure bookkeeping -- no
arithmetic -- kept so that ARSAgent / Estimator signatures work unchanged."""
import numpy as np


class Database:
    def __init__(self):
        self.policies = []
        self.trajectories = []
        self.size = 0

    def load(self, path):
        data = np.load(path)
        if "policies" not in data.files or "trajectories" not in data.files:
            raise AssertionError("The file loaded doesn't contain the array 'policies' and 'trajectories'")
        policies, trajectories = data["policies"], data["trajectories"]
        if len(policies) != len(trajectories):
            raise AssertionError("'policies' and 'trajectories' doesn't have the same length")
        for policy, trajectory in zip(policies, trajectories):
            self.add_trajectory(trajectory, policy)

    def add_trajectory(self, trajectory, policy):
        self.trajectories.append(trajectory)
        self.policies.append(policy)
        self.size += 1

    def save(self, path):
        np.savez(path, policies=self.policies, trajectories=self.trajectories)


def pick_sub_database(data_path, size, sub_data_path):
    """Random sub-sample of a stored database (ars/database.py:40-48)."""
    data = Database()
    data.load(data_path)
    sub = Database()
    for i in np.random.randint(0, data.size, size):
        sub.add_trajectory(data.trajectories[i], data.policies[i])
    sub.save(sub_data_path)
