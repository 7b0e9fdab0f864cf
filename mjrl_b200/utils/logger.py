"""Minimal DataLog with the reference's interface (mjrl/utils/logger.py:16-42): dict of lists + CSV dump.
Only what the agents call (`log_kv`, `get_current_log`, `save_log`) -- observability is out of scope."""
import csv
import os
import pickle


class DataLog:
    def __init__(self):
        self.log = {}
        self.max_len = 0

    def log_kv(self, key, value):
        if key not in self.log:
            self.log[key] = []
        self.log[key].append(value)
        self.max_len = max(self.max_len, len(self.log[key]))

    def get_current_log(self):
        return {k: v[-1] for k, v in self.log.items() if v}

    def save_log(self, save_path):
        os.makedirs(save_path, exist_ok=True)
        with open(os.path.join(save_path, "log.pickle"), "wb") as f:
            pickle.dump(self.log, f)
        keys = sorted(self.log)
        with open(os.path.join(save_path, "log.csv"), "w", newline="") as f:
            wr = csv.DictWriter(f, fieldnames=keys)
            wr.writeheader()
            for i in range(self.max_len):
                wr.writerow({k: self.log[k][i] for k in keys if i < len(self.log[k])})
