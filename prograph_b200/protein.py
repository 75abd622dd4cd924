"""Record of one protein: a ``Sequence`` plus arbitrary labels (prograph/protein.py)."""
import numpy as np


class Protein:
    def __init__(self, Sequence, **labels):
        self.Sequence = Sequence
        self.__dict__.update(labels)

    def __getitem__(self, keys):
        if isinstance(keys, list):
            return tuple(self.__dict__[k] for k in keys)
        return self.__dict__[keys]

    def __len__(self):
        return len(self.Sequence)

    def __eq__(self, other):
        return self.Sequence == other.Sequence

    def __hash__(self):
        return hash(self.Sequence)

    def __repr__(self):
        parts = []
        for key, value in vars(self).items():
            if isinstance(value, np.ndarray):
                parts.append(f"{key}=np.array({list(value)})")
            elif isinstance(value, str):
                parts.append(f"{key}='{value}'")
            else:
                parts.append(f"{key}={value}")
        return "Protein(" + ",".join(parts) + ")"
