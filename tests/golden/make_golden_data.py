"""Generates tests/golden/data_loader.json from the UNMODIFIED reference's pydynet/data.py: the index batches its samplers
produce under seeded shuffling (one numpy.random.permutation per epoch), sequentially, with and without drop_last."""
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
from pydynet.data import DataLoader, Dataset, data_loader  # noqa: E402


class Idx(Dataset):

    def __init__(self, n):
        self.n = n

    def __getitem__(self, index):
        return list(index)

    def __len__(self):
        return self.n


cases = []
for n, bs in ((23, 5), (16, 4), (7, 10), (1, 1)):
    for shuffle in (False, True):
        for drop in (False, True):
            np.random.seed(100 + n)
            dl = DataLoader(Idx(n), bs, shuffle, drop)
            epochs = [[b for b in dl] for _ in range(2)]  # second epoch continues the global NumPy stream
            cases.append({"n": n, "batch_size": bs, "shuffle": shuffle, "drop_last": drop, "seed": 100 + n, "epochs": epochs,
                          "len": len(dl.batch_sampler)})
X, y = np.arange(40).reshape(10, 4), np.arange(10)
np.random.seed(5)
xy = [[bx.tolist(), by.tolist()] for bx, by in data_loader(X, y, 4, True)]
out = {"cases": cases, "xy_seed": 5, "xy": xy}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_loader.json"), "w"))
print(len(cases), "cases")
