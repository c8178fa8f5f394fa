"""Generates tests/golden/sampler_golden.json by importing the REFERENCE's MyBatchSampler
(/root/reference/data/custom_datasets.py:19-78) in the build container (the reference cannot travel to the GPU
box).  Its module imports HuggingFace `datasets` at the top, which is not installed here and is not used by the
sampler: a stub module stands in.  Run:  python tests/golden/make_sampler_golden.py"""
import json
import os
import sys
import types

REF = "/root/reference"
stub = types.ModuleType("datasets")
stub.load_from_disk = stub.concatenate_datasets = None
sys.modules["datasets"] = stub
sys.path.insert(0, REF)
from data.custom_datasets import MyBatchSampler, pad_only_according_data  # noqa: E402

cases = []
for (sizes, bss, world, skipped) in [([10, 7, 9], [2, 1, 3], 1, 0), ([64, 48, 40, 24, 16, 8], [32, 16, 8, 4, 2, 1], 2, 0),
                                     ([64, 48, 40, 24, 16, 8], [8, 4, 2, 2, 1, 1], 4, 3), ([5, 5], [4, 2], 2, 0)]:
    cum, t = [], 0
    for s in sizes:
        t += s
        cum.append(t)
    per_rank = []
    for rank in range(world):
        sm = MyBatchSampler(None, 1, True, cum, bss, skipped_batches=skipped)
        sm.set_world_size(world)
        sm.rank = rank
        per_rank.append({"batches": [list(b) for b in sm], "len": len(sm)})
    cases.append({"cumulative_sizes": cum, "batch_sizes": bss, "world": world, "skipped": skipped, "ranks": per_rank})
feats = [{"input_ids": [5, 6, 7], "labels": [-100, 6, 1], "fixed_len": 6}, {"input_ids": [9], "labels": [1], "fixed_len": 6}]
ids, lab = pad_only_according_data(feats)
out = {"sampler": cases, "collate": {"features": feats, "input_ids": ids.tolist(), "labels": lab.tolist()}}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.json")
json.dump(out, open(path, "w"))
print(path, sum(len(r["batches"]) for c in cases for r in c["ranks"]), "batches")
