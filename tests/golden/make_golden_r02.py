#!/usr/bin/env python
"""Round-2 fixtures from the reference's shipped artefacts (run in the build container only; needs /root/reference):

  net_ck2.npz         weights of a SECOND shipped trained checkpoint (xw_mcts/wandb/run-20201108_220640-m3c7x7l3/temp/
                      temp.pth.tar; net.npz holds run-20201113_144231-15y0rcng) plus the reference BinPackingNNet's own
                      fp32 outputs on net.npz's 48 states — the "previous net" of the arena parity test
  wandb_first30.json  `iter mean reward` / `optimality percentage` of the first 30 iterations of the five archived
                      training runs (xw_mcts/wandb/run-*/wandb-history.jsonl), the learning-curve sanity reference
"""
import glob
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/xw_mcts"
HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("WANDB_MODE", "disabled")
sys.path.insert(0, REF)

from binpacking.pytorch.BinpackingNNet import BinPackingNNet as RefNet  # noqa: E402


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


class _G:
    def getBoardSize(self):
        return (15, 15)

    def getActionSize(self):
        return 150


def main():
    ck = os.path.join(REF, "wandb", "run-20201108_220640-m3c7x7l3", "temp", "temp.pth.tar")
    sd = torch.load(ck, map_location="cpu")["state_dict"]
    net = RefNet(_G(), dotdict(num_items=10, num_bins=1))
    net.load_state_dict(sd)
    net.eval()
    states = np.load(os.path.join(HERE, "net.npz"))["ck_states"]
    with torch.no_grad():
        lp, v = net(torch.from_numpy(states.astype(np.float32)))
    out = {"w." + k: t.numpy().astype(np.float32) for k, t in sd.items()}
    out["pi"] = torch.exp(lp).numpy().astype(np.float32)
    out["v"] = v.view(-1).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "net_ck2.npz"), **out)
    runs = {}
    for f in sorted(glob.glob(os.path.join(REF, "wandb", "run-*", "wandb-history.jsonl"))):
        rows = [json.loads(line) for line in open(f)]
        mean = [r["iter mean reward"] for r in rows if "iter mean reward" in r][:30]
        opt = [r["optimality percentage"] for r in rows if "optimality percentage" in r][:30]
        runs[os.path.basename(os.path.dirname(f))] = {"iter_mean_reward": mean, "optimality_percentage": opt}
    json.dump(runs, open(os.path.join(HERE, "wandb_first30.json"), "w"), indent=1)
    for k, r in runs.items():
        print(k, "first-30 mean of mean reward %.3f, last-10-of-30 %.3f" % (np.mean(r["iter_mean_reward"]),
                                                                       np.mean(r["iter_mean_reward"][20:])))


if __name__ == "__main__":
    main()
