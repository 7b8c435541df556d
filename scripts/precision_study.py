#!/usr/bin/env python
"""CPU study (this container only: reads the reference's shipped checkpoints under /root/reference): which operand
formats does the policy/value forward need, layer by layer, to stay within |d pi| <= 2e-2, |d v| <= 2e-2 of the fp32
reference on the five TRAINED checkpoints (xw_mcts/wandb/run-*/temp/temp.pth.tar)?

Emulates the tensor-core data path: every stored activation and every weight is rounded to the chosen format
(bf16 / fp16 / hi+lo pairs of them), products and sums are exact-ish (float64), bias fp32.
Output: profiles/r02_precision_study.txt
"""
import glob
import itertools
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def q(x, fmt):
    if fmt == "fp32":
        return x.float().double()
    base = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt.replace("x2", "")]
    xf = x.float()
    hi = xf.to(base).float()
    if fmt.endswith("x2"):
        lo = (xf - hi).to(base).float()
        return (hi + lo).double()
    return hi.double()


LAYERS = ["s%d.%s" % (s, k) for s in range(3) for k in ("conv", "r0c0", "r0c1", "r1c0", "r1c1")] + ["hidden", "logits"]


def forward(Wt, x, afmt, wfmt):
    """afmt[l]: format of the activations layer l READS (i.e. how its input was stored); wfmt[l]: weight format.
    residual stream (raw) is stored in the format of the layer that reads it next."""
    x = x.double()
    li = 0
    names = {"conv": "conv", "r0c0": "res_block0.conv0", "r0c1": "res_block0.conv1", "r1c0": "res_block1.conv0",
             "r1c1": "res_block1.conv1"}

    def conv(x, s, k, l):
        p = "conv_seqs.%d.%s." % (s, names[k])
        return F.conv2d(x, q(Wt[p + "weight"], wfmt[l]), Wt[p + "bias"].double(), padding=1)
    for s in range(3):
        x = q(x, afmt[li])                      # input planes are 0/1 at s = 0 (exact); raw of the previous sequence
        t = conv(x, s, "conv", li); li += 1
        t = q(t, afmt[li])                      # T stored, pooled, becomes raw (read by r0c0 and the skip)
        x = F.max_pool2d(t, 3, 2, 1)
        for b in range(2):
            y = conv(F.relu(x), s, "r%dc0" % b, li); li += 1
            y = q(y, afmt[li])                  # actB
            z = conv(F.relu(y), s, "r%dc1" % b, li) + x; li += 1
            nxt = afmt[li] if li < len(afmt) else afmt[-1]
            x = q(z, nxt)                       # raw stored in the next reader's format
    h = F.relu(F.linear(F.relu(torch.flatten(x, 1)), q(Wt["hidden_fc.weight"], wfmt[15]), Wt["hidden_fc.bias"].double()))
    h = q(h, afmt[16])
    lg = F.linear(h, q(Wt["logits_fc.weight"], wfmt[16]), Wt["logits_fc.bias"].double())
    v = torch.tanh(F.linear(h, q(Wt["value_fc.weight"], wfmt[16]), Wt["value_fc.bias"].double()))
    return F.softmax(lg, 1), v.view(-1), lg, h


def main():
    sys.path.insert(0, "/root/reference/xw_mcts")
    cks = sorted(glob.glob("/root/reference/xw_mcts/wandb/run-*/temp/temp.pth.tar"))
    d = np.load(os.path.join(ROOT, "tests/golden/net.npz"))
    states = torch.from_numpy(d["ck_states"].astype(np.float32))
    # more states: random playouts of the oracle game
    from oracle import bpp_oracle as O
    rng = np.random.RandomState(0)
    extra = []
    for seed in range(40):
        gh = int(rng.randint(2, 16))
        items = O.OracleItemsGenerator(15, gh, 10).items_generator(1000 + seed)
        g = O.OracleGame(15, 15, 10, 1)
        board, planes = g.getInitBoard(), g.getInitItems(items)
        while True:
            st = g.getBinItem(board, planes)
            if not g.has_valid_moves(st):
                break
            valid = np.flatnonzero(g.getValidMoves(st))
            extra.append(np.asarray(st, dtype=np.float32))
            board, planes = g.getNextState(board, int(rng.choice(valid)), planes)
    states = torch.cat([states, torch.from_numpy(np.stack(extra))])
    out = []

    def log(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        out.append(s)
    log("states:", tuple(states.shape), "checkpoints:", len(cks))
    nets = []
    for ck in cks:
        sd = torch.load(ck, map_location="cpu")["state_dict"]
        Wt = {k: v.float() for k, v in sd.items()}
        ref = forward(Wt, states, ["fp32"] * 17, ["fp32"] * 17)
        nets.append((os.path.basename(os.path.dirname(os.path.dirname(ck))), Wt, ref))
        log(nets[-1][0], "logits range [%.1f, %.1f]" % (ref[2].min(), ref[2].max()), "max logit per state min %.1f" %
            ref[2].max(1).values.min(), "hidden max %.1f" % ref[3].max(),
            "max|w| conv %.3f fc %.3f" % (max(v.abs().max() for k, v in Wt.items() if "conv" in k and "weight" in k),
                                          max(Wt[k].abs().max() for k in ("hidden_fc.weight", "logits_fc.weight"))))

    def err(afmt, wfmt):
        ep = ev = 0.0
        agree = 1.0
        for _, Wt, ref in nets:
            pi, v, _, _ = forward(Wt, states, afmt, wfmt)
            ep = max(ep, float((pi - ref[0]).abs().max()))
            ev = max(ev, float((v - ref[1]).abs().max()))
            agree = min(agree, float((pi.argmax(1) == ref[0].argmax(1)).float().mean()))
        return ep, ev, agree
    log("\n== uniform formats (activations, weights): max |d pi|, max |d v|, min arg-max agreement over 5 checkpoints")
    fmts = ["bf16", "fp16", "bf16x2", "fp16x2"]
    for a, w in itertools.product(fmts, fmts):
        e = err([a] * 17, [w] * 17)
        log("act %-7s w %-7s  dpi %.2e dv %.2e agree %.3f" % (a, w, *e))
    log("\n== fp16 activations + fp16 weights, ONE layer upgraded to (fp16x2 act, fp16x2 w)")
    for l in range(17):
        a, w = ["fp16"] * 17, ["fp16"] * 17
        a[l] = w[l] = "fp16x2"
        log("%-8s dpi %.2e dv %.2e agree %.3f" % (LAYERS[l], *err(a, w)))
    log("\n== all fp16x2 except ONE layer downgraded to plain fp16 (act, w)")
    for l in range(17):
        a, w = ["fp16x2"] * 17, ["fp16x2"] * 17
        a[l] = w[l] = "fp16"
        log("%-8s dpi %.2e dv %.2e agree %.3f" % (LAYERS[l], *err(a, w)))
    log("\n== fp16 act everywhere, weights fp16x2 from layer k on (2 MMAs per product there)")
    for k in range(0, 18, 2):
        w = ["fp16"] * k + ["fp16x2"] * (17 - k)
        log("k=%2d dpi %.2e dv %.2e agree %.3f" % (k, *err(["fp16"] * 17, w)))
    with open(os.path.join(ROOT, "profiles", "r02_precision_study.txt"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
