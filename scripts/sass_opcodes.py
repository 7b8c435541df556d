"""Static SASS evidence: per kernel of csrc/libbpp_b200.so the code size and the counts of the tensor-core / tensor-memory /
bulk-copy / barrier opcodes (cuobjdump -sass; no GPU needed).   python scripts/sass_opcodes.py > profiles/r02_sass_opcodes_tensor.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "resource_packing_self_play_b200", "csrc", "libbpp_b200.so")
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTCATOMSWS", "LDGSTS", "SYNCS", "REDUX", "DFMA", "DADD", "DMUL", "FFMA",
        "LDL", "STL"]


def demangle(names):
    r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
    out = r.stdout.split("\n") if r.returncode == 0 else names
    def short(o):
        o = o.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        cut = o.rfind(">(")                       # parameter list behind the template arguments ...
        o = o[:cut + 1] if cut >= 0 else o.split("(")[0]   # ... or behind a plain name
        return o.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
    return [short(o) for o in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = {"name": m.group(1), "ops": collections.Counter(), "last": 0}
            kernels.append(cur)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["last"] = max(cur["last"], int(m.group(1), 16))
            cur["ops"][m.group(3)] += 1
    names = demangle([k["name"] for k in kernels])
    print("cuobjdump -sass csrc/libbpp_b200.so (scripts/sass_opcodes.py): static counts of tensor-core / TMEM / bulk-copy / barrier /")
    print("float64 / local-memory opcodes per kernel (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,")
    print("UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops, LDL / STL = local memory;")
    print(" k_net_gr<false|true> = grid-row trunk in bf16 | split-bf16, the kernels bpp_net_forward launches; k_net_role / k_net_forward_tc")
    print(" = round-1 fallback; k_episode / k_expand_search = search kernels; k_lr_stage_* = fused learner stages)\n")
    keep = ("k_net_gr", "k_net_heads_tc", "k_net_role", "k_net_forward_tc", "k_episode", "k_expand_search", "k_search", "k_lr_")
    for k, n in zip(kernels, names):
        if not any(s in n for s in keep):
            continue
        parts = "  ".join(f"{o} {k['ops'][o]}" for o in WANT if k["ops"][o])
        print(f"{n[:62]:62s} code {k['last'] + 16:7d} B  {parts}")


if __name__ == "__main__":
    sys.exit(main())
