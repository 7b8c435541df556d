"""Restatement of numpy's float64 pairwise add-reduction (np.sum over a contiguous 1-D float64 array), the summation
order the reference relies on at MCTS_bpp.py:90,100 (`np.sum(self.Ps[s])`).  TEST INFRASTRUCTURE ONLY.

numpy (numpy/_core/src/umath/loops_utils.h.src, `DOUBLE_pairwise_sum`; numpy is a pinned dependency of the
reference, not vendored) sums blocks of <= 128 elements with 8 interleaved accumulators combined as
((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) followed by a sequential tail, and splits longer inputs recursively at n/2 rounded
down to a multiple of 8.  The CUDA kernels implement exactly this plan (csrc/bpp_device.cuh: np_pairwise_sum); the
test `tests/test_oracle_golden.py::test_pairwise_sum_replica_matches_numpy` pins this restatement against numpy
itself for every length 0..700.
"""
import numpy as np


def pairwise_sum(a):
    a = np.asarray(a, dtype=np.float64)
    n = len(a)
    if n < 8:
        res = np.float64(0.0)
        for i in range(n):
            res = res + a[i]
        return res
    if n <= 128:
        r = [a[j] for j in range(8)]
        lim = n - (n % 8)
        for i in range(8, lim, 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for i in range(lim, n):
            res = res + a[i]
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def sum_plan(n):
    """(leaves [(base, n)], postfix program) exactly as build_sum_plan_rec in csrc/bpp_engine.cu"""
    leaves, prog = [], []

    def rec(base, m):
        if m <= 128:
            prog.append(len(leaves))
            leaves.append((base, m))
        else:
            m2 = m // 2
            m2 -= m2 % 8
            rec(base, m2)
            rec(base + m2, m - m2)
            prog.append(-1)
    rec(0, n)
    return leaves, prog
