"""Exact shortcut for the sequential Float16 accumulation of non-negative deposits (EXACT tallies with
PAIRWISE = FALSE on Float16 decks: `mesh.energydep[c, k] += v`, imc_transport.jl:120).

The reference adds the deposits of a cell one by one in Float16.  The running sum never decreases, and a deposit v changes
it only if fl16(acc + v) != acc, which for a fixed acc is true exactly for v >= vmin(acc) (rounding is monotone).  So the
chain can jump from one effective deposit to the next: "first record after position i with v >= vmin(acc)" is a descent
in a max-tree over the records, and acc takes at most 31 744 values.  This script checks the shortcut against the plain
loop bit for bit on Su-Olson-like data (CPU prototype for the CUDA reducer of a later round).

    python scratch/f16_chain_shortcut.py
"""
import numpy as np


def sequential(v):
    acc = np.float16(0.0)
    for x in v:
        acc = np.float16(acc + x)
    return acc


def vmin(acc):
    """Smallest Float16 v >= 0 with fl16(acc + v) != acc (binary search over the Float16 bit patterns, which are ordered)."""
    lo, hi = 0, 0x7C00                      # +0.0 .. +Inf
    while lo < hi:
        mid = (lo + hi) // 2
        v = np.array([mid], dtype=np.uint16).view(np.float16)[0]
        if np.float16(acc + v) != acc:
            hi = mid
        else:
            lo = mid + 1
    return np.array([lo], dtype=np.uint16).view(np.float16)[0]


class MaxTree:
    def __init__(self, v):
        n = 1
        while n < len(v):
            n *= 2
        self.n = n
        self.t = np.zeros(2 * n, dtype=np.float16)
        self.t[n:n + len(v)] = v
        for i in range(n - 1, 0, -1):
            self.t[i] = max(self.t[2 * i], self.t[2 * i + 1])

    def first_at_least(self, start, thr):
        """Smallest index j >= start with v[j] >= thr, or -1."""
        i = start + self.n
        if self.t[i] >= thr:
            return start
        while True:                           # climb until a right sibling holds a large enough value
            while i & 1:
                i >>= 1
                if i == 1:
                    return -1
            if i == 1:
                return -1
            i += 1
            if self.t[i] >= thr:
                break
        while i < self.n:                     # descend to the leftmost such leaf
            i = 2 * i if self.t[2 * i] >= thr else 2 * i + 1
        return i - self.n


def shortcut(v):
    tree = MaxTree(v)
    acc, i, steps = np.float16(0.0), 0, 0
    while i < len(v):
        j = tree.first_at_least(i, vmin(acc))
        if j < 0:
            break
        acc = np.float16(acc + v[j])
        i = j + 1
        steps += 1
    return acc, steps


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, scale in ((2000, 1.0), (50000, 0.02), (200000, 1e-3), (200000, 30.0)):
        v = (rng.random(n) ** 3 * scale).astype(np.float16)
        a = sequential(v)
        b, steps = shortcut(v)
        print(f"n = {n:7d} scale {scale:g}: sequential {a!r}  shortcut {b!r}  effective additions {steps}")
        assert a.tobytes() == b.tobytes()
    print("identical")
