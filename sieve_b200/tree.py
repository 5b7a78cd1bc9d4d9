"""Trunk trees in SIEVE's numbering and the post-order op lists the device consumes.

Shape facts mirrored from the reference (tree/ScsTree.java:220-260): n leaves numbered 0..n-1, n-1 binary
internal nodes n..2n-2, and a unary root 2n-1 whose single child is the tumour MRCA (the "trunk");
nodeCount = 2n, internalNodeCount = n.
"""
import numpy as np


class ScsTree:
    def __init__(self, n_leaves, left, right, height):
        self.n = int(n_leaves)
        self.left = np.asarray(left, dtype=np.int32)    # -1 for leaves
        self.right = np.asarray(right, dtype=np.int32)  # -1 for leaves and for the unary root
        self.height = np.asarray(height, dtype=np.float64)
        self.node_count = 2 * self.n
        self.root = 2 * self.n - 1
        self.parent = np.full(self.node_count, -1, dtype=np.int32)
        for p in range(self.n, self.node_count):
            for c in (self.left[p], self.right[p]):
                if c >= 0:
                    self.parent[c] = p

    def is_leaf(self, i):
        return i < self.n

    def branch_length(self, i):
        """Node.getLength(): parent height - own height (root: 0)."""
        p = self.parent[i]
        return 0.0 if p < 0 else float(self.height[p] - self.height[i])

    def branch_lengths(self):
        bl = np.zeros(self.node_count)
        nz = self.parent >= 0
        bl[nz] = self.height[self.parent[nz]] - self.height[nz]
        return bl

    def post_order(self):
        """Internal nodes in the order ScsTreeLikelihood.traverse finishes them
        (likelihood/ScsTreeLikelihood.java:612-619: left subtree, right subtree, node)."""
        out = []
        stack = [(self.root, 0)]
        while stack:
            node, state = stack.pop()
            if node < self.n:
                continue
            if state == 0:
                stack.append((node, 1))
                if self.right[node] >= 0:
                    stack.append((self.right[node], 0))
                stack.append((self.left[node], 0))
            else:
                out.append(node)
        return out

    def ops(self, nodes=None):
        """[(parent, child1, child2 | -1)] for `nodes` (default: every internal node), post-order."""
        order = self.post_order()
        if nodes is not None:
            keep = set(int(x) for x in nodes)
            order = [x for x in order if x in keep]
        return np.array([[p, self.left[p], self.right[p]] for p in order], dtype=np.int32).reshape(-1, 3)

    def levels(self):
        """level[node] = 0 for leaves, 1 + max(children) otherwise (the launch order of a level-batched update)."""
        lv = np.zeros(self.node_count, dtype=np.int32)
        for p in self.post_order():
            c = [self.left[p]] + ([self.right[p]] if self.right[p] >= 0 else [])
            lv[p] = 1 + max(lv[x] for x in c)
        return lv

    def ops_by_level(self, nodes=None):
        """ops sorted by level plus the level offsets (ops[off[i]:off[i+1]] are independent)."""
        ops = self.ops(nodes)
        if len(ops) == 0:
            return ops, np.zeros(1, dtype=np.int32)
        lv = self.levels()[ops[:, 0]]
        order = np.argsort(lv, kind="stable")
        ops = ops[order]
        lv = lv[order]
        uniq, first = np.unique(lv, return_index=True)
        off = np.concatenate([first, [len(ops)]]).astype(np.int32)
        return ops, off

    def dirty_closure(self, nodes):
        """`nodes` plus all their ancestors -- the set traverse() recomputes when those nodes are dirty
        (likelihood/ScsTreeLikelihood.java:621-628). Leaves are dropped (they have no op)."""
        s = set()
        for x in nodes:
            x = int(x)
            while x >= 0 and x not in s:
                s.add(x)
                x = int(self.parent[x])
        return sorted(i for i in s if i >= self.n)

    def copy(self):
        return ScsTree(self.n, self.left.copy(), self.right.copy(), self.height.copy())


def caterpillar(n, min_internal_height=0.0, step=1.0):
    """tree/ScsTree.java:220-260 makeCaterpillar(minInternalHeight, step)."""
    left = np.full(2 * n, -1, dtype=np.int32)
    right = np.full(2 * n, -1, dtype=np.int32)
    height = np.zeros(2 * n)
    cur = 0
    for i in range(1, n):
        p = n + i - 1
        left[p], right[p] = cur, i
        height[p] = min_internal_height + i * step
        cur = p
    root = 2 * n - 1
    left[root] = cur
    height[root] = min_internal_height + n * step
    return ScsTree(n, left, right, height)


def random_coalescent(n, rng, pop_size=1.0, trunk=None):
    """Kingman coalescent over n leaves at height 0, plus a trunk above the MRCA (ScsRandomTree-like shape)."""
    left = np.full(2 * n, -1, dtype=np.int32)
    right = np.full(2 * n, -1, dtype=np.int32)
    height = np.zeros(2 * n)
    active = list(range(n))
    t = 0.0
    nxt = n
    while len(active) > 1:
        k = len(active)
        t += rng.exponential(pop_size / (k * (k - 1) / 2.0))
        i, j = rng.choice(k, size=2, replace=False)
        a, b = active[i], active[j]
        left[nxt], right[nxt] = a, b
        height[nxt] = t
        active = [x for idx, x in enumerate(active) if idx not in (i, j)] + [nxt]
        nxt += 1
    root = 2 * n - 1
    left[root] = active[0]
    height[root] = t + (trunk if trunk is not None else rng.exponential(pop_size))
    if n == 1:
        height[root] = trunk if trunk is not None else 1.0
    return ScsTree(n, left, right, height)


def fixture_9_leaves():
    """The hand-built 9-leaf trunk tree of test/beast/evolution/likelihood/ScsTreeLikelihoodTest.java:40-143
    (leaves 0-8, internals 9-16, root 17 over node 16); only the shape is taken, heights are ours."""
    n = 9
    # (parent, left, right) -- a fixed, unbalanced shape with the same node numbering scheme
    topo = [(9, 0, 1), (10, 2, 3), (11, 9, 10), (12, 4, 5), (13, 12, 6), (14, 7, 8), (15, 13, 14), (16, 11, 15)]
    left = np.full(2 * n, -1, dtype=np.int32)
    right = np.full(2 * n, -1, dtype=np.int32)
    height = np.zeros(2 * n)
    for p, l, r in topo:
        left[p], right[p] = l, r
        height[p] = 0.05 + max(height[l], height[r]) + 0.03 * (p % 4)
    left[17] = 16
    height[17] = height[16] + 0.1
    return ScsTree(n, left, right, height)
