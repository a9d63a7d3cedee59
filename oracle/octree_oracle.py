"""CPU restatement of the reference's octree quantiser — TEST INFRASTRUCTURE ONLY (imported by
tests/; the product path is kmg_octree_palette in kmeans-gpu_b200/csrc/kmg_host.cpp).

Follows core/src/octree.rs literally, data structure by data structure, in pure Python (small
inputs: the reference never hands it more than 128 x 128 pixels, core/src/lib.rs:288-316):

  get_color_index      octree.rs:12-26
  ColorTree.add_color  octree.rs:41-63   (child created at loop level L stores level = L)
  ColorTree.reduce     octree.rs:65-110  (deque sorted descending, pop_back, binary searches)
  Node ordering        octree.rs:246-272 (child_count, pixel_count >> level, node id)
  output_color         octree.rs:128-135 (integer division, alpha 255)

Pinned by the reference's own unit test (octree.rs:280-340: 46 colours -> 8) and by hand-checked
small cases in tests/test_octree.py.
"""
from __future__ import annotations

from functools import cmp_to_key


class _Node:
    __slots__ = ("level", "node_id", "color_index", "parent", "children", "child_count", "n", "r", "g", "b")

    def __init__(self, node_id, parent=None, color_index=0, level=0):
        self.level = level
        self.node_id = node_id
        self.color_index = color_index
        self.parent = parent
        self.children = [None] * 8
        self.child_count = 0
        self.n = self.r = self.g = self.b = 0


def _cmp(a: _Node, b: _Node) -> int:
    """Node::partial_cmp (octree.rs:246-266)."""
    if a.node_id == b.node_id:
        return 0
    if a.child_count != b.child_count:
        return -1 if a.child_count < b.child_count else 1
    ac, bc = a.n >> a.level, b.n >> b.level
    if ac != bc:
        return -1 if ac < bc else 1
    return -1 if a.node_id < b.node_id else 1


def _binary_search_by(seq, f):
    """slice::binary_search_by: f(probe) is the ordering of probe relative to the target."""
    lo, hi = 0, len(seq)
    while lo < hi:
        mid = lo + (hi - lo) // 2
        c = f(seq[mid])
        if c == 0:
            return True, mid
        if c < 0:
            lo = mid + 1
        else:
            hi = mid
    return False, lo


def octree_palette(pixels, color_count: int):
    """operations::extract_palette_octree (core/src/operations.rs:90-97): list of (r,g,b,a)."""
    nodes = [_Node(0)]
    for px in pixels:
        r, g, b = int(px[0]), int(px[1]), int(px[2])
        root = 0
        for level in range(8):
            mask = 0x80 >> level
            ci = (4 if r & mask else 0) | (2 if g & mask else 0) | (1 if b & mask else 0)
            if nodes[root].children[ci] is None:
                new_id = len(nodes)
                nodes[root].children[ci] = new_id
                nodes[root].child_count += 1
                nodes.append(_Node(new_id, root, ci, level))
            root = nodes[root].children[ci]
        leaf = nodes[root]
        leaf.r += r
        leaf.g += g
        leaf.b += b
        leaf.n += 1
    if color_count == 0:
        return []
    leaves = [nd for nd in nodes if nd.n > 0]
    leaves.sort(key=cmp_to_key(lambda a, b: -_cmp(a, b)))  # a.cmp(b).reverse()
    while len(leaves) > color_count:
        node = leaves.pop()
        if node.parent is not None:
            parent = nodes[node.parent]
            found, pos = _binary_search_by(leaves, lambda probe: _cmp(parent, probe))
            if found:
                del leaves[pos]
            parent.r += node.r
            parent.g += node.g
            parent.b += node.b
            parent.n += node.n
            parent.child_count -= 1
            parent.children[node.color_index] = None
            node.parent = None
            found, pos = _binary_search_by(leaves, lambda probe: _cmp(parent, probe))
            if not found:
                leaves.insert(pos, parent)
    pal = sorted({(nd.r // nd.n, nd.g // nd.n, nd.b // nd.n, 255) for nd in leaves})
    return pal
