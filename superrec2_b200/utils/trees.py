"""Ancestry queries and polytomy resolution on host-side trees.

Host-side counterpart of ``/root/reference/src/superrec2/utils/trees.py``:

* ``LowestCommonAncestor`` (reference ``utils/trees.py:31-142``) answers the
  same queries (``__call__``, ``is_ancestor_of``, ``is_strict_ancestor_of``,
  ``is_comparable``, ``level``, ``distance``).  Instead of an Euler tour with a
  sparse table it numbers the nodes with entry/exit times, which is also the
  form the device kernels use (ancestor-or-self test = interval containment).
* ``is_binary`` (reference ``:370-374``).
* Polytomy resolutions are *unranked* instead of enumerated: resolution number
  ``r`` of ``binarize()`` (reference ``:377-446``; order restated in SURVEY.md
  Appendix C) is built directly from ``r`` by mixed-radix decomposition, so a
  range of resolution numbers can be handed to any GPU.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Sequence, Tuple

from ..model.tree import Tree, TreeNode


class LowestCommonAncestor:
    """Constant-time ancestry tests and level look-ups on a fixed tree."""

    def __init__(self, tree: Tree):
        self.tree = tree
        self._tin: Dict[TreeNode, int] = {}
        self._tout: Dict[TreeNode, int] = {}
        self._level: Dict[TreeNode, int] = {}
        clock = 0
        stack = [(tree, 0, False)]
        while stack:
            node, level, leaving = stack.pop()
            if leaving:
                self._tout[node] = clock
                continue
            self._tin[node] = clock
            self._level[node] = level
            clock += 1
            stack.append((node, level, True))
            for child in reversed(node.children):
                stack.append((child, level + 1, False))

    def __call__(self, *nodes: TreeNode) -> TreeNode:
        if not nodes:
            raise TypeError("At least one node is needed to compute the LCA")
        low = min(self._tin[node] for node in nodes)
        high = max(self._tin[node] for node in nodes)
        result = nodes[0]
        while not (self._tin[result] <= low and high < self._tout[result]):
            result = result.up
        return result

    def is_ancestor_of(self, first: TreeNode, second: TreeNode) -> bool:
        """``first`` is ``second`` or one of its ancestors."""
        return self._tin[first] <= self._tin[second] < self._tout[first]

    def is_strict_ancestor_of(self, first: TreeNode, second: TreeNode) -> bool:
        return first is not second and self.is_ancestor_of(first, second)

    def is_comparable(self, first: TreeNode, second: TreeNode) -> bool:
        return self.is_ancestor_of(first, second) or self.is_ancestor_of(second, first)

    def level(self, node: TreeNode) -> int:
        return self._level[node]

    def distance(self, first: TreeNode, second: TreeNode) -> int:
        meet = self(first, second)
        return self._level[first] + self._level[second] - 2 * self._level[meet]


def is_binary(tree: Tree) -> bool:
    """Every internal node has exactly two children."""
    return all(len(node.children) in (0, 2) for node in tree.traverse("preorder"))


# --------------------------------------------------------------------------
# Polytomy resolutions by unranking
# --------------------------------------------------------------------------
def double_factorial_odd(k: int) -> int:
    """(2k-3)!! = number of rooted binary trees over k labelled items."""
    count = 1
    for m in range(3, k + 1):
        count *= 2 * m - 3
    return count


def count_resolutions(tree: Tree) -> int:
    """Number of binary resolutions of ``tree`` (1 if already binary)."""
    total = 1
    for node in tree.traverse("preorder"):
        total *= double_factorial_odd(len(node.children))
    return total


def _arrangement(items: Sequence[TreeNode], index: int) -> TreeNode:
    """
    ``index``-th binary arrangement of ``items`` in the reference's order
    (``arrange_leaves``/``graft``, reference ``utils/trees.py:377-420``): start
    from the last item and insert the others from right to left; the digit of
    the item inserted last (``items[0]``) varies fastest and counts positions in
    preorder over the nodes of the partial tree, items being opaque.
    """
    count = len(items)
    if count == 1:
        return items[0]
    digits = []
    for m in range(count - 1):  # items[m] goes into a tree over count-1-m items
        radix = 2 * (count - 1 - m) - 1
        digits.append(index % radix)
        index //= radix
    opaque = {id(item) for item in items}
    root = items[-1]
    for m in range(count - 2, -1, -1):
        # find the node at preorder position digits[m]
        position = digits[m]
        stack = [root]
        target = None
        while stack:
            node = stack.pop()
            if position == 0:
                target = node
                break
            position -= 1
            if id(node) not in opaque:
                stack.extend(reversed(node.children))
        joint = TreeNode()
        parent = target.up
        if parent is not None:
            slot = 0 if parent.children[0] is target else 1
            parent.children[slot] = joint
            joint.up = parent
        else:
            root = joint
        target.up = None
        joint.add_child(items[m])
        joint.add_child(target)
    root.up = None
    return root


def unrank_resolution(tree: Tree, index: int) -> Tree:
    """
    Build resolution number ``index`` of ``tree`` (a fresh tree; ``tree`` is not
    modified).  Order: children's choices vary slowest (first child slowest),
    the node's own arrangement fastest (reference ``utils/trees.py:435-439``).
    """
    counts: Dict[int, int] = {}
    for node in tree.traverse("postorder"):
        total = double_factorial_odd(len(node.children))
        for child in node.children:
            total *= counts[id(child)]
        counts[id(node)] = total
    if not 0 <= index < counts[id(tree)]:
        raise IndexError("resolution index out of range")

    # top-down split of the index, then bottom-up assembly
    share: Dict[int, Tuple[int, int]] = {}
    pending = [(tree, index)]
    order: List[TreeNode] = []
    while pending:
        node, idx = pending.pop()
        order.append(node)
        ways = double_factorial_odd(len(node.children))
        own = idx % ways
        idx //= ways
        share[id(node)] = (own, 0)
        for child in reversed(node.children):  # last child varies fastest
            pending.append((child, idx % counts[id(child)]))
            idx //= counts[id(child)]

    built: Dict[int, TreeNode] = {}
    for node in reversed(order):  # children before parents
        if not node.children:
            twin = TreeNode(name=node.name)
            twin.props = dict(node.props)
        else:
            items = [built[id(child)] for child in node.children]
            twin = _arrangement(items, share[id(node)][0])
            twin.name = node.name
            twin.props = dict(node.props)
        twin.dist, twin.support = node.dist, node.support
        built[id(node)] = twin
    return built[id(tree)]


def binarize(tree: Tree) -> Iterator[Tree]:
    """All binary resolutions of ``tree`` in the reference's order."""
    for index in range(count_resolutions(tree)):
        yield unrank_resolution(tree, index)
