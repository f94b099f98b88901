"""Rooted tree container with newick/NHX input and output.

The reference keeps its trees in ``ete3.Tree`` objects (third-party, not
installed here).  This class offers the small part of that interface which the
reconciliation entry points rely on, so that code written against the reference
(``tree & "name"``, ``node.children``, ``tree.traverse("postorder")``, iterating
a tree to get its leaves, ``tree.write(format=8, ...)``) reads the same:

* reading: ``Tree(newick, format=1)`` as in
  ``/root/reference/src/superrec2/model/reconciliation.py:120-121``;
* writing: ``write(format=8, format_root_node=True, features=["color"])`` as in
  ``model/reconciliation.py:98-107`` (names of every node, no branch lengths,
  NHX ``[&&NHX:color=...]`` annotations passed through);
* default ``traverse()`` order is level order, which is what fixes the
  reference's tie-breaking (``compute/reconciliation.py:79,93,136,286``).

All walks are iterative: trees with thousands of levels do not hit Python's
recursion limit.
"""
from __future__ import annotations

import re
from collections import deque
from typing import Dict, Iterator, List, Optional

# Characters that cannot appear in a newick label; replaced by "_" on output.
_UNSAFE = re.compile(r"[:;(),\[\]\t\n\r=]")
_STRUCT = "(),;"


class TreeError(Exception):
    """Raised for malformed newick input or failed node look-ups."""


class TreeNode:
    """A node of a rooted tree (the tree is identified with its root)."""

    __slots__ = ("name", "children", "up", "props", "dist", "support")

    def __init__(self, newick: Optional[str] = None, format: int = 1,
                 name: str = ""):
        self.name: str = name
        self.children: List["TreeNode"] = []
        self.up: Optional["TreeNode"] = None
        # NHX annotations ("color", ...) carried from input to output
        self.props: Dict[str, str] = {}
        self.dist: float = 1.0
        self.support: float = 1.0
        if newick is not None:
            _parse_into(self, newick)

    # ------------------------------------------------------------------ basics
    def is_leaf(self) -> bool:
        return not self.children

    def is_root(self) -> bool:
        return self.up is None

    def add_child(self, child: Optional["TreeNode"] = None,
                  name: Optional[str] = None) -> "TreeNode":
        if child is None:
            child = TreeNode()
        if name is not None:
            child.name = name
        child.up = self
        self.children.append(child)
        return child

    @property
    def features(self):
        """Names of the annotations present on this node (ete3 spelling)."""
        return {"name", "dist", "support", *self.props}

    def add_feature(self, key: str, value) -> None:
        if key == "name":
            self.name = value
        elif key in ("dist", "support"):
            setattr(self, key, value)
        else:
            self.props[key] = value

    def __getattr__(self, key):
        # only reached when normal attribute look-up fails: NHX annotations
        props = object.__getattribute__(self, "props")
        if key in props:
            return props[key]
        raise AttributeError(key)

    def __bool__(self) -> bool:
        return True

    def __len__(self) -> int:
        return sum(1 for _ in self.iter_leaves())

    def __iter__(self) -> Iterator["TreeNode"]:
        return self.iter_leaves()

    def __and__(self, name: str) -> "TreeNode":
        """First node in level order carrying that name."""
        name = str(name)
        for node in self.traverse():
            if node.name == name:
                return node
        raise TreeError(f"Node not found: {name!r}")

    def __contains__(self, item) -> bool:
        if isinstance(item, TreeNode):
            return any(node is item for node in self.traverse() if node is not self)
        return any(node.name == item for node in self.traverse())

    def __repr__(self) -> str:
        return f"TreeNode({self.name!r}, {len(self.children)} children)"

    # --------------------------------------------------------------- traversal
    def traverse(self, strategy: str = "levelorder") -> Iterator["TreeNode"]:
        if strategy == "levelorder":
            queue = deque((self,))
            while queue:
                node = queue.popleft()
                yield node
                queue.extend(node.children)
        elif strategy == "preorder":
            stack = [self]
            while stack:
                node = stack.pop()
                yield node
                stack.extend(reversed(node.children))
        elif strategy == "postorder":
            stack = [(self, 0)]
            while stack:
                node, pos = stack.pop()
                if pos < len(node.children):
                    stack.append((node, pos + 1))
                    stack.append((node.children[pos], 0))
                else:
                    yield node
        else:
            raise TreeError(f"unknown traversal strategy {strategy!r}")

    def iter_leaves(self) -> Iterator["TreeNode"]:
        for node in self.traverse("preorder"):
            if not node.children:
                yield node

    def get_leaves(self) -> List["TreeNode"]:
        return list(self.iter_leaves())

    def copy(self) -> "TreeNode":
        """Deep copy of the subtree below (and including) this node."""
        twin = {id(self): _clone_one(self)}
        for node in self.traverse("preorder"):
            mine = twin[id(node)]
            for child in node.children:
                kid = _clone_one(child)
                twin[id(child)] = kid
                mine.add_child(kid)
        return twin[id(self)]

    # ------------------------------------------------------------------ output
    def write(self, format: int = 8, format_root_node: bool = True,
              features: Optional[List[str]] = None) -> str:
        """Serialise to newick.  ``format=8``: all names; ``format=9``: leaf names."""
        if format not in (8, 9, 100):
            raise TreeError(f"unsupported output format {format}")
        parts: List[str] = []
        stack = [(self, False)]
        while stack:
            node, closing = stack.pop()
            if closing:
                parts.append(")")
                if node.up is not None or format_root_node:
                    if format == 8:
                        parts.append(_UNSAFE.sub("_", node.name))
                    parts.append(_nhx(node, features))
                continue
            if node is not self and node.up.children[0] is not node:
                parts.append(",")
            if node.children:
                parts.append("(")
                stack.append((node, True))
                stack.extend((child, False) for child in reversed(node.children))
            else:
                if format != 100:
                    parts.append(_UNSAFE.sub("_", node.name))
                parts.append(_nhx(node, features))
        parts.append(";")
        return "".join(parts)


Tree = TreeNode


def _clone_one(node: TreeNode) -> TreeNode:
    twin = TreeNode(name=node.name)
    twin.props = dict(node.props)
    twin.dist = node.dist
    twin.support = node.support
    return twin


def _nhx(node: TreeNode, features: Optional[List[str]]) -> str:
    if features is None:
        return ""
    keys = sorted(node.features) if features == [] else features
    shown = []
    for key in keys:
        if key == "name":
            value = node.name
        elif key in ("dist", "support"):
            value = str(getattr(node, key))
        elif key in node.props:
            value = node.props[key]
        else:
            continue
        shown.append(f"{key}={_UNSAFE.sub('_', str(value))}")
    return "[&&NHX:" + ":".join(shown) + "]" if shown else ""


# ----------------------------------------------------------------------- input
_TOKEN = re.compile(r"\[[^\]]*\]|[(),;]|[^(),;\[]+")


def _set_label(node: TreeNode, text: str, comments: List[str]) -> None:
    name, _, length = text.partition(":")
    name = name.strip()
    if name:
        node.name = name
    length = length.strip()
    if length:
        try:
            node.dist = float(length)
        except ValueError as exc:
            raise TreeError(f"bad branch length {length!r}") from exc
    for comment in comments:
        if comment.startswith("[&&NHX:"):
            for item in comment[7:-1].split(":"):
                key, eq, value = item.partition("=")
                if eq:
                    node.add_feature(key, value)


def _parse_into(root: TreeNode, newick: str) -> None:
    """Token-driven newick reader (names may contain spaces, quotes, '+', ...)."""
    text = re.sub(r"[\n\r\t]+", "", newick).strip()
    if not text.endswith(";"):
        raise TreeError("newick string must end with ';'")
    node = root
    label = ""
    comments: List[str] = []
    depth = 0
    for token in _TOKEN.findall(text):
        if token == "(":
            if label.strip():
                raise TreeError("label before '(' in newick string")
            node = node.add_child(TreeNode())
            depth += 1
        elif token == ",":
            if depth == 0:
                raise TreeError("',' outside parentheses in newick string")
            _set_label(node, label, comments)
            label, comments = "", []
            node = node.up.add_child(TreeNode())
        elif token == ")":
            if depth == 0:
                raise TreeError("unbalanced ')' in newick string")
            _set_label(node, label, comments)
            label, comments = "", []
            node = node.up
            depth -= 1
        elif token == ";":
            _set_label(node, label, comments)
            label, comments = "", []
            break
        elif token.startswith("["):
            comments.append(token)
        else:
            label += token
    if depth != 0 or node is not root:
        raise TreeError("unbalanced '(' in newick string")
