"""Stand-in for the subset of the un-vendored, unpinned third-party ``ete3``
package that the superrec2 reference uses.

TEST INFRASTRUCTURE ONLY (oracle side).  ``ete3`` is declared by the reference
(pyproject.toml:21) but is not installed here and there is no network.  It
supplies only a tree container + newick I/O; no DP arithmetic.  The surface
implemented is the one listed in SURVEY.md Appendix E, following the published
behaviour of ete3 3.1.x (``TreeNode``; ``traverse`` defaults to level-order;
``tree & name`` = first level-order match; iterating a tree yields its leaves;
newick read ``format=1``; write ``format=8/9`` with ``format_root_node`` and
NHX ``features``).  The product code in ``superrec2_b200`` never imports this.
"""
import copy as _copy
import re
from collections import deque
from hashlib import md5

DEFAULT_NAME = ""
DEFAULT_DIST = 1.0
DEFAULT_SUPPORT = 1.0
_ILLEGAL = re.compile(r"[:;(),\[\]\t\n\r=]")


class TreeError(Exception):
    pass


class NewickError(Exception):
    pass


class TreeNode:
    def __init__(self, newick=None, format=0, dist=None, support=None,
                 name=None, quoted_node_names=False):
        self.children = []
        self.up = None
        self.name = name if name is not None else DEFAULT_NAME
        self.dist = dist if dist is not None else DEFAULT_DIST
        self.support = support if support is not None else DEFAULT_SUPPORT
        self.features = {"dist", "support", "name"}
        if newick is not None:
            _read_newick(newick, self, format)

    # -- identity semantics (ete3 nodes hash/compare by identity) ----------
    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __ne__(self, other):
        return self is not other

    def __bool__(self):
        return True

    def __len__(self):
        return sum(1 for _ in self.iter_leaves())

    def __iter__(self):
        return self.iter_leaves()

    def __and__(self, value):
        value = str(value)
        for node in self.traverse():
            if node.name == value:
                return node
        raise TreeError("Node not found")

    def __contains__(self, item):
        if isinstance(item, TreeNode):
            return any(item is n for n in self.get_descendants())
        if isinstance(item, str):
            return any(n.name == item for n in self.traverse())
        return False

    def __repr__(self):
        return "Tree node '%s' (%s)" % (self.name, hex(id(self)))

    # -- structure ----------------------------------------------------------
    def is_leaf(self):
        return not self.children

    def is_root(self):
        return self.up is None

    def add_feature(self, pr_name, pr_value):
        setattr(self, pr_name, pr_value)
        self.features.add(pr_name)

    def add_features(self, **features):
        for key, value in features.items():
            self.add_feature(key, value)

    def add_child(self, child=None, name=None, dist=None, support=None):
        if child is None:
            child = self.__class__()
        if name is not None:
            child.name = name
        if dist is not None:
            child.dist = dist
        if support is not None:
            child.support = support
        self.children.append(child)
        child.up = self
        return child

    def remove_child(self, child):
        self.children.remove(child)
        child.up = None
        return child

    def detach(self):
        if self.up is not None:
            self.up.children.remove(self)
            self.up = None
        return self

    def get_sisters(self):
        if self.up is None:
            return []
        return [n for n in self.up.children if n is not self]

    def iter_ancestors(self):
        node = self
        while node.up is not None:
            node = node.up
            yield node

    def get_ancestors(self):
        return list(self.iter_ancestors())

    def get_tree_root(self):
        node = self
        while node.up is not None:
            node = node.up
        return node

    def copy(self, method="cpickle"):
        parent = self.up
        self.up = None
        try:
            new = _copy.deepcopy(self)
        finally:
            self.up = parent
        return new

    # -- traversal ----------------------------------------------------------
    def traverse(self, strategy="levelorder", is_leaf_fn=None):
        if strategy == "levelorder":
            return self._iter_levelorder()
        if strategy == "preorder":
            return self._iter_preorder()
        if strategy == "postorder":
            return self._iter_postorder()
        raise TreeError("unknown strategy %r" % (strategy,))

    def _iter_levelorder(self):
        queue = deque([self])
        while queue:
            node = queue.popleft()
            yield node
            queue.extend(node.children)

    def _iter_preorder(self):
        stack = [self]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children))

    def _iter_postorder(self):
        stack = [(self, False)]
        while stack:
            node, done = stack.pop()
            if done or not node.children:
                yield node
            else:
                stack.append((node, True))
                for child in reversed(node.children):
                    stack.append((child, False))

    def iter_descendants(self, strategy="levelorder"):
        for node in self.traverse(strategy):
            if node is not self:
                yield node

    def get_descendants(self, strategy="levelorder"):
        return list(self.iter_descendants(strategy))

    def iter_leaves(self):
        for node in self._iter_preorder():
            if not node.children:
                yield node

    def get_leaves(self):
        return list(self.iter_leaves())

    def get_leaf_names(self):
        return [leaf.name for leaf in self.iter_leaves()]

    def iter_search_nodes(self, **conditions):
        for node in self.traverse():
            if all(getattr(node, k, None) == v for k, v in conditions.items()):
                yield node

    def search_nodes(self, **conditions):
        return list(self.iter_search_nodes(**conditions))

    def get_common_ancestor(self, *target_nodes):
        if len(target_nodes) == 1 and isinstance(target_nodes[0], (list, tuple, set)):
            target_nodes = tuple(target_nodes[0])
        nodes = [self & n if isinstance(n, str) else n for n in target_nodes]
        if not nodes:
            raise TreeError("no nodes")
        if len(nodes) == 1:
            nodes = [self] + nodes

        def path(node):
            result = []
            while node is not None:
                result.append(node)
                node = node.up
            return result

        common = None
        for node in nodes:
            ancestors = path(node)
            if common is None:
                common = ancestors
            else:
                ids = {id(a) for a in ancestors}
                common = [a for a in common if id(a) in ids]
        if not common:
            raise TreeError("Nodes are not connected!")
        return common[0]

    def get_topology_id(self, attr="name"):
        """md5 over the sorted leaf-label bipartitions below this node."""
        content = {}
        for node in self._iter_postorder():
            if not node.children:
                content[id(node)] = {node}
            else:
                merged = set()
                for child in node.children:
                    merged |= content[id(child)]
                content[id(node)] = merged
        all_leaves = content[id(self)]
        edge_keys = []
        for side1 in content.values():
            side2 = all_leaves - side1
            key1 = sorted(getattr(e, attr) for e in side1)
            key2 = sorted(getattr(e, attr) for e in side2)
            edge_keys.append(sorted([key1, key2]))
        return md5(str(sorted(edge_keys)).encode("utf-8")).hexdigest()

    # -- newick output -------------------------------------------------------
    def write(self, features=None, outfile=None, format=0, is_leaf_fn=None,
              format_root_node=False, dist_formatter=None,
              support_formatter=None, name_formatter=None,
              quoted_node_names=False):
        text = _write_newick(self, features, format, format_root_node)
        if outfile is not None:
            with open(outfile, "w") as handle:
                handle.write(text)
            return None
        return text

    def get_ascii(self, *args, **kwargs):
        return self.write(format=8, format_root_node=True)


Tree = TreeNode
PhyloTree = TreeNode
PhyloNode = TreeNode


# ---------------------------------------------------------------------------
# newick reader (formats with flexible names: 0,1,8,9,100 read alike here)
# ---------------------------------------------------------------------------
_LABEL = re.compile(r"^\s*([^:\[]*?)\s*(?::\s*([^\[\s]*)\s*)?(?:\[(.*)\])?\s*$", re.S)


def _apply_label(node, text, is_leaf, fmt):
    match = _LABEL.match(text)
    if match is None:
        raise NewickError("cannot parse node label %r" % (text,))
    name, dist, comment = match.groups()
    if name:
        if not is_leaf and fmt in (0, 2):
            try:
                node.support = float(name)
            except ValueError:
                node.name = name
        else:
            node.name = name
    if dist:
        try:
            node.dist = float(dist)
        except ValueError:
            raise NewickError("bad branch length %r" % (dist,))
    if comment and comment.startswith("&&NHX:"):
        for field in comment[len("&&NHX:"):].split(":"):
            if "=" in field:
                key, value = field.split("=", 1)
                node.add_feature(key, value)


def _read_newick(newick, root, fmt):
    text = re.sub(r"[\n\r\t]+", "", newick).strip()
    if not text.endswith(";"):
        raise NewickError("newick string must end with ';'")
    text = text[:-1]
    current = root
    label = []
    depth_comment = 0
    closed = False  # whether `current` just had its children closed
    stack = []
    i = 0
    n = len(text)

    def finish(node, is_leaf):
        _apply_label(node, "".join(label), is_leaf, fmt)
        label.clear()

    node = root
    pending_leaf = True  # `node` has no children (yet)
    while i < n:
        ch = text[i]
        if depth_comment:
            label.append(ch)
            if ch == "]":
                depth_comment -= 1
        elif ch == "[":
            depth_comment += 1
            label.append(ch)
        elif ch == "(":
            if "".join(label).strip():
                raise NewickError("unexpected '(' after label")
            label.clear()
            stack.append(node)
            child = node.add_child(root.__class__())
            node = child
            pending_leaf = True
        elif ch == ",":
            finish(node, pending_leaf)
            if not stack:
                raise NewickError("',' outside parentheses")
            node = stack[-1].add_child(root.__class__())
            pending_leaf = True
        elif ch == ")":
            finish(node, pending_leaf)
            if not stack:
                raise NewickError("unbalanced ')'")
            node = stack.pop()
            pending_leaf = False
        else:
            label.append(ch)
        i += 1
    if stack:
        raise NewickError("unbalanced '('")
    finish(node, pending_leaf)
    return root


# ---------------------------------------------------------------------------
# newick writer
# ---------------------------------------------------------------------------
def _safe(value):
    return _ILLEGAL.sub("_", str(value))


def _fmt_dist(value):
    return "%0.6g" % value


def _format_node(node, is_leaf, fmt):
    if fmt in (8, 100) and fmt == 100:
        return ""
    if fmt == 8:
        return _safe(node.name)
    if fmt == 9:
        return _safe(node.name) if is_leaf else ""
    if fmt == 1:
        return "%s:%s" % (_safe(node.name), _fmt_dist(node.dist))
    if fmt == 0:
        if is_leaf:
            return "%s:%s" % (_safe(node.name), _fmt_dist(node.dist))
        return "%s:%s" % (_fmt_dist(node.support), _fmt_dist(node.dist))
    if fmt == 5:
        return ("%s:%s" % (_safe(node.name), _fmt_dist(node.dist))
                if is_leaf else ":%s" % _fmt_dist(node.dist))
    raise NewickError("unsupported newick format %r in stand-in" % (fmt,))


def _features_string(node, features):
    if features is None:
        return ""
    if features == []:
        features = sorted(node.features)
    parts = []
    for key in features:
        if hasattr(node, key):
            raw = getattr(node, key)
            if isinstance(raw, (list, tuple, set, frozenset)):
                raw = "|".join(map(str, raw))
            elif not isinstance(raw, str):
                raw = str(raw)
            parts.append("%s=%s" % (key, _safe(raw)))
    return "[&&NHX:" + ":".join(parts) + "]" if parts else ""


def _write_newick(root, features, fmt, format_root_node):
    out = []
    # iterative pre/post-order walk
    stack = [(root, False)]
    while stack:
        node, post = stack.pop()
        if post:
            out.append(")")
            if node.up is not None or format_root_node:
                out.append(_format_node(node, False, fmt))
                out.append(_features_string(node, features))
            continue
        if node is not root and node is not node.up.children[0]:
            out.append(",")
        if not node.children:
            out.append(_format_node(node, True, fmt))
            out.append(_features_string(node, features))
        else:
            out.append("(")
            stack.append((node, True))
            for child in reversed(node.children):
                stack.append((child, False))
    out.append(";")
    return "".join(out)
