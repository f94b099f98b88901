"""ctypes binding of the CPU oracle ``libsr2_oracle.so`` (TEST INFRASTRUCTURE ONLY).

Parity pinning: see the header of ``sr2_oracle.c`` and ``tests/test_oracle_golden.py``.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_int, c_int32, c_int64

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsr2_oracle.so")
INF = 0x3FFFFFFF
ERR_KEYERROR = -100
_P32 = POINTER(c_int32)
_P64 = POINTER(c_int64)
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "sr2_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["bash", os.path.join(HERE, "build.sh")], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(LIB)
    return _lib


def _a32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def _p(x):
    return x.ctypes.data_as(_P32) if x is not None else None


def dtl(parent, child0, child1, left, right, leaf_species, costs, strict=True, tables=True):
    """Returns dict(rc, value[G,S], tag[G,S,2], min_cost, root_species, mapping[G])."""
    lib = load()
    parent, child0, child1 = _a32(parent), _a32(child0), _a32(child1)
    left, right, leaf_species, costs = _a32(left), _a32(right), _a32(leaf_species), _a32(costs)
    S, G = len(parent), len(left)
    value = np.empty((G, S), np.int32) if tables else None
    tag = np.empty((G, S, 2), np.int32) if tables else None
    min_cost = c_int32(0)
    root = c_int32(0)
    mapping = np.full(G, -1, np.int32)
    lib.oracle_dtl.restype = c_int
    rc = lib.oracle_dtl(
        c_int(S), _p(parent), _p(child0), _p(child1), c_int(G), _p(left), _p(right), _p(leaf_species),
        _p(costs), c_int(1 if strict else 0), _p(value), _p(tag), ctypes.byref(min_cost),
        ctypes.byref(root), _p(mapping))
    return dict(rc=rc, value=value, tag=tag, min_cost=min_cost.value, root_species=root.value,
                mapping=mapping)


def dtl_batch(parent, child0, child1, node_off, left, right, leaf_species, costs, want_map=True, threads=None):
    """Independent instances over OpenMP threads; ``threads`` overrides OMP_NUM_THREADS (bench.py)."""
    lib = load()
    if threads:
        lib.oracle_set_threads(c_int(int(threads)))
    parent, child0, child1 = _a32(parent), _a32(child0), _a32(child1)
    left, right, leaf_species, costs = _a32(left), _a32(right), _a32(leaf_species), _a32(costs)
    node_off = np.ascontiguousarray(node_off, dtype=np.int64)
    B = len(node_off) - 1
    min_cost = np.empty(B, np.int32)
    root = np.empty(B, np.int32)
    mapping = np.empty(len(left), np.int32) if want_map else None
    lib.oracle_dtl_batch.restype = c_int
    rc = lib.oracle_dtl_batch(
        c_int(len(parent)), _p(parent), _p(child0), _p(child1), c_int(B),
        node_off.ctypes.data_as(_P64), _p(left), _p(right), _p(leaf_species), _p(costs),
        _p(min_cost), _p(root), _p(mapping))
    if rc:
        raise RuntimeError(f"oracle_dtl_batch failed with {rc}")
    return min_cost, root, mapping


def superdtl(parent, child0, child1, left, right, leaf_species, leaf_bits, costs, tables=True):
    """One binary resolution.  leaf_bits: uint32 [G, W].  Returns a dict with gain/lca sets,
    value[G,S,2], tag[G,S,2,4], min_cost, root_species, mapping, kind, syn[G,W]."""
    lib = load()
    parent, child0, child1 = _a32(parent), _a32(child0), _a32(child1)
    left, right, leaf_species, costs = _a32(left), _a32(right), _a32(leaf_species), _a32(costs)
    leaf_bits = np.ascontiguousarray(leaf_bits, dtype=np.uint32)
    S, G, W = len(parent), len(left), leaf_bits.shape[1]
    gain = np.zeros((G, W), np.uint32)
    lca = np.zeros((G, W), np.uint32)
    value = np.empty((G, S, 2), np.int32) if tables else None
    tag = np.empty((G, S, 2, 4), np.int32) if tables else None
    min_cost, root = c_int32(0), c_int32(0)
    mapping = np.full(G, -1, np.int32)
    kind = np.full(G, -1, np.int32)
    syn = np.zeros((G, W), np.uint32)
    u32 = POINTER(ctypes.c_uint32)
    lib.oracle_superdtl.restype = c_int
    rc = lib.oracle_superdtl(
        c_int(S), _p(parent), _p(child0), _p(child1), c_int(G), _p(left), _p(right), _p(leaf_species),
        c_int(W), leaf_bits.ctypes.data_as(u32), _p(costs), gain.ctypes.data_as(u32),
        lca.ctypes.data_as(u32), _p(value), _p(tag), ctypes.byref(min_cost), ctypes.byref(root),
        _p(mapping), _p(kind), syn.ctypes.data_as(u32))
    return dict(rc=rc, gain=gain, lca=lca, value=value, tag=tag, min_cost=min_cost.value,
                root_species=root.value, mapping=mapping, kind=kind, syn=syn)


def superdtl_batch(parent, child0, child1, node_off, left, right, leaf_species, leaf_bits, costs,
                   want_nodes=False, threads=None):
    """Independent binary instances over OpenMP threads -> (min_cost[B], root[B][, mapping, kind, syn])."""
    lib = load()
    if threads:
        lib.oracle_set_threads(c_int(int(threads)))
    parent, child0, child1 = _a32(parent), _a32(child0), _a32(child1)
    left, right, leaf_species, costs = _a32(left), _a32(right), _a32(leaf_species), _a32(costs)
    node_off = np.ascontiguousarray(node_off, dtype=np.int64)
    leaf_bits = np.ascontiguousarray(leaf_bits, dtype=np.uint32).reshape(len(left), -1)
    B, W = len(node_off) - 1, leaf_bits.shape[1]
    min_cost = np.empty(B, np.int32)
    root = np.empty(B, np.int32)
    mapping = np.empty(len(left), np.int32) if want_nodes else None
    kind = np.empty(len(left), np.int32) if want_nodes else None
    syn = np.empty((len(left), W), np.uint32) if want_nodes else None
    u32 = POINTER(ctypes.c_uint32)
    lib.oracle_superdtl_batch.restype = c_int
    rc = lib.oracle_superdtl_batch(
        c_int(len(parent)), _p(parent), _p(child0), _p(child1), c_int(B), node_off.ctypes.data_as(_P64),
        _p(left), _p(right), _p(leaf_species), c_int(W), leaf_bits.ctypes.data_as(u32), _p(costs),
        _p(min_cost), _p(root), _p(mapping), _p(kind),
        syn.ctypes.data_as(u32) if syn is not None else None)
    if rc:
        raise RuntimeError(f"oracle_superdtl_batch failed with {rc}")
    if want_nodes:
        return min_cost, root, mapping, kind, syn
    return min_cost, root
