"""ctypes binding of ``libsr2.so`` (C ABI in ``include/sr2.h``).

The library is built in-tree by ``superrec2_b200/csrc/build.sh`` (or
``__graft_entry__.build()``).  Nothing here falls back to another
implementation: a missing library or GPU raises :class:`Sr2Error`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int16, c_int32, c_int64, c_uint32, c_uint64,
                    c_void_p)

import numpy as np

SR2_INF = 0x3FFFFFFF
SR2_KEEP_TABLES = 1
SR2_POLICY_ANY, SR2_POLICY_ALL = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SR2_LIBRARY", os.path.join(_HERE, "libsr2.so"))  # override for A/B builds


class Sr2Error(RuntimeError):
    """An ``sr2_*`` call failed; ``code`` is the negative ``sr2_status``."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libsr2 error {code}: {message}")
        self.code = code


class Timing(ctypes.Structure):
    _fields_ = [
        ("upload_ms", c_float),
        ("solve_ms", c_float),
        ("waves_ms", c_float),
        ("results_ms", c_float),
        ("n_waves", c_int32),
        ("n_launches", c_int32),
        ("n_cells", c_int64),
    ]


_I32P = POINTER(c_int32)
_I64P = POINTER(c_int64)
_I16P = POINTER(c_int16)
_U32P = POINTER(c_uint32)

#: every symbol declared in include/sr2.h: name -> (restype, argtypes)
SIGNATURES = {
    "sr2_version": (c_int, []),
    "sr2_create": (c_int, [c_int, POINTER(c_void_p)]),
    "sr2_destroy": (None, [c_void_p]),
    "sr2_last_error": (c_char_p, [c_void_p]),
    "sr2_set_stream": (c_int, [c_void_p, c_void_p]),
    "sr2_synchronize": (c_int, [c_void_p]),
    "sr2_set_policy": (c_int, [c_void_p, c_int]),
    "sr2_set_species": (c_int, [c_void_p, c_int32, _I32P, _I32P, _I32P]),
    "sr2_dtl_upload": (c_int, [c_void_p, c_int32, _I64P, _I32P, _I32P, _I32P, _I32P, c_uint32]),
    "sr2_dtl_solve": (c_int, [c_void_p]),
    "sr2_dtl_results": (c_int, [c_void_p, _I32P, _I32P, _I32P]),
    "sr2_dtl_run": (c_int, [c_void_p, c_int32, _I64P, _I32P, _I32P, _I32P, _I32P, c_uint32,
                            _I32P, _I32P, _I32P]),
    "sr2_dtl_run16": (c_int, [c_void_p, c_int32, _I64P, _I16P, _I16P, _I16P, _I32P, c_uint32,
                              _I32P, _I16P, _I16P]),
    "sr2_dtl_tables": (c_int, [c_void_p, c_int32, _I32P, _I32P]),
    "sr2_superdtl_upload": (c_int, [c_void_p, c_int32, _I64P, _I32P, _I32P, _I32P, c_int32, _U32P, _I32P,
                                    c_int64, c_uint32]),
    "sr2_superdtl_set_allowed": (c_int, [c_void_p, _I32P]),
    "sr2_superdtl_solve": (c_int, [c_void_p]),
    "sr2_superdtl_results": (c_int, [c_void_p, _I32P, _I32P, _I32P, _I32P, _U32P]),
    "sr2_superdtl_best": (c_int, [c_void_p, _I32P, _I64P, _I32P, POINTER(c_uint64)]),
    "sr2_superdtl_result_one": (c_int, [c_void_p, c_int32, _I32P, _I32P, _U32P]),
    "sr2_superdtl_run": (c_int, [c_void_p, c_int32, _I64P, _I32P, _I32P, _I32P, c_int32, _U32P, _I32P,
                                 c_uint32, _I32P, _I32P, _I32P, _I32P, _U32P]),
    "sr2_superdtl_tables": (c_int, [c_void_p, c_int32, _I32P, _I32P, _U32P, _U32P]),
    "sr2_resolutions_count": (c_int, [c_int32, _I32P, _I32P, _I64P, _I32P]),
    "sr2_resolutions_tree": (c_int, [c_int32, _I32P, _I32P, c_int64, _I32P, _I32P, _I32P]),
    "sr2_resolutions_upload": (c_int, [c_void_p, c_int32, _I32P, _I32P, _I32P, c_int32, _U32P, _I32P,
                                       c_int64, c_int64, c_uint32]),
    "sr2_last_timing": (c_int, [c_void_p, POINTER(Timing)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load ``libsr2.so`` (once) and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sr2Error(-2, f"{LIB_PATH} is missing: build it with superrec2_b200/csrc/build.sh "
                           "(there is no CPU implementation to fall back to)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def i32(array) -> np.ndarray:
    return np.ascontiguousarray(array, dtype=np.int32)


def i64(array) -> np.ndarray:
    return np.ascontiguousarray(array, dtype=np.int64)


def p32(array: np.ndarray):
    return array.ctypes.data_as(_I32P)


def p64(array: np.ndarray):
    return array.ctypes.data_as(_I64P)


class Context:
    """One ``sr2_ctx``: a device, a stream, a species tree and at most one batch."""

    def __init__(self, device: int = 0):
        self._lib = load()
        handle = c_void_p()
        code = self._lib.sr2_create(device, ctypes.byref(handle))
        if code != 0:
            message = self._lib.sr2_last_error(None).decode()
            raise Sr2Error(code, message)
        self._handle = handle
        self.device = device
        self.n_species = 0

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.sr2_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, code: int):
        if code != 0:
            raise Sr2Error(code, self._lib.sr2_last_error(self._handle).decode())

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.sr2_set_stream(self._handle, c_void_p(cuda_stream_ptr)))

    def set_policy(self, policy: int):
        """RetentionPolicy of later uploads: SR2_POLICY_ALL keeps the whole table for the host-side
        enumeration of all optimal solutions (include/sr2.h)."""
        self._check(self._lib.sr2_set_policy(self._handle, int(policy)))

    def synchronize(self):
        self._check(self._lib.sr2_synchronize(self._handle))

    def set_species(self, parent, child0, child1):
        parent, child0, child1 = i32(parent), i32(child0), i32(child1)
        self._check(self._lib.sr2_set_species(
            self._handle, len(parent), p32(parent), p32(child0), p32(child1)))
        self.n_species = len(parent)

    # ------------------------------------------------------------------ dtl
    def dtl_upload(self, node_off, left, right, leaf_species, costs, flags=0):
        node_off, left, right = i64(node_off), i32(left), i32(right)
        leaf_species, costs = i32(leaf_species), i32(costs)
        self._batch = (len(node_off) - 1, int(node_off[-1]))
        self._check(self._lib.sr2_dtl_upload(
            self._handle, len(node_off) - 1, p64(node_off), p32(left), p32(right),
            p32(leaf_species), p32(costs), flags))

    def dtl_solve(self):
        self._check(self._lib.sr2_dtl_solve(self._handle))

    def dtl_results(self):
        n_inst, n_nodes = self._batch
        min_cost = np.empty(n_inst, np.int32)
        root = np.empty(n_inst, np.int32)
        mapping = np.empty(n_nodes, np.int32)
        self._check(self._lib.sr2_dtl_results(self._handle, p32(min_cost), p32(root), p32(mapping)))
        return min_cost, root, mapping

    def dtl_run(self, node_off, left, right, leaf_species, costs, flags=0, out=None):
        """upload + solve + results with host buffers (the e2e path).  ``out`` = optional
        ``(min_cost[int32 n_inst], root[int32 n_inst], mapping[int32 n_nodes])`` arrays to fill
        (a caller that solves batch after batch reuses its result buffers)."""
        node_off, left, right = i64(node_off), i32(left), i32(right)
        leaf_species, costs = i32(leaf_species), i32(costs)
        n_inst, n_nodes = len(node_off) - 1, int(node_off[-1])
        self._batch = (n_inst, n_nodes)
        if out is not None:
            min_cost, root, mapping = out
            for arr, size in ((min_cost, n_inst), (root, n_inst), (mapping, n_nodes)):
                if arr.dtype != np.int32 or arr.shape != (size,) or not arr.flags.c_contiguous:
                    raise ValueError("out arrays must be contiguous int32 of the batch's sizes")
        else:
            min_cost = np.empty(n_inst, np.int32)
            root = np.empty(n_inst, np.int32)
            mapping = np.empty(n_nodes, np.int32)
        self._check(self._lib.sr2_dtl_run(
            self._handle, n_inst, p64(node_off), p32(left), p32(right), p32(leaf_species),
            p32(costs), flags, p32(min_cost), p32(root), p32(mapping)))
        return min_cost, root, mapping

    def dtl_run16(self, node_off, left, right, leaf_species, costs, flags=0, out=None):
        """``dtl_run`` with int16 child / leaf-species arrays in and int16 root / mapping arrays out
        (``sr2_dtl_run16``: half the bytes over PCIe; trees and species tree of at most 32,767 nodes).
        ``out`` = optional ``(min_cost[int32], root[int16], mapping[int16])`` arrays to fill."""
        node_off = i64(node_off)
        left, right, leaf_species = (np.ascontiguousarray(a, dtype=np.int16) for a in (left, right, leaf_species))
        costs = i32(costs)
        n_inst, n_nodes = len(node_off) - 1, int(node_off[-1])
        self._batch = (n_inst, n_nodes)
        if out is not None:
            min_cost, root, mapping = out
            for arr, size, dtype in ((min_cost, n_inst, np.int32), (root, n_inst, np.int16), (mapping, n_nodes, np.int16)):
                if arr.dtype != dtype or arr.shape != (size,) or not arr.flags.c_contiguous:
                    raise ValueError("out arrays must be contiguous (int32, int16, int16) of the batch's sizes")
        else:
            min_cost = np.empty(n_inst, np.int32)
            root = np.empty(n_inst, np.int16)
            mapping = np.empty(n_nodes, np.int16)
        p16 = lambda a: a.ctypes.data_as(_I16P)  # noqa: E731
        self._check(self._lib.sr2_dtl_run16(
            self._handle, n_inst, p64(node_off), p16(left), p16(right), p16(leaf_species),
            p32(costs), flags, p32(min_cost), p16(root), p16(mapping)))
        return min_cost, root, mapping

    def dtl_tables(self, instance: int, n_nodes: int):
        values = np.empty((n_nodes, self.n_species), np.int32)
        tags = np.empty((n_nodes, self.n_species, 2), np.int32)
        self._check(self._lib.sr2_dtl_tables(self._handle, instance, p32(values), p32(tags)))
        return values, tags

    # ------------------------------------------------------------- superdtl
    def superdtl_upload(self, node_off, left, right, leaf_species, leaf_bits, costs,
                        index_base=0, flags=0):
        node_off, left, right = i64(node_off), i32(left), i32(right)
        leaf_species, costs = i32(leaf_species), i32(costs)
        leaf_bits = np.ascontiguousarray(leaf_bits, dtype=np.uint32).reshape(len(left), -1)
        self._sbatch = (len(node_off) - 1, int(node_off[-1]), leaf_bits.shape[1], node_off.copy())
        self._check(self._lib.sr2_superdtl_upload(
            self._handle, len(node_off) - 1, p64(node_off), p32(left), p32(right), p32(leaf_species),
            leaf_bits.shape[1], leaf_bits.ctypes.data_as(_U32P), p32(costs), index_base, flags))

    def superdtl_set_allowed(self, node_species):
        """Restrict every object node of the uploaded batch to one species rank (-1 = any)."""
        if node_species is None:
            self._check(self._lib.sr2_superdtl_set_allowed(self._handle, None))
            return
        node_species = i32(node_species)
        if len(node_species) != self._sbatch[1]:
            raise ValueError("one entry per object node of the uploaded batch")
        self._check(self._lib.sr2_superdtl_set_allowed(self._handle, p32(node_species)))

    def superdtl_solve(self):
        self._check(self._lib.sr2_superdtl_solve(self._handle))

    def superdtl_results(self, costs_only=False):
        """(min cost, root species[, mapping, kind, synteny bits]) of every instance."""
        n_inst, n_nodes, words, _ = self._sbatch
        min_cost = np.empty(n_inst, np.int32)
        root = np.empty(n_inst, np.int32)
        if costs_only:
            self._check(self._lib.sr2_superdtl_results(self._handle, p32(min_cost), p32(root), None, None, None))
            return min_cost, root
        mapping = np.empty(n_nodes, np.int32)
        kind = np.empty(n_nodes, np.int32)
        syn = np.empty((n_nodes, words), np.uint32)
        self._check(self._lib.sr2_superdtl_results(
            self._handle, p32(min_cost), p32(root), p32(mapping), p32(kind), syn.ctypes.data_as(_U32P)))
        return min_cost, root, mapping, kind, syn

    def superdtl_best(self):
        """(min cost, global instance number, root species, packed key) of the batch."""
        cost, inst, root, key = c_int32(0), c_int64(0), c_int32(0), c_uint64(0)
        self._check(self._lib.sr2_superdtl_best(
            self._handle, ctypes.byref(cost), ctypes.byref(inst), ctypes.byref(root), ctypes.byref(key)))
        return cost.value, inst.value, root.value, key.value

    def superdtl_result_one(self, instance: int):
        _, _, words, node_off = self._sbatch
        n_nodes = node_off if isinstance(node_off, int) else int(node_off[instance + 1] - node_off[instance])
        mapping = np.empty(n_nodes, np.int32)
        kind = np.empty(n_nodes, np.int32)
        syn = np.empty((n_nodes, words), np.uint32)
        self._check(self._lib.sr2_superdtl_result_one(
            self._handle, instance, p32(mapping), p32(kind), syn.ctypes.data_as(_U32P)))
        return mapping, kind, syn

    def superdtl_run(self, node_off, left, right, leaf_species, leaf_bits, costs, flags=0):
        self.superdtl_upload(node_off, left, right, leaf_species, leaf_bits, costs, 0, flags)
        self.superdtl_solve()
        return self.superdtl_results()

    def superdtl_tables(self, instance: int):
        _, _, words, node_off = self._sbatch
        n_nodes = int(node_off[instance + 1] - node_off[instance])
        values = np.empty((n_nodes, self.n_species, 2), np.int32)
        tags = np.empty((n_nodes, self.n_species, 2, 4), np.int32)
        gain = np.empty((n_nodes, words), np.uint32)
        lca = np.empty((n_nodes, words), np.uint32)
        self._check(self._lib.sr2_superdtl_tables(
            self._handle, instance, p32(values), p32(tags), gain.ctypes.data_as(_U32P),
            lca.ctypes.data_as(_U32P)))
        return values, tags, gain, lca

    def resolutions_upload(self, child_off, child_idx, leaf_species, leaf_bits, costs, first, count,
                           flags=0):
        """Unrank resolutions [first, first+count) of a multifurcating tree into a superdtl batch."""
        child_off, child_idx = i32(child_off), i32(child_idx)
        leaf_species, costs = i32(leaf_species), i32(costs)
        leaf_bits = np.ascontiguousarray(leaf_bits, dtype=np.uint32).reshape(len(leaf_species), -1)
        _, binary_nodes = resolutions_count(child_off, child_idx)
        # uniform trees: instance i starts at node i * binary_nodes (no offsets array for a million instances)
        self._sbatch = (count, count * binary_nodes, leaf_bits.shape[1], binary_nodes)
        self._check(self._lib.sr2_resolutions_upload(
            self._handle, len(child_off) - 1, p32(child_off), p32(child_idx), p32(leaf_species),
            leaf_bits.shape[1], leaf_bits.ctypes.data_as(_U32P), p32(costs), first, count, flags))

    def timing(self) -> Timing:
        out = Timing()
        self._check(self._lib.sr2_last_timing(self._handle, ctypes.byref(out)))
        return out


def resolutions_count(child_off, child_idx):
    """(number of resolutions, nodes of one resolved tree) of a multifurcating tree (host only)."""
    lib = load()
    child_off, child_idx = i32(child_off), i32(child_idx)
    count, nodes = c_int64(0), c_int32(0)
    code = lib.sr2_resolutions_count(len(child_off) - 1, p32(child_off), p32(child_idx),
                                     ctypes.byref(count), ctypes.byref(nodes))
    if code:
        raise Sr2Error(code, "malformed multifurcating tree")
    return count.value, nodes.value


def resolutions_tree(child_off, child_idx, index):
    """(left, right, origin) of resolution ``index`` in the library's numbering (host only)."""
    lib = load()
    child_off, child_idx = i32(child_off), i32(child_idx)
    _, nodes = resolutions_count(child_off, child_idx)
    left = np.empty(nodes, np.int32)
    right = np.empty(nodes, np.int32)
    origin = np.empty(nodes, np.int32)
    code = lib.sr2_resolutions_tree(len(child_off) - 1, p32(child_off), p32(child_idx), index,
                                    p32(left), p32(right), p32(origin))
    if code:
        raise Sr2Error(code, "bad resolution index or tree")
    return left, right, origin


_contexts = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context per device (created on first use)."""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx
