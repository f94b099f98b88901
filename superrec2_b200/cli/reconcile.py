"""Compute a minimum-cost (super-)reconciliation of two trees on a B200 GPU."""
import argparse
import inspect
import json
import sys
import textwrap

from ..compute.reconciliation import reconcile_lca, reconcile_thl
from ..compute.unordered_super_reconciliation import (
    usreconcile_base_uspfs,
    usreconcile_extended_uspfs,
)
from ..model.reconciliation import (
    EdgeEvent,
    NodeEvent,
    ReconciliationInput,
    ReconciliationOutput,
    SuperReconciliationInput,
    get_default_cost,
)
from ..utils.dynamic_programming import RetentionPolicy

# Same registry shape as /root/reference/src/superrec2/cli/reconcile.py:32-40, restricted to
# the algorithms this package implements; "dtl" is an alias of the reference's "thl".
algorithms = {
    "lca": reconcile_lca,
    "thl": reconcile_thl,
    "dtl": reconcile_thl,
    "base_uspfs": usreconcile_base_uspfs,
    "superdtl": usreconcile_extended_uspfs,
}

cost_events = {
    NodeEvent.SPECIATION: ("spe", "a speciation"),
    NodeEvent.DUPLICATION: ("dup", "a duplication"),
    NodeEvent.HORIZONTAL_TRANSFER: ("hgt", "an horizontal transfer"),
    EdgeEvent.FULL_LOSS: ("floss", "a full loss"),
    EdgeEvent.SEGMENTAL_LOSS: ("sloss", "a segmental loss"),
}


def parse_cost(text):
    """A cost flag: a non-negative integer (the reference ``eval``s the text, ``:52-54``)."""
    value = int(text)
    if value < 0:
        raise argparse.ArgumentTypeError("costs must be non-negative integers")
    return value


def read_input(args):
    """Read the input JSON and inject the cost flags (reference ``:57-70``)."""
    data = json.load(args.input)
    data["costs"] = {kind: getattr(args, f"cost_{name}") for kind, (name, _) in cost_events.items()}
    if "leaf_syntenies" in data:
        return SuperReconciliationInput.from_dict(data)
    return ReconciliationInput.from_dict(data)


def call_algorithm(args, rec_input):
    """Dispatch exactly as the reference does (``:73-115``), by signature annotations."""
    algo = algorithms[args.algorithm]
    params = list(inspect.signature(algo).parameters.values())
    if params[0].annotation != type(rec_input):
        if params[0].annotation == ReconciliationInput:
            print(f"Warning: '{args.algorithm}' is not a super-reconciliation "
                  "algorithm: declared leaf syntenies will be ignored", file=sys.stderr)
        else:
            print(f"Error: '{args.algorithm}' is a super-reconciliation "
                  "algorithm: you need to provide leaf syntenies", file=sys.stderr)
            return None
    if len(params) == 1:
        output = algo(rec_input)
    elif len(params) == 2 and params[1].annotation == RetentionPolicy:
        output = algo(rec_input, getattr(RetentionPolicy, args.solutions.upper()))
    else:
        return None
    if output is None:
        results = []
    elif isinstance(output, ReconciliationOutput):
        results = [output]
    else:
        results = list(output)
    if not results:
        return None
    print("Minimum cost:", results[0].cost(), file=sys.stderr)
    return results


def dump_results(args, results):
    for result in results:
        json.dump(result.to_dict(), args.output)
        print(file=args.output)


def reconcile(args):
    rec_input = read_input(args)
    results = call_algorithm(args, rec_input)
    if results is None:
        return 1
    return dump_results(args, results)


def add_args(subparsers):
    desc = __doc__ + "\n\navailable algorithms:\n\n"
    indent = 22
    for key, impl in algorithms.items():
        first = textwrap.dedent(impl.__doc__).split("\n\n")[0]
        wrapped = textwrap.indent(textwrap.fill(" ".join(first.split()), 70), " " * indent)
        desc += "  " + key + " " * (indent - 2 - len(key)) + wrapped.strip() + "\n\n"
    parser = subparsers.add_parser(
        "reconcile", description=desc, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--input", metavar="PATH", type=argparse.FileType("r"), default="-",
                        help="path to file defining the reconciliation problem (default: read from stdin)")
    parser.add_argument("--output", metavar="PATH", type=argparse.FileType("w"), default="-",
                        help="path to file where the reconciliation results will be stored "
                             "(default: write to stdout)")
    parser.add_argument("algorithm", metavar="ALGO", choices=set(algorithms),
                        help="reconciliation algorithm to use (see list above)")
    parser.add_argument("--solutions", metavar="POLICY", default="any", choices=("any", "all"),
                        help="whether to print all minimum-cost solutions or any such solution "
                             "(choices: all, any; default: any)")
    for kind, (name, full) in cost_events.items():
        parser.add_argument(f"--cost-{name}", metavar="COST", default=get_default_cost()[kind],
                            type=parse_cost, help=f"cost of {full} event (default: %(default)s)")
    parser.set_defaults(func=reconcile)
