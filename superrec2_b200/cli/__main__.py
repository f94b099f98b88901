import argparse
import sys

from . import reconcile


def run():
    parser = argparse.ArgumentParser(prog="superrec2_b200")
    subparsers = parser.add_subparsers(required=True)
    reconcile.add_args(subparsers)
    args = parser.parse_args()
    return args.func(args)


if __name__ == "__main__":
    sys.exit(run())
