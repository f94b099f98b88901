"""``python -m superrec2_b200.cli <command> ...`` -- one sub-command so far (``reconcile``)."""
import argparse

from .reconcile import add_args as add_reconcile_command


def main(argv=None) -> int:
    top = argparse.ArgumentParser(prog="superrec2_b200", description=__doc__)
    commands = top.add_subparsers(dest="command", required=True, metavar="COMMAND")
    add_reconcile_command(commands)
    options = top.parse_args(argv)
    status = options.func(options)
    return int(status or 0)


if __name__ == "__main__":
    raise SystemExit(main())
