"""TEST INFRASTRUCTURE ONLY — `python -m oracle.minir script.R [args...]`: run an R script under the minimal evaluator."""
import sys

from .interp import Interp, RError


def main(argv):
    pre = []
    while len(argv) > 2 and argv[1] == "--source":                # files sourced first (the package a `library()` call would attach)
        pre.append(argv[2])
        argv = argv[:1] + argv[3:]
    if len(argv) < 2:
        print("usage: python -m oracle.minir [--source package.R]... script.R [args...]", file=sys.stderr)
        return 2
    R = Interp(args=argv[2:])
    R.script_file = argv[1]
    try:
        for f in pre:
            R.source(f)
        R.source(argv[1])
    except RError as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1
    for desc, ok, msg, n in getattr(R, "tests", []):
        print(f"{'ok  ' if ok else 'FAIL'} {desc} ({n} expectations){': ' + msg if msg else ''}")
    return 0 if all(t[1] for t in getattr(R, "tests", [])) else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
