"""`python -m minijl script.jl [args…]` (with tools/ on sys.path) — run a Julia-subset script like `julia script.jl args…`.
Used for the repository's Julia hosts (polymer-stats_b200/julia/*.jl) where no Julia runtime is installed:
`ccall` goes through ctypes (minijl/ffi.py), so the host drives the real libpolymc_b200.so."""
import sys

from .interp import Interp, JlError


def main(argv):
    if not argv:
        print(__doc__, file=sys.stderr)
        return 2
    it = Interp(argv=argv[1:])
    try:
        it.run_main(argv[0])
    except JlError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
