"""`python -m dblink_b200.run <config.conf>` -- Run.main (Run.scala:27-50): parse the HOCON file, run its steps."""
import sys

from .project import Project


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m dblink_b200.run <path to config file>", file=sys.stderr)
        return 2
    proj = Project.from_file(argv[0])
    proj.write_run_txt()  # Run.scala:38-43
    res = proj.execute()
    for k, v in res.items():
        print(k, v)
    return 0


if __name__ == "__main__":
    sys.exit(main())
