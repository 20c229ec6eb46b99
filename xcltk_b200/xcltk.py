"""Command-line dispatcher (surface of xcltk/xcltk.py:40-53 for the commands in scope)."""

import sys

from .config import APP, VERSION
from .rdr.fc.main import fc_main as rdr_basefc


def _usage(fp=sys.stdout):
    fp.write("\n"
             "Program: %s (Toolkit for XClone Preprocessing) -- B200 counting paths\n"
             "Version: %s\n"
             "\n"
             "Usage:   %s <command> [options]\n"
             "\n"
             "Commands:\n"
             "  -- RDR calculation\n"
             "     basefc           Basic feature counting (GPU).\n"
             "\n"
             "  -- BAF calculation\n"
             "     (feature-level allele counting is a Python API: xcltk_b200.baf.fc.main.afc_wrapper;\n"
             "      the `baf` pipeline around it -- cellsnp-lite, Eagle2 -- stays in the reference)\n"
             "\n"
             "  -- Others\n"
             "     -h, --help       Print this message and exit.\n"
             "     -V, --version    Print version and exit.\n"
             "\n" % (APP, VERSION, APP))


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 2:
        _usage()
        sys.exit(0)
    command = argv[1]
    if command == "basefc":
        sys.exit(rdr_basefc(argv) or 0)
    elif command in ("-h", "--help"):
        _usage()
        sys.exit(0)
    elif command in ("-V", "--version"):
        sys.stderr.write("%s\n" % VERSION)
        sys.exit(0)
    elif command in ("baf", "fixref", "convert"):
        sys.stderr.write("Error: command '%s' is outside the scope of xcltk_b200; use the reference.\n" % command)
        sys.exit(1)
    else:
        sys.stderr.write("Error: wrong command '%s'\n" % command)
        sys.exit(1)
