"""Global constants (reference: xcltk/config.py:3-5)."""

APP = "xcltk"
VERSION = "0.5.2"          # reference version whose behaviour is reproduced
B200_VERSION = "0.1"
DEBUG = 0
