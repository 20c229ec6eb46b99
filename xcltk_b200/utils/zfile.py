"""Plain / gzip text input (reader behaviour of xcltk/utils/zfile.py:41-46: a name ending
in .gz or .gzip is read through gzip; bgzip files are gzip members too)."""

import gzip


def zopen(file_name, mode="rt"):
    if not file_name:
        raise OSError()
    fn = file_name.lower()
    if fn.endswith(".gz") or fn.endswith(".gzip"):
        return gzip.open(file_name, mode)
    return open(file_name, mode)
