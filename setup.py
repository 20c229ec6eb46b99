"""Install xcltk_b200 under the reference's name: the `xcltk` console script and the `xcltk.*` import paths
(reference setup.py:58-62) resolve to this package.  The CUDA / C++ library is built in-tree for sm_100a by
xcltk_b200/build.py (nvcc, g++, zlib) before the files are collected."""

import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


class BuildWithLibrary(build_py):
    def run(self):
        from xcltk_b200 import build as b
        b.build()
        super().run()


ns = {}
exec(open(os.path.join(HERE, "xcltk_b200", "config.py")).read(), ns)

setup(
    name="xcltk-b200",
    version=ns["VERSION"],
    description="B200-native basefc / baf feature counting with the xcltk entry points",
    packages=find_packages(include=["xcltk", "xcltk_b200", "xcltk_b200.*"]),
    package_data={"xcltk_b200": ["_lib/*.so", "csrc/*", "data/*"], "": ["../include/*.h"]},
    entry_points={"console_scripts": ["xcltk = xcltk_b200.xcltk:main"]},
    install_requires=["numpy", "scipy"],
    cmdclass={"build_py": BuildWithLibrary},
)
